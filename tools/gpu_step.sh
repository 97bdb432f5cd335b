#!/bin/bash
# Quick verification pass: GPU parity tests, GEMM microbench, short bench, optional ncu capture of the GEMM kernels.
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log
REPS=20 timeout 300 python tools/gemm_bench.py > gpurun_out/${TAG}_gemm.txt 2>&1; echo "gemm=$?"
cat gpurun_out/${TAG}_gemm.txt
BATCH=0 NFE=8 timeout 300 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b0.txt 2>&1; echo "shape0=$?"
head -16 gpurun_out/${TAG}_shape_b0.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
cat gpurun_out/${TAG}_bench.json
if [ -n "$NCU" ]; then
  PB=32 PL=1200 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc --launch-skip 3 -c 8 \
    -o gpurun_out/${TAG}_ncu_gemm python tools/kernels_probe.py > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "ncu=$?"
  ncu -i gpurun_out/${TAG}_ncu_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_gemm.raw.csv 2>/dev/null
fi
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do rm -f "$(ls -S gpurun_out/*.ncu-rep | head -1)"; done
