#!/bin/bash
# Information pass: per-shape device times, stage times, glue profile, bf16-residual error, ncu full captures.
TAG=${1:-r01b}
mkdir -p gpurun_out
BATCH=3 NFE=8 timeout 300 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b3.txt 2>&1; echo "shape3=$?"
BATCH=0 NFE=8 timeout 300 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b0.txt 2>&1; echo "shape0=$?"
NFE=128 timeout 300 python tools/stage_times.py > gpurun_out/${TAG}_stages.txt 2>&1; echo "stages=$?"
timeout 300 python tools/glue_profile.py > gpurun_out/${TAG}_glue.txt 2>&1; echo "glue=$?"
REPS=20 timeout 300 python tools/gemm_bench.py > gpurun_out/${TAG}_gemm.txt 2>&1; echo "gemm=$?"
FLAMED_B200_RESIDUAL=bf16 timeout 600 python -m pytest tests -m gpu -x -q -s -k "bf16 or 128 or velocity" > gpurun_out/${TAG}_pytest_resbf16.log 2>&1; echo "resbf16=$?"
# one small --set full capture per kernel family (2 launches each; reports stay a few MB)
for K in "tapgemm_tc_kernel<256, 5>" "tapgemm_tc_kernel<256, 1>" ln_mod_kernel dwconv_kernel gn_convnext_kernel act1d_kernel; do
  N=$(echo "$K" | tr -c 'a-zA-Z0-9' '_')
  PB=32 PL=1200 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 2 -c 2 \
    -o gpurun_out/${TAG}_ncu_$N python tools/kernels_probe.py > gpurun_out/${TAG}_ncu_$N.log 2>&1; echo "ncu $N=$?"
  ncu -i gpurun_out/${TAG}_ncu_$N.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$N.raw.csv 2>/dev/null
done
du -sh gpurun_out; ls -la gpurun_out
# never exceed the 64 MiB copy-back limit: drop the largest reports first
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do rm -f "$(ls -S gpurun_out/*.ncu-rep | head -1)"; done
