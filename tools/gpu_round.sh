#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list of a short bench run.
# usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref_exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --utterances 64 --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu_exit=$?"
tail -c 600 gpurun_out/${TAG}_pytest_gpu.log
