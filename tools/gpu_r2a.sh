#!/bin/bash
# Round-2 first GPU visit: new + old parity tests (without -x: see every failure), smoke, default bench.
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rA > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -5
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.log | head -30
tail -3 gpurun_out/${TAG}_smoke.log
tail -c 1500 gpurun_out/${TAG}_bench.err
head -c 3000 gpurun_out/${TAG}_bench.json
