#!/bin/bash
# full GPU test suite + smoke + bench (default) + torch-profiler view of one step
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"
tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
timeout 600 python tools/step_profile.py > gpurun_out/${TAG}_step_profile.txt 2>&1; echo "step_profile_exit=$?"
head -40 gpurun_out/${TAG}_step_profile.txt
