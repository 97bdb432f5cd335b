#!/bin/bash
# smoke + bench (default, all legs) + torch-profiler view of one step
TAG=${1:-r2y}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"
tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
tail -c 400 gpurun_out/${TAG}_bench.err
timeout 600 python tools/step_profile.py > gpurun_out/${TAG}_step_profile.txt 2>&1; echo "step_profile_exit=$?"
head -45 gpurun_out/${TAG}_step_profile.txt | cut -c1-150
