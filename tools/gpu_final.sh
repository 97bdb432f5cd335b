#!/bin/bash
# last check of the round: GPU tests + smoke + default bench
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench_exit=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('value %.1f e2e %.1f ms %.1f launches %d roofline %s padding %s eager %s cpu %s clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['gpu_launches'],d['roofline']['frac'],d['padding']['ratio'],d['gpu_eager_baseline']['bf16_autocast']['value'],d['cpu_baseline']['value'],d['clocks']))"
