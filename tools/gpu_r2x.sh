#!/bin/bash
# session check: velocity parity + per-class timing of the default path, full GPU test-suite, row-budget / max-batch sweep
TAG=${1:-r2x}
mkdir -p gpurun_out
MODES=fma bash tools/gpu_dwtc.sh ${TAG}
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest_exit=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
for CFG in "32768 64" "49152 64" "65536 64" "65536 96" "98304 128"; do
  set -- $CFG
  timeout 600 python bench.py --steps 3 --warmup 2 --row-budget $1 --max-batch $2 --no-cpu-baseline --no-eager-baseline --no-profile --no-e2e > gpurun_out/${TAG}_rb$1_mb$2.json 2> gpurun_out/${TAG}_rb$1_mb$2.err; echo "rb$1 mb$2 exit=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_rb$1_mb$2.json').read().strip().splitlines()[-1]);print('rb',$1,'mb',$2, round(d['value'],1), round(d['ms_per_step'],1), d['padding']['ratio'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
