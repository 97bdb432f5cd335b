#!/bin/bash
# back-batch formation: cost of one more batch (rows) in the dynamic programme of parallel.bucket_by_rows
TAG=${1:-lam}
mkdir -p gpurun_out
for CFG in "1000000 32768" "2500 32768" "1200 32768" "2500 49152" "500 32768"; do
  set -- $CFG
  timeout 600 python bench.py --steps 3 --warmup 2 --row-budget $2 --batch-overhead-rows $1 --no-cpu-baseline --no-eager-baseline --no-profile --no-e2e > gpurun_out/${TAG}_l$1_rb$2.json 2> gpurun_out/${TAG}_l$1_rb$2.err; echo "lam$1 rb$2 exit=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_l$1_rb$2.json').read().strip().splitlines()[-1]);print('overhead_rows',$1,'row_budget',$2, 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'padding', d['padding']['ratio'], 'launches', d['gpu_launches'], 'sm_mhz', d['clocks']['sm_mhz'])"
done
