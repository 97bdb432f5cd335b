#!/bin/bash
# back-batch formation: wave-aware cost and the cost of one more batch (rows) in parallel.bucket_by_rows
# usage: gpu_lam.sh TAG "overhead budget wave" ["overhead budget wave" ...]
TAG=$1; shift
mkdir -p gpurun_out
for CFG in "$@"; do
  set -- $CFG
  timeout 600 python bench.py --steps 3 --warmup 2 --row-budget $2 --batch-overhead-rows $1 --wave-rows $3 --no-cpu-baseline --no-eager-baseline --no-profile --no-e2e > gpurun_out/${TAG}_l$1_rb$2_w$3.json 2> gpurun_out/${TAG}_l$1_rb$2_w$3.err; echo "lam$1 rb$2 w$3 exit=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_l$1_rb$2_w$3.json').read().strip().splitlines()[-1]);print('overhead_rows',$1,'row_budget',$2,'wave_rows',$3,'value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'padding', d['padding']['ratio'], 'launches', d['gpu_launches'], 'sm_mhz', d['clocks']['sm_mhz'])"
done
