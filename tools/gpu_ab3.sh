#!/bin/bash
# A/B of an environment switch through the bench (device-resident value only): gpu_ab3.sh TAG VAR VAL_A VAL_B
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for V in "$@"; do
  env $VAR=$V timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-eager-baseline --no-e2e > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err; echo "$VAR=$V exit=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$V.json').read().strip().splitlines()[-1]);print('$VAR=$V', round(d['value'],1), round(d['ms_per_step'],1), d['gpu_launches'], d['clocks']['sm_mhz'], {k:(round(v['ms'],1), v.get('frac')) for k,v in d.get('kernels',{}).items()})"
done
