#!/bin/bash
TAG=${1:-r2l}
mkdir -p gpurun_out
for RB in 16384 20480 24576 40960; do
timeout 600 python bench.py --steps 3 --warmup 2 --row-budget $RB --no-cpu-baseline --no-eager-baseline --no-profile --no-e2e > gpurun_out/${TAG}_rb$RB.json 2> gpurun_out/${TAG}_rb$RB.err; echo "rb$RB exit=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_rb$RB.json').read().strip().splitlines()[-1]);print($RB, round(d['value'],1), round(d['ms_per_step'],1), d['padding']['ratio'], round(d['padding']['value_on_padded_audio'],1), d['gpu_launches'])"
done
