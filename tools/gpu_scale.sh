#!/bin/bash
# default bench at N ranks (weak scaling of one dealt pool): gpu_scale.sh N TAG
N=${1:-2}; TAG=${2:-scale}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-profile > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench_n${N}_exit=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]);print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],1),'padding',d['padding']['ratio'])"
