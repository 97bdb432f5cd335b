#!/bin/bash
N=${1:-8}; TAG=${2:-r2z}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 1200 $RUN tools/config5_sweep.py > gpurun_out/${TAG}_config5_n$N.jsonl 2> gpurun_out/${TAG}_config5_n$N.err; echo "config5_n${N}_exit=$?"
tail -c 300 gpurun_out/${TAG}_config5_n$N.err; cut -c1-200 gpurun_out/${TAG}_config5_n$N.jsonl
