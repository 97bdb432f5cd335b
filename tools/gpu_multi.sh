#!/bin/bash
# multi-GPU visit: N ranks of the default bench (weak scaling of a global pool dealt by cost, PCM gather over
# flm_gather_wav), config 4 (4096-utterance pool, strong scaling), and the N-rank reference-free sanity of the gather.
N=${1:-2}; TAG=${2:-r2k}; EXTRA=${3:-}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench_n${N}_exit=$?"
tail -c 400 gpurun_out/${TAG}_bench_n$N.err; head -c 600 gpurun_out/${TAG}_bench_n$N.json; echo
timeout 1200 $RUN bench.py --gpus $N --workload config4 --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-profile $EXTRA > gpurun_out/${TAG}_config4_n$N.json 2> gpurun_out/${TAG}_config4_n$N.err; echo "config4_n${N}_exit=$?"
tail -c 400 gpurun_out/${TAG}_config4_n$N.err; head -c 600 gpurun_out/${TAG}_config4_n$N.json; echo
if [ "${4:-}" = "sweep" ]; then
timeout 1200 $RUN tools/config5_sweep.py > gpurun_out/${TAG}_config5_n$N.jsonl 2> gpurun_out/${TAG}_config5_n$N.err; echo "config5_n${N}_exit=$?"
tail -c 300 gpurun_out/${TAG}_config5_n$N.err; cut -c1-260 gpurun_out/${TAG}_config5_n$N.jsonl
fi
