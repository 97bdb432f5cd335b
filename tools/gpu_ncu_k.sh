#!/bin/bash
# ncu --set full of one kernel of a single velocity evaluation: gpu_ncu_k.sh TAG KERNEL_REGEX [SKIP] [ENV=VAL ...]
TAG=$1; KER=$2; SKIP=${3:-6}; shift 3
mkdir -p gpurun_out
env "$@" timeout 120 python tools/one_velocity.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KER -s $SKIP -c 1 -o gpurun_out/${TAG} python tools/one_velocity.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}.raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}.source.csv 2>/dev/null
rm -f gpurun_out/${TAG}.ncu-rep
python tools/ncu_key_metrics.py gpurun_out/${TAG}.raw.csv
