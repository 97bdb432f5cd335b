#!/bin/bash
TAG=${1:-tc2}
mkdir -p gpurun_out
timeout 120 python tools/tc2_check.py ${CASES} > gpurun_out/${TAG}_check.txt 2>&1; echo "check=$?"
cat gpurun_out/${TAG}_check.txt | tail -40
if [ -n "$BENCH" ]; then
  REPS=20 timeout 200 python tools/gemm_bench.py > gpurun_out/${TAG}_gemm.txt 2>&1; echo "gemm=$?"; cat gpurun_out/${TAG}_gemm.txt
fi
