#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
for MODE in 2 1; do
FLAMED_B200_FUSED_MODE=$MODE timeout 300 python tools/fused_check.py > gpurun_out/${TAG}_fused_check_mode$MODE.txt 2>&1; echo "fused_check_mode${MODE}_exit=$?"
tail -12 gpurun_out/${TAG}_fused_check_mode$MODE.txt
done
