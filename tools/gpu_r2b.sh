#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
for MODE in 2 1; do
FLAMED_B200_FUSED_MODE=$MODE timeout 300 python tools/fused_check.py > gpurun_out/${TAG}_fused_check_mode$MODE.txt 2>&1; echo "fused_check_mode${MODE}_exit=$?"
tail -12 gpurun_out/${TAG}_fused_check_mode$MODE.txt
done
FLAMED_B200_NO_FUSED=1 timeout 300 python tools/shape_profile.py > gpurun_out/${TAG}_shape.txt 2>&1; echo "shape_exit=$?"
head -45 gpurun_out/${TAG}_shape.txt
