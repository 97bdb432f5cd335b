#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
for PDL in 1; do
FLAMED_B200_PDL=$PDL timeout 300 python tools/fused_check.py > gpurun_out/${TAG}_fused_check_pdl$PDL.txt 2>&1; echo "fused_check_pdl${PDL}_exit=$?"
tail -13 gpurun_out/${TAG}_fused_check_pdl$PDL.txt | cut -c1-220
done
