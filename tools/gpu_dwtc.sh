#!/bin/bash
# tensor-core depthwise conv vs the FMA kernel: one velocity vs the oracle + per-class timing
TAG=${1:-dwtc}
mkdir -p gpurun_out
for MODE in ${MODES:-tc fma}; do
  FLAMED_B200_DWCONV=$MODE timeout 100 python tools/fused_check.py > gpurun_out/${TAG}_fused_$MODE.txt 2>&1; echo "fused_check[$MODE] exit=$?"
  tail -13 gpurun_out/${TAG}_fused_$MODE.txt | cut -c1-250
done
