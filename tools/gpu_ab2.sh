#!/bin/bash
# A/B of the algebraic MLP-branch LayerNorm: one velocity vs the oracle + per-class timing
TAG=${1:-ab2}
mkdir -p gpurun_out
for M in 1 0; do
  FLAMED_B200_MLPLN=$M timeout 100 python tools/fused_check.py > gpurun_out/${TAG}_fused_mlpln$M.txt 2>&1; echo "fused_check[mlpln=$M] exit=$?"
  tail -13 gpurun_out/${TAG}_fused_mlpln$M.txt | cut -c1-250
done
