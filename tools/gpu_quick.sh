#!/bin/bash
# Quick A/B: GPU parity tests (short timeouts: a hang must not eat the GPU budget) + per-shape profile
TAG=${1:-q}
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest_exit=$?"; tail -4 gpurun_out/${TAG}_pytest.log
BATCH=0 NFE=8 timeout 90 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b0.txt 2>&1; echo "shape0=$?"
grep -v "tapgemm_fp32_fma" gpurun_out/${TAG}_shape_b0.txt | head -${HEAD:-24}
if [ -n "$ALT" ]; then
  env $ALT BATCH=0 NFE=8 timeout 90 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b0_alt.txt 2>&1; echo "alt=$?"
  grep "dwconv\|ln_mod\|groupnorm\|snake\|total" gpurun_out/${TAG}_shape_b0_alt.txt
fi
