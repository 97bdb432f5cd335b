#!/bin/bash
# Quick A/B: denoiser parity tests + per-shape profile (optionally with env toggles given as arguments "VAR=1 VAR2=x")
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "velocity or denoiser or bf16" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest_exit=$?"; tail -2 gpurun_out/${TAG}_pytest.log
BATCH=0 NFE=8 timeout 300 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b0.txt 2>&1; echo "shape0=$?"
grep -v "tapgemm_fp32_fma" gpurun_out/${TAG}_shape_b0.txt | head -24
if [ -n "$ALT" ]; then
  env $ALT BATCH=0 NFE=8 timeout 300 python tools/shape_profile.py > gpurun_out/${TAG}_shape_b0_alt.txt 2>&1; echo "alt=$?"
  grep "dwconv\|ln_mod\|groupnorm\|snake\|total" gpurun_out/${TAG}_shape_b0_alt.txt
fi
