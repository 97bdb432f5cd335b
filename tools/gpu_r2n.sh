#!/bin/bash
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -rA -k "attention or prompt_side or prior_decoders" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest_new_exit=$?"
grep -E "passed|failed|error|rel-L2|codes differing" gpurun_out/${TAG}_pytest_new.log | tail -15
grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/${TAG}_pytest_new.log | head -20
timeout 200 python tools/sdpa_bench.py > gpurun_out/${TAG}_sdpa.txt 2>&1; echo "sdpa_exit=$?"; cat gpurun_out/${TAG}_sdpa.txt | tail -6
