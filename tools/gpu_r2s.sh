#!/bin/bash
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "activation1d or codec_decode" > gpurun_out/${TAG}_pytest_act.log 2>&1; echo "pytest_exit=$?"; tail -3 gpurun_out/${TAG}_pytest_act.log
timeout 200 python tools/codec_check.py > gpurun_out/${TAG}_codec_check.txt 2>&1; cat gpurun_out/${TAG}_codec_check.txt | cut -c1-200
