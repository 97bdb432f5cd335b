#!/bin/bash
# tests (short timeouts) + bench summary
TAG=${1:-tb}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 ${PYTEST_ARGS} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest_exit=$?"; tail -3 gpurun_out/${TAG}_pytest.log
bash tools/gpu_bench.sh ${TAG}
