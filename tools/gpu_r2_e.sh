#!/bin/bash
# targeted GPU tests (pytest -k expression in $1), then the 4K bench for each remaining argument (an environment assignment per variant)
mkdir -p gpurun_out
expr="$1"; shift
if [ -n "$expr" ]; then
  timeout 1800 python -m pytest tests -m gpu -q -x -k "$expr" --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  grep -E "passed|failed|^FAILED|^ERROR|^E  +|rc=" gpurun_out/pytest_gpu.log | cut -c1-400 | head -30
fi
SKIP_TESTS=1 bash tools/gpu_ab.sh "$@"
