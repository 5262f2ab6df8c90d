#!/bin/bash
# targeted GPU tests (expression in $1) and env-selected bench variants (remaining args)
mkdir -p gpurun_out
expr="$1"; shift
if [ -n "$expr" ]; then
  timeout 2400 python -m pytest tests -m gpu -q -k "$expr" --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  grep -E "passed|failed|^FAILED|^ERROR|^E  +[A-Za-z]*Error" gpurun_out/pytest_gpu.log | cut -c1-500
fi
SKIP_TESTS=1 bash tools/gpu_ab.sh "$@"
