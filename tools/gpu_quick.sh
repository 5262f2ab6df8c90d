#!/bin/bash
# quick check of a kernel change: per variant (environment assignments, one argument each) the per-pass / tiny / edges parity
# tests, then the 4K bench (8-frame cycle, no CPU legs)
mkdir -p gpurun_out
for v in "$@"; do
  name=$(echo "$v" | tr '= /' '___')
  if [ -z "$SKIP_TESTS" ]; then
    env $v timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "${TESTS:-per_pass or tiny or edges}" > gpurun_out/pytest_q_$name.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q_$name.log
    tail -4 gpurun_out/pytest_q_$name.log | cut -c1-300
  fi
  env $v timeout 600 python bench.py --steps 96 --warmup 16 --frames 8 --no-cpu-baseline --legs none > gpurun_out/bench_q_$name.json 2> gpurun_out/bench_q_$name.err
  python - gpurun_out/bench_q_$name.json <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/frame %.4f' % j['ms_per_step'], ' '.join('%s:%.3f' % (p['pass'], p['ms']) for p in j['passes']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
