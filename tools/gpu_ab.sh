#!/bin/bash
# A/B on the GPU: per-pass parity with each variant (environment variables select them), then the 4K bench per variant
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/frame %.4f' % j['ms_per_step'], 'e2e Mpx/s %.0f' % j['e2e']['value'], ' '.join('%s:%.3f' % (p['pass'], p['ms']) for p in j['passes']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
for v in "$@"; do
  name=$(echo "$v" | tr '= /' '___')
  if [ -z "$SKIP_TESTS" ]; then
    env $v timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "${TESTS:-per_pass or tiny or edges}" > gpurun_out/pytest_$name.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$name.log
    tail -4 gpurun_out/pytest_$name.log | cut -c1-300
  fi
  env $v timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  summ gpurun_out/bench_$name.json
done
