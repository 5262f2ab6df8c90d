#!/bin/bash
# A/B: tests with the default build, then bench variants selected by environment variables
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; fi
tail -15 gpurun_out/pytest_gpu.log
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
  env $v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --frames 4 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  summ gpurun_out/bench_$name.json
done
