#!/bin/bash
# round 2, last session: full GPU tests on the build with programmatic dependent launch, then A/B of the 4K frame with
# ASVGF_PDL=1 / 0 back to back (twice, interleaved, so that box drift shows)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR|^E  +[A-Za-z]*Error" gpurun_out/pytest_gpu.log | cut -c1-400; }
summ() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/frame %.4f' % j['ms_per_step'], 'p50 %.4f' % j.get('frame_ms', {}).get('p50', 0), ' '.join('%s:%.3f' % (p['pass'], p['ms']) for p in j['passes']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
for round in 1 2; do
  for v in 1 0; do
    ASVGF_PDL=$v timeout 600 python bench.py --steps 96 --warmup 16 --frames 8 --no-cpu-baseline --legs none > gpurun_out/bench_pdl${v}_$round.json 2> gpurun_out/bench_pdl${v}_$round.err
    summ gpurun_out/bench_pdl${v}_$round.json
  done
done
