#!/bin/bash
# the other BASELINE configurations on the final build (one GPU): 4K with five a-trous iterations, 1080p, 720p
mkdir -p gpurun_out
for w in 4k-n5 1080p 720p; do
  timeout 600 python bench.py --workload $w --frames 16 --no-cpu-baseline --legs none > gpurun_out/bench_cfg_$w.json 2> gpurun_out/bench_cfg_$w.err
  python - gpurun_out/bench_cfg_$w.json <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/frame %.4f value %.0f' % (j['ms_per_step'], j['value']), j['frame_ms'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
