#!/bin/bash
# round 2, last session: compile-time A/B (temporal pass at 64-thread CTAs, anti-firefly at 5 / 6 CTAs per SM) against the default
# build, then the record run of the default bench line and the launch list of the same command under ncu
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1], 'ms/frame %.4f' % j['ms_per_step'], ' '.join('%s:%.3f' % (p['pass'], p['ms']) for p in j['passes']))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
for v in base tb2 af5 af6 base; do
  lib=""; [ $v != base ] && lib="ASVGF_LIB=$PWD/build_ab/libasvgf_$v.so"
  env $lib timeout 600 python bench.py --steps 96 --warmup 16 --frames 8 --no-cpu-baseline --legs none > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  summ gpurun_out/bench_ab_$v.json
done
timeout 1200 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 3000 gpurun_out/bench_final_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 200 -c 40 --csv --log-file gpurun_out/launches_4k.csv python bench.py --steps 30 --warmup 10 --frames 8 --no-cpu-baseline --legs none > gpurun_out/b_launches.log 2>&1
tail -5 gpurun_out/launches_4k.csv | cut -c1-200
