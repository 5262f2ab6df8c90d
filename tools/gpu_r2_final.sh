#!/bin/bash
# round 2, final build: full GPU tests, the default bench line, the launch list of three steady-state frames and ncu --set full
# (with source) of one steady-state 4K frame (frame 20 of the 32-frame cycle)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR|^E  +[A-Za-z]*Error" gpurun_out/pytest_gpu.log | cut -c1-400
timeout 1200 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; python - <<'PY'
import json
j = json.load(open('gpurun_out/bench_final_n1.json'))
print('ms/frame %.4f value %.0f e2e %.0f roofline %.3f (%s) chain %.3f' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['kernel'], j['roofline']['chain']['frac']), j['frame_ms'])
print(' '.join('%s:%.3f' % (p['pass'], p['ms']) for p in j['passes']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 240 -c 36 --csv --log-file gpurun_out/launches_4k.csv python bench.py --steps 30 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/b_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 240 -c 12 -f -o gpurun_out/prof_frame_r02c python bench.py --steps 30 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/prof_frame_r02c.ncu-rep gpurun_out/launches_4k.csv
