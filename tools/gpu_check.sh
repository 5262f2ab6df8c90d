#!/bin/bash
# one gpurun call: GPU tests, 4K bench, launch list (+ optional full ncu with NCU_FULL=1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_4k.json'))
    print('ms/frame', j['ms_per_step'], 'Mpix/s', j['value'], 'e2e', j['e2e']['value'], 'chain frac', j['roofline']['chain']['frac'])
    for p in j['passes']: print(p['pass'], p['kernel'], p['ms'], p['frac'])
    print(j.get('cpu_baseline'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_4k.err').read()[-3000:])
PY
if [ -n "$NCU_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 36 -c 12 -o gpurun_out/prof_frame python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
fi
