#!/bin/bash
# round 2, call B: GPU tests (all, no -x: collect achieved parity), 4K bench, optional ncu
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_4k.json'))
    print('ms/frame', j['ms_per_step'], 'Mpix/s', j['value'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    for p in j['passes']: print(p['pass'], p['kernel'], p['ms'], p['frac'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_4k.err').read()[-3000:])
PY
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU" -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-5} -o gpurun_out/prof_r2b python bench.py --steps 2 --warmup 2 --frames 4 --no-cpu-baseline > gpurun_out/b_ncu_r2b.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
