#!/bin/bash
# round 2, call A: GPU test suite (incl. the large configs), 4K bench, ncu source-level captures of the two big kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 --frames 8 --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_4k.json'))
    print('ms/frame', j['ms_per_step'], 'Mpix/s', j['value'], 'e2e', j['e2e']['value'])
    for p in j['passes']: print(p['pass'], p['kernel'], p['ms'], p['frac'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_4k.err').read()[-3000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_temporal_fast|k_atrous_tiled' -s 10 -c 5 -o gpurun_out/prof_r2a python bench.py --steps 2 --warmup 2 --frames 4 --no-cpu-baseline > gpurun_out/b_ncu_r2a.log 2>&1
ls -la gpurun_out/*.ncu-rep
