#!/bin/bash
# round 2: full GPU test suite + default bench run (with the band / view-stream legs)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
[ -n "$SKIP_TESTS" ] || { timeout 2400 python -m pytest tests -m gpu -q --durations=8 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR|^E  +[A-Za-z]*Error" gpurun_out/pytest_gpu.log | cut -c1-400; }
N=${GPUS:-1}
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --steps ${STEPS:-100} --warmup 10 ${BENCH_ARGS} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 ${BENCH_ARGS} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads([l for l in open('gpurun_out/bench_n%s.json' % n) if l.startswith('{')][-1])
    print('n_gpus', j['n_gpus'], 'ms/frame %.4f' % j['ms_per_step'], 'Mpix/s %.0f' % j['value'], 'e2e %.0f' % j['e2e']['value'], 'launches', j['gpu_launches'])
    print(' '.join('%s:%.3f' % (p['pass'], p['ms']) for p in j['passes']))
    print('bands', json.dumps(j.get('bands'))[:1200])
    print('view_streams', json.dumps(j.get('view_streams'))[:600])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_n%s.err' % n).read()[-3000:])
PY
