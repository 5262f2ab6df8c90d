#!/bin/bash
# run under gpurun --gpus N: band bit-equality test (both schedules), 8K band strong scaling with both schedules
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bands.py -m gpu -x -q > gpurun_out/pytest_bands.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bands.log; tail -5 gpurun_out/pytest_bands.log
run() { # schedule gpus
  out=gpurun_out/bands_8k_$1_n$2
  if [ "$2" = "1" ]; then timeout 900 python bench.py --workload 8k-bands --band-schedule $1 --steps 30 --warmup 5 --frames 4 --no-cpu-baseline > $out.json 2> $out.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --workload 8k-bands --band-schedule $1 --steps 30 --warmup 5 --frames 4 --no-cpu-baseline > $out.json 2> $out.err; fi
  python - $out.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'n_gpus', j['n_gpus'], 'ms/step %.4f' % j['ms_per_step'], 'Mpx/s %.0f' % j['value'], j['config']['parallelism'][:90])
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
}
run apron 1
g=2; while [ $g -le $N ]; do run apron $g; run per-pass $g; g=$((g*2)); done
