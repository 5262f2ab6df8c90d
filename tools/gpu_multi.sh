#!/bin/bash
# run under gpurun --gpus N: band bit-equality test, band strong scaling, stream weak scaling
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_bands.py -m gpu -x -q > gpurun_out/pytest_bands.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bands.log; tail -5 gpurun_out/pytest_bands.log
run() { # workload gpus
  if [ "$2" = "1" ]; then timeout 900 python bench.py --workload $1 --steps 30 --warmup 5 --frames 4 --no-cpu-baseline > gpurun_out/multi_$1_n$2.json 2> gpurun_out/multi_$1_n$2.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --workload $1 --steps 30 --warmup 5 --frames 4 --no-cpu-baseline > gpurun_out/multi_$1_n$2.json 2> gpurun_out/multi_$1_n$2.err; fi
  python - gpurun_out/multi_$1_n$2.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'n_gpus', j['n_gpus'], 'ms/step %.4f' % j['ms_per_step'], 'Mpx/s %.0f' % j['value'], j['scaling'])
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
}
for w in 8k-bands 4k; do
  g=1; while [ $g -le $N ]; do run $w $g; g=$((g*2)); done
done
