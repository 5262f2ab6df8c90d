#!/bin/bash
# per-instruction profile (SASS + source lines) of selected kernels; CSV pages come back, the .ncu-rep stays on the box
mkdir -p gpurun_out /tmp/ncu
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o /tmp/ncu/src_$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 32 > gpurun_out/b_src_$k.log 2>&1
  ncu -i /tmp/ncu/src_$k.ncu-rep --page source --print-source sass --csv > gpurun_out/src_$k.sass.csv 2>/dev/null
  ncu -i /tmp/ncu/src_$k.ncu-rep --page source --print-source cuda --csv > gpurun_out/src_$k.cuda.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
