#!/bin/bash
# ncu --set full of one steady-state 4K frame (frame 20 of the 32-frame cycle) with the packed-tile a-trous kernel, source view included
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 240 -c 12 -f -o gpurun_out/prof_frame_r02b python bench.py --steps 30 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
