#!/bin/bash
# launch list (three steady-state frames) + ncu --set full of one steady-state frame of the production path at 4K:
# 12 kernels per frame; frame 20 of the 32-frame cycle (history long everywhere but at the disocclusions)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 240 -c 36 --csv --log-file gpurun_out/launches_4k.csv python bench.py --steps 30 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/b_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 240 -c 12 -f -o gpurun_out/prof_frame_r02 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_4k.csv
