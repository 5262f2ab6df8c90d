#!/bin/bash
# ncu --set full (with source) of the four a-trous launches of one steady-state 4K frame, persistent pipelined variant
mkdir -p gpurun_out
ASVGF_ATROUS_PIPE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_atrous -s 80 -c 4 -f -o gpurun_out/prof_pipe python bench.py --steps 30 --warmup 10 --no-cpu-baseline --legs none > gpurun_out/b_ncu_pipe.log 2>&1
ls -la gpurun_out/prof_pipe.ncu-rep
