#!/bin/bash
# one gpurun call: tests, bench, launch list, full ncu capture of one frame (+ the post kernels)
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 36 --csv --log-file gpurun_out/launches_4k.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 24 -c 12 -o gpurun_out/prof_frame python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_lum|k_checker' -c 4 -o gpurun_out/prof_post python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_4k.json
