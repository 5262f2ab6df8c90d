#!/bin/bash
# under gpurun --gpus N: cross-device band-set test, then the default bench under torchrun with the legs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt; nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_bandset_gpu.py -m gpu -q -k "across or identical_to_the_single" > gpurun_out/pytest_bandset_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bandset_n$N.log; tail -4 gpurun_out/pytest_bandset_n$N.log | cut -c1-300
GPUS=$N STEPS=${STEPS:-60} SKIP_TESTS=1 BENCH_ARGS="--no-cpu-baseline --frames 8" bash tools/gpu_r2_c.sh
