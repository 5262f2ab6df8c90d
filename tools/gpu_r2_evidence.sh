#!/bin/bash
# round 2 evidence run: full GPU tests, racecheck of the shared-memory kernels, launch list + ncu --set full of one frame
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || { timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR|^E  +[A-Za-z]*Error" gpurun_out/pytest_gpu.log | cut -c1-400; }
# racecheck: the tiled a-trous (all strides, pair form), anti-firefly, histogram (mbarrier + bulk copy) on small frames
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_parity_gpu.py tests/test_post_passes.py -m gpu -q -x -k "tiny or edges or post_passes" > gpurun_out/racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck.log; tail -8 gpurun_out/racecheck.log | cut -c1-300
# every launch of two frames with its device time (bench per-pass leg excluded: production path only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 200 -c 40 --csv --log-file gpurun_out/launches_4k.csv python bench.py --steps 30 --warmup 10 --frames 8 --no-cpu-baseline --legs none > gpurun_out/b_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 200 -c 12 -o gpurun_out/prof_frame_r02 python bench.py --steps 30 --warmup 10 --frames 8 --no-cpu-baseline --legs none > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_4k.csv
