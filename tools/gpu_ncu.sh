#!/bin/bash
# ncu evidence for profiles/: launch list of two frames, full capture of one frame and of the post kernels.
# The .ncu-rep files stay on the box (size); their raw pages come back as CSV.
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 36 -c 24 --csv --log-file gpurun_out/launches_4k.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_ -s 36 -c 12 -o /tmp/ncu/prof_frame python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
if [ -z "$SKIP_POST" ]; then
timeout 900 ncu --set full --clock-control none -k regex:'k_lum|k_checker' -s 8 -c 4 -o /tmp/ncu/prof_post python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
fi
for f in prof_frame prof_post; do [ -f /tmp/ncu/$f.ncu-rep ] && ncu -i /tmp/ncu/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out /tmp/ncu
