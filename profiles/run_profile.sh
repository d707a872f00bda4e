#!/bin/bash
# Profiling recipe (run under gpurun on one B200).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv
python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu.log 2>&1
# the two render kernels, full set
ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 2 \
    -o gpurun_out/prof_render python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
