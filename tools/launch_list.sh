#!/bin/bash
# ncu launch list of the bench command (per-launch gpu__time_duration; cold-cache, serialised): the kernel's SHARE of a step.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"trace_|gen_|partition8|scan_|frame_|primary" -c 400 --csv \
    --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --passes 2 --no-strong --build-reps 0 > gpurun_out/r2_launches_bench.log 2>&1
