#!/bin/bash
# Development: ncu passes on the 10 M-triangle configuration with in-call physical ray sorting (sort_rays = 2).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"ray_order|trace_" --csv --log-file gpurun_out/r2_soup10m_sorted_launches.csv \
    python tools/run_configs.py soup10m --rays 12500000 --check-rays 0 --sort 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"trace_ww" -c 1 -o gpurun_out/r2_trace_sorted10m \
    python tools/run_configs.py soup10m --rays 12500000 --check-rays 0 --sort 2 > /dev/null 2>&1
ncu -i gpurun_out/r2_trace_sorted10m.ncu-rep --page raw --csv > gpurun_out/r2_trace_sorted10m_raw.csv 2>/dev/null
cut -d, -f5,12- gpurun_out/r2_soup10m_sorted_launches.csv | tail -12
