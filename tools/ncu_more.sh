#!/bin/bash
# ncu --set full captures of the other round-2 kernels (one launch each); raw CSVs land in gpurun_out/.
# (ncu matches the kernel's base name, so template instances are picked by launch order)
python tools/ncu_traffic.py --capture --tag r2_trace_rtao --kernel trace_ww_stackless_kernel --skip 2 --cmd "tools/run_configs.py rtao --reps 2"
python tools/ncu_traffic.py --capture --tag r2_trace_camera --kernel trace_ww_stackless_kernel --skip 2 --cmd "tools/frame_bench.py --reps 3"
ls -la gpurun_out/*_raw.csv
