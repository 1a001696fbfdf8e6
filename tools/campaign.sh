#!/bin/bash
# Round-2 measurement campaign on ONE GPU with the current binary; everything lands in gpurun_out/ (copied into profiles/ afterwards).
#   bash tools/campaign.sh
set -u
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_ref_n1.json 2>> $O/r2_bench_n1.err
python tools/run_configs.py rtao spec_shadow bounce4k > $O/r2_configs_n1.jsonl 2> $O/r2_configs_n1.err
python tools/run_configs.py rtao --bucket 1 >> $O/r2_configs_n1.jsonl 2>> $O/r2_configs_n1.err
python tools/run_configs.py soup10m --check-build >> $O/r2_configs_n1.jsonl 2>> $O/r2_configs_n1.err
python tools/run_configs.py soup10m --sort 0 --rays 12500000 >> $O/r2_configs_n1.jsonl 2>> $O/r2_configs_n1.err
python tools/run_configs.py soup10m --builder 1 --rays 12500000 >> $O/r2_configs_n1.jsonl 2>> $O/r2_configs_n1.err
python tools/trace_bench.py > $O/r2_trace_bench.jsonl 2>&1
python tools/frame_bench.py --octant 1 > $O/r2_frame_bench.json 2>&1
python tools/ncu_traffic.py --capture > $O/r2_ncu_capture.log 2>&1
bash tools/launch_list.sh
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_build_launches_262k.csv python tools/one_build.py > /dev/null 2>&1
tail -c 400 $O/r2_configs_n1.jsonl
