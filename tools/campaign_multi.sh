#!/bin/bash
# Round-2 measurement campaign on N GPUs of one box:  bash tools/campaign_multi.sh N
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 2> $O/r2_bench_n$N.err | grep '^{' > $O/r2_bench_n$N.json
$TR --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>> $O/r2_bench_n$N.err | grep '^{' > $O/r2_bench_ref_n$N.json
$TR --master-port 29523 tools/run_configs.py bounce4k soup10m --bvh broadcast 2> $O/r2_configs_n$N.err | grep '^{' > $O/r2_configs_n$N.jsonl
$TR --master-port 29524 tools/run_configs.py bounce4k --transport peer 2>> $O/r2_configs_n$N.err | grep '^{' >> $O/r2_configs_n$N.jsonl
$TR --master-port 29525 bench.py --gpus $N --steps 5 --warmup 3 --strong-transport gather 2>> $O/r2_bench_n$N.err | grep '^{' > $O/r2_bench_gather_n$N.json
python tools/multi_bench.py 2>> $O/r2_configs_n$N.err | grep '^{' > $O/r2_multi_bench_n$N.json
tail -c 300 $O/r2_multi_bench_n$N.json
