#!/bin/bash
# strong-scaling frame of bench.py at N GPUs for --strong-split 1 2 3 4:  bash tools/exp/strong_split.sh N
N=${1:-1}
mkdir -p gpurun_out
port=29580
for sp in ${SPLITS:-1 2 3 4}; do
  port=$((port+1))
  if [ "$N" = "1" ]; then
    python bench.py --steps 3 --warmup 3 --strong-split $sp 2>gpurun_out/split_${sp}_n$N.err | grep "^{" > gpurun_out/split_${sp}_n$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 3 --warmup 3 --strong-split $sp 2>gpurun_out/split_${sp}_n$N.err | grep "^{" > gpurun_out/split_${sp}_n$N.json
  fi
  python - "$sp" "$N" <<'PY'
import json, sys
sp, n = sys.argv[1], sys.argv[2]
try:
    s = json.load(open(f"gpurun_out/split_{sp}_n{n}.json"))["strong_scaling"]
    print("split", sp, {k: s.get(k) for k in ("ms_per_step", "mrays_s", "final_frame_gather_ms", "transport", "sub_shards_per_rank")}, s["frame_check"]["rays_match_counters"], s.get("parity"))
except Exception as e:
    print("split", sp, "failed", e); print(open(f"gpurun_out/split_{sp}_n{n}.err").read()[-1500:])
PY
done
