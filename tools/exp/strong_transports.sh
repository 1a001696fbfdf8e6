#!/bin/bash
# strong-scaling frame of bench.py at N GPUs with both transports:  bash tools/exp/strong_transports.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29560
for t in peer gather; do
  port=$((port+1))
  $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --strong-transport $t 2>gpurun_out/strong_${t}_n$N.err | grep "^{" > gpurun_out/strong_${t}_n$N.json
  python - "$t" "$N" <<'PY'
import json, sys
t, n = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/strong_{t}_n{n}.json"))
def find(o):
    if isinstance(o, dict):
        if o.get("scaling") == "strong": return o
        for v in o.values():
            r = find(v)
            if r: return r
    return None
s = find(d)
print(t, {k: s.get(k) for k in ("ms_per_step", "mrays_s", "final_frame_gather_ms", "transport", "frame_check")}, s.get("transport_note"))
print("e2e", d["e2e"]["value"], "value", d["value"])
PY
done
