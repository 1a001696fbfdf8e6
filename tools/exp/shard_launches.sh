#!/bin/bash
# ncu launch list of ONE 1/16 shard of the 4K x 8 spp x 4 bounce frame, next to the same kernels of the whole frame
mkdir -p gpurun_out
cat > /tmp/one_shard.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import candela_b200 as cb
from candela_b200 import api, scenes
v, i, m = scenes.make_s260k()
ri = cb.RayIntersector(cb.STACKLESS); ri.AddObject(2, v, i, m); ri.BufferData(); ri.PushEntity(2); ri.BufferEntities()
W, H = 3840, 2160
iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
st = torch.cuda.current_stream().cuda_stream
d = torch.empty(W * H * 32, dtype=torch.uint8, device='cuda')
for shards in (16, 1):
    p = cb.frame_params(iv, ip, W, H, spp=8, bounces=4, seed=4000, shard_index=0, shard_count=shards, out_format=api.FRAME_OUT_PIXEL32)
    for _ in range(2):
        ri.trace_frame_device(p, d.data_ptr(), 0, st)
    torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"trace_|gen_|frame_" -c 200 --csv --log-file gpurun_out/shard_launches.csv python /tmp/one_shard.py > gpurun_out/shard_launches.log 2>&1
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open('gpurun_out/shard_launches.csv')) if len(r) > 10 and r[0].isdigit()]
names = [(re.sub(r'cndl::|<unnamed>::|\(.*', '', r[4])[:50], float(r[-1]), r[8]) for r in rows]
n = len(names) // 4
a, b = names[n:2*n], names[3*n:4*n]     # second frame of each
ta = tb = 0
for (na, xa, ga), (nb, xb, gb) in zip(a, b):
    print(f"{na:50s} shard/16 {xa/1000:9.1f} us   whole {xb/1000:9.1f} us   ratio x16 {16*xa/xb:5.2f}  grid {ga} {gb}")
    ta += xa; tb += xb
print("sum", ta/1000, tb/1000, 16*ta/tb)
PY
