#!/bin/bash
# per-kernel durations (ncu launch list; serialised, cold cache) of one frame with the segmented and with the compacting generator
mkdir -p gpurun_out
cat > /tmp/one_frame.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import candela_b200 as cb
from candela_b200 import api, scenes
v, i, m = scenes.make_s260k()
ri = cb.RayIntersector(cb.STACKLESS); ri.AddObject(2, v, i, m); ri.BufferData(); ri.PushEntity(2); ri.BufferEntities()
W, H = 1920, 1080
iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
st = torch.cuda.current_stream().cuda_stream
for compact in (0, 1):
    for octant in (1, 0):
        p = cb.frame_params(iv, ip, W, H, seed=5, out_format=api.FRAME_OUT_HIT16, octant_order=bool(octant), compact_rays=bool(compact))
        d = torch.empty(ri.frame_records(p) * 16, dtype=torch.uint8, device='cuda')
        for _ in range(3):
            ri.trace_frame_device(p, d.data_ptr(), 0, st)
        torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"trace_|gen_|partition8|p8_|scan|frame_|primary" -c 400 --csv --log-file gpurun_out/frame_launches.csv python /tmp/one_frame.py > gpurun_out/frame_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/frame_launches.csv')) if len(r) > 10 and r[0].isdigit()]
import re
names = [(re.sub(r'cndl::|<unnamed>::|\(.*', '', r[4])[:60], float(r[-1])) for r in rows]
per = 0
for k, (n, t) in enumerate(names):
    print(f"{k:3d} {n:40s} {t/1000:8.1f} us")
PY
