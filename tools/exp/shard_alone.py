"""How long does ONE rank's share of the 4K x 8 spp x 4 bounce frame take on a GPU by itself (no other rank, no transport)?
  python tools/exp/shard_alone.py [--shards 8]"""
import argparse, json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch
import candela_b200 as cb
from candela_b200 import api, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--shards", type=int, default=8)
ap.add_argument("--tile", type=int, default=64)
args = ap.parse_args()
v, i, m = scenes.make_s260k()
ri = cb.RayIntersector(cb.STACKLESS); ri.AddObject(2, v, i, m); ri.BufferData(); ri.PushEntity(2); ri.BufferEntities()
W, H = 3840, 2160
iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
st = torch.cuda.current_stream().cuda_stream
d = torch.empty(W * H * 32, dtype=torch.uint8, device="cuda")
out = []
for s in range(args.shards):
    p = cb.frame_params(iv, ip, W, H, spp=8, bounces=4, seed=4000, tile=args.tile, shard_index=s, shard_count=args.shards, out_format=api.FRAME_OUT_PIXEL32)
    for _ in range(2):
        ri.trace_frame_device(p, d.data_ptr(), 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ri.trace_frame_device(p, d.data_ptr(), 0, st)
    e1.record(); torch.cuda.synchronize()
    out.append((round(e0.elapsed_time(e1) / 5, 3), int(ri.frame_rays_traced(0))))
print(json.dumps(dict(shards=args.shards, tile=args.tile, ms_and_rays=out, max_ms=max(o[0] for o in out), sum_ms=round(sum(o[0] for o in out), 2))))
