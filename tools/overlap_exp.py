"""Experiment: do the camera pass (coherent) and the diffuse pass (incoherent) of two frames overlap usefully when each persistent
kernel takes only part of every SM?  Sequential at 8 CTAs/SM vs concurrent on two streams at a + b CTAs/SM."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import candela_b200 as cb  # noqa: E402
from candela_b200 import api, scenes  # noqa: E402

v, i, m = scenes.make_s260k()
W, H = 1920, 1080
iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)


def make():
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    return ri


ra, rb = make(), make()   # two contexts on the same GPU: independent knobs
s0 = torch.cuda.current_stream().cuda_stream
d_prim = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
d_ph = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
ra.intersect_primary_device(iv, ip, W, H, d_ph.data_ptr(), d_prim.data_ptr(), s0)
d_r = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
n = ra.generate_rays_device(api.GEN_DIFFUSE, d_prim.data_ptr(), d_ph.data_ptr(), W * H, d_r.data_ptr(), seed=1000, bucket_octants=True, stream=s0)
d_h = torch.empty((n, 8), dtype=torch.float32, device="cuda")
d_ph2 = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
reps = 20


def run(a_bps, b_bps, concurrent):
    ra.set_tuning(0, a_bps)
    rb.set_tuning(0, b_bps)
    for _ in range(2):
        ra.intersect_closest_device(d_r.data_ptr(), n, d_h.data_ptr(), api.IGNORE_TRANSPARENT, sa.cuda_stream)
        rb.intersect_closest_device(d_prim.data_ptr(), W * H, d_ph2.data_ptr(), 0, (sb if concurrent else sa).cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_stream(torch.cuda.current_stream())
    sb.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        ra.intersect_closest_device(d_r.data_ptr(), n, d_h.data_ptr(), api.IGNORE_TRANSPARENT, sa.cuda_stream)
        rb.intersect_closest_device(d_prim.data_ptr(), W * H, d_ph2.data_ptr(), 0, (sb if concurrent else sa).cuda_stream)
    torch.cuda.current_stream().wait_stream(sa)
    torch.cuda.current_stream().wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("sequential 8 + 8 CTAs/SM: %.4f ms per (diffuse + camera) pair" % run(8, 8, False))
for a, b in ((4, 4), (5, 3), (6, 2), (3, 5), (8, 8), (6, 6)):
    print("concurrent %d + %d CTAs/SM: %.4f ms" % (a, b, run(a, b, True)))
