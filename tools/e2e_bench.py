"""Developer tool: end-to-end (host buffers) throughput of cndl_intersect_closest vs pipeline chunk count, plus raw PCIe copy rates."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import candela_b200 as cb
from candela_b200 import api, scenes

v, i, m = scenes.make_s260k()
ri = cb.RayIntersector(cb.STACKLESS)
ri.AddObject(2, v, i, m); ri.BufferData(); ri.PushEntity(2); ri.BufferEntities()
W, H = 1920, 1080
iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
nodes, tris, _ = ri.read_buffers()
drays, _ = scenes.bounce_rays(rays, hits, tris, v, seed=2)
R = len(drays)
pr, ph = cb.PinnedBuffer(R, api.RAY_DT), cb.PinnedBuffer(R, api.HIT_DT)
pr.array[:] = drays
# raw copies
a = torch.empty(R * 8, dtype=torch.float32).pin_memory(); d = torch.empty(R * 8, dtype=torch.float32, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(a, non_blocking=True)), ("d2h", lambda: a.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{name} {R*32/dt/1e9:.1f} GB/s ({dt*1e3:.3f} ms for {R*32/1e6:.0f} MB)")
for chunks in (1, 2, 4, 8, 12, 16, 24, 32):
    ri.set_tuning(4, chunks)
    for _ in range(3): ri.IntersectRays(pr.array, ignore_transparent=True, out=ph.array)
    t0 = time.perf_counter()
    for _ in range(20): ri.IntersectRays(pr.array, ignore_transparent=True, out=ph.array)
    dt = (time.perf_counter() - t0) / 20
    print(f"chunks {chunks:3d}: {dt*1e3:.3f} ms  {R/dt/1e6:.1f} Mrays/s")
