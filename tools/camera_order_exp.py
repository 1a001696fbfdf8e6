"""Development experiment: does the order in which the camera rays of a frame reach the traversal kernel matter?
The 1080p camera batch of the bench scene is traced in row-major order (a warp = 32 x 1 pixels), in the frame call's
64 x 64 tile order, and in w x h pixel blocks inside those tiles (a warp = one 8 x 4 block, ...).  Hit records do not
depend on the order; only the time does.

  python tools/camera_order_exp.py [--blocks 8x4,4x8,16x2]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import candela_b200 as cb  # noqa: E402
from candela_b200 import scenes  # noqa: E402


def tile_order(W, H, T, bw, bh):
    """Pixel indices in tile-major order, inside a tile in bw x bh blocks (row-major blocks, row-major inside a block)."""
    ys, xs = np.divmod(np.arange(W * H, dtype=np.int64), W)
    tile = (ys // T) * ((W + T - 1) // T) + xs // T
    ix, iy = xs % T, ys % T
    block = (iy // bh) * (T // bw) + ix // bw
    inner = (iy % bh) * bw + ix % bw
    return np.lexsort((inner, block, tile))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", default="64x1,8x4,4x8,16x2,8x8,16x4")
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    v, i, m = scenes.make_s260k()
    W, H = args.width, args.height
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    ri.set_traversal_mode(2, sort_rays=0)
    d_prim = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    d_ph = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    ri.intersect_primary_device(iv, ip, W, H, d_ph.data_ptr(), d_prim.data_ptr(), stream)
    torch.cuda.synchronize()
    want = d_ph.clone()
    orders = [("row-major", np.arange(W * H))]
    for b in args.blocks.split(","):
        bw, bh = (int(x) for x in b.split("x"))
        orders.append((f"tile64 {bw}x{bh}", tile_order(W, H, 64, bw, bh)))
    for name, order in orders:
        idx = torch.from_numpy(order).cuda()
        rays = d_prim[idx].contiguous()
        hits = torch.empty_like(rays)
        ts = []
        for _ in range(args.reps + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ri.intersect_closest_device(rays.data_ptr(), W * H, hits.data_ptr(), 0, stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        same = bool(torch.equal(hits.view(torch.int32), want[idx].view(torch.int32)))
        print(json.dumps(dict(order=name, ms=round(float(np.median(ts[3:])), 4), same=same)), flush=True)
    ri.close()


if __name__ == "__main__":
    main()
