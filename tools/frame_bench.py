"""Development tool: where the time of one frame-level call goes (device time, CUDA events / ncu launch list).

  python tools/frame_bench.py [--width 1920 --height 1080 --spp 1 --bounces 1 --octant 0 --reps 20]
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
from candela_b200 import api, scenes  # noqa: E402


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--bounces", type=int, default=1)
    ap.add_argument("--octant", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fmt", type=int, default=api.FRAME_OUT_HIT16)
    args = ap.parse_args()
    W, H = args.width, args.height
    v, i, m = scenes.make_s260k()
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    stream = torch.cuda.current_stream().cuda_stream
    n_pix = W * H
    p = cb.frame_params(iv, ip, W, H, spp=args.spp, bounces=args.bounces, seed=5, out_format=args.fmt, octant_order=bool(args.octant))
    rec = 16 if args.fmt == api.FRAME_OUT_HIT16 else 32
    d_out = torch.empty(ri.frame_records(p) * rec, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ri.trace_frame_device(p, d_out.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    out = {"frame_ms": timed(lambda: ri.trace_frame_device(p, d_out.data_ptr(), 0, stream), args.reps), "rays": ri.frame_rays_traced(0)}
    # the pieces through the ordinary device calls
    d_prim = torch.empty((n_pix, 8), dtype=torch.float32, device="cuda")
    d_ph = torch.empty((n_pix, 8), dtype=torch.float32, device="cuda")
    out["primary_ms"] = timed(lambda: ri.intersect_primary_device(iv, ip, W, H, d_ph.data_ptr(), d_prim.data_ptr(), stream), args.reps)
    d_r = torch.empty((n_pix * args.spp, 8), dtype=torch.float32, device="cuda")
    n = [0]

    def gen():
        n[0] = ri.generate_rays_device(api.GEN_DIFFUSE, d_prim.data_ptr(), d_ph.data_ptr(), n_pix, d_r.data_ptr(), spp=args.spp, seed=5, bucket_octants=bool(args.octant),
                                       stream=stream)
    out["generate_ms"] = timed(gen, args.reps)
    d_h = torch.empty((n_pix * args.spp, 8), dtype=torch.float32, device="cuda")
    out["trace_ms"] = timed(lambda: ri.intersect_closest_device(d_r.data_ptr(), n[0], d_h.data_ptr(), api.IGNORE_TRANSPARENT, stream), args.reps)
    pin = cb.PinnedBuffer(len(d_out), np.uint8)
    t = torch.from_numpy(pin.array)
    out["d2h_ms"] = timed(lambda: t.copy_(d_out, non_blocking=True), args.reps)
    out["d2h_GBs"] = round(len(d_out) / out["d2h_ms"] / 1e6, 1)
    print(json.dumps(out))
    # generator variants: one frame on one stream, and frames pipelined through cndl_frame_submit (3 in flight, records to pinned host memory)
    import time
    F = 3
    pins = [cb.PinnedBuffer(ri.frame_records(p), api.HIT16_DT if args.fmt == api.FRAME_OUT_HIT16 else (api.PIXEL_DT if args.fmt == api.FRAME_OUT_PIXEL32 else api.HIT_DT))
            for _ in range(F)]
    for octant in (0, 1):
        for compact in (0, 1):
            mk = lambda seed: cb.frame_params(iv, ip, W, H, spp=args.spp, bounces=args.bounces, seed=seed, out_format=args.fmt, octant_order=bool(octant),
                                              compact_rays=bool(compact))
            pv = mk(5)
            for _ in range(3):
                ri.trace_frame_device(pv, d_out.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            one = timed(lambda: ri.trace_frame_device(pv, d_out.data_ptr(), 0, stream), args.reps)
            n_frames = 600
            for k in range(2 * F):
                ri.frame_submit(mk(k), pins[k % F].array, k % F)
            for sl in range(F):
                ri.frame_wait(sl)
            rays = 0
            t0 = time.perf_counter()
            for k in range(n_frames):
                sl = k % F
                if k >= F:
                    ri.frame_wait(sl)
                    rays += ri.frame_rays_traced(sl)
                ri.frame_submit(mk(100 + k), pins[sl].array, sl)
            for k in range(n_frames - F, n_frames):
                ri.frame_wait(k % F)
                rays += ri.frame_rays_traced(k % F)
            dt = time.perf_counter() - t0
            print(json.dumps(dict(octant=octant, compact_rays=compact, frame_ms=round(one, 4), pipelined_ms_per_frame=round(dt / n_frames * 1e3, 4),
                                  pipelined_mrays_s=round(rays / dt / 1e6, 1))), flush=True)


if __name__ == "__main__":
    main()
