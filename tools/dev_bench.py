"""Developer micro-benchmark (not the driver's bench.py): times traversal variants on the S260k
diffuse batch with device-resident inputs and checks each against the oracle.  Usage on the GPU box:
    python tools/dev_bench.py [--rays-scale 1] [--check]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import candela_b200 as cb  # noqa: E402
from candela_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--modes", default="0,1,2")
    ap.add_argument("--knobs", default="", help="semicolon-separated knob lists for mode 2, in knob-id order: bps,leaf,idle,variant,chunks,stack_leaf,hot_nodes,block_threads")
    ap.add_argument("--fmt", default="stackless")
    ap.add_argument("--sort", type=int, default=0, help="1: bucket the rays by direction octant inside the call (cndl_set_traversal_mode sort_rays)")
    ap.add_argument("--gpu-builder", type=int, default=-1, help="-1: oracle-built buffers; 0: GPU SAH; 1: GPU LBVH")
    ap.add_argument("--presort", type=int, default=0, help="host-side Morton sort of the diffuse rays (experiment): bits per axis")
    ap.add_argument("--presort-mode", type=int, default=0, help="0: origin cell major, octant minor; 1: octant major; 2: direction cell (16x16 on the octahedron) major")
    args = ap.parse_args()
    from oracle import binding as ob
    fmt = ob.STACKLESS if args.fmt == "stackless" else ob.STACK
    v, i, m = scenes.make_s260k()
    t0 = time.time()
    b = ob.build(fmt, v, i, m)
    print(f"oracle build {1e3 * (time.time() - t0):.0f} ms, nodes {len(b.nodes)}", flush=True)
    ri = cb.RayIntersector(fmt)
    if args.gpu_builder >= 0:
        ri.AddObject(2, v, i, m, builder=args.gpu_builder)
        print(f"gpu build {ri.last_build_ms:.3f} ms")
        class _B: pass
        sah = b
        b = _B()
        b.nodes, b.tris, _ = ri.read_buffers()
    else:
        ri.AddPrebuiltObject(2, b.nodes, b.tris, v)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    W, H = args.width, args.height
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
    drays, _ = scenes.bounce_rays(rays, hits, b.tris, v, seed=2)
    if args.presort:
        o = drays["o"]; d = drays["d"]
        lo, hi = o.min(0), o.max(0)
        q = np.clip(((o - lo) / (hi - lo) * (2 ** args.presort - 1)).astype(np.uint64), 0, 2 ** args.presort - 1)
        def spread(x):
            r = np.zeros_like(x)
            for b in range(args.presort):
                r |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
            return r
        key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
        octant = ((d[:, 0] > 0).astype(np.uint64) | ((d[:, 1] > 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] > 0).astype(np.uint64) << np.uint64(2)))
        if args.presort_mode >= 100:   # global stable buckets by direction class
            ad = np.abs(d); dom = np.argmax(ad, axis=1).astype(np.uint64)
            sgn_dom = (np.take_along_axis(d, dom[:, None].astype(np.int64), 1)[:, 0] > 0).astype(np.uint64)
            cls = {100: (d[:, 0] > 0).astype(np.uint64),                       # sign of x
                   101: (d[:, 1] > 0).astype(np.uint64),                       # sign of y
                   102: dom * np.uint64(2) + sgn_dom,                          # dominant axis + sign (6)
                   103: octant * np.uint64(3) + dom,                           # octant x dominant axis (24)
                   104: (octant & np.uint64(3)),                               # signs of x,y (4)
                   105: (octant & np.uint64(5)),                               # signs of x,z (4)
                   }[args.presort_mode]
            key = cls
        elif args.presort_mode >= 10:   # chunk-local octant grouping: chunks of 2^mode consecutive rays, stable
            key = ((np.arange(len(d), dtype=np.uint64) >> np.uint64(args.presort_mode)) << np.uint64(3)) | octant
        elif args.presort_mode == 0:
            key = (key << np.uint64(3)) | octant
        elif args.presort_mode == 1:
            key = key | (octant << np.uint64(3 * args.presort))
        else:
            n1 = d / np.abs(d).sum(1, keepdims=True)
            u, w = n1[:, 0].copy(), n1[:, 1].copy()
            neg = n1[:, 2] < 0
            u2 = (1 - np.abs(w)) * np.sign(u + 1e-30); w2 = (1 - np.abs(n1[:, 0])) * np.sign(w + 1e-30)
            u[neg], w[neg] = u2[neg], w2[neg]
            cu = np.clip(((u * 0.5 + 0.5) * 16).astype(np.uint64), 0, 15); cw = np.clip(((w * 0.5 + 0.5) * 16).astype(np.uint64), 0, 15)
            key = key | ((cu * np.uint64(16) + cw) << np.uint64(3 * args.presort))
        drays = drays[np.argsort(key, kind="stable")]
    R = len(drays)
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(b.nodes))
    ref = None
    t0 = time.time()
    ref, cnt = ob.trace(fmt, ob.CLOSEST, b.nodes, b.tris, v, ents, drays, nthreads=ob.hardware_threads())
    cpu_s = time.time() - t0
    nn, nt = cnt["node_iters"] / R, cnt["tri_tests"] / R
    bray = nn * (32 if fmt == ob.STACKLESS else 64) + nt * 64
    print(f"rays {R} Nn {nn:.2f} Nt {nt:.2f} B_ray {bray:.0f} cpu {R / cpu_s / 1e6:.2f} Mrays/s on {ob.hardware_threads()} threads", flush=True)
    d_rays = torch.from_numpy(drays.view(np.float32).reshape(-1, 8)).cuda()
    d_prim = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.empty((max(R, len(rays)), 8), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for name, dr, n in (("diffuse", d_rays, R), ("primary", d_prim, len(rays))):
        variants = []
        for mode in [int(x) for x in args.modes.split(",")]:
            if mode == 2 and args.knobs:
                variants += [(2, tuple(int(v) for v in k.split(","))) for k in args.knobs.split(";")]
            else:
                variants.append((mode, None))
        for mode, knobs in variants:
            ri.set_traversal_mode(mode, bool(args.sort))
            if knobs:
                for kid, val in enumerate(knobs):
                    ri.set_tuning(kid, val)
                if len(knobs) > 6:
                    ri.BufferData()     # the staged-node count takes effect at commit
            for _ in range(3):
                ri.intersect_closest_device(dr.data_ptr(), n, d_hits.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ri.intersect_closest_device(dr.data_ptr(), n, d_hits.data_ptr(), 0, stream)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            line = dict(batch=name, mode=mode, sort=args.sort, knobs=knobs, fmt=args.fmt, ms=round(ms, 4), mrays=round(n / ms / 1e3, 1))
            if name == "diffuse":
                line["roofline_frac"] = round(n / (ms * 1e-3) * bray / 6550.1e9, 4)
                if args.check:
                    got = d_hits[:R].cpu().numpy().view(ob.HIT_DT).reshape(-1)
                    line["bit_identical"] = bool(got.tobytes() == ref.tobytes())
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
