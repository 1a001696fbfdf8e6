"""Development tool: traversal time of the generator-made diffuse batch (the bench.py batch) for node formats, batch orders and
tuning knobs; every variant is checked bit for bit against the oracle (hit records of the plain order, permuted).

  python tools/trace_bench.py --fmt stackless,stack --octant 0,1 --knobs "8,14,10,0;8,10,10,0"
knob list order: blocks_per_sm, leaf_threshold, idle_threshold, variant[, host_chunks, stack_leaf_threshold, hot_nodes, block_threads]
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

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="stackless,stack")
    ap.add_argument("--octant", default="0,1")
    ap.add_argument("--knobs", default="")
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--camera", action="store_true", help="time the camera pass (cndl_intersect_primary_device) instead of the diffuse batch")
    args = ap.parse_args()
    from oracle import binding as ob
    v, i, m = scenes.make_s260k()
    W, H = args.width, args.height
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for fmt_name in args.fmt.split(","):
        fmt, ofmt, node_bytes = (cb.STACKLESS, ob.STACKLESS, 32) if fmt_name == "stackless" else (cb.STACK, ob.STACK, 64)
        ri = cb.RayIntersector(fmt)
        ri.AddObject(2, v, i, m)
        ri.BufferData()
        ri.PushEntity(2)
        ri.BufferEntities()
        nodes, tris, _ = ri.read_buffers()
        ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
        d_prim = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
        d_ph = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
        ri.intersect_primary_device(iv, ip, W, H, d_ph.data_ptr(), d_prim.data_ptr(), stream)
        if args.camera:
            knob_sets = [None] + [[int(x) for x in k.split(",")] for k in args.knobs.split(";") if k]
            for ks in knob_sets:
                if ks:
                    for kid, val in enumerate(ks):
                        ri.set_tuning(kid, val)
                ts = []
                for _ in range(args.reps + 3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ri.intersect_primary_device(iv, ip, W, H, d_ph.data_ptr(), d_prim.data_ptr(), stream)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                print(json.dumps(dict(fmt=fmt_name, camera=True, knobs=ks, ms=round(float(np.median(ts[3:])), 4))), flush=True)
            ri.close()
            continue
        want = None
        for octant in [int(x) for x in args.octant.split(",")]:
            d_r = torch.empty((W * H * args.spp, 8), dtype=torch.float32, device="cuda")
            d_ids = torch.empty(W * H * args.spp, dtype=torch.int32, device="cuda")
            n = ri.generate_rays_device(api.GEN_DIFFUSE, d_prim.data_ptr(), d_ph.data_ptr(), W * H, d_r.data_ptr(), spp=args.spp, seed=1000, bucket_octants=bool(octant),
                                        d_ids_out=d_ids.data_ptr(), stream=stream)
            d_h = torch.empty((n, 8), dtype=torch.float32, device="cuda")
            ids = d_ids[:n].cpu().numpy().view(np.uint32)
            if want is None and not args.no_check:
                rays = d_r[:n].cpu().numpy().view(api.RAY_DT).reshape(-1)
                w_hits, cnt = ob.trace(ofmt, ob.CLOSEST_IGNORE_TRANSPARENT, nodes, tris, v, ents, rays, nthreads=ob.hardware_threads())
                want = dict(zip(ids.tolist(), range(n))), w_hits, cnt
                b_ray = (cnt["node_iters"] * node_bytes + cnt["tri_tests"] * 64.0) / n
            knob_sets = [None] + [[int(x) for x in k.split(",")] for k in args.knobs.split(";") if k]
            for ks in knob_sets:
                if ks:
                    for kid, val in enumerate(ks):
                        ri.set_tuning(kid, val)
                for _ in range(3):
                    ri.intersect_closest_device(d_r.data_ptr(), n, d_h.data_ptr(), api.IGNORE_TRANSPARENT, stream)
                ts = []
                for _ in range(args.reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ri.intersect_closest_device(d_r.data_ptr(), n, d_h.data_ptr(), api.IGNORE_TRANSPARENT, stream)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                ok = None
                if want is not None:
                    got = d_h.cpu().numpy().view(api.HIT_DT).reshape(-1)
                    pos = np.array([want[0][k] for k in ids.tolist()])
                    ok = bool(got.tobytes() == want[1][pos].tobytes())
                print(json.dumps(dict(fmt=fmt_name, octant=octant, knobs=ks, rays=n, ms=round(ms, 4), mrays_s=round(n / ms / 1e3, 1),
                                      frac=None if want is None else round(n / (ms * 1e-3) * b_ray / (PEAK * 1e9), 4), bit_identical=ok)), flush=True)
            if knob_sets[-1]:
                for kid, val in enumerate((8, 14, 10, 0, 0, 12)):
                    ri.set_tuning(kid, val)
        ri.close()


if __name__ == "__main__":
    main()
