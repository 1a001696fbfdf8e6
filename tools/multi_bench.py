"""Development tool: the multi-device handle (cndl_multi_*) in ONE process over all visible GPUs.

  * scene replication device to device (build once on device 0, cudaMemcpy over NVLink to the others), 262k and ~2M triangles
  * the BASELINE configs[3] frame (3840x2160 x 8 spp x 4 bounces) with tiles dealt round-robin, every device's resolve kernel storing
    its pixels into device 0's frame through peer memory; host wall clock around cndl_multi_trace_frame (one D2H of the frame
    included), and the pipelined submit/wait rate
  * the frame against the single-device frame, bit for bit
"""
from __future__ import annotations

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
from candela_b200 import api, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--big-grid", type=int, default=1001, help="heightfield grid for the replication figure (1001 -> 2.0 M triangles)")
    args = ap.parse_args()
    n_dev = torch.cuda.device_count()
    devices = tuple(range(n_dev))
    out = {"devices": n_dev}
    # replication of a ~2M-triangle scene
    v, i, m = scenes.make_heightfield(args.big_grid)
    mb = cb.MultiRayIntersector(cb.STACKLESS, devices)
    mb.AddObject(2, v, i, m)
    mb.AddObject(3, v, i, m)          # second call: allocations are warm
    out["replicate_2M_tris_ms"] = round(mb.last_replicate_ms, 3)
    out["replicate_2M_tris"] = int(len(i) // 3)
    a = mb.context(0).read_buffers()
    same = all(all(x.tobytes() == y.tobytes() for x, y in zip(a, mb.context(k).read_buffers())) for k in range(1, n_dev))
    out["replicas_byte_identical"] = bool(same)
    mb.close()

    v, i, m = scenes.make_s260k()
    mi = cb.MultiRayIntersector(cb.STACKLESS, devices)
    mi.AddObject(2, v, i, m)
    out["replicate_262k_tris_ms"] = round(mi.last_replicate_ms, 3)
    mi.BufferData()
    mi.PushEntity(2)
    mi.BufferEntities()
    W, H = args.width, args.height
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    p = cb.frame_params(iv, ip, W, H, spp=args.spp, bounces=args.bounces, seed=4000, out_format=api.FRAME_OUT_PIXEL32, octant_order=False)   # 8 spp: pixel order keeps a pixel's samples together (see profiles/r2_experiments.md)
    pins = [cb.PinnedBuffer(W * H, api.PIXEL_DT) for _ in range(2)]
    for transport, label in ((api.TRANSPORT_PEER_STORES, "peer_stores"), (api.TRANSPORT_STAGED_COPY, "staged_copy")):
        if n_dev == 1 and transport == api.TRANSPORT_STAGED_COPY:
            continue
        mi.set_transport(transport)
        mi.TraceFrame(p, pins[0].array)
        mi.TraceFrame(p, pins[0].array)
        for k in range(4):   # both slots allocate their scratch (and, for staged copies, their staging) outside the timed loops
            mi.frame_submit(p, pins[k & 1].array, k & 1)
        mi.frame_wait(0)
        mi.frame_wait(1)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            mi.TraceFrame(p, pins[0].array)
        dt = (time.perf_counter() - t0) / args.reps
        rays = mi.frame_rays_traced(0)
        out[f"frame_ms_{label}"] = round(1e3 * dt, 3)
        out[f"mrays_s_{label}"] = round(rays / dt / 1e6, 1)
        # two frames in flight
        t0 = time.perf_counter()
        for k in range(args.reps * 2):
            mi.frame_submit(p, pins[k & 1].array, k & 1)
        mi.frame_wait(0)
        mi.frame_wait(1)
        dt = (time.perf_counter() - t0) / (args.reps * 2)
        out[f"frame_ms_{label}_pipelined"] = round(1e3 * dt, 3)
        out[f"mrays_s_{label}_pipelined"] = round(rays / dt / 1e6, 1)
    out["rays_per_frame"] = int(rays)
    frame = pins[0].array.copy()
    single = mi.context(0).TraceFrame(p)
    out["equals_single_device_frame"] = bool(single.tobytes() == frame.tobytes())
    t0 = time.perf_counter()
    for _ in range(2):
        mi.context(0).TraceFrame(p, pins[1].array)
    out["single_device_frame_ms"] = round(1e3 * (time.perf_counter() - t0) / 2, 3)
    print(json.dumps(out))
    mi.close()


if __name__ == "__main__":
    main()
