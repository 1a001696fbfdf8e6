#!/usr/bin/env python
"""Headline benchmark (driver contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`).

Workload (BASELINE.json configs[1]): 1920x1080 cosine-hemisphere diffuse-GI rays, 1 spp, single
bounce, closest hit (IntersectRayIgnoreTransparent, DiffuseTrace.glsl:484) on the ~260k-triangle
scene, stackless node format.  One "step" = one pass of the hot path over one such ray batch.

  value      : Mrays/s with rays and hit records resident in HBM (CUDA events on the launch stream,
               L2 flushed between steps, summed over ranks / max-over-ranks time).
  e2e        : the same batch through the C-ABI host call cndl_intersect_closest with pinned HOST
               buffers: H2D copy of the rays and D2H copy of the hit records inside the timed region.
  roofline   : algorithmic bytes per ray (oracle counters on the same rays and BVH)
               x rays / kernel time, against the measured HBM bandwidth in MEASURED_PEAKS.json.
  cpu_baseline: the reference's GLSL traversal compiled against its glm (oracle/_ref; else the oracle port) on all host cores
                (rank 0, N=1 only).

--impl reference times the reference's CPU path (the oracle port: the reference's traversal is GLSL
and has no CPU implementation; its builder is timed through oracle/_ref when that library exists).
With N > 1 every rank traces its own batch (weak scaling; BVH replicated per GPU, no data-path
collective); hit records are gathered once over NCCL after the timed region ("final frame").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WIDTH, HEIGHT = 1920, 1080
METRIC = "Mrays/s (incoherent closest-hit, diffuse-GI batch, ~260k-tri scene)"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """DRAM bytes per launch of the traversal kernel from the committed ncu --set full capture."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(rank: int):
    """Scene + camera (inputs shared by both arms)."""
    from candela_b200 import scenes
    verts, indices, mesh_ids = scenes.make_s260k()
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=WIDTH, height=HEIGHT)
    return scenes, verts, indices, mesh_ids, iv, ip


def reference_tracer(ob, threads):
    """(trace function, kind, description): the reference's own GLSL traversal compiled against its glm (oracle/_ref, built in the
    container where /root/reference exists and shipped with the snapshot) when that library loads, else the oracle port."""
    if ob.REF_LIB_PATH.exists():
        try:
            ob.ref_lib()

            def trace(kind, nodes, tris, verts, ents, rays):
                return ob.ref_glsl_trace(ob.STACKLESS, kind, nodes, tris, verts, ents, rays, nthreads=threads)
            return trace, "reference", "the reference's TraverseBVHStackless.glsl compiled against its vendored glm (oracle/_ref), std::thread over ray ranges"
        except OSError:
            pass

    def trace(kind, nodes, tris, verts, ents, rays):
        return ob.trace(ob.STACKLESS, kind, nodes, tris, verts, ents, rays, nthreads=threads)[0]
    return trace, "port", "oracle port of the GLSL traversal, std::thread over ray ranges"


def run_reference(args):
    """The reference's CPU path on the host cores: its GLSL traversal compiled (oracle/_ref) or the oracle port [+ oracle/_ref builder]."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    scenes, verts, indices, mesh_ids, iv, ip = build_workload(0)
    threads = ob.hardware_threads()
    t0 = time.perf_counter()
    b = ob.build(ob.STACKLESS, verts, indices, mesh_ids)
    port_build_ms = 1e3 * (time.perf_counter() - t0)
    ref_build_ms = None
    if ob.REF_LIB_PATH.exists():
        try:
            t0 = time.perf_counter()
            ob.ref_build(ob.STACKLESS, [(verts, indices.reshape(-1, 3), 0)])
            ref_build_ms = 1e3 * (time.perf_counter() - t0)
        except Exception:
            ref_build_ms = None
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(b.nodes))
    prim = ob.primary_rays(iv, ip, WIDTH, HEIGHT)
    hits, _ = ob.trace(ob.STACKLESS, ob.CLOSEST, b.nodes, b.tris, verts, ents, prim, nthreads=threads)
    rays, _ = scenes.bounce_rays(prim, hits, b.tris, verts, seed=1000)
    R = len(rays)
    trace, kind, how = reference_tracer(ob, threads)
    for _ in range(args.warmup):
        trace(ob.CLOSEST_IGNORE_TRANSPARENT, b.nodes, b.tris, verts, ents, rays)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        trace(ob.CLOSEST_IGNORE_TRANSPARENT, b.nodes, b.tris, verts, ents, rays)
    dt = time.perf_counter() - t0
    mrays = R * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mrays, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "diffuse_gi_1080p_1spp_closest_hit", "scene": "S260k stand-in (262,624 triangles)", "node_format": "stackless",
                   "rays_per_step": R, "resolution": [WIDTH, HEIGHT]},
        "cpu_baseline": {"value": round(mrays, 3), "unit": "Mrays/s", "cores": threads, "kind": kind,
                         "sample": f"the full {R}-ray diffuse batch per step; {how}",
                         "build_ms_port_1thread": round(port_build_ms, 1), "build_ms_reference_builder_1thread": None if ref_build_ms is None else round(ref_build_ms, 1)},
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import candela_b200 as cb
    from candela_b200 import api, sharding
    numa_bound = sharding.bind_to_gpu_numa_node(local_rank)   # before any pinned allocation: e2e stages rays through host memory
    scenes, verts, indices, mesh_ids, iv, ip = build_workload(rank)

    # ---- scene: GPU build behind RayIntersector::AddObject, BVH replicated per rank ----
    ri = cb.RayIntersector(cb.STACKLESS, device=local_rank)
    ri.set_traversal_mode(args.mode, bool(args.sort))
    build_ms = []
    for rep in range(args.build_reps):
        tmp = cb.RayIntersector(cb.STACKLESS, device=local_rank)
        tmp.AddObject(2, verts, indices, mesh_ids)
        build_ms.append(tmp.last_build_ms)
        tmp.close()
    ri.AddObject(2, verts, indices, mesh_ids)
    build_ms.append(ri.last_build_ms)
    ri.BufferData(True)
    ri.PushEntity(2)
    ri.BufferEntities()
    nodes, tris, _ = ri.read_buffers()

    # ---- this rank's ray batch: primary hits -> cosine-hemisphere diffuse rays (seed differs per rank) ----
    hits0, prim = ri.IntersectPrimary(iv, ip, WIDTH, HEIGHT, return_rays=True)
    rays, _ = scenes.bounce_rays(prim, hits0, tris, verts, seed=1000 + rank)
    R = len(rays)
    flags = api.IGNORE_TRANSPARENT

    pin_rays = cb.PinnedBuffer(R, api.RAY_DT)
    pin_hits = cb.PinnedBuffer(R, api.HIT_DT)
    pin_rays.array[:] = rays
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.empty((R, 8), dtype=torch.float32, device="cuda")
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        ri.intersect_closest_device(d_rays.data_ptr(), R, d_hits.data_ptr(), flags, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- warm-up (the clock sampler starts here so that it is running during the timed region) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---- timed: K steps, device time per step with CUDA events, L2 flushed (untimed) between steps ----
    launches0 = ri.launch_count
    evs = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = ri.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))

    # ---- end to end through the host C-ABI call (pinned host buffers; copies inside the timed region) ----
    for _ in range(2):
        ri.IntersectRays(pin_rays.array, ignore_transparent=True, out=pin_hits.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ri.IntersectRays(pin_rays.array, ignore_transparent=True, out=pin_hits.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_s * 1e3, float(R)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms_max, e2e_ms_max, rays_all = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        total_ms_max, e2e_ms_max, rays_all = total_ms, e2e_s * 1e3, float(R)

    # ---- final-frame gather of hit records over NCCL (outside the timed region) ----
    gather_ms = None
    if world > 1:
        n_pad = int(torch.tensor([R], device="cuda").max().item())
        rmax = torch.tensor([R], dtype=torch.int64, device="cuda")
        dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
        n_pad = int(rmax.item())
        send = torch.zeros((n_pad, 8), dtype=torch.float32, device="cuda")
        send[:R] = d_hits
        recv = torch.empty((world * n_pad, 8), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(recv, send)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)

    if rank == 0:
        value = rays_all * args.steps / (total_ms_max * 1e-3) / 1e6
        e2e_value = rays_all * args.steps / (e2e_ms_max * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(total_ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "diffuse_gi_1080p_1spp_closest_hit", "scene": "S260k stand-in (262,624 triangles, GPU-built binned SAH)",
                       "node_format": "stackless", "rays_per_gpu_per_step": R, "resolution": [WIDTH, HEIGHT], "bvh": "replicated per GPU",
                       "l2": "flushed between steps (192 MiB memset, untimed)", "traversal_mode": args.mode, "sort_rays": bool(args.sort)},
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": int(rays_all) * 32, "d2h_bytes_per_step": int(rays_all) * 32,
                    "timing": "host wall clock around the synchronous C-ABI call, max over ranks",
                    "host_buffers": "pinned; each rank bound to its GPU's NUMA node" if numa_bound else "pinned"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "build": {"gpu_ms": round(min(build_ms), 3), "triangles": int(len(indices) // 3), "nodes": int(len(nodes)),
                      "builder": "binned SAH, byte-identical to BVH::BuildBVH"},
            "wall_s_timed_region": round(wall, 3),
        }
        if gather_ms is not None:
            line["final_gather_ms"] = round(gather_ms, 3)
        # ---- oracle: parity of this batch, algorithmic bytes per ray, CPU baseline (N=1 only) ----
        from oracle import binding as ob
        ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
        threads = ob.hardware_threads()
        want, cnt = ob.trace(ob.STACKLESS, ob.CLOSEST_IGNORE_TRANSPARENT, nodes, tris, verts, ents, rays, nthreads=threads)
        got = d_hits.cpu().numpy().view(api.HIT_DT).reshape(-1)
        line["parity"] = {"checked_rays": R, "bit_identical_to_oracle": bool(got.tobytes() == want.tobytes()),
                          "index_mismatches": int(np.count_nonzero((got["tri"] != want["tri"]) | (got["mesh"] != want["mesh"]) | (got["entity"] != want["entity"])))}
        nn, nt = cnt["node_iters"] / R, cnt["tri_tests"] / R
        b_ray = nn * 32.0 + nt * 64.0
        peak, peak_src = measured_peak()
        kernel_ms = statistics.mean(step_ms)  # one traversal launch per step on this rank
        achieved = b_ray * R / (kernel_ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                            "traffic": recorded_traffic(), "peak_source": peak_src, "kernel": {0: "trace_simple_kernel", 1: "trace_persistent_stackless_kernel", 2: "trace_ww_stackless_kernel"}[args.mode],
                            "bytes_per_ray": round(b_ray, 1), "node_iters_per_ray": round(nn, 3), "tri_tests_per_ray": round(nt, 3),
                            "ray_io_bytes_per_ray_not_included": 64, "kernel_ms": round(kernel_ms, 4)}
        if world == 1:
            cpu_trace, cpu_kind, cpu_how = reference_tracer(ob, threads)
            reps, t_cpu = 0, 0.0
            while t_cpu < 1.5 and reps < 20:
                t0 = time.perf_counter()
                cpu_hits = cpu_trace(ob.CLOSEST_IGNORE_TRANSPARENT, nodes, tris, verts, ents, rays)
                t_cpu += time.perf_counter() - t0
                reps += 1
            if cpu_kind == "reference":   # the benchmarked batch against the compiled reference shaders themselves
                line["parity"]["bit_identical_to_compiled_reference_glsl"] = bool(cpu_hits.tobytes() == got.tobytes())
            t0 = time.perf_counter()
            ob.build(ob.STACKLESS, verts, indices, mesh_ids)
            cpu_build_ms = 1e3 * (time.perf_counter() - t0)
            line["cpu_baseline"] = {"value": round(R * reps / t_cpu / 1e6, 3), "unit": "Mrays/s", "cores": threads, "kind": cpu_kind,
                                    "sample": f"{reps} passes over the full {R}-ray batch ({t_cpu * threads:.0f} core-seconds); {cpu_how}",
                                    "build_ms_1thread": round(cpu_build_ms, 1)}
        print(json.dumps(line), flush=True)

    pin_rays.free()
    pin_hits.free()
    ri.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", type=int, default=2, help="traversal kernel: 0 one thread per ray, 1 persistent warps, 2 persistent while-while")
    ap.add_argument("--sort", type=int, default=0)
    ap.add_argument("--build-reps", type=int, default=3, help="extra timed GPU builds (0 keeps the ncu launch list short)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
