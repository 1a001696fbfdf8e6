"""Runs the BASELINE.json configurations other than the bench.py headline (configs[2..4]) and prints one
JSON line per configuration.  Single process or under torchrun (one rank per GPU).

  rtao      configs[2]: any-hit, tmax 2.4, 4 spp at 1920x1080, stackless                      (1 GPU)
  bounce4k  configs[3]: 3840x2160, 8 spp, multi-bounce closest-hit, screen tiles dealt to the ranks
  soup10m   configs[4]: ~10M-triangle synthetic scene: GPU build + 100M random rays split over ranks

Every configuration checks a sample (or all) of its results bit-for-bit against the CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import candela_b200 as cb  # noqa: E402
from candela_b200 import api, scenes, sharding  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def dev_rays(rays: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)).cuda()


def timed(fn, reps=3, flush=None):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def setup_s260k(local_rank, fmt=cb.STACKLESS):
    v, i, m = scenes.make_s260k()
    ri = cb.RayIntersector(fmt, device=local_rank)
    ri.AddObject(2, v, i, m)
    ri.BufferData(True)
    ri.PushEntity(2)
    ri.BufferEntities()
    return ri, v, i, m


def oracle_scene(ri, verts):
    from oracle import binding as ob
    nodes, tris, _ = ri.read_buffers()
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
    return ob, nodes, tris, ents


def cfg_rtao(args, rank, world, local_rank):
    if rank != 0:
        return None
    W, H, spp = 1920, 1080, 4
    ri, v, i, m = setup_s260k(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    d_prim = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    d_hits = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    ri.intersect_primary_device(iv, ip, W, H, d_hits.data_ptr(), d_prim.data_ptr(), stream)
    d_ao = torch.empty((W * H * spp, 8), dtype=torch.float32, device="cuda")
    n = ri.generate_rays_device(api.GEN_DIFFUSE, d_prim.data_ptr(), d_hits.data_ptr(), W * H, d_ao.data_ptr(), spp=spp, offset=0.05, tmax=2.4, seed=7,
                                bucket_octants=bool(args.bucket), stream=stream)
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ri.intersect_any_device(d_ao.data_ptr(), n, d_t.data_ptr(), stream), reps=args.reps, flush=flush)
    ob, nodes, tris, ents = oracle_scene(ri, v)
    rays = d_ao[:n].cpu().numpy().view(api.RAY_DT).reshape(-1)
    t0 = time.perf_counter()
    want, cnt = ob.trace(ob.STACKLESS, ob.ANY, nodes, tris, v, ents, rays, nthreads=ob.hardware_threads())
    cpu_s = time.perf_counter() - t0
    got = d_t.cpu().numpy()
    b_ray = (cnt["node_iters"] * 32.0 + cnt["tri_tests"] * 64.0) / n
    return dict(config="rtao_anyhit_1080p_4spp_tmax2.4", octant_bucketed=bool(args.bucket), rays=n, ms=round(ms, 4), mrays_s=round(n / ms / 1e3, 1), occluded_frac=round(float((got > 0).mean()), 4),
                bit_identical_to_oracle=bool(got.tobytes() == want.tobytes()), bytes_per_ray=round(b_ray, 1),
                roofline_frac=round(n / (ms * 1e-3) * b_ray / (PEAK * 1e9), 4), node_iters_per_ray=round(cnt["node_iters"] / n, 2),
                cpu_mrays_s=round(n / cpu_s / 1e6, 2), cpu_threads=ob.hardware_threads())


def cfg_spec_shadow(args, rank, world, local_rank):
    """The other two ray batches the engine traces: specular reflection rays (SpecularTrace.glsl, closest hit) and shadow rays
    towards a directional light (any hit), both generated on the GPU from the 1080p primary hits."""
    if rank != 0:
        return None
    W, H = 1920, 1080
    ri, v, i, m = setup_s260k(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    d_prim = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    d_hits = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    ri.intersect_primary_device(iv, ip, W, H, d_hits.data_ptr(), d_prim.data_ptr(), stream)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    ob, nodes, tris, ents = oracle_scene(ri, v)
    out = []
    for label, kind, kw, any_hit in (("specular_rough0.3_1080p_closest_hit", api.GEN_SPECULAR, dict(roughness=0.3, offset=-1.0, seed=21), False),
                                     ("specular_mirror_1080p_closest_hit", api.GEN_SPECULAR, dict(roughness=0.0, offset=-1.0, seed=22), False),
                                     ("shadow_sun_1080p_any_hit", api.GEN_SHADOW, dict(light_dir=(0.3244, 0.8111, 0.4867), light_cone=0.02, tmax=200.0, offset=0.02, seed=23), True)):
        d_r = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
        n = ri.generate_rays_device(kind, d_prim.data_ptr(), d_hits.data_ptr(), W * H, d_r.data_ptr(), stream=stream, **kw)
        if any_hit:
            d_o = torch.empty(n, dtype=torch.float32, device="cuda")
            ms = timed(lambda: ri.intersect_any_device(d_r.data_ptr(), n, d_o.data_ptr(), stream), reps=args.reps, flush=flush)
        else:
            d_o = torch.empty((n, 8), dtype=torch.float32, device="cuda")
            ms = timed(lambda: ri.intersect_closest_device(d_r.data_ptr(), n, d_o.data_ptr(), 0, stream), reps=args.reps, flush=flush)
        rays = d_r[:n].cpu().numpy().view(api.RAY_DT).reshape(-1)
        want, cnt = ob.trace(ob.STACKLESS, ob.ANY if any_hit else ob.CLOSEST, nodes, tris, v, ents, rays, nthreads=ob.hardware_threads())
        got = d_o.cpu().numpy() if any_hit else d_o.cpu().numpy().view(api.HIT_DT).reshape(-1)
        b_ray = (cnt["node_iters"] * 32.0 + cnt["tri_tests"] * 64.0) / n
        out.append(dict(config=label, rays=n, ms=round(ms, 4), mrays_s=round(n / ms / 1e3, 1), bit_identical_to_oracle=bool(got.tobytes() == want.tobytes()),
                        bytes_per_ray=round(b_ray, 1), roofline_frac=round(n / (ms * 1e-3) * b_ray / (PEAK * 1e9), 4),
                        node_iters_per_ray=round(cnt["node_iters"] / n, 2), hit_frac=round(float(((got > 0) if any_hit else (got["t"] > 0)).mean()), 4)))
    return out


def cfg_bounce4k(args, rank, world, local_rank):
    """BASELINE configs[3] through the frame-level call: tiles dealt round-robin to the ranks, every rank's resolved 32-byte pixels
    gathered into rank 0's row-major frame (sharding.gather_frame: one NCCL gather of the tile-major shards + the untile kernel)."""
    from oracle import frame as of
    W, H, spp, bounces = args.width or 3840, args.height or 2160, 8, 4
    ri, v, i, m = setup_s260k(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)

    def params(r, local=True):
        return cb.frame_params(iv, ip, W, H, spp=spp, bounces=bounces, seed=4000, shard_index=r, shard_count=world, out_format=api.FRAME_OUT_PIXEL32,
                               octant_order=bool(args.bucket), local_layout=local)
    shard = torch.zeros((ri.frame_records(params(rank)), 32), dtype=torch.uint8, device="cuda")

    def untile(sh, r, frame):
        ri.frame_untile_device(params(r), sh.data_ptr(), frame.data_ptr(), stream=stream)

    peer = None
    if world > 1 and args.transport == "peer":   # every rank's resolve kernel stores into rank 0's frame over NVLink (CUDA IPC mapping)
        peer = sharding.PeerFrame(ri, W * H * 32, dst=0)

    def step():
        if peer is not None:
            ri.trace_frame_device(params(rank, local=False), peer.ptr, slot=0, stream=stream)
            peer.complete()
            return peer.tensor() if rank == 0 else None
        ri.trace_frame_device(params(rank), shard.data_ptr(), slot=0, stream=stream)
        return sharding.gather_frame(shard, W, H, 64, untile=untile)
    for _ in range(2):
        frame = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        frame = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    t_max, rays_all = sharding.reduce_timing(ms, float(ri.frame_rays_traced(0)), device="cuda")
    px = frame.cpu().numpy().view(api.PIXEL_DT).reshape(-1) if rank == 0 else None
    if peer is not None:
        peer.close()
    if rank != 0:
        return None
    # one tile row of the frame against the oracle (every pixel of 64 image rows)
    ob, nodes, tris, ents = oracle_scene(ri, v)
    pixels = np.arange(W * 64 * 8, W * 64 * 9, dtype=np.uint32)
    want, _ = of.trace_frame(ob.STACKLESS, nodes, tris, v, ents, iv, ip, W, H, spp=spp, bounces=bounces, seed=4000, out_format=of.OUT_PIXEL32, pixels=pixels)
    return dict(config=f"multibounce_{W}x{H}_8spp_4bounces_tiles", octant_major=bool(args.bucket), n_gpus=world, transport=("peer" if args.transport == "peer" and world > 1 else "gather"),
                rays_all_ranks=int(rays_all),
                ms_max_over_ranks_gather_included=round(t_max, 3), mrays_s=round(rays_all / t_max / 1e3, 1), frame_pixels=int(len(px)),
                rays_in_frame=int(px["rays"].sum()), sample_pixels_checked=int(len(pixels)),
                sample_bit_identical_to_oracle=bool(px[pixels].tobytes() == want[pixels].tobytes()))


def cfg_soup10m(args, rank, world, local_rank):
    n_grid = args.grid or 2237
    n_rays = args.rays or 100_000_000
    v, i, m = scenes.make_heightfield(n_grid)
    T = len(i) // 3
    ri = cb.RayIntersector(cb.STACKLESS, device=local_rank)
    t0 = time.perf_counter()
    bcast_ms = None
    if args.bvh == "broadcast" and world > 1:
        # SURVEY.md §8e baseline: build on one GPU, replicate the reference-layout buffers over NVLink (NCCL broadcast)
        if rank == 0:
            ri.AddObject(2, v, i, m, builder=args.builder)
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        sharding.broadcast_scene(ri, [2], src=0, device="cuda")
        torch.cuda.synchronize()
        dist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t1)
    else:
        warm = cb.RayIntersector(cb.STACKLESS, device=local_rank)   # the first build of a process also loads the kernels: report the second
        warm.AddObject(2, v, i, m, builder=args.builder)
        first_build_ms = warm.last_build_ms
        warm.close()
        t0 = time.perf_counter()
        ri.AddObject(2, v, i, m, builder=args.builder)
    wall_build = time.perf_counter() - t0
    build_ms = ri.last_build_ms
    if args.knobs:
        for kid, val in enumerate(int(x) for x in args.knobs.split(",")):
            ri.set_tuning(kid, val)
    ri.set_traversal_mode(2, args.sort)
    ri.BufferData(True)
    ri.PushEntity(2)
    ri.BufferEntities()
    lo, hi = shard_lo_hi = sharding.shard_range(n_rays, world, rank)
    n = hi - lo
    pos = v["position"][:, :3]
    box_lo, box_hi = pos.min(0), pos.max(0) + np.array([0, 10, 0], np.float32)
    stream = torch.cuda.current_stream().cuda_stream
    chunk = 12_500_000
    total_ms, done, sample = 0.0, 0, None
    d_hits = torch.empty((min(chunk, n), 8), dtype=torch.float32, device="cuda")
    while done < n:
        c = min(chunk, n - done)
        rays = scenes.random_rays(box_lo, box_hi, c, seed=1 + rank * 1000 + done // chunk)
        if args.presort:      # experiment: rays in Morton order of their origin cell (bits per axis), direction octant minor / major
            o, d = rays["o"], rays["d"]
            q = np.clip(((o - box_lo) / (box_hi - box_lo) * (2 ** args.presort - 1)).astype(np.uint64), 0, 2 ** args.presort - 1)
            def spread(x):
                r = np.zeros_like(x)
                for b in range(args.presort):
                    r |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
                return r
            key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
            octant = (d[:, 0] > 0).astype(np.uint64) | ((d[:, 1] > 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] > 0).astype(np.uint64) << np.uint64(2))
            key = (key << np.uint64(3)) | octant if args.presort_mode == 0 else key | (octant << np.uint64(3 * args.presort))
            rays = rays[np.argsort(key, kind="stable")]
        d_r = dev_rays(rays)
        ri.intersect_closest_device(d_r.data_ptr(), c, d_hits.data_ptr(), 0, stream)  # warm-up / page-in
        torch.cuda.synchronize()
        total_ms += timed(lambda: ri.intersect_closest_device(d_r.data_ptr(), c, d_hits.data_ptr(), 0, stream), reps=1)
        if sample is None and rank == 0:
            k = min(c, args.check_rays)
            sample = (rays[:k].copy(), d_hits[:k].cpu().numpy().view(api.HIT_DT).reshape(-1).copy())
        done += c
    t_max, rays_all = sharding.reduce_timing(total_ms, float(n), device="cuda")
    if rank != 0:
        return None
    out = dict(config=f"heightfield_{T}_tris_random_rays", n_gpus=world, triangles=T, nodes=ri.node_count, gpu_build_ms=round(build_ms, 2),
               build_wall_ms=round(1e3 * wall_build, 1), builder="sah_exact" if args.builder == 0 else "lbvh", rays_all_ranks=int(rays_all),
               trace_ms_max_over_ranks=round(t_max, 2), mrays_s=round(rays_all / t_max / 1e3, 1), sort_rays=args.sort, bvh=args.bvh,
               bvh_broadcast_wall_ms=None if bcast_ms is None else round(bcast_ms, 1))
    if args.check_rays:
        from oracle import binding as ob
        nodes, tris, _ = ri.read_buffers()
        ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
        t0 = time.perf_counter()
        want, cnt = ob.trace(ob.STACKLESS, ob.CLOSEST, nodes, tris, v, ents, sample[0], nthreads=ob.hardware_threads())
        cpu_s = time.perf_counter() - t0
        k = len(sample[0])
        b_ray = (cnt["node_iters"] * 32.0 + cnt["tri_tests"] * 64.0) / k
        out.update(sample_rays=k, sample_bit_identical_to_oracle=bool(want.tobytes() == sample[1].tobytes()), bytes_per_ray=round(b_ray, 1),
                   roofline_frac_per_gpu=round(rays_all / world / (t_max * 1e-3) * b_ray / (PEAK * 1e9), 4), hit_frac=round(float((want["t"] > 0).mean()), 3),
                   cpu_mrays_s=round(k / cpu_s / 1e6, 2), cpu_threads=ob.hardware_threads())
        if args.check_build:
            t0 = time.perf_counter()
            ref = ob.build(ob.STACKLESS, v, i, m)
            out.update(cpu_build_ms=round(1e3 * (time.perf_counter() - t0)), build_byte_identical=bool(ref.nodes.tobytes() == nodes.tobytes() and ref.tris.tobytes() == tris.tobytes()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=["rtao", "spec_shadow", "bounce4k", "soup10m"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--builder", type=int, default=0)
    ap.add_argument("--presort", type=int, default=0, help="soup10m experiment: host-side Morton sort of each ray chunk, bits per axis")
    ap.add_argument("--presort-mode", type=int, default=0, help="0: origin cell major; 1: direction octant major")
    ap.add_argument("--sort", type=int, default=4, help="soup10m: cndl_set_traversal_mode sort_rays (0 off, 1 octant buckets, 2 octant + origin Morton order with the rays moved, 3 through an index list, 4 automatic)")
    ap.add_argument("--knobs", default="", help="comma-separated tuning knobs in knob-id order (see dev_bench.py)")
    ap.add_argument("--transport", default="gather", choices=["gather", "peer"],
                    help="bounce4k at N > 1: NCCL gather of tile-major shards + untile, or every rank storing into rank 0's frame over NVLink (CUDA IPC)")
    ap.add_argument("--bvh", default="build", choices=["build", "broadcast"], help="soup10m at N > 1: every rank builds, or rank 0 builds and broadcasts")
    ap.add_argument("--bucket", type=int, default=0, help="1: the generator emits each batch octant-major (CNDL_GEN_BUCKET_OCTANTS)")
    ap.add_argument("--check-rays", type=int, default=1_000_000)
    ap.add_argument("--check-build", action="store_true")
    args = ap.parse_args()
    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    for c in args.configs:
        res = {"rtao": cfg_rtao, "spec_shadow": cfg_spec_shadow, "bounce4k": cfg_bounce4k, "soup10m": cfg_soup10m}[c](args, rank, world, local_rank)
        if rank == 0 and res is not None:
            for line in (res if isinstance(res, list) else [res]):
                print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
