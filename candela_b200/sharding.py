"""Multi-GPU host logic: the path shards with no exchange step.  Screen tiles are dealt
round-robin to ranks, every rank traces its own tiles against its replica of the BVH, and hit
records are gathered once for the final frame.  Works with any torch.distributed backend (NCCL on
the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def tile_ids(width: int, height: int, tile: int = 64) -> np.ndarray:
    """Tile index of every pixel (row-major pixels, row-major tiles)."""
    ys, xs = np.divmod(np.arange(width * height, dtype=np.int64), width)
    return (ys // tile) * ((width + tile - 1) // tile) + (xs // tile)


def shard_pixels(width: int, height: int, world: int, rank: int, tile: int = 64) -> np.ndarray:
    """Pixel indices owned by `rank`: tiles dealt round-robin by tile index, pixels in tile-major order."""
    tid = tile_ids(width, height, tile)
    mine = np.nonzero(tid % world == rank)[0]
    return mine[np.argsort(tid[mine], kind="stable")]


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous split of n rays (config 5)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_frame(local_records, pixel_index, n_pixels: int, group=None):
    """Assembles the final frame on every rank: `local_records` [n_local, K] (torch tensor on the
    backend's device) belong to pixels `pixel_index` [n_local] (int64 tensor).  Returns [n_pixels, K].
    One all_gather of padded shards; ranks may own different numbers of pixels."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = torch.full((n_pixels, local_records.shape[1]), -1, dtype=local_records.dtype, device=local_records.device)
        out[pixel_index] = local_records
        return out
    n_local = torch.tensor([local_records.shape[0]], dtype=torch.int64, device=local_records.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    n_max = int(max(c.item() for c in counts))
    pad_r = torch.zeros((n_max, local_records.shape[1]), dtype=local_records.dtype, device=local_records.device)
    pad_i = torch.full((n_max,), -1, dtype=torch.int64, device=local_records.device)
    pad_r[: local_records.shape[0]] = local_records
    pad_i[: local_records.shape[0]] = pixel_index
    all_r = [torch.empty_like(pad_r) for _ in range(world)]
    all_i = [torch.empty_like(pad_i) for _ in range(world)]
    dist.all_gather(all_r, pad_r, group=group)
    dist.all_gather(all_i, pad_i, group=group)
    out = torch.full((n_pixels, local_records.shape[1]), -1, dtype=local_records.dtype, device=local_records.device)
    for r, i, c in zip(all_r, all_i, counts):
        n = int(c.item())
        out[i[:n]] = r[:n]
    return out


def reduce_timing(ms_local: float, units_local: float, device="cpu", group=None) -> tuple[float, float]:
    """(max over ranks of the device time, sum over ranks of the units processed)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ms_local, units_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(u.item())


def broadcast_arrays(arrays, src: int = 0, device="cpu", group=None):
    """Broadcasts a list of numpy arrays (any dtypes / shapes) from rank `src`; the other ranks pass None and
    receive copies.  Sizes travel first, then the raw bytes of each array (torch.distributed.broadcast:
    NCCL over NVLink on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return arrays
    rank = dist.get_rank(group)
    n = torch.tensor([len(arrays) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src, group=group)
    sizes = torch.zeros(int(n.item()), dtype=torch.int64, device=device)
    if rank == src:
        sizes = torch.tensor([a.nbytes for a in arrays], dtype=torch.int64, device=device)
    dist.broadcast(sizes, src, group=group)
    out = []
    for k, nbytes in enumerate(int(x) for x in sizes.tolist()):
        if rank == src:
            buf = torch.from_numpy(np.ascontiguousarray(arrays[k]).view(np.uint8).reshape(-1).copy()).to(device)
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if nbytes:
            dist.broadcast(buf, src, group=group)
        out.append(arrays[k] if rank == src else buf.cpu().numpy())
    return out


def broadcast_scene(ri, object_ids, src: int = 0, device="cuda", group=None):
    """BVH distribution of SURVEY.md §8e: rank `src` has built (or loaded) the objects `object_ids` in `ri`; every
    other rank passes an empty intersector of the same node format and receives them as prebuilt objects
    (reference-layout nodes / triangles / vertices), so the scene is built once and replicated over NVLink.
    Call BufferData() afterwards on every rank."""
    import torch.distributed as dist
    from . import api
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    rank = dist.get_rank(group)
    arrays = None
    if rank == src:
        nodes, tris, verts = ri.read_buffers()
        arrays = []
        for oid in object_ids:
            o = ri.object_data(oid)
            ends = sorted(ri.object_data(x)["tri_offset"] for x in object_ids) + [len(tris)]
            vends = sorted(ri.object_data(x)["vert_offset"] for x in object_ids) + [len(verts)]
            t1 = min(e for e in ends if e > o["tri_offset"])
            v1 = min(e for e in vends if e > o["vert_offset"])
            t = tris[o["tri_offset"]:t1].copy()
            t["v"] -= o["vert_offset"]          # AddPrebuiltObject takes object-local vertex indices (Intersector.h:190-197 rebases them)
            arrays += [nodes[o["node_offset"]:o["node_offset"] + o["node_count"]], t, verts[o["vert_offset"]:v1]]
    arrays = broadcast_arrays(arrays, src, device, group)
    if rank != src:
        for k, oid in enumerate(object_ids):
            n, t, v = arrays[3 * k:3 * k + 3]
            ri.AddPrebuiltObject(oid, n.view(ri.node_dtype), t.view(api.TRIANGLE_DT), v.view(api.VERTEX_DT))


def bind_to_gpu_numa_node(gpu_index: int) -> bool:
    """Restricts this process to the CPU cores NVML reports as local to GPU `gpu_index`, so that pinned host buffers
    allocated afterwards (first touch) and the copy-issuing thread sit on the GPU's own NUMA node.  With one process per
    GPU on a two-socket box, half of the ranks otherwise stage their rays through the remote socket.  Returns False
    (and changes nothing) when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False

