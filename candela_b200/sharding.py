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


def shard_slots(width: int, height: int, world: int, rank: int, tile: int = 64) -> np.ndarray:
    """The tile map of the frame-level C calls (cndl_frame_params: tiles t = ty * tiles_x + tx dealt round-robin, shard `rank` owns
    t % world == rank in ascending order): pixel index of every slot of the shard's LOCAL layout (local_tile * tile^2 +
    y_in_tile * tile + x_in_tile), -1 for the padding slots of edge tiles.  len() == cndl_frame_shard_records for one record per pixel."""
    tiles_x, tiles_y = (width + tile - 1) // tile, (height + tile - 1) // tile
    tiles = np.arange(rank, tiles_x * tiles_y, world, dtype=np.int64)
    ty, tx = np.divmod(tiles, tiles_x)
    iy, ix = np.divmod(np.arange(tile * tile, dtype=np.int64), tile)
    y = (ty[:, None] * tile + iy[None, :]).ravel()
    x = (tx[:, None] * tile + ix[None, :]).ravel()
    return np.where((x < width) & (y < height), y * width + x, -1)


def gather_frame(local_records, width: int, height: int, tile: int = 64, group=None, untile=None):
    """Final-frame gather without an index payload: every rank passes its shard in the local tile-major layout
    ([slots, K] tensor, slots = len(shard_slots(...)), on the backend's device); rank 0 receives the shards (one gather of
    equal-size buffers: shard 0 is the largest, the others are padded by at most one tile) and scatters them into the row-major
    frame — the tile map is a pure function of (width, height, tile, world), so no pixel indices travel.  Returns the
    [width * height, K] frame on rank 0 and None elsewhere.  `untile(shard_tensor, rank, frame_tensor)`: device-side scatter
    (cndl_frame_untile_device); default: the numpy mapping above (CPU tensors)."""
    import torch
    import torch.distributed as dist
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    n_max = len(shard_slots(width, height, world, 0, tile))
    send = local_records
    if local_records.shape[0] < n_max:
        send = torch.zeros((n_max,) + tuple(local_records.shape[1:]), dtype=local_records.dtype, device=local_records.device)
        send[: local_records.shape[0]] = local_records
    shards = [send]
    if multi:
        shards = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
        dist.gather(send, shards, dst=0, group=group)
    if rank != 0:
        return None
    frame = torch.full((width * height,) + tuple(local_records.shape[1:]), -1, dtype=local_records.dtype, device=local_records.device)
    for r, sh in enumerate(shards):
        if untile is not None:
            untile(sh, r, frame)
        else:
            slots = torch.from_numpy(shard_slots(width, height, world, r, tile)).to(frame.device)
            ok = slots >= 0
            frame[slots[ok]] = sh[: len(slots)][ok]
    return frame


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


class _DevicePtr:
    """A raw device pointer as a torch-importable object (__cuda_array_interface__), so that NCCL can send straight from / into
    the intersector's own buffers."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def broadcast_scene(ri, object_ids=None, src: int = 0, device="cuda", group=None):
    """BVH distribution of SURVEY.md §8e, device to device: rank `src` has built (or loaded) its objects in `ri`; every other rank
    passes an EMPTY intersector of the same node format and receives ALL of them, in insertion order, as prebuilt objects:
    NCCL broadcasts straight out of rank src's reference-layout buffers (cndl_object_device_view) into device tensors, which
    cndl_add_prebuilt_object_device appends with a device-to-device copy — nothing crosses the host bus.  Leaf packs embed global
    triangle offsets, so a subset or a different order cannot be replicated: the receiver checks every object's offset and fails
    loudly.  `object_ids`, when given, must be the complete list in insertion order (checked).  Call BufferData() afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        ids = sorted(ri.object_ids(), key=lambda oid: ri.object_data(oid)["node_offset"])
        if object_ids is not None and list(object_ids) != ids:
            raise ValueError(f"broadcast_scene replicates the whole scene in insertion order {ids}; got {list(object_ids)}")
        meta = [[(oid, ri.object_data(oid), {k: v for k, v in ri.object_device_view(oid).items() if k.startswith("n_")}) for oid in ids]]
    dist.broadcast_object_list(meta, src=src, group=group, device=torch.device(device))
    node_size = 32 if ri.node_format == 0 else 64
    for oid, od, view in meta[0]:
        sizes = (view["n_nodes"] * node_size, view["n_tris"] * 16, view["n_verts"] * 32)
        if rank == src:
            v = ri.object_device_view(oid)
            bufs = [torch.as_tensor(_DevicePtr(p, n), device=device) for p, n in zip((v["d_nodes"], v["d_tris"], v["d_verts"]), sizes)]
        else:
            bufs = [torch.empty(n, dtype=torch.uint8, device=device) for n in sizes]
        for b in bufs:
            dist.broadcast(b, src, group=group)
        if rank != src:
            torch.cuda.synchronize()
            ri.AddPrebuiltObjectDevice(oid, bufs[0].data_ptr(), view["n_nodes"], bufs[1].data_ptr(), view["n_tris"], bufs[2].data_ptr(), view["n_verts"],
                                       vertex_index_base=od["vert_offset"], leaf_triangle_offset=od["tri_offset"])


class PeerFrame:
    """The final frame of SURVEY.md §8e without a gather: `nbytes` of device memory in rank `dst`'s process that every rank of the
    group maps (cudaIpc; NVLink peer access) and passes as the output of its frame call in the ROW-MAJOR layout, so each rank's
    resolve kernel stores its own tiles' records straight into the one frame.  `ptr` is this process's address of the frame.
    After a frame: complete() = a one-element all-reduce on the current stream — when it has passed on rank dst, every rank's
    stores have landed.  Needs one process per GPU on one node; raises CandelaError where CUDA IPC is not available."""

    def __init__(self, ri, nbytes: int, dst: int = 0, device="cuda", group=None):
        import torch
        import torch.distributed as dist
        self.ri, self.dst, self.group, self.nbytes = ri, dst, group, nbytes
        self.multi = dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.multi else 0
        self.owner = self.rank == dst
        box = [None]
        err = None
        if self.owner:
            try:
                self.ptr, handle = ri.ipc_alloc(nbytes)
                box = [handle]
            except Exception as e:  # every rank must learn about it, or the others would wait in the broadcast forever
                err = e
                box = [None]
        if self.multi:
            dist.broadcast_object_list(box, src=dst, group=group, device=torch.device(device))
        if box[0] is None:
            raise err if err is not None else RuntimeError("PeerFrame: the owning rank could not export its frame (CUDA IPC unavailable)")
        ok = 1
        if not self.owner:
            try:
                self.ptr = ri.ipc_open(box[0])
            except Exception as e:
                err, ok, self.ptr = e, 0, 0
        self._flag = torch.ones(1, dtype=torch.int32, device=device) * ok if self.multi else None
        if self.multi:
            dist.all_reduce(self._flag, op=dist.ReduceOp.MIN, group=group)
            if int(self._flag.item()) == 0:
                self.close()
                raise err if err is not None else RuntimeError("PeerFrame: another rank could not map the frame (CUDA IPC unavailable)")
            self._flag.fill_(1)

    def complete(self):
        """Enqueue the completion barrier on the current stream."""
        if self.multi:
            import torch.distributed as dist
            dist.all_reduce(self._flag, op=dist.ReduceOp.MAX, group=self.group)

    def tensor(self):
        """The frame as a uint8 tensor (any rank; rank dst reads the assembled frame from it)."""
        import torch
        return torch.as_tensor(_DevicePtr(self.ptr, self.nbytes), device="cuda")

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.multi:
            import torch.distributed as dist
            if not self.owner and getattr(self, "ptr", 0):
                self.ri.ipc_close(self.ptr)
            dist.barrier(group=self.group)       # every peer has unmapped before the owner frees
        if self.owner and getattr(self, "ptr", 0):
            self.ri.ipc_free(self.ptr)
        self.ptr = 0


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

