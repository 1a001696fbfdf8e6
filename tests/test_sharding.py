"""Host logic of the N > 1 path on CPU: tile sharding, final-frame gather and max-over-ranks timing
with torch.distributed (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from candela_b200 import sharding


def test_tiles_partition_the_screen():
    W, H, world = 200, 130, 3
    parts = [sharding.shard_pixels(W, H, world, r, tile=64) for r in range(world)]
    allpix = np.concatenate(parts)
    assert len(allpix) == W * H and len(np.unique(allpix)) == W * H
    tid = sharding.tile_ids(W, H, 64)
    for r, p in enumerate(parts):
        assert np.all(tid[p] % world == r)
        assert np.all(np.diff(tid[p]) >= 0)          # tile-major order keeps rays of a tile together
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 2 * 64 * 64


def test_contiguous_ranges_cover_everything():
    n, world = 1000003, 8
    r = [sharding.shard_range(n, world, k) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pix = sharding.shard_pixels(W, H, world, rank, tile=16)
    # a stand-in for this rank's hit records: 8 floats per pixel derived from the pixel index
    rec = torch.from_numpy(np.stack([pix * 1.0 + k for k in range(8)], 1).astype(np.float32))
    frame = sharding.gather_frame(rec, torch.from_numpy(pix), W * H)
    t, u = sharding.reduce_timing(10.0 + rank, float(len(pix)))
    if rank == 0:
        out.put((frame.numpy(), t, u))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_frame_and_timing_world2():
    W, H, world = 96, 50, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame, t, u = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.arange(W * H) * 1.0 + k for k in range(8)], 1).astype(np.float32)
    assert np.array_equal(frame, want)
    assert t == 11.0 and u == W * H


def test_single_process_gather_is_identity():
    pix = sharding.shard_pixels(40, 30, 1, 0)
    rec = torch.arange(len(pix) * 8, dtype=torch.float32).reshape(-1, 8)
    frame = sharding.gather_frame(rec, torch.from_numpy(pix), 40 * 30)
    assert torch.equal(frame[torch.from_numpy(pix)], rec)


def _bcast_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    src = [rng.random((1000, 8)).astype(np.float32), rng.integers(0, 99, size=(333, 4)).astype(np.int32), np.zeros(0, np.uint8),
           rng.integers(0, 255, size=77).astype(np.uint8)]
    got = sharding.broadcast_arrays(src if rank == 0 else None, src=0)
    if rank == 1:
        out.put([np.asarray(a).tobytes() for a in got])
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_arrays_world2():
    """BVH distribution: byte-exact replication of rank 0's buffers."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    want = [rng.random((1000, 8)).astype(np.float32), rng.integers(0, 99, size=(333, 4)).astype(np.int32), np.zeros(0, np.uint8),
            rng.integers(0, 255, size=77).astype(np.uint8)]
    assert got == [a.tobytes() for a in want]

