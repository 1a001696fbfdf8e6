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


def test_shard_slots_match_the_c_tile_map():
    """sharding.shard_slots is the Python statement of frame.cu's TileMap: sizes agree with cndl_frame_shard_records, the shards of any
    world size partition the pixels, shard 0 is the largest, and the slot order is tile-major / row-major inside a tile."""
    import ctypes as C
    import candela_b200 as cb
    from candela_b200 import api
    L = api.load_library()
    for W, H, tile in ((200, 130, 64), (1920, 1080, 64), (97, 33, 16), (64, 64, 64), (5, 300, 7)):
        for world in (1, 2, 3, 8):
            seen = []
            sizes = []
            for r in range(world):
                sl = sharding.shard_slots(W, H, world, r, tile)
                p = cb.frame_params(np.eye(4), np.eye(4), W, H, spp=1, tile=tile, shard_index=r, shard_count=world, out_format=api.FRAME_OUT_PIXEL32, local_layout=True)
                assert len(sl) == L.cndl_frame_shard_records(C.byref(p))
                sizes.append(len(sl))
                pix = sl[sl >= 0]
                seen.append(pix)
                tid = sharding.tile_ids(W, H, tile)[pix]
                assert np.all(tid % world == r) and np.all(np.diff(tid) >= 0)
            allpix = np.concatenate(seen)
            assert len(allpix) == W * H and len(np.unique(allpix)) == W * H
            assert sizes[0] == max(sizes) and max(sizes) - min(sizes) <= tile * tile
    p = cb.frame_params(np.eye(4), np.eye(4), 3840, 2160, spp=8, out_format=api.FRAME_OUT_HIT16)
    assert L.cndl_frame_records(C.byref(p)) == 3840 * 2160 * 8 and L.cndl_frame_record_bytes(api.FRAME_OUT_HIT16) == 16


def _worker(rank, world, port, W, H, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slots = sharding.shard_slots(W, H, world, rank, tile=16)
    # a stand-in for this rank's records in the local tile-major layout: 8 floats per slot derived from the pixel index (padding: -7)
    rec = torch.from_numpy(np.stack([np.where(slots >= 0, slots * 1.0 + k, -7.0) for k in range(8)], 1).astype(np.float32))
    frame = sharding.gather_frame(rec, W, H, tile=16)
    t, u = sharding.reduce_timing(10.0 + rank, float(np.count_nonzero(slots >= 0)))
    if rank == 0:
        out.put((frame.numpy(), t, u))
    else:
        assert frame is None
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
    slots = sharding.shard_slots(40, 30, 1, 0, tile=16)
    rec = torch.from_numpy(np.stack([np.where(slots >= 0, slots * 2.0, -1.0)] * 3, 1).astype(np.float32))
    frame = sharding.gather_frame(rec, 40, 30, tile=16)
    assert torch.equal(frame[:, 0], torch.arange(40 * 30, dtype=torch.float32) * 2)


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



def test_sub_shards_partition_a_ranks_tiles():
    """bench.py / INTEGRATION.md §3b: rank r of N cuts its tiles into `split` sub-shards (r + k N, split N) traced by separate frame
    calls on separate streams.  Tile t belongs to rank t mod N, and t mod (split N) is r + k N for exactly one k: the sub-shards
    partition the rank's pixels, so the calls store disjoint records into the one frame."""
    from candela_b200 import sharding
    for (W, H, tile) in ((200, 120, 32), (1920, 1080, 64), (97, 45, 7)):
        for world in (1, 2, 3, 8):
            for split in (2, 3, 4):
                for rank in range(world):
                    whole = sharding.shard_slots(W, H, world, rank, tile)
                    whole = np.sort(whole[whole >= 0])
                    parts = [sharding.shard_slots(W, H, world * split, rank + k * world, tile) for k in range(split)]
                    parts = [p[p >= 0] for p in parts]
                    joined = np.sort(np.concatenate(parts))
                    assert np.array_equal(joined, whole) and len(np.unique(joined)) == len(joined)
