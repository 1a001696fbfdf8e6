"""GPU tests of the frame-level calls (cndl_trace_frame, cndl_frame_submit / _wait, cndl_trace_frame_device,
cndl_frame_untile_device), scene replication and the multi-device handle (cndl_multi_*), all through the C ABI and all
bit-for-bit against oracle/frame.py — the same frame assembled on the CPU from the oracle's camera rays, ray generator
and traversal."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _camera(W, H):
    from candela_b200 import scenes
    return scenes.camera((6.0, 7.0, 8.0), (0.0, 4.0, 0.0), W, H)


@pytest.fixture(scope="module")
def dragon2(cb, ob):
    """The dragon twice: an identity entity and a scaled, shifted, slightly translucent copy (so the first bounce's
    IgnoreTransparent differs from the later bounces' IntersectRay)."""
    from candela_b200 import scenes
    P, F = scenes.load_dragon()
    V = cb.make_vertices(P)
    mids = np.zeros(len(F), np.int32)
    m2 = np.eye(4, dtype=np.float32)
    m2[:3, 3] = (4.0, 0.5, -3.0)
    m2[0, 0] = m2[1, 1] = m2[2, 2] = 0.8

    def fill(ri):
        ri.PushEntity(2)
        ri.PushEntity(2, m2, 0.0, 0.4)
        ri.BufferEntities()
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, V, F.ravel(), mids)
    ri.BufferData()
    fill(ri)
    nodes, tris, _ = ri.read_buffers()
    ents = np.concatenate([ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes)), ob.make_entity(m2, 0, len(nodes), 0.0, 0.4)])
    yield dict(ri=ri, V=V, F=F, mids=mids, nodes=nodes, tris=tris, ents=ents, fill=fill)
    ri.close()


def _oracle_frame(ob, sc, iv, ip, W, H, **kw):
    from oracle import frame as of
    return of.trace_frame(ob.STACKLESS, sc["nodes"], sc["tris"], sc["V"], sc["ents"], iv, ip, W, H, **kw)


def test_frame_hit_formats_equal_the_oracle(cb, ob, dragon2):
    from candela_b200 import api
    ri = dragon2["ri"]
    W, H, spp = 200, 120, 2                                  # not a multiple of the tile: edge tiles carry padding slots
    iv, ip = _camera(W, H)
    want32, traced = _oracle_frame(ob, dragon2, iv, ip, W, H, spp=spp, seed=5, out_format=0)
    assert 0.1 * W * H * spp < traced < W * H * spp and np.count_nonzero(want32["entity"] == 1) == 0   # the translucent copy is skipped (:484 IgnoreTransparent)
    for octant in (False, True):
        for tile in (32, 64, 7):
            for compact in (False, True):                    # the one-pass segmented generator (default) and the three-pass compacting one
                p = cb.frame_params(iv, ip, W, H, spp=spp, seed=5, tile=tile, out_format=api.FRAME_OUT_HIT32, octant_order=octant, compact_rays=compact)
                got = ri.TraceFrame(p)
                assert got.tobytes() == want32.tobytes(), (octant, tile, compact)
                assert ri.frame_rays_traced(0) == traced
    want16, _ = _oracle_frame(ob, dragon2, iv, ip, W, H, spp=spp, seed=5, out_format=1)
    got16 = ri.TraceFrame(cb.frame_params(iv, ip, W, H, spp=spp, seed=5, out_format=api.FRAME_OUT_HIT16))
    assert got16.dtype.itemsize == 16 and got16.tobytes() == want16.tobytes()
    # u is derivable from the compact record exactly
    h = got16["t"] > 0
    assert np.array_equal((np.float32(1.0) - got16["v"][h]) - got16["w"][h], want32["u"][h])


def test_frame_pixels_multibounce_equal_the_oracle(cb, ob, dragon2):
    from candela_b200 import api
    ri = dragon2["ri"]
    W, H = 160, 96
    iv, ip = _camera(W, H)
    for spp, bounces, seed in ((3, 3, 11), (1, 1, 2), (8, 4, 77)):
        want, traced = _oracle_frame(ob, dragon2, iv, ip, W, H, spp=spp, bounces=bounces, seed=seed, out_format=2)
        for octant, compact in ((False, False), (True, False), (False, True), (True, True)):
            got = ri.TraceFrame(cb.frame_params(iv, ip, W, H, spp=spp, bounces=bounces, seed=seed, tile=32, out_format=api.FRAME_OUT_PIXEL32, octant_order=octant,
                                                compact_rays=compact))
            assert got.tobytes() == want.tobytes(), (spp, bounces, octant, compact)
            assert ri.frame_rays_traced(0) == traced == int(got["rays"].sum())
    assert got["ao"].min() >= 0.0 and got["ao"].max() <= 1.0 and np.count_nonzero(got["escaped"]) > 0


def test_frame_shards_assemble_to_the_unsharded_frame(cb, ob, dragon2):
    """Tiles dealt round-robin: the shards of any shard count write disjoint pixels of one buffer and together
    give the unsharded frame bit for bit; the local (tile-major) layout untiles to the same frame."""
    import torch
    from candela_b200 import api
    ri = dragon2["ri"]
    W, H, spp = 200, 120, 2
    iv, ip = _camera(W, H)
    stream = torch.cuda.current_stream().cuda_stream
    for fmt, kw in ((api.FRAME_OUT_HIT16, dict(spp=spp, bounces=1)), (api.FRAME_OUT_PIXEL32, dict(spp=spp, bounces=3))):
        whole = ri.TraceFrame(cb.frame_params(iv, ip, W, H, seed=9, tile=32, out_format=fmt, **kw))
        rec = whole.dtype.itemsize
        for shards in (2, 3, 8, 40):
            frame = torch.full((len(whole) * rec,), 0xEE, dtype=torch.uint8, device="cuda")
            frame2 = torch.full((len(whole) * rec,), 0xEE, dtype=torch.uint8, device="cuda")
            total = 0
            for s in range(shards):
                p = cb.frame_params(iv, ip, W, H, seed=9, tile=32, shard_index=s, shard_count=shards, out_format=fmt, **kw)
                ri.trace_frame_device(p, frame.data_ptr(), slot=s & 1, stream=stream)
                torch.cuda.synchronize()
                total += ri.frame_rays_traced(s & 1)
                pl = cb.frame_params(iv, ip, W, H, seed=9, tile=32, shard_index=s, shard_count=shards, out_format=fmt, local_layout=True, **kw)
                local = torch.full((max(ri.frame_records(pl), 1) * rec,), 0xDD, dtype=torch.uint8, device="cuda")
                ri.trace_frame_device(pl, local.data_ptr(), slot=0, stream=stream)
                ri.frame_untile_device(pl, local.data_ptr(), frame2.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            assert frame.cpu().numpy().tobytes() == whole.tobytes(), (fmt, shards)
            assert frame2.cpu().numpy().tobytes() == whole.tobytes(), (fmt, shards, "local layout")
            assert total == (int(whole["rays"].sum()) if fmt == api.FRAME_OUT_PIXEL32 else total)


def test_frames_in_flight_on_both_slots(cb, ob, dragon2):
    """cndl_frame_submit / cndl_frame_wait: two frames in flight, pinned host buffers, results as the synchronous call."""
    from candela_b200 import api
    ri = dragon2["ri"]
    W, H = 256, 128
    iv, ip = _camera(W, H)
    ps = [cb.frame_params(iv, ip, W, H, spp=2, seed=s, out_format=api.FRAME_OUT_HIT16) for s in (1, 2, 3)]
    want = [ri.TraceFrame(p).copy() for p in ps]
    bufs = [cb.PinnedBuffer(W * H * 2, api.HIT16_DT) for _ in range(2)]
    ri.frame_submit(ps[0], bufs[0].array, 0)
    ri.frame_submit(ps[1], bufs[1].array, 1)
    ri.frame_wait(0)
    assert bufs[0].array.tobytes() == want[0].tobytes()
    ri.frame_submit(ps[2], bufs[0].array, 0)
    ri.frame_wait(1)
    assert bufs[1].array.tobytes() == want[1].tobytes()
    ri.frame_wait(0)
    assert bufs[0].array.tobytes() == want[2].tobytes()
    for b in bufs:
        b.free()
    with pytest.raises(cb.CandelaError):
        ri.TraceFrame(cb.frame_params(iv, ip, W, H, spp=1, bounces=2, out_format=api.FRAME_OUT_HIT16))   # hit formats hold one bounce


def test_scene_replication_device_to_device(cb, ob, dragon2):
    """cndl_clone_scene / cndl_add_prebuilt_object_device: byte-identical buffers, identical frames; a replica that does not start
    at the offset its leaf packs embed is refused."""
    from candela_b200 import api
    src = dragon2["ri"]
    dst = cb.RayIntersector(cb.STACKLESS)
    dst.CloneSceneFrom(src)
    dst.BufferData()
    dragon2["fill"](dst)
    a, b = src.read_buffers(), dst.read_buffers()
    assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
    W, H = 128, 96
    iv, ip = _camera(W, H)
    p = cb.frame_params(iv, ip, W, H, spp=2, seed=4, out_format=api.FRAME_OUT_HIT32)
    assert dst.TraceFrame(p).tobytes() == src.TraceFrame(p).tobytes()
    v = src.object_device_view(2)
    with pytest.raises(cb.CandelaError, match="leaf packs embed"):
        dst.AddPrebuiltObjectDevice(3, v["d_nodes"], v["n_nodes"], v["d_tris"], v["n_tris"], v["d_verts"], v["n_verts"], 0, 0)
    dst.close()


def test_multi_device_handle(cb, ob, dragon2):
    """cndl_multi_*: one handle over several GPUs; the frame equals the single-GPU frame bit for bit."""
    import torch
    from candela_b200 import api
    n_dev = torch.cuda.device_count()
    devices = tuple(range(n_dev)) if n_dev >= 2 else (0, 0)    # two contexts on one GPU still exercise sharding, peer stores and the gather
    mi = cb.MultiRayIntersector(cb.STACKLESS, devices)
    mi.AddObject(2, dragon2["V"], dragon2["F"].ravel(), dragon2["mids"])
    mi.BufferData()
    dragon2["fill"](mi)
    for i in range(len(devices)):
        assert all(x.tobytes() == y.tobytes() for x, y in zip(mi.context(i).read_buffers(), dragon2["ri"].read_buffers()))
    W, H = 200, 120
    iv, ip = _camera(W, H)
    for transport in (api.TRANSPORT_PEER_STORES, api.TRANSPORT_STAGED_COPY):
        mi.set_transport(transport)
        for fmt, kw in ((api.FRAME_OUT_HIT16, dict(spp=2, bounces=1)), (api.FRAME_OUT_PIXEL32, dict(spp=4, bounces=3))):
            p = cb.frame_params(iv, ip, W, H, seed=21, tile=32, out_format=fmt, **kw)
            want = dragon2["ri"].TraceFrame(p)
            got = mi.TraceFrame(p)
            assert got.tobytes() == want.tobytes(), (fmt, transport)
            assert mi.frame_rays_traced(0) == dragon2["ri"].frame_rays_traced(0)
    mi.close()


def test_frame_call_on_a_stack_format_context(cb, ob, dragon2):
    """The frame pipeline over FlattenedStackNode buffers (the stack-format traversal kernels read the generator's segmented batches and
    octant lists too): the oracle's frame for the stack format — same hits, different `iters` — in all three output formats."""
    from candela_b200 import api
    from oracle import frame as of
    sc = dragon2
    ri = cb.RayIntersector(cb.STACK)
    ri.AddObject(2, sc["V"], sc["F"].ravel(), sc["mids"])
    ri.BufferData()
    sc["fill"](ri)
    nodes, tris, _ = ri.read_buffers()
    ents = sc["ents"].copy()
    ents["node_count"] = len(nodes)
    W, H = 160, 96
    iv, ip = _camera(W, H)
    for fmt, kw in ((0, dict(spp=2, seed=5)), (1, dict(spp=2, seed=5)), (2, dict(spp=3, bounces=3, seed=11))):
        want, traced = of.trace_frame(ob.STACK, nodes, tris, sc["V"], ents, iv, ip, W, H, out_format=fmt, **kw)
        for octant, compact in ((False, False), (True, False), (True, True)):
            got = ri.TraceFrame(cb.frame_params(iv, ip, W, H, tile=32, out_format=fmt, octant_order=octant, compact_rays=compact, **kw))
            assert got.tobytes() == want.tobytes(), (fmt, octant, compact)
            assert ri.frame_rays_traced(0) == traced
    ri.close()
