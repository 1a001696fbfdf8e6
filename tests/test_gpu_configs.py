"""BASELINE.json configs[2..4] under `pytest -m gpu`, at (or sampled from) their full sizes, bit for bit against the oracle:

  configs[2]  the full 8.29 M-ray RTAO any-hit batch (tmax 2.4, 4 spp at 1920x1080)
  configs[3]  one shard in sixteen of the 3840x2160 x 8 spp frame: every first-bounce hit record, and the resolved pixels of
              the 4-bounce frame
  configs[4]  a 2 M-triangle scene: GPU build byte-identical to the oracle's, 1 M random rays
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_full_rtao_batch_bit_identical(cb, ob, s260k):
    import torch
    from candela_b200 import api, scenes
    ri = s260k["ri"]
    W, H, spp = 1920, 1080, 4
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    stream = torch.cuda.current_stream().cuda_stream
    d_prim = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    d_hits = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    ri.intersect_primary_device(iv, ip, W, H, d_hits.data_ptr(), d_prim.data_ptr(), stream)
    d_ao = torch.empty((W * H * spp, 8), dtype=torch.float32, device="cuda")
    n = ri.generate_rays_device(api.GEN_DIFFUSE, d_prim.data_ptr(), d_hits.data_ptr(), W * H, d_ao.data_ptr(), spp=spp, offset=0.05, tmax=2.4, seed=7,
                                bucket_octants=True, stream=stream)
    assert n > 8_200_000
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    ri.intersect_any_device(d_ao.data_ptr(), n, d_t.data_ptr(), stream)
    torch.cuda.synchronize()
    rays = d_ao[:n].cpu().numpy().view(api.RAY_DT).reshape(-1)
    # the batch itself is the oracle generator's (octant-major), and every one of its 8.29 M any-hit distances is the oracle's
    prim = d_prim.cpu().numpy().view(api.RAY_DT).reshape(-1)
    phits = d_hits.cpu().numpy().view(api.HIT_DT).reshape(-1)
    want_rays, _, _ = ob.generate_rays(prim, phits, s260k["tris"], s260k["v"], s260k["ents"], kind=ob.GEN_DIFFUSE, spp=spp, seed=7, offset=0.05, tmax=2.4,
                                       bucket_octants=True)
    assert rays.tobytes() == want_rays.tobytes()
    want, _ = ob.trace(ob.STACKLESS, ob.ANY, s260k["nodes"], s260k["tris"], s260k["v"], s260k["ents"], rays, nthreads=ob.hardware_threads())
    got = d_t.cpu().numpy()
    assert got.tobytes() == want.tobytes()
    assert 0.05 < float((got > 0).mean()) < 0.95


def test_config3_one_shard_in_sixteen_of_the_4k_frame(cb, ob, s260k):
    from candela_b200 import api, scenes, sharding
    from oracle import frame as of
    ri = s260k["ri"]
    W, H, spp, shards, shard = 3840, 2160, 8, 16, 5
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    slots = sharding.shard_slots(W, H, shards, shard, 64)
    pixels = slots[slots >= 0]
    assert len(pixels) >= W * H // 16 - 64 * 64
    args = (ob.STACKLESS, s260k["nodes"], s260k["tris"], s260k["v"], s260k["ents"], iv, ip, W, H)
    # first bounce: all 8 spp hit records of the shard's pixels
    p = cb.frame_params(iv, ip, W, H, spp=spp, bounces=1, seed=4000, shard_index=shard, shard_count=shards, out_format=api.FRAME_OUT_HIT32, octant_order=True,
                        local_layout=True)
    got = ri.TraceFrame(p).reshape(-1, spp)
    want, traced = of.trace_frame(*args, spp=spp, bounces=1, seed=4000, out_format=of.OUT_HIT32, pixels=pixels)
    assert traced > 4_000_000 and ri.frame_rays_traced(0) == traced
    assert got[slots >= 0].tobytes() == want.reshape(-1, spp)[pixels].tobytes()
    assert np.all(got[slots < 0]["t"] == -1.0)
    # the 4-bounce frame resolved to pixels
    p = cb.frame_params(iv, ip, W, H, spp=spp, bounces=4, seed=4000, shard_index=shard, shard_count=shards, out_format=api.FRAME_OUT_PIXEL32, octant_order=True,
                        local_layout=True)
    got = ri.TraceFrame(p)
    want, traced = of.trace_frame(*args, spp=spp, bounces=4, seed=4000, out_format=of.OUT_PIXEL32, pixels=pixels)
    assert ri.frame_rays_traced(0) == traced > 15_000_000
    assert got[slots >= 0].tobytes() == want[pixels].tobytes()


def test_config4_two_million_triangles_build_and_trace(cb, ob):
    from candela_b200 import scenes
    v, i, m = scenes.make_heightfield(1001)
    assert len(i) // 3 == 2_000_000
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    nodes, tris, _ = ri.read_buffers()
    ref = ob.build(ob.STACKLESS, v, i, m)
    assert nodes.tobytes() == ref.nodes.tobytes() and tris.tobytes() == ref.tris.tobytes()
    pos = v["position"][:, :3]
    rays = scenes.random_rays(pos.min(0), pos.max(0) + np.array([0, 10, 0], np.float32), 1_000_000, seed=3)
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
    want, _ = ob.trace(ob.STACKLESS, ob.CLOSEST, nodes, tris, v, ents, rays, nthreads=ob.hardware_threads())
    for sort in (0, 2, 3, 4):       # no ordering, rays moved into (octant, origin cell) order, index list, automatic (= 3 here: the scene exceeds the L2)
        ri.set_traversal_mode(2, sort)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), sort
    assert float((want["t"] > 0).mean()) > 0.2
    ri.close()
