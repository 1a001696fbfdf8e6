"""GPU tests of the pieces around the traversal kernels: wavefront bounce-ray generation (compaction),
size-independent properties at BASELINE sizes, tuning knobs (which must never change a result), and the
host-buffer pipeline."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen_inputs(ri, W, H, cam=None, knock_out=0):
    import torch
    from candela_b200 import scenes
    iv, ip = scenes.camera(*cam, W, H) if cam else scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
    hits = hits.copy()
    if knock_out:
        hits["t"][::knock_out] = -1.0                                          # force misses into the batch
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.from_numpy(hits.view(np.float32).reshape(-1, 8)).cuda()
    return rays, hits, d_rays, d_hits


def _run_gen(ri, kind, d_rays, d_hits, R, d_ids=None, **kw):
    import torch
    from candela_b200 import api
    spp = kw.get("spp", 1)
    d_out = torch.zeros((R * spp, 8), dtype=torch.float32, device="cuda")
    d_par = torch.zeros(R * spp, dtype=torch.int32, device="cuda")
    d_ido = torch.zeros(R * spp, dtype=torch.int32, device="cuda")
    n = ri.generate_rays_device(kind, d_rays.data_ptr(), d_hits.data_ptr(), R, d_out.data_ptr(), d_parent_out=d_par.data_ptr(),
                                d_ids_in=0 if d_ids is None else d_ids.data_ptr(), d_ids_out=d_ido.data_ptr(), **kw)
    return d_out[:n].cpu().numpy().view(api.RAY_DT).reshape(-1), d_par[:n].cpu().numpy().view(np.uint32), d_ido[:n].cpu().numpy().view(np.uint32)


def test_generated_rays_are_bit_identical_to_the_oracle(cb, ob, s260k):
    """cndl_generate_rays_device == oracle/oracle_raygen.cpp (itself pinned against the compiled shader functions,
    tests/test_raygen_oracle.py) for every kind, with misses in the batch, several samples per hit, the octant-major
    order and caller-supplied stream ids: rays, parents and ids, bit for bit."""
    import torch
    from candela_b200 import api
    ri = s260k["ri"]
    W, H = 640, 360
    rays, hits, d_rays, d_hits = _gen_inputs(ri, W, H, cam=((-18.0, 5.0, 0.7), (10.0, 30.0, -0.4)), knock_out=7)
    R = W * H
    L = np.array([0.3, 0.9, 0.2], np.float32)
    L /= np.linalg.norm(L)
    light = tuple(float(x) for x in L)
    perm_ids = torch.from_numpy(np.random.default_rng(2).permutation(R).astype(np.int32)).cuda()
    cases = [dict(kind=api.GEN_DIFFUSE, spp=3, offset=0.05, tmax=2.4, seed=9),
             dict(kind=api.GEN_DIFFUSE, spp=1, offset=0.02, seed=31, bucket_octants=True),
             dict(kind=api.GEN_DIFFUSE, spp=2, seed=5, ids=True),
             dict(kind=api.GEN_SPECULAR, roughness=0.0, offset=-1.0, seed=11),
             dict(kind=api.GEN_SPECULAR, roughness=0.3, offset=-1.0, seed=11),
             dict(kind=api.GEN_SPECULAR, roughness=0.8, offset=-1.0, seed=12, bucket_octants=True),
             dict(kind=api.GEN_SHADOW, light_dir=light, light_cone=0.0, tmax=200.0, offset=0.02),
             dict(kind=api.GEN_SHADOW, light_dir=light, light_cone=0.05, spp=2, seed=4)]
    for c in cases:
        c = dict(c)
        use_ids = c.pop("ids", False)
        kind = c.pop("kind")
        got = _run_gen(ri, kind, d_rays, d_hits, R, d_ids=perm_ids if use_ids else None, **c)
        want = ob.generate_rays(rays, hits, s260k["tris"], s260k["v"], s260k["ents"], kind=kind,
                                ids=perm_ids.cpu().numpy().view(np.uint32) if use_ids else None, **c)
        assert len(got[0]) == len(want[0]) > 0, c
        assert got[0].tobytes() == want[0].tobytes(), c
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), c
    # compaction in input order; deterministic; the generated rays trace identically on GPU and oracle
    out, par, _ = _run_gen(ri, api.GEN_DIFFUSE, d_rays, d_hits, R, spp=3, offset=0.05, tmax=2.4, seed=9)
    ok = hits["t"] > 0
    assert len(out) == 3 * int(ok.sum()) and np.array_equal(par, np.repeat(np.nonzero(ok)[0], 3))
    assert _run_gen(ri, api.GEN_DIFFUSE, d_rays, d_hits, R, spp=3, offset=0.05, tmax=2.4, seed=9)[0].tobytes() == out.tobytes()
    want, _ = ob.trace(ob.STACKLESS, ob.ANY, s260k["nodes"], s260k["tris"], s260k["v"], s260k["ents"], out, nthreads=ob.hardware_threads())
    assert ri.IntersectRaysAny(out).tobytes() == want.tobytes()


def test_probe_rays_equal_the_compiled_shader(cb, ob, s260k):
    """cndl_generate_probe_rays_device for the reference's 48 x 24 x 48 probe grid (Macros.h:22-24): bit-identical to the
    oracle, and its directions to the outputs of the compiled ImportanceSample (tests/golden/raygen_golden.npz)."""
    import torch
    from candela_b200 import api
    from conftest import GOLDEN
    from golden.make_raygen_golden import PROBE_RES, PROBE_SEED
    ri = s260k["ri"]
    n = PROBE_RES[0] * PROBE_RES[1] * PROBE_RES[2]
    d = torch.zeros((n, 8), dtype=torch.float32, device="cuda")
    org, size = (0.5, 6.0, -0.25), (24.0, 12.0, 24.0)                         # PROBE_GRID_SIZE, Macros.h:26-29
    ri.generate_probe_rays_device(org, size, PROBE_RES, PROBE_SEED, d.data_ptr())
    torch.cuda.synchronize()
    got = d.cpu().numpy().view(api.RAY_DT).reshape(-1)
    want = ob.probe_rays(org, size, PROBE_RES, PROBE_SEED)
    assert got.tobytes() == want.tobytes()
    gold = np.load(GOLDEN / "raygen_golden.npz")["probe_grid_directions"]
    assert np.ascontiguousarray(got["d"][::9]).tobytes() == gold.tobytes()
    # and they trace identically
    hits = ri.IntersectRays(got)
    ref, _ = ob.trace(ob.STACKLESS, ob.CLOSEST, s260k["nodes"], s260k["tris"], s260k["v"], s260k["ents"], got, nthreads=ob.hardware_threads())
    assert hits.tobytes() == ref.tobytes()


def test_knobs_and_modes_never_change_results(cb, ob, s260k):
    from helpers import rays_in_box
    ri = s260k["ri"]
    rays = rays_in_box((-20, 0, -9), (20, 14, 9), 300000, 21)
    want, _ = ob.trace(ob.STACKLESS, ob.CLOSEST, s260k["nodes"], s260k["tris"], s260k["v"], s260k["ents"], rays, nthreads=ob.hardware_threads())
    for mode in (0, 1, 2):
        ri.set_traversal_mode(mode)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), mode
    for knobs in ((7, 1, 1, 1), (8, 32, 32, 4), (10, 3, 17, 3), (12, 8, 8, 9), (8, 8, 8, 10), (8, 8, 8, 2)):
        for k, val in enumerate(knobs):
            ri.set_tuning(k, val)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), knobs
    # the shared-memory staged kernel (hot-first derived node layout): variants 33..35, staged-node counts, CTA sizes
    for hot, block, variant in ((2048, 1024, 34), (1, 256, 33), (511, 512, 35), (7168, 1024, 34), (100000, 256, 34)):
        ri.set_tuning(6, hot)
        ri.set_tuning(7, block)
        ri.set_tuning(3, variant)
        ri.BufferData()
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), (hot, block, variant)
    # two rays per lane over the derived layout (variants 41..43), 4 / 5 / 6 CTAs per SM
    for bps, park, idle, variant in ((5, 16, 16, 42), (4, 1, 1, 41), (6, 24, 40, 43), (5, 64, 64, 42)):
        for k, val in enumerate((bps, park, idle, variant)):
            ri.set_tuning(k, val)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), (bps, park, idle, variant)
    for k, val in enumerate((8, 12, 8, 18)):
        ri.set_tuning(k, val)
    # ray bucketing by direction octant inside the call (every kernel family reads its rays through the bucket lists)
    # (sort_rays 1) or sorted by octant + origin Morton code (sort_rays 2)
    for variant, sort in ((18, 1), (34, 1), (42, 1), (18, 2), (34, 2), (18, 3), (42, 3)):
        ri.set_tuning(3, variant)
        ri.set_traversal_mode(2, sort)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), ("reordered", variant, sort)
        short = rays.copy()
        short["tmax"] = 3.0
        t_b = ri.IntersectRaysAny(short)
        ri.set_traversal_mode(2, 0)
        assert t_b.tobytes() == ri.IntersectRaysAny(short).tobytes(), ("reordered any-hit", variant, sort)
    ri.set_tuning(3, 18)
    for chunks in (1, 3, 16):
        ri.set_tuning(4, chunks)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), chunks
    ri.set_tuning(4, 0)


def test_stack_kernel_knobs_and_bucketing(cb, ob):
    """Stack format: the while-while kernel's thresholds / steps and ray bucketing never change a result."""
    from candela_b200 import scenes
    from helpers import rays_in_box
    P, F = scenes.load_dragon()
    V = cb.make_vertices(P)
    ri = cb.RayIntersector(cb.STACK)
    ri.AddObject(2, V, F.ravel(), np.zeros(len(F), np.int32))
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    nodes, tris, _ = ri.read_buffers()
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
    rays = rays_in_box(P.min(0), P.max(0), 200000, 33)
    want, _ = ob.trace(ob.STACK, ob.CLOSEST, nodes, tris, V, ents, rays, nthreads=ob.hardware_threads())
    for park, idle, steps, sort in ((12, 8, 1, 0), (1, 1, 2, 0), (32, 32, 2, 1), (5, 20, 1, 2)):
        ri.set_tuning(5, park)
        ri.set_tuning(2, idle)
        ri.set_tuning(3, steps)
        ri.set_traversal_mode(2, sort)
        assert ri.IntersectRays(rays).tobytes() == want.tobytes(), (park, idle, steps, sort)
    ri.close()


def test_full_size_properties_rtao(cb, s260k):
    """BASELINE configs[2] at full size (8.3 M any-hit rays): determinism, permutation equivariance,
    consistency of any-hit with closest-hit, monotonicity in tmax."""
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
    n = ri.generate_bounce_rays_device(d_prim.data_ptr(), d_hits.data_ptr(), W * H, d_ao.data_ptr(), spp=spp, tmax=2.4, seed=3, stream=stream)
    assert n > 8_000_000
    d_ao = d_ao[:n]
    t1 = torch.empty(n, dtype=torch.float32, device="cuda")
    t2 = torch.empty(n, dtype=torch.float32, device="cuda")
    ri.intersect_any_device(d_ao.data_ptr(), n, t1.data_ptr(), stream)
    ri.intersect_any_device(d_ao.data_ptr(), n, t2.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(t1, t2)                                                 # idempotent / deterministic
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    d_p = d_ao[perm].contiguous()
    ri.intersect_any_device(d_p.data_ptr(), n, t2.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(t2, t1[perm])                                           # a ray's result does not depend on its neighbours
    assert bool(((t1 > 0) <= (t1 < 2.4)).all()) and bool(((t1 == -1) | (t1 > 0)).all())
    # closest hit with t < 2.4  <=>  some hit with t < 2.4 (the walks differ only in how TMax evolves)
    hc = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    ri.intersect_closest_device(d_ao.data_ptr(), n, hc.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    tc = hc[:, 0]
    agree = ((tc > 0) & (tc < 2.4)) == (t1 > 0)
    assert float(agree.float().mean()) > 0.99999
    assert bool((tc[(t1 > 0) & agree] <= t1[(t1 > 0) & agree]).all())          # the closest hit is never farther than the first found
    # a longer ray can only find more
    d_long = d_ao.clone()
    d_long[:, 7] = 10.0
    ri.intersect_any_device(d_long.data_ptr(), n, t2.data_ptr(), stream)
    torch.cuda.synchronize()
    assert bool(((t1 > 0) <= (t2 > 0)).all())


def test_two_devices_in_one_process(cb, ob):
    """One context per device, both in this process: kernel attributes (carve-out, dynamic shared memory) are configured per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from candela_b200 import scenes
    from helpers import rays_in_box
    P, F = scenes.load_dragon()
    V = cb.make_vertices(P)
    rays = rays_in_box(P.min(0), P.max(0), 150000, 44)
    outs = []
    for dev in (0, 1):
        ri = cb.RayIntersector(cb.STACKLESS, device=dev)
        ri.AddObject(2, V, F.ravel(), np.zeros(len(F), np.int32))
        ri.BufferData()
        ri.PushEntity(2)
        ri.BufferEntities()
        res = [ri.IntersectRays(rays).tobytes()]
        for variant in (34, 42):          # staged (dynamic shared memory) and paired kernels on this device too
            ri.set_tuning(3, variant)
            res.append(ri.IntersectRays(rays).tobytes())
        assert res[0] == res[1] == res[2]
        outs.append(res[0])
        ri.close()
    assert outs[0] == outs[1]

