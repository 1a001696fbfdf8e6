"""GPU tests of the pieces around the traversal kernels: wavefront bounce-ray generation (compaction),
size-independent properties at BASELINE sizes, tuning knobs (which must never change a result), and the
host-buffer pipeline."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def s260k(cb, ob):
    from candela_b200 import scenes
    v, i, m = scenes.make_s260k()
    ri = cb.RayIntersector(cb.STACKLESS)
    ri.AddObject(2, v, i, m)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    nodes, tris, _ = ri.read_buffers()
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
    yield dict(ri=ri, v=v, nodes=nodes, tris=tris, ents=ents)
    ri.close()


def test_bounce_ray_generation_compacts_and_is_deterministic(cb, ob, s260k):
    import torch
    from candela_b200 import api, scenes
    ri = s260k["ri"]
    W, H, spp = 640, 360, 3
    iv, ip = scenes.camera((-18.0, 5.0, 0.7), (10.0, 30.0, -0.4), W, H)     # looks partly at the ceiling/out: some misses
    hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
    hits = hits.copy()
    hits["t"][::7] = -1.0                                                     # force misses into the batch
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.from_numpy(hits.view(np.float32).reshape(-1, 8)).cuda()
    outs = []
    for _ in range(2):
        d_out = torch.zeros((W * H * spp, 8), dtype=torch.float32, device="cuda")
        d_par = torch.zeros(W * H * spp, dtype=torch.int32, device="cuda")
        n = ri.generate_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), W * H, d_out.data_ptr(), spp=spp, offset=0.05, tmax=2.4, seed=9,
                                           d_parent_out=d_par.data_ptr())
        outs.append((n, d_out[:n].cpu().numpy().view(api.RAY_DT).reshape(-1), d_par[:n].cpu().numpy()))
    n, out, par = outs[0]
    ok = hits["t"] > 0
    assert n == spp * int(ok.sum())
    assert np.array_equal(par, np.repeat(np.nonzero(ok)[0], spp))             # compacted in input order
    assert outs[1][0] == n and outs[1][1].tobytes() == out.tobytes()          # deterministic
    assert np.all(out["tmax"] == np.float32(2.4))
    assert np.allclose(np.linalg.norm(out["d"], axis=1), 1.0, atol=1e-5)
    # origin = hit point + 0.05 * normal, direction in the normal's hemisphere
    P = rays["o"][par] + rays["d"][par] * hits["t"][par][:, None]
    N = (out["o"] - P) / 0.05
    assert np.allclose(np.linalg.norm(N, axis=1), 1.0, atol=2e-3)
    assert np.all(np.sum(N * out["d"], axis=1) > -1e-3)
    assert np.all(np.sum(N * rays["d"][par], axis=1) < 1e-3)                  # turned against the incoming ray
    # the generated rays trace identically on GPU and oracle
    want, _ = ob.trace(ob.STACKLESS, ob.ANY, s260k["nodes"], s260k["tris"], s260k["v"], s260k["ents"], out, nthreads=ob.hardware_threads())
    assert ri.IntersectRaysAny(out).tobytes() == want.tobytes()


def _pcg(v):
    v = np.asarray(v, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    s = (v * np.uint64(747796405) + np.uint64(2891336453)) & np.uint64(0xFFFFFFFF)
    w = (((s >> ((s >> np.uint64(28)) + np.uint64(4))) ^ s) * np.uint64(277803737)) & np.uint64(0xFFFFFFFF)
    return ((w >> np.uint64(22)) ^ w) & np.uint64(0xFFFFFFFF)


def _u01(h):
    return ((h >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def _specular_reference(I, N, rough_pbr, seed, idx):
    """StochasticReflectionDirection restated (SpecularTrace.glsl:102-135 + Include/Sampling.glsl:63-83), spp = 1,
    with the generator's counter-based xi instead of the shader's fract(sin()) hash."""
    rough = np.float32(rough_pbr) * np.float32(0.9)
    k = _pcg(np.uint64(seed) ^ _pcg(idx))
    M = N.copy()
    if rough >= 0.01:
        a2 = np.float32(rough * rough) ** 2
        up = np.where((np.abs(N[:, 2]) < 0.999)[:, None], np.array([0, 0, 1], np.float32), np.array([1, 0, 0], np.float32))
        T = np.cross(up, N)
        T /= np.linalg.norm(T, axis=1, keepdims=True)
        B = np.cross(N, T)
        found = np.zeros(len(N), bool)
        for t in range(12):
            x1 = _u01(_pcg(k + np.uint64((2 * t + 1) * 0x9E3779B9))) * np.float32(0.8)
            x2 = _u01(_pcg(k + np.uint64((2 * t + 2) * 0x9E3779B9))) * np.float32(0.7)
            phi = np.float32(2 * np.pi) * x1
            ct = np.sqrt((1 - x2) / (1 + (a2 - 1) * x2))
            st = np.sqrt(np.maximum(1 - ct * ct, 0))
            S = T * (np.cos(phi) * st)[:, None] + B * (np.sin(phi) * st)[:, None] + N * ct[:, None]
            S /= np.linalg.norm(S, axis=1, keepdims=True)
            ok = (np.sum(S * N, axis=1) > 0.001) & ~found
            M[ok] = S[ok]
            found |= ok
    D = I - 2 * np.sum(M * I, axis=1, keepdims=True) * M
    return D / np.linalg.norm(D, axis=1, keepdims=True)


def test_specular_shadow_and_bucketed_generation(cb, ob, s260k):
    """cndl_generate_rays_device: specular rays against a numpy restatement of the reference's sampler, shadow rays,
    and the octant-major emission order (a stable partition of the plain order)."""
    import torch
    from candela_b200 import api, scenes
    ri = s260k["ri"]
    W, H = 640, 360
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.from_numpy(hits.view(np.float32).reshape(-1, 8)).cuda()
    R = W * H
    ok = np.nonzero(hits["t"] > 0)[0]

    def run(kind, **kw):
        spp = kw.get("spp", 1)
        d_out = torch.zeros((R * spp, 8), dtype=torch.float32, device="cuda")
        d_par = torch.zeros(R * spp, dtype=torch.int32, device="cuda")
        n = ri.generate_rays_device(kind, d_rays.data_ptr(), d_hits.data_ptr(), R, d_out.data_ptr(), d_parent_out=d_par.data_ptr(), **kw)
        return d_out[:n].cpu().numpy().view(api.RAY_DT).reshape(-1), d_par[:n].cpu().numpy()

    # geometric normal and hit point, as the generator defines them
    tv = s260k["tris"]["v"][hits["tri"][ok]]
    Pv = s260k["v"]["position"][:, :3]
    N = np.cross(Pv[tv[:, 1]] - Pv[tv[:, 0]], Pv[tv[:, 2]] - Pv[tv[:, 0]]).astype(np.float32)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    I = rays["d"][ok]
    flip = np.sum(N * I, axis=1) > 0
    N[flip] = -N[flip]
    P = rays["o"][ok] + I * hits["t"][ok][:, None]

    for rough in (0.0, 0.3, 0.8):
        out, par = run(api.GEN_SPECULAR, roughness=rough, offset=-1.0, seed=11)
        assert len(out) == len(ok) and np.array_equal(par, ok)
        want = _specular_reference(I, N, rough, 11, ok.astype(np.uint64))
        assert np.abs(out["d"] - want).max() < 2e-4, rough
        off = 0.05 + 0.05 * min(max(rough * 1.4, 0.0), 1.0)
        assert np.abs(out["o"] - (P + N * np.float32(off))).max() < 1e-4
    mirror, _ = run(api.GEN_SPECULAR, roughness=0.0, offset=0.05)
    assert np.abs(np.sum(mirror["d"] * N, axis=1) + np.sum(I * N, axis=1)).max() < 1e-5        # angle out = angle in

    L = np.array([0.3, 0.9, 0.2], np.float32)
    L /= np.linalg.norm(L)
    sh, par = run(api.GEN_SHADOW, light_dir=tuple(float(x) for x in L), light_cone=0.0, tmax=200.0, offset=0.02)
    lit = np.sum(N * L, axis=1) > 0
    assert np.array_equal(par, ok[lit]) and np.abs(sh["d"] - L).max() < 1e-6 and np.all(sh["tmax"] == np.float32(200.0))
    soft, _ = run(api.GEN_SHADOW, light_dir=tuple(float(x) for x in L), light_cone=0.05, spp=2, seed=4)
    cosang = np.sum(soft["d"] * L, axis=1)
    assert len(soft) == 2 * int(lit.sum()) and cosang.min() > np.cos(np.arcsin(0.05)) - 1e-4 and cosang.std() > 0

    # octant-major emission = stable partition of the plain emission by the direction signs
    plain, ppar = run(api.GEN_DIFFUSE, spp=3, seed=9)
    buck, bpar = run(api.GEN_DIFFUSE, spp=3, seed=9, bucket_octants=True)
    octant = (plain["d"][:, 0] > 0).astype(np.int64) | ((plain["d"][:, 1] > 0).astype(np.int64) << 1) | ((plain["d"][:, 2] > 0).astype(np.int64) << 2)
    order = np.argsort(octant, kind="stable")
    assert buck.tobytes() == plain[order].tobytes() and np.array_equal(bpar, ppar[order])
    # and the traversal of the bucketed batch gives the same records, permuted
    a = ri.IntersectRays(plain)
    b = ri.IntersectRays(buck)
    assert b.tobytes() == a[order].tobytes()


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
    for variant, sort in ((18, 1), (34, 1), (42, 1), (18, 2), (34, 2)):
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

