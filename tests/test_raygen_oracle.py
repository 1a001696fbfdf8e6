"""The ray generators' oracle (oracle/oracle_raygen.cpp) against the reference's shader functions.

* live (where /root/reference exists): the restatement == the shader functions compiled from the reference
  (oracle/_ref, namespace ref_raygen), bit for bit;
* everywhere: the restatement == tests/golden/raygen_golden.npz (outputs of those compiled functions);
* the DEFINED sin / cos / acos / pow are correctly rounded to well within 1 ulp;
* orc_generate_rays' ordering rules (compaction in input order, octant-major = stable partition, id streams).
"""
import numpy as np
import pytest

from conftest import GOLDEN

GOLD = np.load(GOLDEN / "raygen_golden.npz")
CASES = [("cos_hemisphere", 0, ("normals", "xi"), 0.0), ("ggx_vndf_r0.27", 1, ("normals", "xi"), 0.27), ("ggx_vndf_r0.72", 1, ("normals", "xi"), 0.72),
         ("stochastic_reflection_r0.005", 2, ("normals", "incident", "keys"), 0.005), ("stochastic_reflection_r0.27", 2, ("normals", "incident", "keys"), 0.27),
         ("stochastic_reflection_r0.72", 2, ("normals", "incident", "keys"), 0.72), ("sample_cone_c0.98", 3, ("normals", "xi"), 0.98),
         ("probe_importance_sample", 4, ("keys",), 0.0)]


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("name,which,args,rough", CASES)
def test_samplers_equal_the_compiled_shader_outputs(ob, name, which, args, rough):
    kw = {a: GOLD[a] for a in args}
    mine = ob.sample_directions(which, roughness=rough, **kw)
    assert same_bits(mine, GOLD[name])
    if ob.REFERENCE_ROOT.exists():   # and against the compiled functions themselves, on fresh inputs
        rng = np.random.default_rng(hash(name) & 0xFFFF)
        n = 20000
        v = rng.normal(size=(n, 3)).astype(np.float32)
        fresh = dict(normals=(v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32),
                     incident=np.roll(v, 1, axis=0) / np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32),
                     xi=rng.random((n, 2), dtype=np.float32), keys=rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
        kw = {a: fresh[a] for a in args}
        assert same_bits(ob.sample_directions(which, roughness=rough, **kw), ob.ref_sample_directions(which, roughness=rough, **kw))


def test_probe_grid_directions_equal_the_compiled_importance_sample(ob):
    from golden.make_raygen_golden import PROBE_RES, PROBE_SEED
    rays = ob.probe_rays((1.0, 2.0, 3.0), (24.0, 12.0, 24.0), PROBE_RES, PROBE_SEED)
    assert len(rays) == 48 * 24 * 48                                  # PROBE_GRID, Macros.h:22-24
    assert same_bits(rays["d"][::9], GOLD["probe_grid_directions"])
    # origins: u_BoxOrigin + (Pixel / u_Resolution * 2 - 1) * u_Size, x fastest (UpdateRadianceProbes.glsl:412-417)
    idx = np.arange(len(rays))
    x, y, z = idx % 48, (idx // 48) % 24, idx // (48 * 24)
    want = np.stack([np.float32(1.0) + (x.astype(np.float32) / np.float32(48) * np.float32(2) - np.float32(1)) * np.float32(24.0),
                     np.float32(2.0) + (y.astype(np.float32) / np.float32(24) * np.float32(2) - np.float32(1)) * np.float32(12.0),
                     np.float32(3.0) + (z.astype(np.float32) / np.float32(48) * np.float32(2) - np.float32(1)) * np.float32(24.0)], axis=1).astype(np.float32)
    assert same_bits(rays["o"], want)
    assert np.all(rays["tmax"] == np.float32(1.0e6)) and np.allclose(np.linalg.norm(rays["d"], axis=1), 1.0, atol=1e-6)


def test_defined_transcendentals_are_correctly_rounded(ob):
    rng = np.random.default_rng(3)

    def ulp_err(got, exact):
        return np.abs(got.astype(np.float64) - exact) / np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
    x = (rng.random(400000, dtype=np.float32) * np.float32(6.2831855)).astype(np.float32)
    assert ulp_err(ob.xmath(0, x), np.sin(x.astype(np.float64))).max() < 0.5001
    assert ulp_err(ob.xmath(1, x), np.cos(x.astype(np.float64))).max() < 0.5001
    c = (rng.random(400000, dtype=np.float32) * 2 - 1).astype(np.float32)
    c[:5] = [-1.0, 1.0, 0.0, 0.5, -0.5]
    assert ulp_err(ob.xmath(2, c), np.arccos(c.astype(np.float64))).max() < 0.5001
    u = rng.random(400000, dtype=np.float32)
    y = np.full_like(u, np.float32(1) / np.float32(3))
    got = ob.xmath(3, u, y)
    ok = u > 0
    assert ulp_err(got[ok], np.power(u[ok].astype(np.float64), y[ok].astype(np.float64))).max() < 0.5001
    assert ob.xmath(3, np.float32([0.0]), np.float32([0.3]))[0] == 0.0
    # the counter stream: xi in [0, 1), exactly representable multiples of 2^-24
    s = ob.hash2_stream(int(ob.stream_keys(5, [17])[0]), 64)
    assert np.all((s >= 0) & (s < 1)) and np.all(s * np.float32(2 ** 24) == np.floor(s * np.float32(2 ** 24))) and len(np.unique(s)) > 60


@pytest.fixture(scope="module")
def small_scene(ob, golden_meshes):
    P, F = golden_meshes["dragon"]
    V = ob.make_vertices(P)
    sc = ob.Scene(ob.STACKLESS)
    sc.add_object(2, V, F.ravel(), np.zeros(len(F), np.int32))
    sc.push_entity(2)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (0.3, 0.0, 0.2)
    m[0, 0] = m[1, 1] = m[2, 2] = 1.5
    sc.push_entity(2, m)
    from helpers import rays_in_box
    lo, hi = P.min(0) - 0.2, P.max(0) * 1.5 + 0.4
    rays = rays_in_box(lo, hi, 6000, 11)
    hits, _ = sc.trace(ob.CLOSEST, rays)
    return sc, rays, hits


def test_generate_rays_ordering_rules(ob, small_scene):
    sc, rays, hits = small_scene
    ok = np.nonzero(hits["t"] > 0)[0]
    assert 500 < len(ok) < len(rays)
    out, par, ids = ob.generate_rays(rays, hits, sc.tris, sc.verts, sc.entities, kind=ob.GEN_DIFFUSE, spp=3, seed=9)
    assert len(out) == 3 * len(ok) and np.array_equal(par, np.repeat(ok, 3)) and np.array_equal(ids, np.repeat(ok, 3) * 3 + np.tile(np.arange(3), len(ok)))
    assert np.allclose(np.linalg.norm(out["d"], axis=1), 1.0, atol=1e-6) and np.all(out["tmax"] == np.float32(1.0e6))
    P = rays["o"][par] + rays["d"][par] * hits["t"][par][:, None]
    N = (out["o"] - P) / np.float32(0.05)
    assert np.allclose(np.linalg.norm(N, axis=1), 1.0, atol=5e-3)
    assert np.all(np.sum(N * out["d"], axis=1) > -1e-3) and np.all(np.sum(N * rays["d"][par], axis=1) < 1e-3)
    # octant-major = stable partition of the plain order
    b_out, b_par, b_ids = ob.generate_rays(rays, hits, sc.tris, sc.verts, sc.entities, kind=ob.GEN_DIFFUSE, spp=3, seed=9, bucket_octants=True)
    octant = (out["d"][:, 0] > 0).astype(np.int64) | ((out["d"][:, 1] > 0).astype(np.int64) << 1) | ((out["d"][:, 2] > 0).astype(np.int64) << 2)
    order = np.argsort(octant, kind="stable")
    assert b_out.tobytes() == out[order].tobytes() and np.array_equal(b_par, par[order]) and np.array_equal(b_ids, ids[order])
    # id streams: a permuted batch with ids naming the original slots produces the same rays, permuted
    perm = np.random.default_rng(1).permutation(len(rays))
    p_out, p_par, p_ids = ob.generate_rays(rays[perm], hits[perm], sc.tris, sc.verts, sc.entities, kind=ob.GEN_DIFFUSE, spp=3, seed=9, ids=perm.astype(np.uint32))
    back = np.argsort(p_ids, kind="stable")
    assert p_out[back].tobytes() == out.tobytes() and np.array_equal(p_ids[back], ids)
    # shadow rays: only lit hits emit; cone 0 gives the light direction itself
    L = np.array([0.3, 0.9, 0.2], np.float32)
    L /= np.linalg.norm(L)
    s_out, s_par, _ = ob.generate_rays(rays, hits, sc.tris, sc.verts, sc.entities, kind=ob.GEN_SHADOW, light_dir=tuple(float(v) for v in L), light_cone=0.0, tmax=200.0)
    assert 0 < len(s_out) < len(ok) and np.abs(s_out["d"] - L).max() < 1e-6 and np.all(s_out["tmax"] == np.float32(200.0))
    # mirror reflection: angle out = angle in
    m_out, m_par, _ = ob.generate_rays(rays, hits, sc.tris, sc.verts, sc.entities, kind=ob.GEN_SPECULAR, roughness=0.0, offset=0.05)
    Nm = (m_out["o"] - (rays["o"][m_par] + rays["d"][m_par] * hits["t"][m_par][:, None])) / np.float32(0.05)
    assert np.abs(np.sum(m_out["d"] * Nm, axis=1) + np.sum(rays["d"][m_par] * Nm, axis=1)).max() < 2e-2
