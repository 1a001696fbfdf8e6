"""GPU parity tests (run with -m gpu on a B200): every query goes through the C ABI of
libcandela_b200.so and is compared with the CPU oracle on the same buffers and rays.
Bar (BASELINE.json): hit triangle / mesh / entity indices bit-exact; t and barycentrics within 1e-5
relative.  The kernels are written to be bit-identical, and that stronger property is asserted too."""
import numpy as np
import pytest

from cases import build_cases
from conftest import GOLDEN
from helpers import assert_hits_equal, bit_identical

pytestmark = pytest.mark.gpu

FORMATS = ["stackless", "stack"]


def fmt_id(ob, name):
    return ob.STACKLESS if name == "stackless" else ob.STACK


def gpu_scene(cb, sc):
    """Hands an oracle-built scene to the GPU path as prebuilt reference-layout buffers."""
    ri = cb.RayIntersector(sc.format)
    for oid, o in sc.objects.items():
        n0, n1 = o["node_offset"], o["node_offset"] + o["node_count"]
        t0, t1 = o["tri_offset"], o["tri_offset"] + o["tri_count"]
        v0, v1 = o["vert_offset"], o["vert_offset"] + o["vert_count"]
        tris = sc.tris[t0:t1].copy()
        tris["v"] -= v0                                  # back to object-local indices
        ri.AddPrebuiltObject(oid, sc.nodes[n0:n1], tris, sc.verts[v0:v1])
    ri.BufferData(True)
    ri.PushEntityRecords(sc.entities)
    ri.BufferEntities()
    return ri


@pytest.fixture(scope="module")
def all_cases(ob, golden_meshes):
    return {f: build_cases(ob, golden_meshes, fmt_id(ob, f)) for f in FORMATS}


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("fmt", FORMATS)
def test_closest_and_any_parity_all_cases(cb, ob, all_cases, fmt, mode):
    for c in all_cases[fmt]:
        sc, rays = c["scene"], c["rays"]
        ri = gpu_scene(cb, sc)
        ri.set_traversal_mode(mode)
        nodes, tris, verts = ri.read_buffers()
        assert nodes.tobytes() == sc.nodes.tobytes() and tris.tobytes() == sc.tris.tobytes() and verts.tobytes() == sc.verts.tobytes(), c["name"]
        for kind, ign in ((ob.CLOSEST, False), (ob.CLOSEST_IGNORE_TRANSPARENT, True)):
            ref, _ = sc.trace(kind, rays, nthreads=8)
            got = ri.IntersectRays(rays, ignore_transparent=ign)
            assert_hits_equal(got, ref)
            assert bit_identical(got, ref), (c["name"], fmt, kind)
        for tmax in (0.0, 2.4):
            r2 = rays.copy()
            r2["tmax"] = tmax
            ref_t, _ = sc.trace(ob.ANY, r2, nthreads=8)
            got_t = ri.IntersectRaysAny(r2)
            assert np.array_equal(got_t > 0, ref_t > 0), (c["name"], fmt, tmax)
            assert got_t.tobytes() == ref_t.tobytes(), (c["name"], fmt, tmax)
        ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_random_scene_sweep_parity(cb, ob, fmt):
    """cases.random_scene (random meshes, mirrored / sheared / scaled instances, translucent entities, unnormalised directions,
    rays from far outside) — the scenes tests/test_oracle_traversal.py pins to the compiled reference GLSL: every record
    bit-identical to the oracle, default kernels and the one-thread-per-ray kernels."""
    import cases
    n_hit = 0
    for seed in range(6):
        sc, rays = cases.random_scene(ob, fmt_id(ob, fmt), seed)
        ri = gpu_scene(cb, sc)
        for mode in (2, 0):
            ri.set_traversal_mode(mode)
            for kind, tmax in cases.random_scene_queries(ob, seed):
                r = rays.copy()
                r["tmax"] = tmax
                ref, _ = sc.trace(kind, r, nthreads=8)
                if kind == ob.ANY:
                    got = ri.IntersectRaysAny(r)
                else:
                    got = ri.IntersectRays(r, ignore_transparent=(kind == ob.CLOSEST_IGNORE_TRANSPARENT))
                    assert_hits_equal(got, ref)
                    n_hit += int(np.count_nonzero(ref["t"] > 0))
                assert got.tobytes() == ref.tobytes(), (seed, fmt, mode, kind, tmax)
        ri.close()
    assert n_hit > 5000


@pytest.mark.parametrize("fmt", FORMATS)
def test_committed_golden_vectors(cb, ob, golden_meshes, fmt):
    z = np.load(GOLDEN / "traversal_golden.npz")
    P, F = golden_meshes["dragon"]
    sc = ob.Scene(fmt_id(ob, fmt))
    sc.add_object(2, ob.make_vertices(P), F.ravel(), np.zeros(len(F), np.int32))
    sc.push_entity(2)
    ri = gpu_scene(cb, sc)
    assert ri.IntersectRays(z["rays"]).tobytes() == z[f"{fmt}_hits"].tobytes()
    assert ri.IntersectRaysAny(z["rays"]).tobytes() == z[f"{fmt}_any"].tobytes()
    ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_gpu_matches_committed_reference_outputs(cb, ob, golden_meshes, fmt):
    """The CUDA path against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/reference_traversal_golden.npz was produced by the
    reference's GLSL traversal files compiled against its own glm (tests/golden/make_reference_traversal_golden.py)."""
    import reference_cases
    z = np.load(GOLDEN / "reference_traversal_golden.npz")
    for c in reference_cases.build(ob, golden_meshes, fmt_id(ob, fmt)):
        ri = gpu_scene(cb, c["scene"])
        rays = c["rays"]
        for kname, kind, tmax in reference_cases.KINDS:
            r = reference_cases.with_tmax(rays, tmax)
            if kind == 2:
                got = ri.IntersectRaysAny(r)
            else:
                got = ri.IntersectRays(r, ignore_transparent=(kind == 1))
            assert got.tobytes() == z[f"{fmt}/{c['name']}/{kname}"].tobytes(), (fmt, c["name"], kname)
        if fmt == "stackless":
            hits = ri.IntersectRays(rays)
            got = ri.GetData(hits).view(np.float32).reshape(-1, 8)
            want = z[f"{fmt}/{c['name']}/get_data"]
            # these fixture meshes carry zero normals: normalize(0) is NaN on both sides, with different payloads
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(got), nan) and got[~nan].tobytes() == want[~nan].tobytes(), (c["name"], "GetData")
        ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_primary_rays_parity(cb, ob, golden_meshes, fmt):
    from candela_b200 import scenes
    P, F = golden_meshes["dragon"]
    sc = ob.Scene(fmt_id(ob, fmt))
    sc.add_object(2, ob.make_vertices(P), F.ravel(), np.zeros(len(F), np.int32))
    sc.push_entity(2)
    ri = gpu_scene(cb, sc)
    W, H = 320, 180
    iv, ip = scenes.camera((-9.0, 6.0, 7.0), (0.0, 4.0, 0.0), W, H)
    hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
    ref_rays = ob.primary_rays(iv, ip, W, H)
    assert rays.tobytes() == ref_rays.tobytes()
    ref, _ = sc.trace(ob.CLOSEST, ref_rays, nthreads=8)
    assert hits.tobytes() == ref.tobytes()
    assert (hits["t"] > 0).mean() > 0.05
    ri.close()


def test_errors_match_the_reference(cb):
    ri = cb.RayIntersector(cb.STACKLESS)
    with pytest.raises(cb.CandelaError, match="parent object hasn't been added"):
        ri.PushEntity(5)
    with pytest.raises(cb.CandelaError):
        ri.IntersectRays(np.zeros(4, dtype=cb.api.RAY_DT))   # nothing committed
    ri.close()


def test_empty_batch_and_ragged_sizes(cb, ob, golden_meshes):
    P, F = golden_meshes["soup400"]
    sc = ob.Scene(ob.STACKLESS)
    sc.add_object(2, ob.make_vertices(P), F.ravel(), None)
    sc.push_entity(2)
    ri = gpu_scene(cb, sc)
    assert len(ri.IntersectRays(np.zeros(0, dtype=cb.api.RAY_DT))) == 0
    from helpers import rays_in_box
    for n in (1, 31, 33, 127, 129, 1000):
        rays = rays_in_box(P.min(0), P.max(0), n, n)
        ref, _ = sc.trace(ob.CLOSEST, rays)
        assert ri.IntersectRays(rays).tobytes() == ref.tobytes()
    ri.close()


def test_full_size_diffuse_batch_s260k(cb, ob):
    """BASELINE config 2 at full size: 1920x1080 primary rays -> cosine-hemisphere diffuse rays on the
    ~260k scene; the oracle runs multi-threaded on the same rays (a few seconds)."""
    from candela_b200 import scenes
    v, i, m = scenes.make_s260k()
    sc = ob.Scene(ob.STACKLESS)
    sc.add_object(2, v, i, m)
    sc.push_entity(2)
    ri = gpu_scene(cb, sc)
    W, H = 1920, 1080
    iv, ip = scenes.camera(**scenes.S260K_CAMERA, width=W, height=H)
    hits, rays = ri.IntersectPrimary(iv, ip, W, H, return_rays=True)
    nodes, tris, verts = ri.read_buffers()
    drays, _ = scenes.bounce_rays(rays, hits, tris, verts, seed=2)
    assert len(drays) > 0.99 * W * H
    got = ri.IntersectRays(drays, ignore_transparent=True)
    ref, c = sc.trace(ob.CLOSEST_IGNORE_TRANSPARENT, drays, nthreads=ob.hardware_threads())
    assert_hits_equal(got, ref)
    assert bit_identical(got, ref)
    short = drays.copy()
    short["tmax"] = 2.4
    ref_t, _ = sc.trace(ob.ANY, short, nthreads=ob.hardware_threads())
    assert ri.IntersectRaysAny(short).tobytes() == ref_t.tobytes()
    ri.close()
