"""CPU tests of the traversal restatement.  The reference's traversal is GLSL with no tests and cannot
run here, so these are closed-form known answers, a brute-force scan, stack-vs-stackless agreement
and the reference quirks listed in SURVEY.md §7.3/§8a — and the committed oracle vectors."""
import numpy as np
import pytest

from cases import build_cases
from conftest import GOLDEN


@pytest.fixture(scope="module")
def cases_sl(ob, golden_meshes):
    return {c["name"]: c for c in build_cases(ob, golden_meshes, ob.STACKLESS)}


@pytest.fixture(scope="module")
def cases_st(ob, golden_meshes):
    return {c["name"]: c for c in build_cases(ob, golden_meshes, ob.STACK)}


@pytest.mark.parametrize("fmt_name", ["stackless", "stack"])
def test_known_answers(ob, cases_sl, cases_st, fmt_name):
    c = (cases_sl if fmt_name == "stackless" else cases_st)["known_answers"]
    hits, _ = c["scene"].trace(ob.CLOSEST, c["rays"])
    sc = c["scene"]
    # which sorted slot holds the z = 5 triangle?
    target = int(np.nonzero(sc.tris["v"][:, 0] == 0)[0][0])
    other = 1 - target
    h = hits[0] if target != 0 else hits[1]
    if target != 0:
        assert h["tri"] == target and h["mesh"] == 11 and h["entity"] == 0
        assert h["t"] == np.float32(5.0) and (h["u"], h["v"], h["w"]) == (0.25, 0.25, 0.5)
        # the other triangle is global triangle 0: accepted by the walk, reported as a miss (…Stackless.glsl:300)
        b = hits[1]
        assert b["tri"] == 0 and b["mesh"] == 12 and b["entity"] == 0 and b["t"] == -1 and b["u"] == -1
    else:
        assert hits[0]["tri"] == 0 and hits[0]["t"] == -1
        assert hits[1]["tri"] == other and hits[1]["t"] == np.float32(9.0)
    # pointing away / starting behind the triangle: nothing accepted
    for k in (2, 3):
        assert hits[k]["t"] == -1 and hits[k]["tri"] == -1 and hits[k]["mesh"] == -1 and hits[k]["entity"] == -1
    any_t, _ = sc.trace(ob.ANY, c["rays"])
    assert any_t[2] == -1 and any_t[3] == -1 and any_t[0] == 5.0


def test_brute_force_agreement(ob, cases_sl):
    c = cases_sl["dragon_random"]
    rays = c["rays"][:1500]
    sc = c["scene"]
    hits, _ = sc.trace(ob.CLOSEST, rays)
    bf = ob.brute_force(sc.tris, sc.verts, sc.entities, rays, nthreads=8)
    hit = bf["t"] > 0
    # the BVH walk may only lose a hit through its box tests; on this seed it loses none
    assert np.array_equal(hit, (hits["tri"] >= 0))
    assert np.array_equal(bf["t"][hit], np.where(hits["tri"][hit] > 0, hits["t"][hit], bf["t"][hit]))
    same = bf["tri"][hit] == hits["tri"][hit]
    assert same.mean() > 0.999  # exact ties may pick the other triangle


@pytest.mark.parametrize("name", ["dragon_random", "multi_entity", "coplanar_grid", "signed_zero"])
def test_stack_and_stackless_agree(ob, cases_sl, cases_st, name):
    a, _ = cases_sl[name]["scene"].trace(ob.CLOSEST, cases_sl[name]["rays"], nthreads=4)
    b, _ = cases_st[name]["scene"].trace(ob.CLOSEST, cases_st[name]["rays"], nthreads=4)
    agree = (a["tri"] == b["tri"]) & (a["t"] == b["t"]) & (a["entity"] == b["entity"])
    assert agree.mean() > 0.999, agree.mean()


def test_iteration_cap(ob, cases_sl):
    c = cases_sl["iteration_cap"]
    hits, cnt = c["scene"].trace(ob.CLOSEST, c["rays"])
    assert cnt["capped"] >= 1 and hits["iters"].max() == 1024


def test_ignore_transparent_and_any(ob, cases_sl):
    c = cases_sl["multi_entity"]
    sc, rays = c["scene"], c["rays"]
    full, _ = sc.trace(ob.CLOSEST, rays, nthreads=4)
    opaque, _ = sc.trace(ob.CLOSEST_IGNORE_TRANSPARENT, rays, nthreads=4)
    assert (full["entity"] == 2).any() and not (opaque["entity"] == 2).any()   # entity 2 has alpha 0.5
    assert (opaque["entity"] == 4).any()                                         # alpha 0.995 >= 0.99 stays
    any_t, _ = sc.trace(ob.ANY, rays, nthreads=4)
    assert np.array_equal(any_t > 0, full["tri"] >= 0)
    short = rays.copy()
    short["tmax"] = 2.4
    any_s, _ = sc.trace(ob.ANY, short, nthreads=4)
    assert np.all(any_s[any_s > 0] < 2.4) and (any_s > 0).sum() < (any_t > 0).sum()
    # closest t below the bound <=> some hit below the bound
    closest_t = np.where(full["tri"] >= 0, 1, 0)
    assert np.all((any_s > 0) <= (closest_t > 0))


def test_entity_without_object_raises(ob):
    sc = ob.Scene(ob.STACKLESS)
    with pytest.raises(KeyError, match="parent object"):
        sc.push_entity(99)


def test_make_entity_inverse(ob):
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = [[0.5, 0.1, 0], [0, 2, 0.3], [0.2, 0, 1.5]]
    m[:3, 3] = (1, -2, 3)
    e = ob.make_entity(m, 3, 9, emissive=1.5, translucency=0.25)[0]
    inv = e["inverse"].reshape(4, 4).T
    assert np.allclose(inv @ m, np.eye(4), atol=1e-6)
    assert e["node_offset"] == 3 and e["node_count"] == 9
    assert e["data"][:2].view(np.float32).tolist() == [1.5, 0.75]


def test_primary_rays_match_formula(ob):
    from candela_b200 import scenes
    iv, ip = scenes.camera((1, 2, 3), (4, 2, -1), 64, 36)
    r = ob.primary_rays(iv, ip, 64, 36)
    assert np.allclose(r["o"], (1, 2, 3)) and np.allclose(np.linalg.norm(r["d"], axis=1), 1, atol=1e-6)
    x, y = 40, 9
    clip = np.array([x / 64 * 2 - 1, y / 36 * 2 - 1, -1, 1], np.float64)
    eye = ip.astype(np.float64) @ clip
    d = iv.astype(np.float64) @ np.array([eye[0], eye[1], -1, 0])
    d = d[:3] / np.linalg.norm(d[:3])
    assert np.allclose(r["d"][y * 64 + x], d, atol=1e-6)


def test_committed_traversal_vectors(ob, golden_meshes):
    z = np.load(GOLDEN / "traversal_golden.npz")
    P, F = golden_meshes["dragon"]
    for fmt, label in ((ob.STACKLESS, "stackless"), (ob.STACK, "stack")):
        sc = ob.Scene(fmt)
        sc.add_object(2, ob.make_vertices(P), F.ravel(), np.zeros(len(F), np.int32))
        sc.push_entity(2)
        hits, c = sc.trace(ob.CLOSEST, z["rays"], nthreads=2)
        assert hits.tobytes() == z[f"{label}_hits"].tobytes()
        any_t, _ = sc.trace(ob.ANY, z["rays"])
        assert any_t.tobytes() == z[f"{label}_any"].tobytes()
        assert [c["node_iters"], c["tri_tests"], c["capped"], c["hits"]] == z[f"{label}_counters"].tolist()


# ---- the reference's own GLSL, compiled (oracle/_ref): this is what pins the restatement ----------------------------

def _reference_cases(ob, golden_meshes):
    import reference_cases
    for fmt, label in ((ob.STACKLESS, "stackless"), (ob.STACK, "stack")):
        for c in reference_cases.build(ob, golden_meshes, fmt):
            yield fmt, label, c


def test_oracle_matches_committed_reference_outputs(ob, golden_meshes):
    """tests/golden/reference_traversal_golden.npz holds outputs of the reference's GLSL traversal itself (compiled against
    its glm by oracle/ref_shim; made by tests/golden/make_reference_traversal_golden.py).  Runs everywhere, also on the GPU box."""
    import reference_cases
    z = np.load(GOLDEN / "reference_traversal_golden.npz")
    n = 0
    for fmt, label, c in _reference_cases(ob, golden_meshes):
        sc, rays = c["scene"], c["rays"]
        assert rays.view(np.float32).reshape(-1, 8).tobytes() == z[f"{label}/{c['name']}/rays"].tobytes(), "fixture made from other rays"
        for kname, kind, tmax in reference_cases.KINDS:
            got, _ = sc.trace(kind, reference_cases.with_tmax(rays, tmax), nthreads=2)
            assert got.tobytes() == z[f"{label}/{c['name']}/{kname}"].tobytes(), (label, c["name"], kname)
            n += len(rays)
        if fmt == ob.STACKLESS:
            hits, _ = sc.trace(ob.CLOSEST, rays)
            assert ob.get_data(sc.tris, sc.verts, sc.entities, hits).tobytes() == z[f"{label}/{c['name']}/get_data"].tobytes(), (c["name"], "GetData")
    assert n > 70000


def test_oracle_vs_compiled_reference_glsl_live(ob, golden_meshes):
    """Every case of cases.py at full size against the compiled reference shaders (only where /root/reference exists)."""
    if not ob.REFERENCE_ROOT.exists():
        pytest.skip("/root/reference is not present (GPU box): the committed fixture covers this")
    import cases
    n = 0
    for fmt in (ob.STACKLESS, ob.STACK):
        for c in cases.build_cases(ob, golden_meshes, fmt):
            sc = c["scene"]
            rays = c["rays"][:6000]
            for kind, tmax in ((ob.CLOSEST, 0.0), (ob.CLOSEST_IGNORE_TRANSPARENT, 0.0), (ob.ANY, 0.0), (ob.ANY, 2.4)):
                r = rays.copy()
                r["tmax"] = tmax
                mine, _ = sc.trace(kind, r, nthreads=ob.hardware_threads())
                ref = ob.ref_glsl_trace(fmt, kind, sc.nodes, sc.tris, sc.verts, sc.entities, r)
                assert mine.tobytes() == ref.tobytes(), (fmt, c["name"], kind, tmax)
                n += len(r)
    assert n > 200000


def test_primary_rays_and_entity_inverse_vs_compiled_reference(ob):
    """Primary-ray generation against the shader's own GetRayDirectionAt + main() arithmetic (Intersectors/TraverseBVHStack.glsl:
    133-138, :414-421) and PushEntity's glm::inverse (Intersector.h:209), both compiled from the reference (oracle/_ref)."""
    if not ob.REFERENCE_ROOT.exists():
        pytest.skip("/root/reference is not present (GPU box)")
    from candela_b200 import scenes
    for pos, look, W, H in (((-18.0, 5.0, 0.7), (10.0, 30.0, -0.4), 320, 180), ((0, 1, 2), (3, -2, 5), 257, 131), ((5, 5, 5), (0, 0, 0), 64, 64)):
        iv, ip = scenes.camera(pos, look, W, H)
        assert ob.primary_rays(iv, ip, W, H).tobytes() == ob.ref_primary_rays(iv, ip, W, H).tobytes()
    rng = np.random.default_rng(0)
    for _ in range(300):
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = rng.normal(size=(3, 3)).astype(np.float32) * np.float32(rng.uniform(0.1, 5))
        m[:3, 3] = rng.normal(size=3) * 10
        e = ob.make_entity(m, 0, 1)
        assert np.asarray(e["inverse"][0]).reshape(4, 4).T.tobytes() == ob.ref_glm_inverse(m).tobytes()


def test_direction_samplers_vs_compiled_reference(ob):
    """What DEFINING sin / cos costs: the generator oracle's CosWeightedHemisphere / SampleGGXVNDF (built-ins = exact_math_ref.h)
    against the same shader functions compiled with glm's libm-backed sin / cos (oracle/_ref, namespace ref_sampling) — they
    agree to an ulp or two, i.e. the definition sits inside the latitude GLSL gives an implementation.  (Bit-identity with the
    functions compiled over the DEFINED built-ins is tests/test_raygen_oracle.py.)  Also the host-side numpy sampler of scenes.py."""
    if not ob.REFERENCE_ROOT.exists():
        pytest.skip("/root/reference is not present (GPU box)")
    from candela_b200 import scenes
    rng = np.random.default_rng(1)
    N = rng.normal(size=(20000, 3)).astype(np.float32)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    N[:4] = [[0, 0, 1], [0, 0, -1], [0, 1, 0], [1, 0, 0]]
    xi = rng.random((20000, 2), dtype=np.float32)
    libm = ob.ref_sample(0, N, xi)
    assert np.abs(scenes.cos_weighted_hemisphere(N, xi) - libm).max() < 5e-7
    assert np.abs(ob.sample_directions(ob.SAMPLE_COS_HEMISPHERE, normals=N, xi=xi) - libm).max() < 5e-7
    for rough in (0.27, 0.72):
        assert np.abs(ob.sample_directions(ob.SAMPLE_GGX_VNDF, normals=N, xi=xi, roughness=rough) - ob.ref_sample(1, N, xi, rough)).max() < 5e-7



def test_oracle_vs_compiled_reference_glsl_random_scene_sweep(ob):
    """The pin of the traversal restatement on scenes nobody looked at (cases.random_scene: seeded random meshes, several
    objects, random affine instances incl. mirrored and sheared ones, translucent and emissive entities, rays from inside and
    outside the scene, unnormalised directions): all four query kinds, both node formats, against the compiled reference GLSL."""
    if not ob.REFERENCE_ROOT.exists():
        pytest.skip("/root/reference is not present (GPU box): the committed fixture covers this")
    import cases
    n = n_hit = 0
    for seed in range(6):
        for fmt in (ob.STACKLESS, ob.STACK):
            sc, rays = cases.random_scene(ob, fmt, seed)
            for kind, tmax in cases.random_scene_queries(ob, seed):
                r = rays.copy()
                r["tmax"] = tmax
                mine, _ = sc.trace(kind, r, nthreads=ob.hardware_threads())
                ref = ob.ref_glsl_trace(fmt, kind, sc.nodes, sc.tris, sc.verts, sc.entities, r, nthreads=ob.hardware_threads())
                assert mine.tobytes() == ref.tobytes(), (seed, fmt, kind, tmax)
                n += len(r)
            hits, _ = sc.trace(ob.CLOSEST, rays)
            n_hit += int(np.count_nonzero(hits["t"] > 0))
    assert n >= 6 * 2 * 4 * 4000 and n_hit > 5000, "the sweep must actually hit things"
