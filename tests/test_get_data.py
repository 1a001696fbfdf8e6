"""GetData (…/Include/TraverseBVHStackless.glsl:370-408, without the texture fetch): the step right after
the path.  CPU tests pin the oracle's restatement with closed-form answers; the
GPU test compares cndl_get_data with the oracle bit for bit on traced hit records."""
import numpy as np
import pytest


def _sphere_like_attributes(P, seed):
    rng = np.random.default_rng(seed)
    n = P - P.mean(0)
    n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-6)
    uv = rng.random((len(P), 2), dtype=np.float32) * 4.0 - 1.0
    return n.astype(np.float32), uv.astype(np.float32)


def test_oracle_get_data_known_answers(ob):
    P = np.array([[-1, -1, 5], [1, -1, 5], [0, 1, 5], [9, 9, 9], [10, 9, 9], [9, 10, 9]], np.float32)
    N = np.array([[0, 0, -1]] * 3 + [[1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    UV = np.array([[0, 0], [1, 0], [0, 1], [0.25, 0.5], [0.75, 0.5], [0.25, 1.0]], np.float32)
    V = ob.pack_vertices(P, N, UV)
    tris = np.zeros(2, ob.TRIANGLE_DT)
    tris["v"] = [[0, 1, 2], [3, 4, 5]]
    tris["mesh"] = [4, 9]
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, 1, emissive=2.5, translucency=0.25)
    hits = np.zeros(5, ob.HIT_DT)
    hits[0] = (3.0, 0.5, 0.25, 0.25, 4, 0, 0, 7)          # inside triangle 0
    hits[1] = (1.0, 0.0, 1.0, 0.0, 9, 1, 0, 3)            # exactly vertex B of triangle 1
    hits[2] = (-1.0, -1.0, -1.0, -1.0, -1, -1, -1, 0)     # miss
    hits[3] = (-1.0, -1.0, -1.0, -1.0, 4, 0, 0, 2)        # the triangle-0 blind spot reports t = -1 -> treated as a miss
    hits[4] = (2.0, 0.2, 0.3, 0.5, 9, 1, 0, 3)
    a = ob.get_data(tris, V, ents, hits)
    assert np.array_equal(a["normal"][0], [0, 0, -1]) and np.array_equal(a["uv"][0], [0.25, 0.25])   # UV = uvA*u + uvB*v + uvC*w
    assert np.array_equal(a["normal"][1], [0, 1, 0]) and np.array_equal(a["uv"][1], [0.75, 0.5])
    for k in (2, 3):
        assert np.array_equal(a["normal"][k], [-1, -1, -1]) and np.all(a["uv"][k] == 0) and a["emissivity"][k] == 0 and a["alpha"][k] == 0
    n4 = np.array([0.2, 0.3, 0.5], np.float64)
    assert np.allclose(a["normal"][4], n4 / np.linalg.norm(n4), atol=2e-7)
    assert np.allclose(np.linalg.norm(a["normal"][[0, 1, 4]].astype(np.float64), axis=1), 1.0, atol=1e-6)
    assert np.all(a["emissivity"][[0, 1, 4]] == np.float32(2.5)) and np.all(a["alpha"][[0, 1, 4]] == np.float32(0.75))
    assert list(a["mesh"]) == [4, 9, -1, 4, 9]


def test_oracle_half_unpack_is_exact_for_every_finite_pattern(ob):
    """unpackHalf2x16 restated: all 63,488 finite binary16 patterns (zeros, subnormals, normals) through the
    low and high half of the packed UV word, with weights (1, 0, 0) so that the output is the value itself."""
    bits = np.array([b for b in range(65536) if (b >> 10) & 0x1F != 0x1F], dtype=np.uint32)
    n = len(bits)
    V = np.zeros(3 * n, ob.VERTEX_DT)
    V["texcoords"][0::3] = bits | (bits[::-1] << 16)
    tris = np.zeros(n, ob.TRIANGLE_DT)
    tris["v"] = np.arange(3 * n, dtype=np.int32).reshape(-1, 3)
    hits = np.zeros(n, ob.HIT_DT)
    hits["t"], hits["u"], hits["tri"] = 1.0, 1.0, np.arange(n)
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, 1)
    a = ob.get_data(tris, V, ents, hits)
    want = bits.astype(np.uint16).view(np.float16).astype(np.float32)
    assert np.array_equal(a["uv"][:, 0], want + np.float32(0))
    assert np.array_equal(a["uv"][:, 1], want[::-1] + np.float32(0))


def test_oracle_get_data_vs_compiled_reference_glsl(ob, golden_meshes):
    """GetData of the reference's shader file itself, compiled against its glm (oracle/_ref), on traced hit records of a
    scene with real packed normals / UVs and three entities."""
    if not ob.REFERENCE_ROOT.exists():
        pytest.skip("/root/reference is not present (GPU box): tests/golden/reference_traversal_golden.npz covers GetData there")
    from cases import scale_rot, translate
    from helpers import rays_in_box
    P, F = golden_meshes["dragon"]
    N, UV = _sphere_like_attributes(P, 3)
    V = ob.pack_vertices(P, N, UV)
    sc = ob.Scene(ob.STACKLESS)
    sc.add_object(2, V, F.ravel(), (np.arange(len(F)) % 5).astype(np.int32))
    sc.push_entity(2, emissive=1.5)
    sc.push_entity(2, model=scale_rot(0.5, 30.0, (3.0, 0.5, 0.0)), translucency=0.4)
    sc.push_entity(2, model=translate(-3.0, 0.0, 1.0), emissive=7.0, translucency=1.0)
    rays = rays_in_box(P.min(0) - (3.5, 0.5, 0.5), P.max(0) + (3.5, 0.5, 1.5), 20000, 11)
    hits, _ = sc.trace(ob.CLOSEST, rays, nthreads=ob.hardware_threads())
    assert hits.tobytes() == ob.ref_glsl_trace(ob.STACKLESS, ob.CLOSEST, sc.nodes, sc.tris, sc.verts, sc.entities, rays).tobytes()
    mine = ob.get_data(sc.tris, sc.verts, sc.entities, hits)
    ref = ob.ref_glsl_get_data(sc.tris, sc.verts, sc.entities, hits)
    hit = hits["t"] > 0
    assert hit.sum() > 3000
    # the shader leaves Alpha unwritten on a miss (the record says 0 there); everything else is compared bit for bit
    assert mine[hit].tobytes() == ref[hit].tobytes()
    for f in ("normal", "uv", "emissivity", "mesh"):
        assert mine[f][~hit].tobytes() == ref[f][~hit].tobytes(), f


@pytest.mark.gpu
def test_gpu_get_data_bit_identical_to_oracle(cb, ob, golden_meshes):
    from cases import scale_rot, translate
    from helpers import rays_in_box
    P, F = golden_meshes["dragon"]
    N, UV = _sphere_like_attributes(P, 3)
    V = ob.pack_vertices(P, N, UV)
    mids = (np.arange(len(F)) % 5).astype(np.int32)
    for fmt, ofmt in ((cb.STACKLESS, ob.STACKLESS), (cb.STACK, ob.STACK)):
        sc = ob.Scene(ofmt)
        sc.add_object(2, V, F.ravel(), mids)
        sc.push_entity(2, emissive=1.5)
        sc.push_entity(2, model=scale_rot(0.5, 30.0, (3.0, 0.5, 0.0)), translucency=0.4)
        sc.push_entity(2, model=translate(-3.0, 0.0, 1.0), emissive=7.0, translucency=1.0)
        ri = cb.RayIntersector(fmt)
        ri.AddPrebuiltObject(2, sc.nodes, sc.tris, V)
        ri.BufferData()
        ri.PushEntityRecords(sc.entities)
        ri.BufferEntities()
        lo, hi = P.min(0) - (3.5, 0.5, 0.5), P.max(0) + (3.5, 0.5, 1.5)
        rays = rays_in_box(lo, hi, 60000, 11)
        hits = ri.IntersectRays(rays)
        want_hits, _ = sc.trace(ob.CLOSEST, rays, nthreads=ob.hardware_threads())
        assert hits.tobytes() == want_hits.tobytes()
        assert 0.2 < float((hits["t"] > 0).mean()) < 0.999 and len(np.unique(hits["entity"][hits["t"] > 0])) == 3
        got = ri.GetData(hits)
        want = ob.get_data(sc.tris, sc.verts, sc.entities, hits)
        assert got.tobytes() == want.tobytes()
        hit = hits["t"] > 0
        assert np.allclose(np.linalg.norm(got["normal"][hit].astype(np.float64), axis=1), 1.0, atol=1e-5)
        assert np.all(got["normal"][~hit] == -1)
        # device-pointer entry point, same records
        import torch
        d_h = torch.from_numpy(hits.view(np.float32).reshape(-1, 8)).cuda()
        d_a = torch.zeros((len(hits), 8), dtype=torch.float32, device="cuda")
        ri.get_data_device(d_h.data_ptr(), len(hits), d_a.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert d_a.cpu().numpy().tobytes() == want.tobytes()
        assert ri.GetData(hits[:0]).shape == (0,)
        ri.close()
