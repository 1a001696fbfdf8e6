"""GPU parity tests of the builder: cndl_add_object (binned SAH on the GPU) must produce node,
triangle and vertex buffers byte-identical to the oracle's restatement of BVH::BuildBVH — which is
itself pinned byte-for-byte to the compiled reference builder (tests/test_oracle_builder.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORMATS = ["stackless", "stack"]


def fmt_id(ob, name):
    return ob.STACKLESS if name == "stackless" else ob.STACK


def first_diff(a, b):
    x, y = np.frombuffer(a.tobytes(), np.uint32), np.frombuffer(b.tobytes(), np.uint32)
    if len(x) != len(y):
        return f"length {len(x)} vs {len(y)}"
    d = np.nonzero(x != y)[0]
    return f"{len(d)} words differ, first at word {d[:4]}" if len(d) else "equal"


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("name", ["dragon", "peach_castle", "zelda_market", "coplanar_grid", "duplicates", "soup400", "collinear", "signed_zero"])
def test_gpu_build_is_byte_identical(cb, ob, golden_meshes, name, fmt):
    P, F = golden_meshes[name]
    V = ob.make_vertices(P)
    mids = (np.arange(len(F)) % 5).astype(np.int32)
    ref = ob.build(fmt_id(ob, fmt), V, F.ravel(), mids)
    ri = cb.RayIntersector(fmt_id(ob, fmt))
    ri.AddObject(2, V, F.ravel(), mids)
    nodes, tris, verts = ri.read_buffers()
    assert len(nodes) == len(ref.nodes)
    assert tris.tobytes() == ref.tris.tobytes(), first_diff(tris, ref.tris)
    assert nodes.tobytes() == ref.nodes.tobytes(), first_diff(nodes, ref.nodes)
    assert verts.tobytes() == V.tobytes()
    ri.close()


@pytest.mark.parametrize("kind", ["soup", "quantised", "flat", "slivers", "clusters", "repeats"])
def test_gpu_build_random_mesh_sweep(cb, ob, kind):
    """Seeded random meshes (tests/cases.py: bin-edge ties, a zero-extent axis, repeated triangles, slivers, clusters) at sizes
    either side of the builder's size classes (64 / 512 / 16384 references per range): both formats byte-identical to the oracle,
    which tests/test_oracle_builder.py pins to the compiled reference builder on the same meshes."""
    import cases
    for n, T in enumerate([100, 513, 2049, 4500, 17000, 40000]):
        P, F = cases.random_mesh(kind, T, 1000 * n + 7)
        V = ob.make_vertices(P)
        mids = (np.arange(T) % 3).astype(np.int32)
        for fmt in FORMATS:
            ref = ob.build(fmt_id(ob, fmt), V, F.ravel(), mids)
            ri = cb.RayIntersector(fmt_id(ob, fmt))
            ri.AddObject(2, V, F.ravel(), mids)
            nodes, tris, _ = ri.read_buffers()
            assert tris.tobytes() == ref.tris.tobytes(), (kind, T, fmt, first_diff(tris, ref.tris))
            assert nodes.tobytes() == ref.nodes.tobytes(), (kind, T, fmt, first_diff(nodes, ref.nodes))
            ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("T", [1, 2, 3, 4, 5, 7, 50, 129])
def test_tiny_meshes(cb, ob, T, fmt):
    rng = np.random.default_rng(T)
    P = rng.uniform(-1, 1, size=(3 * T, 3)).astype(np.float32)
    F = np.arange(3 * T, dtype=np.uint32).reshape(-1, 3)
    V = ob.make_vertices(P)
    ref = ob.build(fmt_id(ob, fmt), V, F.ravel())
    ri = cb.RayIntersector(fmt_id(ob, fmt))
    ri.AddObject(2, V, F.ravel())
    nodes, tris, _ = ri.read_buffers()
    assert tris.tobytes() == ref.tris.tobytes() and nodes.tobytes() == ref.nodes.tobytes(), first_diff(nodes, ref.nodes)
    ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_multi_object_scene_offsets_and_flips(cb, ob, golden_meshes, fmt):
    """Several AddObject calls: triangle offsets in leaf packs, vertex index rebasing, object table
    (Intersector.h:170-198) and hashed child flips."""
    sc = ob.Scene(fmt_id(ob, fmt))
    ri = cb.RayIntersector(fmt_id(ob, fmt))
    for oid, name, seed in ((2, "peach_castle", 5), (3, "soup400", 0), (7, "dragon", 9)):
        P, F = golden_meshes[name]
        V = ob.make_vertices(P)
        mids = np.full(len(F), oid, np.int32)
        pol = ob.SWAP_HASHED if seed else ob.SWAP_NONE
        sc.add_object(oid, V, F.ravel(), mids, swap_policy=pol, swap_seed=seed)
        ri.AddObject(oid, V, F.ravel(), mids, swap_policy=pol, swap_seed=seed)
        o = ri.object_data(oid)
        r = sc.objects[oid]
        assert (o["node_offset"], o["node_count"], o["tri_offset"], o["vert_offset"]) == (r["node_offset"], r["node_count"], r["tri_offset"], r["vert_offset"])
    nodes, tris, verts = ri.read_buffers()
    assert tris.tobytes() == sc.tris.tobytes() and verts.tobytes() == sc.verts.tobytes()
    assert nodes.tobytes() == sc.nodes.tobytes(), first_diff(nodes, sc.nodes)
    ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_s260k_build_and_trace_end_to_end(cb, ob, fmt):
    """The ~260k-triangle scene built on the GPU, byte-compared, then traced through the GPU-built buffers."""
    from candela_b200 import scenes
    from helpers import rays_in_box
    v, i, m = scenes.make_s260k()
    ref = ob.build(fmt_id(ob, fmt), v, i, m)
    ri = cb.RayIntersector(fmt_id(ob, fmt))
    ri.AddObject(2, v, i, m)
    nodes, tris, _ = ri.read_buffers()
    assert tris.tobytes() == ref.tris.tobytes(), first_diff(tris, ref.tris)
    assert nodes.tobytes() == ref.nodes.tobytes(), first_diff(nodes, ref.nodes)
    assert ri.last_build_ms > 0
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    rays = rays_in_box((-20, 0, -9), (20, 14, 9), 200000, 8)
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(ref.nodes))
    want, _ = ob.trace(fmt_id(ob, fmt), ob.CLOSEST, ref.nodes, ref.tris, v, ents, rays, nthreads=ob.hardware_threads())
    assert ri.IntersectRays(rays).tobytes() == want.tobytes()
    ri.close()


def test_heightfield_sorted_input(cb, ob):
    """A grid in scanline order: centroids are monotone along x, the worst case for the partition emulation."""
    from candela_b200 import scenes
    v, i, m = scenes.make_heightfield(150)
    for fmt in (ob.STACKLESS, ob.STACK):
        ref = ob.build(fmt, v, i, m)
        ri = cb.RayIntersector(fmt)
        ri.AddObject(2, v, i, m)
        nodes, tris, _ = ri.read_buffers()
        assert tris.tobytes() == ref.tris.tobytes() and nodes.tobytes() == ref.nodes.tobytes(), first_diff(nodes, ref.nodes)
        ri.close()


@pytest.mark.parametrize("split_node", [64, 300, 5000, 1 << 26])
def test_split_node_threshold_never_changes_the_buffers(cb, ob, golden_meshes, split_node):
    """Ranges above CNDL_KNOB_BUILD_SPLIT_NODE go through the multi-CTA level step (one CTA per 512 or 2048 references, global bins,
    ranks and the Lomuto chain across CTAs), the others through the one-CTA-per-node step: same bytes for every threshold,
    from "nearly every level split" (64) to "never" (2^26), on meshes with split failures, duplicates and signed zeros."""
    from candela_b200 import scenes
    cases = [(n,) + tuple(golden_meshes[n]) for n in ("dragon", "coplanar_grid", "duplicates", "signed_zero", "soup400")]
    v, i, m = scenes.make_heightfield(120)
    for fmt in (ob.STACKLESS, ob.STACK):
        for name, P, F in cases:
            V = ob.make_vertices(P)
            ref = ob.build(fmt, V, F.ravel())
            ri = cb.RayIntersector(fmt)
            ri.set_tuning(8, split_node)
            ri.AddObject(2, V, F.ravel())
            nodes, tris, _ = ri.read_buffers()
            assert tris.tobytes() == ref.tris.tobytes(), (name, first_diff(tris, ref.tris))
            assert nodes.tobytes() == ref.nodes.tobytes(), (name, first_diff(nodes, ref.nodes))
            ri.close()
        ref = ob.build(fmt, v, i, m, swap_policy=ob.SWAP_HASHED, swap_seed=7)
        ri = cb.RayIntersector(fmt)
        ri.set_tuning(8, split_node)
        ri.AddObject(2, v, i, m, swap_policy=ob.SWAP_HASHED, swap_seed=7)
        nodes, tris, _ = ri.read_buffers()
        assert tris.tobytes() == ref.tris.tobytes() and nodes.tobytes() == ref.nodes.tobytes(), ("heightfield", first_diff(nodes, ref.nodes))
        ri.close()


def test_packed_tiny_ranges_never_change_the_buffers(cb, ob, golden_meshes):
    """CNDL_KNOB_BUILD_PACK_MIN = 1: every level handles its ranges of 3..64 references four to a warp (those of <= 8 references side by
    side on eight lanes each, the longer ones in turn) — the default only does so for levels of >= 131072 such ranges.  Same bytes, on
    meshes with split failures, duplicates, signed zeros and degenerate boxes, with and without the hashed child swaps."""
    from candela_b200 import scenes
    names = [n for n in ("dragon", "coplanar_grid", "duplicates", "signed_zero", "soup400", "peach_castle") if n in golden_meshes]
    v, i, m = scenes.make_heightfield(120)
    for fmt in (ob.STACKLESS, ob.STACK):
        for name in names:
            P, F = golden_meshes[name]
            V = ob.make_vertices(P)
            ref = ob.build(fmt, V, F.ravel())
            for split_node in (0, 64):
                ri = cb.RayIntersector(fmt)
                ri.set_tuning(9, 1)
                ri.set_tuning(8, split_node)
                ri.AddObject(2, V, F.ravel())
                nodes, tris, _ = ri.read_buffers()
                assert tris.tobytes() == ref.tris.tobytes(), (name, split_node, first_diff(tris, ref.tris))
                assert nodes.tobytes() == ref.nodes.tobytes(), (name, split_node, first_diff(nodes, ref.nodes))
                ri.close()
        ref = ob.build(fmt, v, i, m, swap_policy=ob.SWAP_HASHED, swap_seed=7)
        ri = cb.RayIntersector(fmt)
        ri.set_tuning(9, 1)
        ri.AddObject(2, v, i, m, swap_policy=ob.SWAP_HASHED, swap_seed=7)
        nodes, tris, _ = ri.read_buffers()
        assert tris.tobytes() == ref.tris.tobytes() and nodes.tobytes() == ref.nodes.tobytes(), ("heightfield", first_diff(nodes, ref.nodes))
        ri.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_build_bvh_free_function_with_triangle_offset(cb, ob, golden_meshes, fmt):
    """cndl_build_bvh == BVH::BuildBVH(object, nodes, vertices, triangles, t_offset) (BVHConstructor.h:86-87): the leaf packs carry
    t_offset, the triangle records keep object-local vertex indices."""
    P, F = golden_meshes["peach_castle"]
    V = ob.make_vertices(P)
    mids = (np.arange(len(F)) % 3).astype(np.int32)
    for t_offset in (0, 12345):
        ref = ob.build(fmt_id(ob, fmt), V, F.ravel(), mids, t_offset=t_offset)
        nodes, tris, ms = cb.BuildBVH(fmt_id(ob, fmt), V, F.ravel(), mids, t_offset=t_offset)
        assert ms > 0 and len(nodes) == len(ref.nodes)
        assert tris.tobytes() == ref.tris.tobytes(), first_diff(tris, ref.tris)
        assert nodes.tobytes() == ref.nodes.tobytes(), first_diff(nodes, ref.nodes)
    with pytest.raises(cb.CandelaError):
        cb.BuildBVH(fmt_id(ob, fmt), V, F.ravel()[:-1], None)


def test_bad_geometry_rejected(cb):
    ri = cb.RayIntersector(cb.STACKLESS)
    V = cb.make_vertices(np.zeros((3, 3), np.float32))
    with pytest.raises(cb.CandelaError):
        ri.AddObject(2, V, np.array([0, 1], np.uint32))
    with pytest.raises(cb.CandelaError):
        ri.AddObject(2, V, np.array([0, 1, 3], np.uint32))
    ri.close()


# ---- LBVH builder (CNDL_BUILDER_LBVH): same layouts, different tree -> parity level P2 ----------------------
def _lbvh_scene(cb, ob, fmt, V, F, mids):
    ri = cb.RayIntersector(fmt)
    ri.AddObject(2, V, F, mids, builder=cb.BUILDER_LBVH)
    ri.BufferData()
    ri.PushEntity(2)
    ri.BufferEntities()
    return ri


@pytest.mark.parametrize("name", ["dragon", "peach_castle", "soup400", "duplicates", "coplanar_grid", "collinear"])
def test_lbvh_structure_and_traversal(cb, ob, golden_meshes, name):
    from helpers import rays_in_box
    from test_oracle_builder import check_stackless_invariants
    P, F = golden_meshes[name]
    V = ob.make_vertices(P)
    mids = (np.arange(len(F)) % 3).astype(np.int32)
    rays = rays_in_box(P.min(0) - 0.5, P.max(0) + 0.5, 20000, 12)
    sah = ob.Scene(ob.STACKLESS)
    sah.add_object(2, V, F.ravel(), mids)
    sah.push_entity(2)
    want, _ = sah.trace(ob.CLOSEST, rays, nthreads=8)
    for fmt in (ob.STACKLESS, ob.STACK):
        ri = _lbvh_scene(cb, ob, fmt, V, F.ravel(), mids)
        nodes, tris, verts = ri.read_buffers()
        # the triangle buffer is a permutation of the input with mesh ids attached
        got_set = sorted(map(tuple, np.concatenate([tris["v"], tris["mesh"][:, None]], 1).tolist()))
        ref_set = sorted(map(tuple, np.concatenate([F.astype(np.int32), mids[:, None]], 1).tolist()))
        assert got_set == ref_set
        if fmt == ob.STACKLESS:
            check_stackless_invariants(nodes, tris, len(F))
        # P1 on the GPU-built buffers: GPU traversal == oracle traversal over the very same buffers
        ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(nodes))
        same_buf, _ = ob.trace(fmt, ob.CLOSEST, nodes, tris, V, ents, rays, nthreads=8)
        got = ri.IntersectRays(rays)
        assert got.tobytes() == same_buf.tobytes()
        # P2 against the reference tree: same hit triangle (identity, not buffer index), same t
        hit_w, hit_g = want["tri"] >= 0, got["tri"] >= 0
        assert (hit_w == hit_g).mean() > 0.9995
        both = hit_w & hit_g
        ident_w = np.concatenate([sah.tris["v"][want["tri"][both]], sah.tris["mesh"][want["tri"][both], None]], 1)
        ident_g = np.concatenate([tris["v"][got["tri"][both]], tris["mesh"][got["tri"][both], None]], 1)
        same = np.all(ident_w == ident_g, axis=1)
        if not both.any():                               # degenerate geometry: nothing can be hit
            pass
        elif name in ("duplicates", "coplanar_grid"):    # coincident triangles: any of the duplicates may win
            assert np.array_equal(got["t"][both], want["t"][both]) or same.mean() > 0.9
        else:
            assert same.mean() > 0.999, same.mean()
            sel = np.nonzero(both)[0][same]
            ok = (got["tri"][sel] > 0) & (want["tri"][sel] > 0)    # the triangle-0 blind spot differs between trees
            assert np.array_equal(got["t"][sel][ok], want["t"][sel][ok])
        ri.close()


def test_lbvh_tiny_and_build_time(cb, ob):
    from candela_b200 import scenes
    for T in (1, 2, 3, 5, 8):
        rng = np.random.default_rng(T)
        P = rng.uniform(-1, 1, size=(3 * T, 3)).astype(np.float32)
        F = np.arange(3 * T, dtype=np.uint32)
        for fmt in (ob.STACKLESS, ob.STACK):
            ri = cb.RayIntersector(fmt)
            ri.AddObject(2, ob.make_vertices(P), F, builder=cb.BUILDER_LBVH)
            nodes, tris, _ = ri.read_buffers()
            assert len(tris) == T and len(nodes) == 2 * ((T + 1) // 2) - 1
            ri.close()
    v, i, m = scenes.make_s260k()
    ri = cb.RayIntersector(ob.STACKLESS)
    ri.AddObject(2, v, i, m, builder=cb.BUILDER_LBVH)
    assert 0 < ri.last_build_ms < 50
    ri.close()
