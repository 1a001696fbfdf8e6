"""The per-mesh material table (BVH::TextureReferences, Intersector.h:32-37, :367-410), GetData's Albedo decision
(…/Include/TraverseBVHStackless.glsl:393-404), and the loader's materials / tangents (ModelFileLoader.cpp:31-99, :138-152, :243-252)."""
import json

import numpy as np
import pytest


def _half3(v):
    """(tangent.x, tangent.y, tangent.z) of a packed vertex."""
    return np.array([int(v["normal_tangent"][1]) >> 16, int(v["normal_tangent"][2]) & 0xFFFF, int(v["normal_tangent"][2]) >> 16], dtype=np.uint16).view(np.float16).astype(np.float32)


# ---- GenerateMeshTextureReferences -------------------------------------------------------------------------------------------

MATS = [  # (albedo handle, found, normal handle, found, ModelColor) — worked by hand from Intersector.h:386-402
    (0x1111, True, 0x2222, True, (0.5, 0.25, 1.0)),      # new, new          -> 0, 1
    (0x1111, True, 0x1111, False, (1.0, 1.0, 1.0)),      # seen; an unknown path returns the first cached handle (Texture.cpp:177-181) -> 0, -1
    (0x9999, False, 0x2222, True, (0.0, 0.0, 0.0)),      # an invalid handle still takes index 2; -> -1, 1
    (0x3333, True, 0x9999, False, (0.1, 0.2, 0.3)),      # 3; 0x9999 is known (2) but invalid -> 3, -1
    (0x9999, True, 0x4444, True, (0.6, 0.6, 0.6)),       # the same handle, this time valid -> 2; new -> 4
]
EXPECTED = [(0, 1), (0, -1), (-1, 1), (3, -1), (2, 4)]
EXPECTED_HANDLES = [0x1111, 0x2222, 0x9999, 0x3333, 0x4444]


def test_texture_references_restatement_known_answer(ob):
    table, handles = ob.texture_references(MATS)
    assert [(int(r["albedo"]), int(r["normal"])) for r in table] == EXPECTED and list(handles) == EXPECTED_HANDLES
    assert np.array_equal(table["model_color"][:, :3], np.array([m[4] for m in MATS], np.float32)) and np.all(table["model_color"][:, 3] == 1.0)
    assert np.all(table["pad"] == 0) and table.dtype.itemsize == 32


def test_product_texture_references_match_the_restatement(cb, ob):
    """Host arithmetic of the C ABI (no GPU needed)."""
    table, handles = cb.api.generate_texture_references(MATS)
    want, want_handles = ob.texture_references(MATS)
    assert table.tobytes() == want.tobytes() and np.array_equal(handles, want_handles)
    rng = np.random.default_rng(3)
    mats = [(int(rng.integers(0, 40)), bool(rng.integers(0, 2)), int(rng.integers(0, 40)), bool(rng.integers(0, 2)), tuple(rng.random(3))) for _ in range(500)]
    table, handles = cb.api.generate_texture_references(mats)
    want, want_handles = ob.texture_references(mats)
    assert table.tobytes() == want.tobytes() and np.array_equal(handles, want_handles)
    t0, h0 = cb.api.generate_texture_references([])
    assert len(t0) == 0 and len(h0) == 0


# ---- GetData with the table: restatement pinned against the compiled reference shader --------------------------------------------

def _scene_and_hits(ob, golden_meshes, n_mesh=6, seed=5):
    """A small two-entity scene whose triangles carry mesh numbers 0..n_mesh-1, packed normals / UVs, and hit records from the oracle."""
    from helpers import rays_in_box
    P, F = golden_meshes["soup400"]
    rng = np.random.default_rng(seed)
    verts = ob.make_vertices(P)
    N = rng.normal(size=P.shape).astype(np.float32)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    UV = rng.random((len(P), 2), dtype=np.float32)
    pk = lambda a, b: (np.asarray(a, np.float16).view(np.uint16).astype(np.uint32) | (np.asarray(b, np.float16).view(np.uint16).astype(np.uint32) << 16))
    verts["normal_tangent"][:, 0] = pk(N[:, 0], N[:, 1])
    verts["normal_tangent"][:, 1] = pk(N[:, 2], np.zeros(len(P)))
    verts["texcoords"] = pk(UV[:, 0], UV[:, 1])
    mesh_ids = (np.arange(len(F)) * n_mesh // len(F)).astype(np.int32)
    return verts, F.astype(np.uint32), mesh_ids, rays_in_box(P.min(0), P.max(0), 4000, seed=seed)


def _table(n_mesh, seed=11):
    rng = np.random.default_rng(seed)
    t = np.zeros(n_mesh, dtype=[("model_color", "<f4", 4), ("albedo", "<i4"), ("normal", "<i4"), ("pad", "<i4", 2)])
    t["model_color"][:, :3] = rng.random((n_mesh, 3), dtype=np.float32)
    t["model_color"][:, 3] = 1.0
    t["albedo"] = [-1, 0, 7, -1, 511, 3][:n_mesh]     # textured and untextured meshes; 511 = the last sampler of Textures[512]
    t["normal"] = [-1, 1, -1, 2, -1, 4][:n_mesh]
    return t


def _build_oracle_scene(ob, verts, idx, mesh_ids):
    res = ob.build(0, verts, idx, mesh_ids)
    ents = ob.make_entity(np.eye(4, dtype=np.float32), 0, len(res.nodes), emissive=2.5, translucency=0.25)
    return res, ents


def test_get_data_material_restatement_is_the_reference_shader(ob, golden_meshes):
    verts, idx, mesh_ids, rays = _scene_and_hits(ob, golden_meshes)
    res, ents = _build_oracle_scene(ob, verts, idx, mesh_ids)
    hits = ob.trace(0, ob.CLOSEST, res.nodes, res.tris, verts, ents, rays)[0]
    hits = np.concatenate([hits, hits[:8]])
    hits["t"][-8:-4] = 0.0           # t == 0 is not a miss (SL:377) but fails `TUVW.x > 0.` (SL:397): the ModelColor branch
    hits["mesh"][-4:] = -1           # Mesh < 0: the early return
    table = _table(6)
    got = ob.get_data_material(res.tris, verts, ents, table, hits)
    ref, tex_uv = ob.ref_glsl_get_data_material(res.tris, verts, ents, table, hits)
    assert (hits["t"] > 0).sum() > 500 and (hits["t"] < 0).sum() > 100
    sampled = ref["albedo_ref"] >= 0
    assert sampled.sum() > 100 and (~sampled & (hits["t"] > 0)).sum() > 50
    assert got.tobytes() == ref.tobytes()
    # the UV the shader handed to texture() is the UV the record reports
    assert np.array_equal(tex_uv[sampled].view(np.uint32), ref["uv"][sampled].view(np.uint32))
    # and the first 32 bytes are the plain GetData record
    plain = ob.get_data(res.tris, verts, ents, hits)
    assert np.array_equal(got[["normal", "uv", "emissivity", "alpha", "mesh"]].tobytes(), plain.tobytes()) or \
        all(np.array_equal(got[f].view(np.uint32), plain[f].view(np.uint32)) for f in ("normal", "uv", "emissivity", "alpha", "mesh"))
    # a mesh outside the table is marked, never guessed
    short = ob.get_data_material(res.tris, verts, ents, table[:3], hits)
    out = (hits["t"] >= 0) & (hits["mesh"] >= 3)
    assert out.any() and np.all(short["albedo_ref"][out] == -2) and np.all(short["albedo"][out] == 0) and short[~out].tobytes() == got[~out].tobytes()


@pytest.mark.gpu
def test_gpu_get_data_material_matches_the_oracle(cb, ob, golden_meshes):
    verts, idx, mesh_ids, rays = _scene_and_hits(ob, golden_meshes)
    res, ents = _build_oracle_scene(ob, verts, idx, mesh_ids)
    for fmt in (cb.STACKLESS, cb.STACK):
        ri = cb.RayIntersector(fmt)
        ri.AddObject(1, verts, idx, mesh_ids)
        ri.BufferData()
        ri.PushEntity(1, emissive=2.5, translucency=0.25)
        ri.BufferEntities()
        hits = ri.IntersectRays(rays)
        hits = np.concatenate([hits, hits[:8]])
        hits["t"][-8:-4] = 0.0
        hits["mesh"][-4:] = -1
        with pytest.raises(cb.CandelaError, match="no texture-reference table"):
            ri.GetDataMaterial(hits)
        table = _table(6)
        ri.SetTextureReferences(table)
        assert ri.texture_reference_count() == 6
        _, tris, gverts = ri.read_buffers()
        want = ob.get_data_material(tris, gverts, ents, table, hits)
        got = ri.GetDataMaterial(hits)
        assert got.tobytes() == want.tobytes()
        assert (got["albedo_ref"] >= 0).sum() > 100 and ((got["albedo_ref"] == -1) & (hits["t"] > 0)).sum() > 50
        plain = ri.GetData(hits)
        assert all(np.array_equal(got[f].view(np.uint32), plain[f].view(np.uint32)) for f in ("normal", "uv", "emissivity", "alpha", "mesh"))
        # a shorter table: the host call refuses to return guessed materials
        ri.SetTextureReferences(table[:3])
        with pytest.raises(cb.CandelaError, match="outside the texture-reference table"):
            ri.GetDataMaterial(hits)
        # the table from the reference's generator, end to end
        handles = ri.GenerateMeshTextureReferences([(0x10 + k, k % 2 == 0, 0x20 + k, True, (k / 8, 0.5, 1.0)) for k in range(6)])
        assert len(handles) == 12
        want_table, _ = ob.texture_references([(0x10 + k, k % 2 == 0, 0x20 + k, True, (k / 8, 0.5, 1.0)) for k in range(6)])
        assert ri.GetDataMaterial(hits).tobytes() == ob.get_data_material(tris, gverts, ents, want_table, hits).tobytes()
        ri.close()


# ---- loader: materials and tangents ---------------------------------------------------------------------------------------------

def test_obj_materials_and_tangents(cb, tmp_path):
    d = tmp_path / "assets"
    d.mkdir()
    (d / "room.mtl").write_text(
        "newmtl brick\nKa 0 0 0\nKd 0.8 0.4 0.2\nmap_Kd -bm 1.0 tex/brick_d.png\nmap_Bump tex/brick_h.png\nnorm tex/brick_n.png\n"
        "newmtl plain\nKd 0.1 0.9 0.3\n"
        "newmtl bare\nmap_Kn n_only.png\n")
    # a unit quad in the xz plane (normal +y) with u along +x and v along -z, then the same quad with u along +z
    (d / "room.obj").write_text(
        "mtllib room.mtl\n"
        "v 0 0 0\nv 1 0 0\nv 1 0 -1\nv 0 0 -1\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
        "vn 0 1 0\n"
        "usemtl brick\nf 1/1/1 2/2/1 3/3/1 4/4/1\n"
        "usemtl plain\nf 1/1/1 4/2/1 3/3/1\n"
        "usemtl bare\nf 1//1 2//1 3//1\n"
        "usemtl missing\nf 1 2 3\n")
    verts, idx, mids, names, mats = cb.api.load_model(d / "room.obj", first_mesh_number=3, materials=True)
    assert names == ["brick", "plain", "bare", "missing"] and list(mids) == [3, 3, 4, 5, 6]
    base = str(d)
    assert mats[0] == {"albedo": base + "/tex/brick_d.png", "normal": base + "/tex/brick_n.png", "color": pytest.approx((0.8, 0.4, 0.2))}
    assert mats[1] == {"albedo": "", "normal": "", "color": pytest.approx((0.1, 0.9, 0.3))}
    assert mats[2] == {"albedo": "", "normal": base + "/n_only.png", "color": pytest.approx((0.6, 0.6, 0.6))}      # Assimp's default diffuse colour
    assert mats[3] == {"albedo": "", "normal": "", "color": pytest.approx((0.6, 0.6, 0.6))}
    # the quad: four joined vertices, every tangent = +x (dP/du), orthogonal to the normal
    quad = verts[idx[:6]]
    assert len(np.unique(idx[:6])) == 4
    for v in quad:
        assert np.array_equal(_half3(v), np.array([1, 0, 0], np.float32))
    # second mesh: u runs from (0,0,0) to (0,0,-1): tangent = -z
    for v in verts[idx[6:9]]:
        assert np.array_equal(_half3(v), np.array([0, 0, -1], np.float32))
    # no UV channel: CalcTangentSpace skips the mesh, the reference packs zeros
    for v in verts[idx[9:]]:
        assert np.array_equal(_half3(v), np.zeros(3, np.float32))


def test_tangent_smoothing_and_seams(cb, tmp_path):
    """Two triangles meeting at an edge with slightly different UV gradients are averaged (within 45 degrees, same normal);
    a mirrored UV island (tangent flipped: 180 degrees apart) keeps its own tangents and its corners are NOT joined."""
    p = tmp_path / "s.obj"
    p.write_text(
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 2 0 0\nv 2 1 0\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 1 0.9\n"
        "vn 0 0 1\n"
        "f 1/1/1 2/2/1 3/3/1\n"          # tangent +x
        "f 1/1/1 3/5/1 4/4/1\n"          # a sheared mapping: tangent a few degrees off +x
        "f 2/2/1 5/1/1 6/4/1\n"          # mirrored: u decreases along +x -> tangent -x
        "f 2/2/1 6/4/1 3/3/1\n")
    verts, idx, mids, names = cb.api.load_model(p)
    t = np.array([_half3(verts[i]) for i in idx])
    assert np.all(np.abs(np.linalg.norm(t, axis=1) - 1) < 2e-3) and np.all(t[:, 2] == 0)
    # corner 0 of triangles 0 and 1 (vertex 1, same uv): one smoothed tangent, one joined vertex
    assert idx[0] == idx[3] and 0.9 < t[0, 0] < 1.0 and abs(t[0, 1]) > 1e-3
    # vertex 2 is used by triangle 0 (tangent +x) and by the mirrored triangles (tangent -x): two vertices in the output
    assert idx[1] != idx[6] and t[1, 0] > 0.9 and t[6, 0] < -0.9
    assert np.array_equal(verts[idx[1]]["position"], verts[idx[6]]["position"])


def test_gltf_materials_and_given_tangents(cb, tmp_path):
    import base64
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    N = np.array([[0, 0, 1]] * 3, np.float32)
    UV = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    TAN = np.array([[0, 1, 0, 1]] * 3, np.float32)   # deliberately NOT dP/du: a TANGENT attribute is imported as it stands
    blob = P.tobytes() + N.tobytes() + UV.tobytes() + TAN.tobytes()
    views = [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 36}, {"buffer": 0, "byteOffset": 72, "byteLength": 24},
             {"buffer": 0, "byteOffset": 96, "byteLength": 48}]
    acc = [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC3"},
           {"bufferView": 2, "componentType": 5126, "count": 3, "type": "VEC2"}, {"bufferView": 3, "componentType": 5126, "count": 3, "type": "VEC4"}]
    doc = {"asset": {"version": "2.0"}, "buffers": [{"uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode(), "byteLength": len(blob)}],
           "bufferViews": views, "accessors": acc,
           "images": [{"uri": "maps/base.png"}, {"bufferView": 0, "mimeType": "image/png"}], "textures": [{"source": 0}, {"source": 1}],
           "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.4, 0.6, 1.0], "baseColorTexture": {"index": 0}}, "normalTexture": {"index": 1}},
                         {"pbrMetallicRoughness": {}}],
           "meshes": [{"name": "m", "primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2, "TANGENT": 3}, "material": 0},
                                                     {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "material": 1},
                                                     {"attributes": {"POSITION": 0, "NORMAL": 1}}]}]}
    path = tmp_path / "m.gltf"
    path.write_text(json.dumps(doc))
    verts, idx, mids, names, mats = cb.api.load_model(path, materials=True)
    assert list(mids) == [0, 1, 2] and len(mats) == 3
    base = str(tmp_path)
    assert mats[0] == {"albedo": base + "/maps/base.png", "normal": base + "/*1", "color": pytest.approx((0.2, 0.4, 0.6))}
    assert mats[1] == {"albedo": "", "normal": "", "color": (1.0, 1.0, 1.0)} and mats[2] == mats[1]
    assert all(np.array_equal(_half3(verts[i]), np.array([0, 1, 0], np.float32)) for i in idx[:3])      # the file's TANGENT
    # computed: FlipUVs makes v run along -y, u still along +x: tangent +x
    assert all(np.array_equal(_half3(verts[i]), np.array([1, 0, 0], np.float32)) for i in idx[3:6])
    assert all(np.array_equal(_half3(verts[i]), np.zeros(3, np.float32)) for i in idx[6:9])             # no UVs: zeros
