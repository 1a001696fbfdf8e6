"""Scene ingest (OBJ -> Vertex packing of ModelFileLoader.cpp:101-185) and the flat-buffer BVH cache —
SURVEY.md §8f ranks 3 and 4a.  The half-float packing is PINNED against glm::packHalf2x16 of the reference's
vendored glm, compiled into oracle/_ref."""
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _glm_like_pack(x):
    """numpy restatement of detail::toFloat16 for normal-range inputs: round to nearest, ties UP in magnitude."""
    x = np.asarray(x, np.float32)
    i = x.view(np.int32).astype(np.int64)
    s = (i >> 16) & 0x8000
    e = ((i >> 23) & 0xFF) - 112
    m = i & 0x7FFFFF
    up = (m & 0x1000) != 0
    m = np.where(up, m + 0x2000, m)
    carry = (m & 0x800000) != 0
    m = np.where(carry, 0, m)
    e = np.where(carry, e + 1, e)
    return (s | (e << 10) | (m >> 13)).astype(np.uint32)


def test_pack_half_matches_the_reference_glm(cb, ob):
    L = cb.api.load_library()
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(size=4000).astype(np.float32), (rng.random(2000, dtype=np.float32) * 2 - 1) * np.float32(1e-6),
                           (rng.random(2000, dtype=np.float32) * 2 - 1) * np.float32(7e4),
                           np.array([0.0, -0.0, 1.0, -1.0, 65504.0, 65519.9, 65520.0, 1e9, np.inf, -np.inf, np.nan, 6.1e-5, 5.96e-8, 2.98e-8, 1e-10], np.float32),
                           np.arange(1024, dtype=np.uint32).astype(np.uint32).__lshift__(13).__add__(0x3F801000).view(np.float32)])  # exact ties
    mine = np.array([L.cndl_pack_half2x16(float(v), float(-v)) for v in vals], np.uint32)
    normal = np.isfinite(vals) & (np.abs(vals) > 1e-4) & (np.abs(vals) < 6e4)
    assert np.array_equal(mine[normal] & 0xFFFF, _glm_like_pack(vals[normal]))
    with np.errstate(over="ignore"):
        rne = vals.astype(np.float16).view(np.uint16).astype(np.uint32)
    ties = np.arange(1024, dtype=np.uint32)
    assert np.mean((mine[normal] & 0xFFFF) == rne[normal]) > 0.9 and np.any((mine[-1024:] & 0xFFFF) != rne[-1024:])   # differs from IEEE RNE exactly on ties
    if ob.REFERENCE_ROOT.exists():
        ref = np.array([ob.ref_pack_half2x16(float(v), float(-v)) for v in vals], np.uint32)
        assert np.array_equal(mine, ref), "cndl_pack_half2x16 differs from the reference's glm::packHalf2x16"


def _write_obj(path, P, F, N=None, UV=None, groups=None):
    with open(path, "w") as f:
        f.write("# test asset\n")
        for p in P:
            f.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
        if UV is not None:
            for t in UV:
                f.write(f"vt {t[0]:.9g} {t[1]:.9g}\n")
        if N is not None:
            for n in N:
                f.write(f"vn {n[0]:.9g} {n[1]:.9g} {n[2]:.9g}\n")
        groups = groups or [("all", 0, len(F))]
        for name, lo, hi in groups:
            f.write(f"usemtl {name}\n")
            for a, b, c in F[lo:hi]:
                if N is not None and UV is not None:
                    f.write(f"f {a+1}/{a+1}/{a+1} {b+1}/{b+1}/{b+1} {c+1}/{c+1}/{c+1}\n")
                elif N is not None:
                    f.write(f"f {a+1}//{a+1} {b+1}//{b+1} {c+1}//{c+1}\n")
                else:
                    f.write(f"f {a+1} {b+1} {c+1}\n")


def test_obj_loader_reproduces_the_vertex_packing(cb, golden_meshes, tmp_path):
    P, F = golden_meshes["soup400"]
    rng = np.random.default_rng(1)
    N = rng.normal(size=P.shape).astype(np.float32)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    UV = rng.random((len(P), 2), dtype=np.float32) * 3 - 1
    path = tmp_path / "soup.obj"
    groups = [("stone", 0, 150), ("wood", 150, 151), ("glass", 151, len(F))]
    _write_obj(path, P, F, N, UV, groups)
    verts, idx, mids, names = cb.api.load_obj(path, first_mesh_number=7)
    assert names == ["stone", "wood", "glass"] and len(idx) == 3 * len(F)
    assert np.array_equal(mids, np.repeat([7, 8, 9], [150, 1, len(F) - 151]))
    # geometry survives: the same triangles, corner for corner (%.9g round-trips float32)
    assert np.array_equal(verts["position"][idx.reshape(-1, 3)][..., :3], P[F])
    assert np.all(verts["position"][:, 3] == 1.0)
    # indices of a mesh stay inside its own vertex range (per-mesh join, offset applied like BVHConstructor.cpp:981-1002)
    tri_first = idx.reshape(-1, 3).min(1)
    assert tri_first[150] > idx.reshape(-1, 3)[:150].max() and tri_first[151:].min() > idx.reshape(-1, 3)[150].max()
    # packing: data.x = packHalf2x16(n.xy), data.y = packHalf2x16(n.z, tan.x), data.z = packHalf2x16(tan.yz), texcoords = packHalf2x16(uv)
    L = cb.api.load_library()
    src = F.reshape(-1)
    halves = lambda u: np.array([u & 0xFFFF, u >> 16], dtype=np.uint16).view(np.float16).astype(np.float32)
    for k in rng.integers(0, len(idx), 300):
        v, s = verts[idx[k]], src[k]
        assert v["normal_tangent"][0] == L.cndl_pack_half2x16(float(N[s, 0]), float(N[s, 1]))
        assert (v["normal_tangent"][1] & 0xFFFF) == (L.cndl_pack_half2x16(float(N[s, 2]), 0.0) & 0xFFFF)
        assert v["texcoords"] == L.cndl_pack_half2x16(float(UV[s, 0]), float(np.float32(1.0) - UV[s, 1]))   # aiProcess_FlipUVs
        # aiProcess_CalcTangentSpace: a unit tangent orthogonal to the corner's normal (to half precision)
        t = np.array([halves(int(v["normal_tangent"][1]))[1], *halves(int(v["normal_tangent"][2]))])
        assert abs(np.linalg.norm(t) - 1.0) < 2e-3 and abs(float(t @ N[s])) < 2e-3
    # quads and negative indices; no normals / UVs -> zeros
    q = tmp_path / "quad.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nf -5 -4 -3 -2\ng top\nf 1 2 5\n")
    verts, idx, mids, names = cb.api.load_obj(q)
    assert len(idx) == 9 and list(mids) == [0, 0, 1] and names == ["default", "top"] and np.all(verts["texcoords"] == 0)
    assert np.all(verts["normal_tangent"][:, 2] == 0) and np.all(verts["normal_tangent"][:, 1] >> 16 == 0)   # no UV channel: no tangents (ModelFileLoader.cpp:138-143)
    # aiProcess_GenNormals: no `vn` in the file -> the quad's corners carry its face normal (0, 0, 1), the other face (0, -1, 0);
    # the quad's four corners are joined across its two triangles
    L = cb.api.load_library()
    assert len(verts) == 7 and np.all(verts["normal_tangent"][:4, 0] == L.cndl_pack_half2x16(0.0, 0.0)) and np.all(verts["normal_tangent"][:4, 1] == L.cndl_pack_half2x16(1.0, 0.0))
    assert np.all(verts["normal_tangent"][4:, 0] == L.cndl_pack_half2x16(0.0, -1.0)) and np.all(verts["normal_tangent"][4:, 1] == L.cndl_pack_half2x16(0.0, 0.0))
    assert np.array_equal(verts["position"][idx[:6]][:, :3], np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(cb.CandelaError, match="out of range"):
        cb.api.load_obj(bad)
    with pytest.raises(cb.CandelaError, match="cannot open"):
        cb.api.load_obj(tmp_path / "missing.obj")


def _write_gltf(tmp_path, kind):
    """Two meshes under a small node tree (root -> [child A (mesh 1), child B (mesh 0)], root has mesh 0 too):
    mesh 0 = indexed quad with normals + UVs (u16 indices, interleaved vertex buffer), mesh 1 = one unindexed triangle
    without normals.  kind: 'bin' (external buffer), 'uri' (base64) or 'glb'."""
    import base64
    import json
    import struct
    quad_p = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    quad_n = np.array([[0, 0, 1]] * 4, np.float32)
    quad_uv = np.array([[0, 0], [1, 0], [1, 0.25], [0, 0.75]], np.float32)
    inter = np.concatenate([quad_p, quad_n, quad_uv], 1).astype(np.float32)          # stride 32
    quad_i = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    tri_p = np.array([[0, 0, 2], [0, 3, 2], [4, 0, 2]], np.float32)
    blob = inter.tobytes() + quad_i.tobytes() + tri_p.tobytes()
    o_i, o_t = inter.nbytes, inter.nbytes + quad_i.nbytes
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"mesh": 0, "children": [1, 2], "translation": [100, 0, 0]}, {"mesh": 1, "name": "A"}, {"mesh": 0, "name": "B"}],
        "meshes": [{"name": "quad", "primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3}]},
                   {"name": "tri", "primitives": [{"attributes": {"POSITION": 4}}]}],
        "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": inter.nbytes, "byteStride": 32},
                        {"buffer": 0, "byteOffset": o_i, "byteLength": quad_i.nbytes}, {"buffer": 0, "byteOffset": o_t, "byteLength": tri_p.nbytes}],
        "accessors": [{"bufferView": 0, "byteOffset": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 4, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 24, "componentType": 5126, "count": 4, "type": "VEC2"},
                      {"bufferView": 1, "componentType": 5123, "count": 6, "type": "SCALAR"},
                      {"bufferView": 2, "componentType": 5126, "count": 3, "type": "VEC3"}],
    }
    if kind == "bin":
        (tmp_path / "geo.bin").write_bytes(blob)
        doc["buffers"] = [{"uri": "geo.bin", "byteLength": len(blob)}]
        path = tmp_path / "scene.gltf"
        path.write_text(json.dumps(doc))
    elif kind == "uri":
        doc["buffers"] = [{"uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode(), "byteLength": len(blob)}]
        path = tmp_path / "scene_uri.gltf"
        path.write_text(json.dumps(doc, indent=2))
    else:
        doc["buffers"] = [{"byteLength": len(blob)}]
        js = json.dumps(doc).encode()
        js += b" " * (-len(js) % 4)
        bn = blob + b"\0" * (-len(blob) % 4)
        path = tmp_path / "scene.glb"
        path.write_bytes(b"glTF" + struct.pack("<II", 2, 12 + 8 + len(js) + 8 + len(bn)) + struct.pack("<I", len(js)) + b"JSON" + js +
                         struct.pack("<I", len(bn)) + b"BIN\0" + bn)
    return path, quad_p, quad_uv, tri_p


@pytest.mark.parametrize("kind", ["bin", "uri", "glb"])
def test_gltf_loader(cb, tmp_path, kind):
    path, quad_p, quad_uv, tri_p = _write_gltf(tmp_path, kind)
    verts, idx, mids, names = cb.api.load_model(path, first_mesh_number=5)
    L = cb.api.load_library()
    # node walk: root (quad), child A (tri), child B (quad again); transforms are NOT applied, as in ProcessAssimpNode
    assert names == ["quad", "tri", "quad"] and list(mids) == [5, 5, 6, 7, 7]
    assert len(verts) == 4 + 3 + 4 and list(idx) == [0, 1, 2, 0, 2, 3, 4, 5, 6, 7, 8, 9, 7, 9, 10]
    assert np.array_equal(verts["position"][:4, :3], quad_p) and np.array_equal(verts["position"][4:7, :3], tri_p) and np.array_equal(verts["position"][7:, :3], quad_p)
    for k in range(4):
        assert verts["texcoords"][k] == L.cndl_pack_half2x16(float(quad_uv[k, 0]), float(np.float32(1) - quad_uv[k, 1]))      # FlipUVs
        assert verts["normal_tangent"][k, 0] == L.cndl_pack_half2x16(0.0, 0.0) and (verts["normal_tangent"][k, 1] & 0xFFFF) == L.cndl_pack_half2x16(1.0, 0.0)
        # aiProcess_CalcTangentSpace: a unit tangent in the quad's plane (normal = +z), pointing along +u = +x for this mapping
        tan = np.array([int(verts["normal_tangent"][k, 1]) >> 16, int(verts["normal_tangent"][k, 2]) & 0xFFFF, int(verts["normal_tangent"][k, 2]) >> 16],
                       dtype=np.uint16).view(np.float16).astype(np.float32)
        assert abs(np.linalg.norm(tan) - 1.0) < 2e-3 and tan[2] == 0.0 and tan[0] > 0.5
    # the bare triangle gets its face normal: cross((0,3,0), (4,0,0)) = (0,0,-12) -> (0,0,-1)
    assert np.all(verts["normal_tangent"][4:7, 1] == L.cndl_pack_half2x16(-1.0, 0.0)) and np.all(verts["texcoords"][4:7] == L.cndl_pack_half2x16(0.0, 0.0))


def test_gltf_errors(cb, tmp_path):
    (tmp_path / "a.gltf").write_text("{ not json")
    with pytest.raises(cb.CandelaError, match="JSON"):
        cb.api.load_model(tmp_path / "a.gltf")
    (tmp_path / "b.gltf").write_text('{"asset":{"version":"2.0"},"buffers":[{"uri":"nope.bin","byteLength":4}],"bufferViews":[],"accessors":[],"meshes":[]}')
    with pytest.raises(cb.CandelaError, match="cannot open buffer"):
        cb.api.load_model(tmp_path / "b.gltf")
    (tmp_path / "c.gltf").write_text('{"asset":{"version":"2.0"},"buffers":[{"uri":"data:application/octet-stream;base64,AAAA","byteLength":3}],'
                                     '"bufferViews":[{"buffer":0,"byteLength":3}],"accessors":[{"bufferView":0,"componentType":5126,"count":9,"type":"VEC3"}],'
                                     '"meshes":[{"primitives":[{"attributes":{"POSITION":0}}]}]}')
    with pytest.raises(cb.CandelaError, match="past its buffer"):
        cb.api.load_model(tmp_path / "c.gltf")
    # crafted sizes that would wrap a naive `off + (count - 1) * stride + elem` check: negative offsets / counts, a huge count, a huge stride
    tmpl = ('{"asset":{"version":"2.0"},"buffers":[{"uri":"data:application/octet-stream;base64,' + "A" * 64 + '","byteLength":48}],'
            '"bufferViews":[{"buffer":0,"byteLength":48%s}],"accessors":[{"bufferView":0,"componentType":5126,"count":%s,"type":"VEC3"%s}],'
            '"meshes":[{"primitives":[{"attributes":{"POSITION":0}}]}]}')
    for k, (view_extra, count, acc_extra, msg) in enumerate(((', "byteOffset": -16', "3", "", "negative"), ("", "-1", "", "negative"), ("", "3", ', "byteOffset": -4', "negative"),
                                                             ("", "1537228672809129302", "", "past its buffer"), (', "byteStride": -12', "3", "", "negative"),
                                                             (', "byteStride": 4611686018427387904', "5", "", "past its buffer"), ("", "9223372036854775807", "", "past its buffer|negative"))):
        f = tmp_path / f"crafted{k}.gltf"
        f.write_text(tmpl % (view_extra, count, acc_extra))
        with pytest.raises(cb.CandelaError, match=msg):
            cb.api.load_model(f)


@pytest.mark.gpu
def test_loaded_model_builds_and_cache_round_trips(cb, ob, golden_meshes, tmp_path):
    from helpers import rays_in_box
    P, F = golden_meshes["dragon"]
    path = tmp_path / "dragon.obj"
    _write_obj(path, P, F, groups=[("a", 0, 9000), ("b", 9000, len(F))])
    verts, idx, mids, _ = cb.api.load_obj(path, first_mesh_number=3)
    Ps, Fs = golden_meshes["soup400"]
    rays = rays_in_box(P.min(0) - 1, P.max(0) + 1, 50000, 8)
    for fmt, ofmt in ((cb.STACKLESS, ob.STACKLESS), (cb.STACK, ob.STACK)):
        ri = cb.RayIntersector(fmt)
        ri.AddObject(2, verts, idx, mids)
        ri.AddObject(3, cb.make_vertices(Ps), Fs.ravel(), np.full(len(Fs), 9, np.int32))
        ri.BufferData()
        ri.PushEntity(2)
        ri.PushEntity(3, model=np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.2], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32))
        ri.BufferEntities()
        sc = ob.Scene(ofmt)
        sc.add_object(2, verts, idx, mids)
        sc.add_object(3, ob.make_vertices(Ps), Fs.ravel(), np.full(len(Fs), 9, np.int32))
        nodes, tris, vv = ri.read_buffers()
        assert nodes.tobytes() == sc.nodes.tobytes() and tris.tobytes() == sc.tris.tobytes()
        want = ri.IntersectRays(rays)
        assert set(np.unique(want["mesh"])) >= {3, 4}
        cache = tmp_path / f"scene_{fmt}.cndl"
        ri.Save(cache)
        r2 = cb.RayIntersector(fmt)
        r2.Load(cache)
        r2.BufferData()
        r2.PushEntity(2)
        r2.PushEntity(3, model=np.array([[1, 0, 0, 0.5], [0, 1, 0, 0.2], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32))
        r2.BufferEntities()
        n2, t2, v2 = r2.read_buffers()
        assert n2.tobytes() == nodes.tobytes() and t2.tobytes() == tris.tobytes() and v2.tobytes() == vv.tobytes()
        assert r2.object_data(3) == ri.object_data(3)
        assert r2.IntersectRays(rays).tobytes() == want.tobytes()
        with pytest.raises(cb.CandelaError, match="empty context"):
            r2.Load(cache)
        # a corrupt cache file is refused at BufferData instead of being traversed: a vertex index outside the vertex buffer, a leaf range
        # outside the triangle buffer
        raw = bytearray(cache.read_bytes())
        head = 8 + 8 + 4 * 8                                             # magic, version + format, four 64-bit counts
        n_obj, n_nodes, n_tris, n_verts = np.frombuffer(bytes(raw[16:48]), dtype=np.uint64)
        node_size = 32 if fmt == cb.STACKLESS else 64
        tri_at = head + int(n_obj) * 28 + int(n_nodes) * node_size
        for what, at, value in (("vertex", tri_at + 16 * 7 + 4, int(n_verts) + 5), ("leaf", None, None)):
            bad = bytearray(raw)
            if what == "vertex":
                bad[at:at + 4] = np.int32(value).tobytes()
            else:
                nodes_at = head + int(n_obj) * 28
                nb = np.frombuffer(bytes(bad[nodes_at:nodes_at + int(n_nodes) * node_size]), dtype=np.int32).reshape(-1, node_size // 4).copy()
                packs = nb[:, 3]                                          # Min.w (stackless) / left child's Min.w (stack)
                k = int(np.nonzero((packs != -1) & (packs > 0))[0][5])
                nb[k, 3] = ((int(n_tris) + 100) << 4) | 2
                bad[nodes_at:nodes_at + int(n_nodes) * node_size] = nb.tobytes()
            cf = tmp_path / f"corrupt_{what}_{fmt}.cndl"
            cf.write_bytes(bytes(bad))
            r3 = cb.RayIntersector(fmt)
            r3.Load(cf)
            with pytest.raises(cb.CandelaError, match="outside the"):
                r3.BufferData()
            with pytest.raises(cb.CandelaError):
                r3.IntersectRays(rays[:10])                              # nothing committed: no traversal of the corrupt buffers
            r3.close()
        other = cb.RayIntersector(cb.STACK if fmt == cb.STACKLESS else cb.STACKLESS)
        with pytest.raises(cb.CandelaError, match="node format"):
            other.Load(cache)
        other.close(); r2.close(); ri.close()


def test_loader_survives_fuzzing_under_asan(tmp_path):
    """The host-side readers under AddressSanitizer + UBSan (tools/fuzz/): byte-level mutations of valid OBJ / glTF / GLB seeds and
    syntactically valid glTF documents whose fields lie (counts, offsets, strides, indices, node cycles, deep chains, container
    lengths).  Every case must load or be refused with an error string — any sanitizer report fails the run.  (Short run; the
    tools run hundreds of thousands of cases.)"""
    import shutil
    import subprocess
    import sys
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    probe = subprocess.run(["g++", "-fsanitize=address,undefined", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){return 0;}", text=True,
                           capture_output=True)
    if probe.returncode != 0:
        pytest.skip("g++ has no sanitizer runtime here")
    r = subprocess.run(["bash", str(ROOT / "tools" / "fuzz" / "run_fuzz_loader.sh"), "400", "5", str(tmp_path / "mut")], capture_output=True, text=True)
    assert r.returncode == 0 and "no sanitizer report" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "fuzz" / "craft_gltf.py"), "250", "5", str(tmp_path / "craft")], capture_output=True, text=True)
    assert r.returncode == 0 and "no sanitizer report" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_obj_lines_of_any_length(cb, tmp_path):
    """A statement is never cut in two: a 70,000-character comment whose tail looks like a vertex statement adds no vertex, and a
    polygon whose `f` line is longer than any fixed buffer is fan-triangulated whole."""
    n = 20000
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    lines = ["# " + "x" * 65533 + "v 9 9 9 and more comment"]           # the tail starts exactly where a 64 KiB buffer would have ended
    lines += [f"v {np.cos(a):.7f} {np.sin(a):.7f} 0" for a in ang]
    lines.append("f " + " ".join(str(k + 1) for k in range(n)))           # one convex polygon, ~129 KB of text
    p = tmp_path / "long.obj"
    p.write_text("\n".join(lines))                                         # no newline at the end of the file either
    verts, idx, mids, names = cb.api.load_obj(p)
    assert len(idx) == 3 * (n - 2) and len(mids) == n - 2
    assert len(verts) == n
    assert float(np.abs(verts["position"][:, :3]).max()) <= 1.0 + 1e-6      # the comment's "v 9 9 9" never became a vertex
