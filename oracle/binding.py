"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/oracle.h).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``candela_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_LIB_PATH = HERE / "_ref" / "libcandela_ref.so"
REFERENCE_ROOT = Path("/root/reference/Source")

STACKLESS, STACK = 0, 1
SWAP_NONE, SWAP_HASHED = 0, 1
CLOSEST, CLOSEST_IGNORE_TRANSPARENT, ANY = 0, 1, 2

VERTEX_DT = np.dtype([("position", "<f4", 4), ("normal_tangent", "<u4", 3), ("texcoords", "<u4")])
TRIANGLE_DT = np.dtype([("v", "<i4", 3), ("mesh", "<i4")])
NODE32_DT = np.dtype([("min", "<f4", 4), ("max", "<f4", 4)])
NODE64_DT = np.dtype([("lmin", "<f4", 4), ("lmax", "<f4", 4), ("rmin", "<f4", 4), ("rmax", "<f4", 4)])
ENTITY_DT = np.dtype([("model", "<f4", 16), ("inverse", "<f4", 16), ("node_offset", "<i4"), ("node_count", "<i4"), ("data", "<i4", 14)])
RAY_DT = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
ATTR_DT = np.dtype([("normal", "<f4", 3), ("uv", "<f4", 2), ("emissivity", "<f4"), ("alpha", "<f4"), ("mesh", "<i4")])
TEXREF_DT = np.dtype([("model_color", "<f4", 4), ("albedo", "<i4"), ("normal", "<i4"), ("pad", "<i4", 2)])   # BVH::TextureReferences, Intersector.h:32-37
MATERIAL_DT = np.dtype([("normal", "<f4", 3), ("uv", "<f4", 2), ("emissivity", "<f4"), ("alpha", "<f4"), ("mesh", "<i4"), ("albedo", "<f4", 3), ("albedo_ref", "<i4")])
HIT_DT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("w", "<f4"), ("mesh", "<i4"), ("tri", "<i4"), ("entity", "<i4"), ("iters", "<i4")])
assert (VERTEX_DT.itemsize, TRIANGLE_DT.itemsize, NODE32_DT.itemsize, NODE64_DT.itemsize, ENTITY_DT.itemsize, RAY_DT.itemsize, HIT_DT.itemsize) == (32, 16, 32, 64, 192, 32, 32)


def build_library(force: bool = False) -> Path:
    """Compiles oracle/liboracle.so (and oracle/_ref when /root/reference exists)."""
    srcs = [HERE / "oracle_build.cpp", HERE / "oracle_trace.cpp", HERE / "oracle_raygen.cpp", HERE / "oracle.h", HERE / "exact_math_ref.h"]
    stale = force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if stale:
        subprocess.run(["make", "-C", str(HERE), "all"] + (["-B"] if force else []), check=True, capture_output=True)
    if REFERENCE_ROOT.exists():  # make decides whether oracle/_ref is stale (it depends on the shim sources)
        subprocess.run(["make", "-C", str(HERE), "ref"] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_library()
        L = C.CDLL(str(LIB_PATH))
        vp, u64, i32, u32p, i32p = C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
        L.orc_build.restype = vp
        L.orc_build.argtypes = [C.c_int, vp, u64, vp, u64, vp, i32, C.c_int, u64]
        L.orc_bvh_free.argtypes = [vp]
        L.orc_bvh_node_count.restype = u64
        L.orc_bvh_node_count.argtypes = [vp]
        L.orc_bvh_tri_count.restype = u64
        L.orc_bvh_tri_count.argtypes = [vp]
        L.orc_bvh_stats.argtypes = [vp, vp]
        L.orc_bvh_fetch.argtypes = [vp, vp, vp]
        L.orc_bvh_fetch_order.argtypes = [vp, vp]
        L.orc_bvh_adopt_flips.restype = C.c_int64
        L.orc_bvh_adopt_flips.argtypes = [vp, vp, u64]
        L.orc_bvh_sah_cost.restype = C.c_double
        L.orc_bvh_sah_cost.argtypes = [vp]
        L.orc_concat_meshes.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        L.orc_rebase_triangles.argtypes = [vp, u64, C.c_uint32]
        L.orc_make_entity.argtypes = [vp, i32, i32, C.c_float, C.c_float, vp]
        L.orc_primary_rays.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        L.orc_trace.argtypes = [C.c_int, C.c_int, vp, u64, vp, vp, vp, i32, vp, u64, vp, vp, vp, C.c_int]
        L.orc_brute_force.argtypes = [vp, u64, vp, vp, i32, vp, u64, vp, C.c_int]
        L.orc_get_data.argtypes = [vp, vp, vp, vp, u64, vp]
        L.orc_get_data_material.argtypes = [vp, vp, vp, vp, u64, vp, u64, vp]
        L.orc_collide_boxes.argtypes = [vp, u64, vp, vp, vp, i32, vp, u64, vp]
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


def make_vertices(positions: np.ndarray) -> np.ndarray:
    """32-byte Vertex records (Utils/Vertex.h:7-12) with w = 1 and zero packed attributes."""
    positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    v = np.zeros(len(positions), dtype=VERTEX_DT)
    v["position"][:, :3] = positions
    v["position"][:, 3] = 1.0
    return v


class BuildResult:
    def __init__(self, fmt, nodes, tris, order, stats, sah_cost):
        self.format = fmt
        self.nodes = nodes
        self.tris = tris
        self.order = order
        self.stats = dict(zip(("created", "leaves", "split_fails", "max_stack", "stack_slots", "depth"), (int(x) for x in stats)))
        self.sah_cost = sah_cost


def build(fmt: int, verts: np.ndarray, indices: np.ndarray, mesh_ids: np.ndarray | None = None, t_offset: int = 0,
          swap_policy: int = SWAP_NONE, swap_seed: int = 0, adopt_flips_from: np.ndarray | None = None) -> BuildResult:
    L = lib()
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    indices = np.ascontiguousarray(indices, dtype=np.uint32).ravel()
    if mesh_ids is not None:
        mesh_ids = np.ascontiguousarray(mesh_ids, dtype=np.int32)
    h = L.orc_build(fmt, _p(verts), len(verts), _p(indices), len(indices), _p(mesh_ids), t_offset, swap_policy, swap_seed)
    if not h:
        raise ValueError("orc_build rejected the input")
    try:
        flips = None
        if adopt_flips_from is not None:
            ref = np.ascontiguousarray(adopt_flips_from, dtype=NODE32_DT)
            flips = L.orc_bvh_adopt_flips(h, _p(ref), len(ref))
            if flips < 0:
                raise ValueError("reference node buffer does not describe the same tree")
        n, t = L.orc_bvh_node_count(h), L.orc_bvh_tri_count(h)
        nodes = np.zeros(n, dtype=NODE32_DT if fmt == STACKLESS else NODE64_DT)
        tris = np.zeros(t, dtype=TRIANGLE_DT)
        order = np.zeros(t, dtype=np.int32)
        stats = np.zeros(6, dtype=np.uint64)
        L.orc_bvh_fetch(h, _p(nodes), _p(tris))
        L.orc_bvh_fetch_order(h, _p(order))
        L.orc_bvh_stats(h, _p(stats))
        res = BuildResult(fmt, nodes, tris, order, stats, float(L.orc_bvh_sah_cost(h)))
        res.flips = flips
        return res
    finally:
        L.orc_bvh_free(h)


def concat_meshes(meshes):
    """meshes: list of (verts[VERTEX_DT], indices[u32], global_mesh_number). BVHConstructor.cpp:981-1002."""
    L = lib()
    verts = np.ascontiguousarray(np.concatenate([m[0] for m in meshes]), dtype=VERTEX_DT)
    idx = np.ascontiguousarray(np.concatenate([np.asarray(m[1], dtype=np.uint32).ravel() for m in meshes]))
    vc = np.array([len(m[0]) for m in meshes], dtype=np.uint64)
    ic = np.array([np.asarray(m[1]).size for m in meshes], dtype=np.uint64)
    nums = np.array([m[2] for m in meshes], dtype=np.int32)
    ov = np.zeros(len(verts), dtype=VERTEX_DT)
    oi = np.zeros(len(idx), dtype=np.uint32)
    om = np.zeros(len(idx) // 3, dtype=np.int32)
    L.orc_concat_meshes(len(meshes), _p(verts), _p(vc), _p(idx), _p(ic), _p(nums), _p(ov), _p(oi), _p(om))
    return ov, oi, om


def make_entity(model: np.ndarray, node_offset: int, node_count: int, emissive: float = 0.0, translucency: float = 0.0) -> np.ndarray:
    """model: 4x4 in the mathematical (row, column) convention; stored column-major like glm."""
    m = np.ascontiguousarray(np.asarray(model, dtype=np.float32).reshape(4, 4).T).ravel()
    out = np.zeros(1, dtype=ENTITY_DT)
    lib().orc_make_entity(_p(m), node_offset, node_count, emissive, translucency, _p(out))
    return out


class Scene:
    """RayIntersector<T> restated (Intersector.h:170-216): concatenated buffers of all objects plus entities."""

    def __init__(self, fmt: int):
        self.format = fmt
        self.nodes = np.zeros(0, dtype=NODE32_DT if fmt == STACKLESS else NODE64_DT)
        self.verts = np.zeros(0, dtype=VERTEX_DT)
        self.tris = np.zeros(0, dtype=TRIANGLE_DT)
        self.entities = np.zeros(0, dtype=ENTITY_DT)
        self.objects = {}
        self.builds = {}

    def add_object(self, object_id: int, verts, indices, mesh_ids=None, swap_policy=SWAP_NONE, swap_seed=0):
        b = build(self.format, verts, indices, mesh_ids, t_offset=len(self.tris), swap_policy=swap_policy, swap_seed=swap_seed)
        self.objects[object_id] = dict(node_offset=len(self.nodes), node_count=len(b.nodes), tri_offset=len(self.tris),
                                       tri_count=len(b.tris), vert_offset=len(self.verts), vert_count=len(verts))
        tris = b.tris.copy()
        lib().orc_rebase_triangles(_p(tris), len(tris), len(self.verts))
        self.nodes = np.concatenate([self.nodes, b.nodes])
        self.tris = np.concatenate([self.tris, tris])
        self.verts = np.concatenate([self.verts, np.ascontiguousarray(verts, dtype=VERTEX_DT)])
        self.builds[object_id] = b
        return b

    def push_entity(self, object_id: int, model=None, emissive=0.0, translucency=0.0):
        if object_id not in self.objects:
            raise KeyError("Trying to push entity whose parent object hasn't been added to global BVH")
        o = self.objects[object_id]
        e = make_entity(np.eye(4, dtype=np.float32) if model is None else model, o["node_offset"], o["node_count"], emissive, translucency)
        self.entities = np.concatenate([self.entities, e])

    def trace(self, kind, rays, nthreads=1):
        return trace(self.format, kind, self.nodes, self.tris, self.verts, self.entities, rays, nthreads)


def trace(fmt, kind, nodes, tris, verts, entities, rays, nthreads=1):
    """Returns (hits or any_t, counters dict)."""
    L = lib()
    rays = np.ascontiguousarray(rays, dtype=RAY_DT)
    nodes = np.ascontiguousarray(nodes)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    counters = np.zeros(4, dtype=np.uint64)
    R = len(rays)
    if kind == ANY:
        out = np.zeros(R, dtype=np.float32)
        L.orc_trace(fmt, kind, _p(nodes), len(nodes), _p(tris), _p(verts), _p(entities), len(entities), _p(rays), R, None, _p(out), _p(counters), nthreads)
    else:
        out = np.zeros(R, dtype=HIT_DT)
        L.orc_trace(fmt, kind, _p(nodes), len(nodes), _p(tris), _p(verts), _p(entities), len(entities), _p(rays), R, _p(out), None, _p(counters), nthreads)
    c = dict(node_iters=int(counters[0]), tri_tests=int(counters[1]), capped=int(counters[2]), hits=int(counters[3]), rays=R)
    return out, c


def brute_force(tris, verts, entities, rays, nthreads=1):
    rays = np.ascontiguousarray(rays, dtype=RAY_DT)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    out = np.zeros(len(rays), dtype=HIT_DT)
    lib().orc_brute_force(_p(tris), len(tris), _p(verts), _p(entities), len(entities), _p(rays), len(rays), _p(out), nthreads)
    return out


def get_data(tris, verts, entities, hits):
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    hits = np.ascontiguousarray(hits, dtype=HIT_DT)
    out = np.zeros(len(hits), dtype=ATTR_DT)
    lib().orc_get_data(_p(tris), _p(verts), _p(entities), _p(hits), len(hits), _p(out))
    return out


def get_data_material(tris, verts, entities, refs, hits):
    """GetData with the BVHTextureReferences table (oracle_trace.cpp: orc_get_data_material)."""
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    refs = np.ascontiguousarray(refs, dtype=TEXREF_DT)
    hits = np.ascontiguousarray(hits, dtype=HIT_DT)
    out = np.zeros(len(hits), dtype=MATERIAL_DT)
    lib().orc_get_data_material(_p(tris), _p(verts), _p(entities), _p(refs), len(refs), _p(hits), len(hits), _p(out))
    return out


def texture_references(materials):
    """RayIntersector::GenerateMeshTextureReferences (Intersector.h:367-402) restated.  materials: one
    (albedo_handle, albedo_valid, normal_handle, normal_valid, (r, g, b)) per mesh, i.e. what GetTextureCachedDataForPath
    (Texture.cpp:168-188) returns for _MeshMaterialData::Albedo / Normal plus its ModelColor.  Returns (table, handles) where
    handles[i] is the handle bound to Textures[i] (m_TextureHandleReferenceMap, :423-427)."""
    data_map = {}                                   # :370-372  DataMap (handle -> index), cleared on every call
    last = 0                                        # :375
    out = np.zeros(len(materials), dtype=TEXREF_DT)
    for i, (a, va, b, vb, color) in enumerate(materials):
        if a not in data_map:                       # :390-392: an index is taken whether or not the path was valid
            data_map[a] = last
            last += 1
        if b not in data_map:                       # :394-396
            data_map[b] = last
            last += 1
        out[i]["model_color"] = (color[0], color[1], color[2], 1.0)   # :401 glm::vec4(ModelColor, 1.0f)
        out[i]["albedo"] = data_map[a] if va else -1                  # :398
        out[i]["normal"] = data_map[b] if vb else -1                  # :399
    handles = np.zeros(last, dtype=np.uint64)
    for h, k in data_map.items():
        handles[k] = h
    return out, handles


BOX_DT = np.dtype([("min", "<f4", 3), ("pad0", "<f4"), ("max", "<f4", 3), ("pad1", "<f4")])
COLLISION_DT = np.dtype([("collided", "<i4"), ("mesh", "<i4"), ("tri", "<i4"), ("entity", "<i4")])


def make_boxes(mins, maxs):
    b = np.zeros(len(mins), dtype=BOX_DT)
    b["min"], b["max"] = np.asarray(mins, np.float32), np.asarray(maxs, np.float32)
    return b


def collide_boxes(nodes, tris, verts, entities, boxes):
    """Physics::CollideBox restated (oracle_trace.cpp) for a batch of boxes; stackless buffers only."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE32_DT)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DT)
    out = np.zeros(len(boxes), dtype=COLLISION_DT)
    lib().orc_collide_boxes(_p(nodes), len(nodes), _p(tris), _p(verts), _p(entities), len(entities), _p(boxes), len(boxes), _p(out))
    return out


def pack_vertices(positions, normals, uvs):
    """Vertex records with packed half-float normal / uv like ModelFileLoader.cpp:133-155 (tangent = 0)."""
    v = make_vertices(positions)
    n16 = np.asarray(normals, dtype=np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)
    t16 = np.asarray(uvs, dtype=np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)
    v["normal_tangent"][:, 0] = n16[:, 0] | (n16[:, 1] << 16)
    v["normal_tangent"][:, 1] = n16[:, 2]
    v["texcoords"] = t16[:, 0] | (t16[:, 1] << 16)
    return v


def primary_rays(inv_view: np.ndarray, inv_proj: np.ndarray, W: int, H: int) -> np.ndarray:
    """inv_view / inv_proj: 4x4 (row, column) matrices; passed column-major like glm uniforms."""
    iv = np.ascontiguousarray(np.asarray(inv_view, dtype=np.float32).reshape(4, 4).T).ravel()
    ip = np.ascontiguousarray(np.asarray(inv_proj, dtype=np.float32).reshape(4, 4).T).ravel()
    rays = np.zeros(W * H, dtype=RAY_DT)
    lib().orc_primary_rays(_p(iv), _p(ip), W, H, _p(rays))
    return rays


# --- the unmodified reference builder (only where /root/reference exists) ---------------------------------------

_ref = None


def reference_available() -> bool:
    return REF_LIB_PATH.exists() or REFERENCE_ROOT.exists()


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        if not REF_LIB_PATH.exists():
            build_library()
        R = C.CDLL(str(REF_LIB_PATH))
        vp = C.c_void_p
        R.ref_sizes.argtypes = [vp]
        R.ref_build.restype = C.c_int
        R.ref_build.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp]
        R.ref_fetch.argtypes = [vp, vp, vp]
        R.ref_glsl_trace.argtypes = [C.c_int, C.c_int, vp, C.c_uint64, vp, vp, vp, C.c_int32, vp, C.c_uint64, vp]
        R.ref_glsl_trace_mt.argtypes = [C.c_int, C.c_int, vp, C.c_uint64, vp, vp, vp, C.c_int32, vp, C.c_uint64, vp, C.c_int]
        R.ref_glsl_primary_rays.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        R.ref_glm_inverse.argtypes = [vp, vp]
        R.ref_glsl_sample.argtypes = [C.c_int, vp, vp, C.c_float, C.c_uint64, vp]
        R.ref_glsl_get_data.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_uint64, vp]
        R.ref_glsl_get_data_material.argtypes = [vp, vp, vp, C.c_int32, vp, vp, C.c_uint64, vp, vp]
        R.ref_pack_half2x16.argtypes = [C.c_float, C.c_float]
        R.ref_pack_half2x16.restype = C.c_uint32
        R.ref_collide_box.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, vp, C.c_uint64, vp]
        _ref = R
    return _ref


def ref_sizes():
    out = np.zeros(4, dtype=np.int32)
    ref_lib().ref_sizes(_p(out))
    return tuple(int(x) for x in out)


def ref_build(fmt: int, meshes, t_offset: int = 0):
    """Runs Candela::BVH::BuildBVH itself. meshes as in concat_meshes. Returns (nodes, tris, verts)."""
    R = ref_lib()
    verts = np.ascontiguousarray(np.concatenate([m[0] for m in meshes]), dtype=VERTEX_DT)
    idx = np.ascontiguousarray(np.concatenate([np.asarray(m[1], dtype=np.uint32).ravel() for m in meshes]))
    vc = np.array([len(m[0]) for m in meshes], dtype=np.uint64)
    ic = np.array([np.asarray(m[1]).size for m in meshes], dtype=np.uint64)
    nums = np.array([m[2] for m in meshes], dtype=np.int32)
    n = np.zeros(3, dtype=np.uint64)
    rc = R.ref_build(fmt, len(meshes), _p(verts), _p(vc), _p(idx), _p(ic), _p(nums), t_offset,
                     n[0:1].ctypes.data_as(C.c_void_p), n[1:2].ctypes.data_as(C.c_void_p), n[2:3].ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError("the reference builder crashes on meshes with fewer than 100 triangles")
    nodes = np.zeros(int(n[0]), dtype=NODE32_DT if fmt == STACKLESS else NODE64_DT)
    tris = np.zeros(int(n[1]), dtype=TRIANGLE_DT)
    ov = np.zeros(int(n[2]), dtype=VERTEX_DT)
    R.ref_fetch(_p(nodes), _p(tris), _p(ov))
    return nodes, tris, ov


def ref_collide_boxes(nodes, tris, verts, entities, boxes) -> np.ndarray:
    """Physics::CollideBox of the compiled reference itself (oracle/_ref): one 0/1 answer per box."""
    nodes = np.ascontiguousarray(nodes, dtype=NODE32_DT)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DT)
    out = np.zeros(len(boxes), dtype=np.int32)
    ref_lib().ref_collide_box(_p(nodes), len(nodes), _p(tris), len(tris), _p(verts), len(verts), _p(entities), len(entities), _p(boxes), len(boxes), _p(out))
    return out


def ref_pack_half2x16(x: float, y: float) -> int:
    """glm::packHalf2x16 of the reference's vendored glm (oracle/_ref)."""
    return int(ref_lib().ref_pack_half2x16(float(x), float(y)))


def ref_glsl_trace(fmt, kind, nodes, tris, verts, entities, rays, nthreads=1):
    """The reference's own GLSL traversal (TraverseBVHStackless.glsl / TraverseBVHStack.glsl), transliterated
    syntactically and compiled against its vendored glm (oracle/_ref): hit records, or any-hit distances."""
    rays = np.ascontiguousarray(rays, dtype=RAY_DT)
    nodes = np.ascontiguousarray(nodes)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    out = np.zeros(len(rays), dtype=np.float32 if kind == ANY else HIT_DT)
    ref_lib().ref_glsl_trace_mt(fmt, kind, _p(nodes), len(nodes), _p(tris), _p(verts), _p(entities), len(entities), _p(rays), len(rays), _p(out), nthreads)
    return out


def ref_glsl_get_data(tris, verts, entities, hits):
    """GetData of the compiled reference shader (texture references with Albedo = -1)."""
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    hits = np.ascontiguousarray(hits, dtype=HIT_DT)
    out = np.zeros(len(hits), dtype=ATTR_DT)
    ref_lib().ref_glsl_get_data(_p(tris), _p(verts), _p(entities), len(entities), _p(hits), len(hits), _p(out))
    return out


def ref_glsl_get_data_material(tris, verts, entities, refs, hits):
    """GetData of the compiled reference shader with a BVHTextureReferences table.  Returns (records, tex_uv): `albedo_ref` is the index
    of the sampler the shader called texture() with (-1: it did not), tex_uv the UV it passed.  Every mesh must be inside the table."""
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    refs = np.ascontiguousarray(refs, dtype=TEXREF_DT)
    hits = np.ascontiguousarray(hits, dtype=HIT_DT)
    assert len(hits) == 0 or int(hits["mesh"].max()) < len(refs)
    out = np.zeros(len(hits), dtype=MATERIAL_DT)
    tex_uv = np.zeros((len(hits), 2), dtype=np.float32)
    ref_lib().ref_glsl_get_data_material(_p(tris), _p(verts), _p(entities), len(entities), _p(refs), _p(hits), len(hits), _p(out), _p(tex_uv))
    return out, tex_uv


def ref_primary_rays(inv_view, inv_proj, W, H) -> np.ndarray:
    """Primary rays formed by the reference shader's own GetRayDirectionAt + main() arithmetic (oracle/_ref)."""
    iv = np.ascontiguousarray(np.asarray(inv_view, dtype=np.float32).reshape(4, 4).T).ravel()
    ip = np.ascontiguousarray(np.asarray(inv_proj, dtype=np.float32).reshape(4, 4).T).ravel()
    rays = np.zeros(W * H, dtype=RAY_DT)
    ref_lib().ref_glsl_primary_rays(_p(iv), _p(ip), W, H, _p(rays))
    return rays


def ref_glm_inverse(model) -> np.ndarray:
    """glm::inverse of the reference's vendored glm; model and result are 4x4 (row, column) matrices."""
    m = np.ascontiguousarray(np.asarray(model, dtype=np.float32).reshape(4, 4).T).ravel()
    out = np.zeros(16, dtype=np.float32)
    ref_lib().ref_glm_inverse(_p(m), _p(out))
    return out.reshape(4, 4).T.copy()


def ref_sample(which: int, normals, xi, roughness: float = 0.0) -> np.ndarray:
    """The reference's own direction samplers, compiled (Include/Sampling.glsl): 0 CosWeightedHemisphere, 1 SampleGGXVNDF."""
    n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    x = np.ascontiguousarray(xi, dtype=np.float32).reshape(-1, 2)
    out = np.zeros_like(n)
    ref_lib().ref_glsl_sample(which, _p(n), _p(x), roughness, len(n), _p(out))
    return out



# --- ray generators (oracle_raygen.cpp) ----------------------------------------------------------------------------------

GEN_DIFFUSE, GEN_SPECULAR, GEN_SHADOW = 0, 1, 2
GEN_BUCKET_OCTANTS = 1
SAMPLE_COS_HEMISPHERE, SAMPLE_GGX_VNDF, SAMPLE_STOCHASTIC_REFLECTION, SAMPLE_CONE, SAMPLE_PROBE = 0, 1, 2, 3, 4


class RaygenParams(C.Structure):
    """orc_raygen_params (same fields as cndl_raygen_params up to `light_cone`)."""
    _fields_ = [("kind", C.c_int32), ("spp", C.c_int32), ("seed", C.c_uint32), ("flags", C.c_uint32), ("offset", C.c_float), ("tmax", C.c_float),
                ("roughness", C.c_float), ("light_dir", C.c_float * 3), ("light_cone", C.c_float)]


def _raygen_protos():
    L = lib()
    if getattr(L, "_raygen_ready", False):
        return L
    vp, u64 = C.c_void_p, C.c_uint64
    for f in ("orc_xsin", "orc_xcos", "orc_xacos"):
        getattr(L, f).argtypes = [C.c_float]
        getattr(L, f).restype = C.c_float
    L.orc_xpow.argtypes = [C.c_float, C.c_float]
    L.orc_xpow.restype = C.c_float
    L.orc_xmath_batch.argtypes = [C.c_int, vp, vp, u64, vp]
    L.orc_sample_directions.argtypes = [C.c_int, vp, vp, vp, vp, C.c_float, u64, vp]
    L.orc_hash2_stream.argtypes = [C.c_uint32, C.c_uint32, vp]
    L.orc_stream_key.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_stream_key.restype = C.c_uint32
    L.orc_generate_rays.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp]
    L.orc_generate_rays.restype = u64
    L.orc_probe_rays.argtypes = [vp, vp, vp, C.c_uint32, vp]
    L._raygen_ready = True
    return L


def xmath(which: int, x, y=None) -> np.ndarray:
    """The DEFINED sin (0) / cos (1) / acos (2) / pow (3) of exact_math_ref.h on float32 arrays."""
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    y = None if y is None else np.ascontiguousarray(y, dtype=np.float32).ravel()
    out = np.zeros_like(x)
    _raygen_protos().orc_xmath_batch(which, _p(x), _p(y), len(x), _p(out))
    return out


def stream_keys(seed: int, elements) -> np.ndarray:
    L = _raygen_protos()
    return np.array([L.orc_stream_key(seed, int(e)) for e in elements], dtype=np.uint32)


def hash2_stream(key: int, m: int) -> np.ndarray:
    out = np.zeros(m, dtype=np.float32)
    _raygen_protos().orc_hash2_stream(int(key), m, _p(out))
    return out


def _sample_args(normals, incident, xi, keys):
    n = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    i = None if incident is None else np.ascontiguousarray(incident, dtype=np.float32).reshape(-1, 3)
    x = None if xi is None else np.ascontiguousarray(xi, dtype=np.float32).reshape(-1, 2)
    k = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint32).ravel()
    count = len(n) if n is not None else len(k)
    return n, i, x, k, count


def sample_directions(which: int, normals=None, incident=None, xi=None, keys=None, roughness: float = 0.0) -> np.ndarray:
    """The oracle's restatement of the shaders' direction samplers (SAMPLE_*)."""
    n, i, x, k, count = _sample_args(normals, incident, xi, keys)
    out = np.zeros((count, 3), dtype=np.float32)
    _raygen_protos().orc_sample_directions(which, _p(n), _p(i), _p(x), _p(k), roughness, count, _p(out))
    return out


def ref_sample_directions(which: int, normals=None, incident=None, xi=None, keys=None, roughness: float = 0.0) -> np.ndarray:
    """The shader functions themselves, compiled (oracle/_ref, namespace ref_raygen): hash2() and the implementation-defined
    built-ins bound to the oracle's definitions, everything else the shader's own text over the reference's glm."""
    R = ref_lib()
    R.ref_glsl_sample_directions.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p]
    n, i, x, k, count = _sample_args(normals, incident, xi, keys)
    out = np.zeros((count, 3), dtype=np.float32)
    R.ref_glsl_sample_directions(which, _p(n), _p(i), _p(x), _p(k), roughness, count, _p(out))
    return out


def generate_rays(rays, hits, tris, verts, entities, kind=GEN_DIFFUSE, spp=1, seed=1, offset=0.05, tmax=1.0e6, roughness=0.0,
                  light_dir=(0.0, 1.0, 0.0), light_cone=0.0, bucket_octants=False, ids=None):
    """cndl_generate_rays_device restated: returns (rays_out, parent, ids_out)."""
    L = _raygen_protos()
    rays = np.ascontiguousarray(rays, dtype=RAY_DT)
    hits = np.ascontiguousarray(hits, dtype=HIT_DT)
    tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DT)
    ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
    p = RaygenParams(kind, spp, seed, GEN_BUCKET_OCTANTS if bucket_octants else 0, offset, tmax, roughness, (C.c_float * 3)(*light_dir), light_cone)
    out = np.zeros(len(rays) * spp, dtype=RAY_DT)
    parent = np.zeros(len(rays) * spp, dtype=np.uint32)
    ids_out = np.zeros(len(rays) * spp, dtype=np.uint32)
    n = int(L.orc_generate_rays(C.byref(p), _p(rays), _p(hits), _p(ids), len(rays), _p(tris), _p(verts), _p(entities), _p(out), _p(parent), _p(ids_out)))
    return out[:n].copy(), parent[:n].copy(), ids_out[:n].copy()


def probe_rays(box_origin, size, res, seed: int) -> np.ndarray:
    """Probe-update rays (UpdateRadianceProbes.glsl:408-427) for a res[0] x res[1] x res[2] grid."""
    o = np.ascontiguousarray(box_origin, dtype=np.float32)
    s = np.ascontiguousarray(size, dtype=np.float32)
    r = np.ascontiguousarray(res, dtype=np.int32)
    out = np.zeros(int(r[0]) * int(r[1]) * int(r[2]), dtype=RAY_DT)
    _raygen_protos().orc_probe_rays(_p(o), _p(s), _p(r), seed, _p(out))
    return out
