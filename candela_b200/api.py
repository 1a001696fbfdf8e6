"""Host-side mirror of Candela's ``RayIntersector<T>`` (Source/Core/BVH/Intersector.h:60-124) over
the C ABI in include/candela_b200.h.  Method names and argument meaning follow the reference;
where the reference throws a string literal this raises ``CandelaError`` with the same text.

This is ctypes glue only: every query runs in libcandela_b200.so on the GPU.  There is no CPU
path; importing on a machine without the compiled library or without a B200 fails loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import _build

STACKLESS, STACK = 0, 1  # BVH::StacklessTraversalNode / BVH::StackTraversalNode (Intersector.h:39-40)
BUILDER_SAH_EXACT, BUILDER_LBVH = 0, 1
SWAP_NONE, SWAP_HASHED = 0, 1
IGNORE_TRANSPARENT = 1

VERTEX_DT = np.dtype([("position", "<f4", 4), ("normal_tangent", "<u4", 3), ("texcoords", "<u4")])
TRIANGLE_DT = np.dtype([("v", "<i4", 3), ("mesh", "<i4")])
NODE_DT = np.dtype([("min", "<f4", 4), ("max", "<f4", 4)])
STACK_NODE_DT = np.dtype([("lmin", "<f4", 4), ("lmax", "<f4", 4), ("rmin", "<f4", 4), ("rmax", "<f4", 4)])
ENTITY_DT = np.dtype([("model", "<f4", 16), ("inverse", "<f4", 16), ("node_offset", "<i4"), ("node_count", "<i4"), ("data", "<i4", 14)])
RAY_DT = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4")])
HIT_DT = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("w", "<f4"), ("mesh", "<i4"), ("tri", "<i4"), ("entity", "<i4"), ("iters", "<i4")])
BOX_DT = np.dtype([("min", "<f4", 3), ("pad0", "<f4"), ("max", "<f4", 3), ("pad1", "<f4")])
COLLISION_DT = np.dtype([("collided", "<i4"), ("mesh", "<i4"), ("tri", "<i4"), ("entity", "<i4")])
ATTR_DT = np.dtype([("normal", "<f4", 3), ("uv", "<f4", 2), ("emissivity", "<f4"), ("alpha", "<f4"), ("mesh", "<i4")])
TEXREF_DT = np.dtype([("model_color", "<f4", 4), ("albedo", "<i4"), ("normal", "<i4"), ("pad", "<i4", 2)])     # cndl_texture_reference
MESH_MATERIAL_DT = np.dtype([("albedo_handle", "<u8"), ("normal_handle", "<u8"), ("albedo_valid", "<i4"), ("normal_valid", "<i4"),
                             ("model_color", "<f4", 3), ("pad", "<f4")])                                           # cndl_mesh_material
MATERIAL_DT = np.dtype([("normal", "<f4", 3), ("uv", "<f4", 2), ("emissivity", "<f4"), ("alpha", "<f4"), ("mesh", "<i4"), ("albedo", "<f4", 3),
                        ("albedo_ref", "<i4")])                                                                    # cndl_hit_material

EXPORTS = [
    "cndl_abi_version", "cndl_create", "cndl_destroy", "cndl_last_error", "cndl_add_object", "cndl_add_prebuilt_object",
    "cndl_node_count", "cndl_triangle_count", "cndl_vertex_count", "cndl_get_object", "cndl_commit", "cndl_read_buffers",
    "cndl_device_buffers", "cndl_push_entity", "cndl_push_entity_records", "cndl_buffer_entities", "cndl_entity_count",
    "cndl_intersect_closest", "cndl_intersect_any", "cndl_intersect_closest_device", "cndl_intersect_any_device",
    "cndl_intersect_primary", "cndl_intersect_primary_device", "cndl_generate_bounce_rays_device", "cndl_host_alloc", "cndl_host_alloc_write_combined", "cndl_host_free",
    "cndl_set_traversal_mode", "cndl_set_tuning", "cndl_launch_count", "cndl_last_build_ms", "cndl_get_data", "cndl_get_data_device", "cndl_generate_rays_device", "cndl_collide_boxes", "cndl_collide_boxes_device",
    "cndl_model_load_obj", "cndl_model_load_gltf", "cndl_model_load", "cndl_model_free", "cndl_model_vertex_count", "cndl_model_index_count", "cndl_model_mesh_count", "cndl_model_vertices",
    "cndl_model_indices", "cndl_model_mesh_ids", "cndl_model_mesh_name", "cndl_add_model", "cndl_pack_half2x16", "cndl_save", "cndl_load", "cndl_build_bvh",
    "cndl_generate_probe_rays_device",
    "cndl_frame_records", "cndl_frame_shard_records", "cndl_frame_record_bytes", "cndl_trace_frame_device", "cndl_frame_submit", "cndl_frame_wait",
    "cndl_trace_frame", "cndl_frame_rays_traced", "cndl_frame_untile_device",
    "cndl_multi_create", "cndl_multi_destroy", "cndl_multi_device_count", "cndl_multi_context", "cndl_multi_last_error", "cndl_multi_add_object",
    "cndl_multi_commit", "cndl_multi_push_entity", "cndl_multi_buffer_entities", "cndl_multi_frame_submit", "cndl_multi_frame_wait",
    "cndl_multi_trace_frame", "cndl_multi_frame_rays_traced", "cndl_multi_last_replicate_ms", "cndl_clone_scene", "cndl_add_prebuilt_object_device",
    "cndl_object_device_view", "cndl_multi_set_transport", "cndl_object_count", "cndl_object_ids",
    "cndl_ipc_alloc", "cndl_ipc_open", "cndl_ipc_close", "cndl_ipc_free",
    "cndl_generate_texture_references", "cndl_set_texture_references", "cndl_texture_reference_count", "cndl_get_data_material", "cndl_get_data_material_device",
    "cndl_model_mesh_albedo_path", "cndl_model_mesh_normal_path", "cndl_model_mesh_color",
]


class CandelaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


GEN_DIFFUSE, GEN_SPECULAR, GEN_SHADOW = 0, 1, 2
GEN_BUCKET_OCTANTS = 1


class RaygenParams(C.Structure):
    """cndl_raygen_params (include/candela_b200.h)."""
    _fields_ = [("kind", C.c_int32), ("spp", C.c_int32), ("seed", C.c_uint32), ("flags", C.c_uint32), ("offset", C.c_float), ("tmax", C.c_float),
                ("roughness", C.c_float), ("light_dir", C.c_float * 3), ("light_cone", C.c_float), ("d_ids_in", C.c_void_p), ("d_ids_out", C.c_void_p)]


FRAME_OUT_HIT32, FRAME_OUT_HIT16, FRAME_OUT_PIXEL32 = 0, 1, 2
FRAME_OCTANT_ORDER, FRAME_LOCAL_LAYOUT, FRAME_COMPACT_RAYS = 1, 2, 4
TRANSPORT_PEER_STORES, TRANSPORT_STAGED_COPY = 0, 1
HIT16_DT = np.dtype([("t", "<f4"), ("tri", "<i4"), ("v", "<f4"), ("w", "<f4")])
PIXEL_DT = np.dtype([("t", "<f4"), ("tri", "<i4"), ("v", "<f4"), ("w", "<f4"), ("ao", "<f4"), ("t_mean", "<f4"), ("rays", "<i4"), ("escaped", "<i4")])
FRAME_RECORD_DT = {FRAME_OUT_HIT32: None, FRAME_OUT_HIT16: HIT16_DT, FRAME_OUT_PIXEL32: PIXEL_DT}   # HIT32 -> HIT_DT (defined above)


class FrameParams(C.Structure):
    """cndl_frame_params (include/candela_b200.h)."""
    _fields_ = [("inv_view", C.c_float * 16), ("inv_proj", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32),
                ("bounces", C.c_int32), ("seed", C.c_uint32), ("tile", C.c_int32), ("shard_index", C.c_int32), ("shard_count", C.c_int32),
                ("out_format", C.c_int32), ("flags", C.c_uint32)]


def frame_params(inv_view, inv_proj, width: int, height: int, spp: int = 1, bounces: int = 1, seed: int = 1, tile: int = 0, shard_index: int = 0,
                 shard_count: int = 1, out_format: int = FRAME_OUT_HIT16, octant_order: bool = False, local_layout: bool = False,
                 compact_rays: bool = False) -> FrameParams:
    """inv_view / inv_proj: 4x4 (row, column) matrices, as IntersectPrimary takes them.  compact_rays: the three-pass compacting
    generator instead of the one-pass segmented one (same results)."""
    iv, ip = _colmajor(inv_view), _colmajor(inv_proj)
    flags = (FRAME_OCTANT_ORDER if octant_order else 0) | (FRAME_LOCAL_LAYOUT if local_layout else 0) | (FRAME_COMPACT_RAYS if compact_rays else 0)
    return FrameParams((C.c_float * 16)(*iv), (C.c_float * 16)(*ip), width, height, spp, bounces, seed, tile, shard_index, shard_count, out_format, flags)


def frame_record_dtype(out_format: int):
    return HIT_DT if out_format == FRAME_OUT_HIT32 else FRAME_RECORD_DT[out_format]


class BuildOpts(C.Structure):
    _fields_ = [("builder", C.c_int32), ("swap_policy", C.c_int32), ("swap_seed", C.c_uint64)]


_lib = None


def library_path() -> Path:
    return _build.LIB


def load_library() -> C.CDLL:
    """Loads libcandela_b200.so. Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not _build.LIB.exists():
        raise ImportError(f"{_build.LIB} is missing: run `python -m candela_b200._build` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(str(_build.LIB))
    vp, sz, i32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_int32)
    L.cndl_abi_version.restype = C.c_int
    L.cndl_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    L.cndl_destroy.argtypes = [vp]
    L.cndl_destroy.restype = None
    L.cndl_last_error.argtypes = [vp]
    L.cndl_last_error.restype = C.c_char_p
    L.cndl_add_object.argtypes = [vp, C.c_uint32, vp, sz, vp, sz, vp, C.POINTER(BuildOpts)]
    L.cndl_build_bvh.argtypes = [C.c_int, C.c_int, vp, sz, vp, sz, vp, C.c_int32, C.POINTER(BuildOpts), vp, sz, C.POINTER(sz), vp, C.POINTER(C.c_float)]
    L.cndl_build_bvh.restype = C.c_int
    L.cndl_add_prebuilt_object.argtypes = [vp, C.c_uint32, vp, sz, vp, sz, vp, sz]
    for f in ("cndl_node_count", "cndl_triangle_count", "cndl_vertex_count", "cndl_entity_count"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = sz
    L.cndl_get_object.argtypes = [vp, C.c_uint32, i32p, i32p, i32p, i32p]
    L.cndl_commit.argtypes = [vp, C.c_int]
    L.cndl_read_buffers.argtypes = [vp, vp, vp, vp]
    L.cndl_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.cndl_push_entity.argtypes = [vp, C.c_uint32, vp, C.c_float, C.c_float]
    L.cndl_push_entity_records.argtypes = [vp, vp, sz]
    L.cndl_buffer_entities.argtypes = [vp]
    L.cndl_intersect_closest.argtypes = [vp, vp, sz, C.c_int, vp]
    L.cndl_intersect_any.argtypes = [vp, vp, sz, vp]
    L.cndl_intersect_closest_device.argtypes = [vp, vp, sz, C.c_int, vp, vp]
    L.cndl_intersect_any_device.argtypes = [vp, vp, sz, vp, vp]
    L.cndl_intersect_primary.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.cndl_intersect_primary_device.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.cndl_generate_bounce_rays_device.argtypes = [vp, vp, vp, sz, C.c_int, C.c_float, C.c_float, C.c_uint32, vp, vp, C.POINTER(sz), vp]
    L.cndl_get_data.argtypes = [vp, vp, sz, vp]
    L.cndl_get_data_device.argtypes = [vp, vp, sz, vp, vp]
    L.cndl_ipc_alloc.argtypes = [vp, sz, C.POINTER(vp), vp]
    L.cndl_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.cndl_ipc_close.argtypes = [vp, vp]
    L.cndl_ipc_free.argtypes = [vp, vp]
    L.cndl_get_data_material.argtypes = [vp, vp, sz, vp]
    L.cndl_get_data_material_device.argtypes = [vp, vp, sz, vp, vp]
    L.cndl_set_texture_references.argtypes = [vp, vp, sz]
    L.cndl_texture_reference_count.argtypes = [vp]
    L.cndl_texture_reference_count.restype = sz
    L.cndl_generate_texture_references.argtypes = [vp, sz, vp, vp, sz, C.POINTER(sz)]
    L.cndl_generate_rays_device.argtypes = [vp, C.POINTER(RaygenParams), vp, vp, sz, vp, vp, C.POINTER(sz), vp]
    L.cndl_generate_probe_rays_device.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, vp]
    fp = C.POINTER(FrameParams)
    for f in ("cndl_frame_records", "cndl_frame_shard_records"):
        getattr(L, f).argtypes = [fp]
        getattr(L, f).restype = sz
    L.cndl_frame_record_bytes.argtypes = [C.c_int]
    L.cndl_frame_record_bytes.restype = sz
    L.cndl_trace_frame_device.argtypes = [vp, fp, vp, C.c_int, vp]
    L.cndl_frame_submit.argtypes = [vp, fp, vp, C.c_int]
    L.cndl_frame_wait.argtypes = [vp, C.c_int]
    L.cndl_trace_frame.argtypes = [vp, fp, vp]
    L.cndl_frame_rays_traced.argtypes = [vp, C.c_int]
    L.cndl_frame_rays_traced.restype = C.c_uint64
    L.cndl_frame_untile_device.argtypes = [vp, fp, vp, vp, vp]
    L.cndl_multi_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int), C.c_int]
    L.cndl_multi_destroy.argtypes = [vp]
    L.cndl_multi_destroy.restype = None
    L.cndl_multi_device_count.argtypes = [vp]
    L.cndl_multi_context.argtypes = [vp, C.c_int]
    L.cndl_multi_context.restype = vp
    L.cndl_multi_last_error.argtypes = [vp]
    L.cndl_multi_last_error.restype = C.c_char_p
    L.cndl_multi_add_object.argtypes = [vp, C.c_uint32, vp, sz, vp, sz, vp, C.POINTER(BuildOpts)]
    L.cndl_multi_commit.argtypes = [vp]
    L.cndl_multi_push_entity.argtypes = [vp, C.c_uint32, vp, C.c_float, C.c_float]
    L.cndl_multi_buffer_entities.argtypes = [vp]
    L.cndl_multi_frame_submit.argtypes = [vp, fp, vp, C.c_int]
    L.cndl_multi_frame_wait.argtypes = [vp, C.c_int]
    L.cndl_multi_trace_frame.argtypes = [vp, fp, vp]
    L.cndl_multi_frame_rays_traced.argtypes = [vp, C.c_int]
    L.cndl_multi_frame_rays_traced.restype = C.c_uint64
    L.cndl_multi_last_replicate_ms.argtypes = [vp]
    L.cndl_multi_last_replicate_ms.restype = C.c_float
    L.cndl_multi_set_transport.argtypes = [vp, C.c_int]
    L.cndl_object_count.argtypes = [vp]
    L.cndl_object_count.restype = sz
    L.cndl_object_ids.argtypes = [vp, C.POINTER(C.c_uint32), sz]
    L.cndl_clone_scene.argtypes = [vp, vp]
    L.cndl_add_prebuilt_object_device.argtypes = [vp, C.c_uint32, vp, sz, vp, sz, vp, sz, C.c_int32, C.c_int32]
    L.cndl_object_device_view.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]
    L.cndl_collide_boxes.argtypes = [vp, vp, sz, vp]
    L.cndl_collide_boxes_device.argtypes = [vp, vp, sz, vp, vp]
    for f in ("cndl_model_load_obj", "cndl_model_load_gltf", "cndl_model_load"):
        getattr(L, f).argtypes = [C.c_char_p, C.c_int32, C.POINTER(vp), C.c_char_p, sz]
    L.cndl_model_free.argtypes = [vp]
    L.cndl_model_free.restype = None
    for f in ("cndl_model_vertex_count", "cndl_model_index_count", "cndl_model_mesh_count"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = sz
    for f in ("cndl_model_vertices", "cndl_model_indices", "cndl_model_mesh_ids"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = vp
    L.cndl_model_mesh_name.argtypes = [vp, sz]
    L.cndl_model_mesh_name.restype = C.c_char_p
    for f in ("cndl_model_mesh_albedo_path", "cndl_model_mesh_normal_path"):
        getattr(L, f).argtypes = [vp, sz]
        getattr(L, f).restype = C.c_char_p
    L.cndl_model_mesh_color.argtypes = [vp, sz, vp]
    L.cndl_add_model.argtypes = [vp, C.c_uint32, vp, C.POINTER(BuildOpts)]
    L.cndl_pack_half2x16.argtypes = [C.c_float, C.c_float]
    L.cndl_pack_half2x16.restype = C.c_uint32
    L.cndl_save.argtypes = [vp, C.c_char_p]
    L.cndl_load.argtypes = [vp, C.c_char_p]
    L.cndl_host_alloc.argtypes = [sz]
    L.cndl_host_alloc.restype = vp
    L.cndl_host_alloc_write_combined.argtypes = [sz]
    L.cndl_host_alloc_write_combined.restype = vp
    L.cndl_host_free.argtypes = [vp]
    L.cndl_host_free.restype = None
    L.cndl_set_traversal_mode.argtypes = [vp, C.c_int, C.c_int]
    L.cndl_set_tuning.argtypes = [vp, C.c_int, C.c_int]
    L.cndl_launch_count.argtypes = [vp]
    L.cndl_launch_count.restype = C.c_uint64
    L.cndl_last_build_ms.argtypes = [vp]
    L.cndl_last_build_ms.restype = C.c_float
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _colmajor(m) -> np.ndarray:
    """4x4 (row, column) matrix -> 16 floats column-major, glm's storage."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(4, 4).T).ravel()


def make_vertices(positions) -> np.ndarray:
    """32-byte Vertex records (Utils/Vertex.h:7-12) with w = 1 and zero packed attributes."""
    positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    v = np.zeros(len(positions), dtype=VERTEX_DT)
    v["position"][:, :3] = positions
    v["position"][:, 3] = 1.0
    return v


def make_rays(origins, directions, tmax=0.0) -> np.ndarray:
    o = np.asarray(origins, dtype=np.float32).reshape(-1, 3)
    r = np.zeros(len(o), dtype=RAY_DT)
    r["o"] = o
    r["d"] = np.asarray(directions, dtype=np.float32).reshape(-1, 3)
    r["tmax"] = tmax
    return r


def BuildBVH(node_format: int, verts, indices, mesh_ids=None, t_offset: int = 0, device: int = 0, builder: int = BUILDER_SAH_EXACT,
             swap_policy: int = SWAP_NONE, swap_seed: int = 0):
    """BVH::BuildBVH (BVHConstructor.cpp:951-1108) as a stand-alone call on the GPU: returns (nodes, triangles, build_ms) as the
    reference leaves them in FlattenedNodes / FlattenedTris (leaf packs include t_offset, vertex indices object-local)."""
    L = load_library()
    verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
    indices = np.ascontiguousarray(indices, dtype=np.uint32).ravel()
    if mesh_ids is not None:
        mesh_ids = np.ascontiguousarray(mesh_ids, dtype=np.int32)
    T = len(indices) // 3
    nodes = np.zeros(max(2 * T - 1, 1), dtype=NODE_DT if node_format == STACKLESS else STACK_NODE_DT)
    tris = np.zeros(T, dtype=TRIANGLE_DT)
    n, ms = C.c_size_t(0), C.c_float(0.0)
    opts = BuildOpts(builder, swap_policy, swap_seed)
    rc = L.cndl_build_bvh(node_format, device, _p(verts), len(verts), _p(indices), len(indices), _p(mesh_ids), t_offset, C.byref(opts), _p(nodes),
                          len(nodes), C.byref(n), _p(tris), C.byref(ms))
    if rc != 0:
        raise CandelaError(rc, "cndl_build_bvh rejected the geometry or found no usable device")
    return nodes[: n.value].copy(), tris, ms.value


def load_model(path, first_mesh_number: int = 0, materials: bool = False):
    """cndl_model_load (.obj / .gltf / .glb) -> (vertices[VERTEX_DT], indices[u32], mesh_ids[i32 per triangle], mesh names): what
    ModelFileLoader.cpp:101-185 hands to the intersector, without Assimp.  Host only.  materials=True appends the per-mesh
    _MeshMaterialData list (ModelFileLoader.h:21-25) as dicts {albedo, normal, color}."""
    L = load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = L.cndl_model_load(str(path).encode(), first_mesh_number, C.byref(h), err, len(err))
    if rc != 0:
        raise CandelaError(rc, err.value.decode())
    try:
        nv, ni, nm = L.cndl_model_vertex_count(h), L.cndl_model_index_count(h), L.cndl_model_mesh_count(h)
        verts = np.frombuffer((C.c_char * (nv * 32)).from_address(L.cndl_model_vertices(h)), dtype=VERTEX_DT).copy()
        idx = np.frombuffer((C.c_char * (ni * 4)).from_address(L.cndl_model_indices(h)), dtype=np.uint32).copy()
        mids = np.frombuffer((C.c_char * (ni // 3 * 4)).from_address(L.cndl_model_mesh_ids(h)), dtype=np.int32).copy()
        names = [L.cndl_model_mesh_name(h, k).decode() for k in range(nm)]
        mats = []
        for k in range(nm):
            rgb = (C.c_float * 3)()
            L.cndl_model_mesh_color(h, k, rgb)
            mats.append({"albedo": L.cndl_model_mesh_albedo_path(h, k).decode(), "normal": L.cndl_model_mesh_normal_path(h, k).decode(),
                         "color": (rgb[0], rgb[1], rgb[2])})
    finally:
        L.cndl_model_free(h)
    return (verts, idx, mids, names, mats) if materials else (verts, idx, mids, names)


def generate_texture_references(materials):
    """cndl_generate_texture_references = RayIntersector::GenerateMeshTextureReferences (Intersector.h:367-402) as host arithmetic.
    materials: MESH_MATERIAL_DT records (or (albedo_handle, albedo_valid, normal_handle, normal_valid, (r, g, b)) tuples).
    Returns (table[TEXREF_DT], handles[u64]) with handles[i] = the handle to bind to Textures[i]."""
    if not (isinstance(materials, np.ndarray) and materials.dtype == MESH_MATERIAL_DT):
        rec = np.zeros(len(materials), dtype=MESH_MATERIAL_DT)
        for i, (a, va, b, vb, color) in enumerate(materials):
            rec[i] = (a, b, int(bool(va)), int(bool(vb)), tuple(color), 0.0)
        materials = rec
    materials = np.ascontiguousarray(materials)
    L = load_library()
    out = np.zeros(len(materials), dtype=TEXREF_DT)
    handles = np.zeros(2 * len(materials), dtype=np.uint64)
    n = C.c_size_t(0)
    rc = L.cndl_generate_texture_references(_p(materials), len(materials), _p(out), _p(handles), len(handles), C.byref(n))
    if rc != 0:
        raise CandelaError(rc, "cndl_generate_texture_references")
    return out, handles[: n.value].copy()


load_obj = load_model


class PinnedBuffer:
    """cndl_host_alloc()'d memory viewed as a numpy array (for ray / hit batches)."""

    def __init__(self, count: int, dtype, write_combined: bool = False):
        self._lib = load_library()
        self.dtype = np.dtype(dtype)
        self.nbytes = max(1, count * self.dtype.itemsize)
        self.ptr = (self._lib.cndl_host_alloc_write_combined if write_combined else self._lib.cndl_host_alloc)(self.nbytes)
        if not self.ptr:
            raise MemoryError("cndl_host_alloc failed")
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=count)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.cndl_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class RayIntersector:
    """``Candela::RayIntersector<T>`` with ``T`` given by ``node_format``."""

    def __init__(self, node_format: int = STACKLESS, device: int = 0):
        self._lib = load_library()
        if node_format not in (STACKLESS, STACK):
            # Intersector.h:146-148
            raise CandelaError(-1, "Template <T> Passed to RayIntersector can only be of type BVH::FlattenedStackNode or BVH::FlattenedNode>!")
        h = C.c_void_p()
        rc = self._lib.cndl_create(C.byref(h), node_format, device)
        if rc != 0:
            raise CandelaError(rc, "cndl_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")
        self._h = h
        self.node_format = node_format
        self.node_dtype = NODE_DT if node_format == STACKLESS else STACK_NODE_DT

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):   # a MultiRayIntersector owns its per-device contexts
                self._lib.cndl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise CandelaError(rc, self._lib.cndl_last_error(self._h).decode())

    def Initialize(self):
        """Intersector.h:154 loads the trace shader; kernels here are precompiled — nothing to do."""

    # -- scene ----------------------------------------------------------------------------------
    def AddObject(self, object_id: int, verts, indices, mesh_ids=None, builder: int = BUILDER_SAH_EXACT,
                  swap_policy: int = SWAP_NONE, swap_seed: int = 0):
        """Intersector.h:170-198 with the BVH built on the GPU. `indices`: object-local, per-mesh offset applied."""
        verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
        indices = np.ascontiguousarray(indices, dtype=np.uint32).ravel()
        if mesh_ids is not None:
            mesh_ids = np.ascontiguousarray(mesh_ids, dtype=np.int32)
        opts = BuildOpts(builder, swap_policy, swap_seed)
        self._check(self._lib.cndl_add_object(self._h, object_id, _p(verts), len(verts), _p(indices), len(indices), _p(mesh_ids), C.byref(opts)))

    def AddPrebuiltObject(self, object_id: int, nodes, tris, verts):
        """AddObject for buffers the engine built with its own BVH::BuildBVH (object-local vertex indices)."""
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == self.node_dtype.itemsize, "node records do not match the context's format"
        tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DT)
        verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
        self._check(self._lib.cndl_add_prebuilt_object(self._h, object_id, _p(nodes), len(nodes), _p(tris), len(tris), _p(verts), len(verts)))

    def Save(self, path):
        """Flat-buffer cache of every object's reference-layout buffers (cndl_save)."""
        self._check(self._lib.cndl_save(self._h, str(path).encode()))

    def Load(self, path):
        """Restores a cndl_save file into this (empty) intersector; call BufferData() afterwards."""
        self._check(self._lib.cndl_load(self._h, str(path).encode()))

    def BufferData(self, ClearCPUData: bool = True):
        """Intersector.h:322-351."""
        self._check(self._lib.cndl_commit(self._h, int(ClearCPUData)))

    def PushEntity(self, object_id: int, model=None, emissive: float = 0.0, translucency: float = 0.0):
        """Intersector.h:201-216. `model` is a 4x4 (row, column) matrix."""
        m = _colmajor(np.eye(4, dtype=np.float32) if model is None else model)
        self._check(self._lib.cndl_push_entity(self._h, object_id, _p(m), emissive, translucency))

    def PushEntities(self, entities):
        """Intersector.h:219-224. entities: iterable of (object_id, model, emissive, translucency)."""
        for e in entities:
            self.PushEntity(*e)

    def PushEntityRecords(self, records):
        records = np.ascontiguousarray(records, dtype=ENTITY_DT)
        self._check(self._lib.cndl_push_entity_records(self._h, _p(records), len(records)))

    def BufferEntities(self):
        """Intersector.h:227-239."""
        self._check(self._lib.cndl_buffer_entities(self._h))

    def object_ids(self):
        """Ids of the objects added so far."""
        n = self._lib.cndl_object_count(self._h)
        ids = (C.c_uint32 * max(n, 1))()
        self._lib.cndl_object_ids(self._h, ids, n)
        return [int(ids[k]) for k in range(n)]

    def object_data(self, object_id: int) -> dict:
        v = [C.c_int32() for _ in range(4)]
        self._check(self._lib.cndl_get_object(self._h, object_id, *[C.byref(x) for x in v]))
        return dict(node_offset=v[0].value, node_count=v[1].value, tri_offset=v[2].value, vert_offset=v[3].value)

    @property
    def node_count(self) -> int:
        return int(self._lib.cndl_node_count(self._h))

    @property
    def triangle_count(self) -> int:
        return int(self._lib.cndl_triangle_count(self._h))

    @property
    def vertex_count(self) -> int:
        return int(self._lib.cndl_vertex_count(self._h))

    @property
    def entity_count(self) -> int:
        return int(self._lib.cndl_entity_count(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.cndl_launch_count(self._h))

    @property
    def last_build_ms(self) -> float:
        return float(self._lib.cndl_last_build_ms(self._h))

    def read_buffers(self):
        """m_BVHNodes, m_BVHTriangles, m_BVHVertices (Intersector.h:96-98) copied back from the device."""
        nodes = np.zeros(self.node_count, dtype=self.node_dtype)
        tris = np.zeros(self.triangle_count, dtype=TRIANGLE_DT)
        verts = np.zeros(self.vertex_count, dtype=VERTEX_DT)
        self._check(self._lib.cndl_read_buffers(self._h, _p(nodes), _p(tris), _p(verts)))
        return nodes, tris, verts

    def set_traversal_mode(self, mode: int, sort_rays: int = 4):
        """sort_rays: 0 off, 1 octant buckets, 2 octant + origin Morton order (rays moved), 3 the same through an index list,
        4 automatic (3 for large batches on scenes beyond the L2, else off; the library's default)."""
        self._check(self._lib.cndl_set_traversal_mode(self._h, mode, int(sort_rays)))

    def set_tuning(self, knob: int, value: int):
        self._check(self._lib.cndl_set_tuning(self._h, knob, value))

    # -- queries: host buffers --------------------------------------------------------------------
    def IntersectRays(self, rays, ignore_transparent: bool = False, out=None) -> np.ndarray:
        """IntersectRay / IntersectRayIgnoreTransparent for a host batch -> hit records."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DT)
        hits = np.zeros(len(rays), dtype=HIT_DT) if out is None else out
        self._check(self._lib.cndl_intersect_closest(self._h, _p(rays), len(rays), IGNORE_TRANSPARENT if ignore_transparent else 0, _p(hits)))
        return hits

    def IntersectRaysAny(self, rays, out=None) -> np.ndarray:
        """float IntersectRay(o, d) for a host batch -> first accepted t or -1."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DT)
        t = np.zeros(len(rays), dtype=np.float32) if out is None else out
        self._check(self._lib.cndl_intersect_any(self._h, _p(rays), len(rays), _p(t)))
        return t

    def IntersectPrimary(self, inv_view, inv_proj, Width: int, Height: int, return_rays: bool = False):
        """Intersector.h:241-266; matrices are 4x4 (row, column)."""
        iv, ip = _colmajor(inv_view), _colmajor(inv_proj)
        hits = np.zeros(Width * Height, dtype=HIT_DT)
        rays = np.zeros(Width * Height, dtype=RAY_DT) if return_rays else None
        self._check(self._lib.cndl_intersect_primary(self._h, _p(iv), _p(ip), Width, Height, _p(hits), _p(rays)))
        return (hits, rays) if return_rays else hits

    def GetData(self, hits) -> np.ndarray:
        """GetData (…/Include/TraverseBVHStackless.glsl:370-408) without the texture fetch, for a host batch of hit records."""
        hits = np.ascontiguousarray(hits, dtype=HIT_DT)
        out = np.zeros(len(hits), dtype=ATTR_DT)
        self._check(self._lib.cndl_get_data(self._h, _p(hits), len(hits), _p(out)))
        return out

    def ipc_alloc(self, nbytes: int):
        """cndl_ipc_alloc -> (device pointer, 64 handle bytes): memory other processes of the node can map (ipc_open) and store into."""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._check(self._lib.cndl_ipc_alloc(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._lib.cndl_ipc_open(self._h, buf, C.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, ptr: int):
        self._check(self._lib.cndl_ipc_close(self._h, ptr))

    def ipc_free(self, ptr: int):
        self._check(self._lib.cndl_ipc_free(self._h, ptr))

    def SetTextureReferences(self, refs):
        """Uploads the BVHTextureReferences table (m_BVHTextureReferencesSSBO, Intersector.h:404-409)."""
        refs = np.ascontiguousarray(refs, dtype=TEXREF_DT)
        self._check(self._lib.cndl_set_texture_references(self._h, _p(refs), len(refs)))

    def GenerateMeshTextureReferences(self, materials):
        """Intersector.h:367-410: builds the table from the per-mesh materials (see generate_texture_references) and uploads it.
        Returns the handles to bind to Textures[0..] (m_TextureHandleReferenceMap)."""
        refs, handles = generate_texture_references(materials)
        self.SetTextureReferences(refs)
        return handles

    def texture_reference_count(self) -> int:
        return int(self._lib.cndl_texture_reference_count(self._h))

    def GetDataMaterial(self, hits) -> np.ndarray:
        """GetData with the Albedo decision (…Stackless.glsl:393-404): MATERIAL_DT records; albedo_ref > -1 = sample Textures[albedo_ref] at uv."""
        hits = np.ascontiguousarray(hits, dtype=HIT_DT)
        out = np.zeros(len(hits), dtype=MATERIAL_DT)
        self._check(self._lib.cndl_get_data_material(self._h, _p(hits), len(hits), _p(out)))
        return out

    def get_data_material_device(self, d_hits: int, n: int, d_out: int, stream: int = 0):
        self._check(self._lib.cndl_get_data_material_device(self._h, d_hits, n, d_out, stream or None))

    def CollideBoxes(self, mins, maxs) -> np.ndarray:
        """Physics::CollideBox (Physics.cpp:203-228) for a batch of boxes -> records {collided, mesh, tri, entity}."""
        mins = np.asarray(mins, dtype=np.float32).reshape(-1, 3)
        boxes = np.zeros(len(mins), dtype=BOX_DT)
        boxes["min"], boxes["max"] = mins, np.asarray(maxs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(boxes), dtype=COLLISION_DT)
        self._check(self._lib.cndl_collide_boxes(self._h, _p(boxes), len(boxes), _p(out)))
        return out

    def CollideBox(self, Min, Max) -> bool:
        return bool(self.CollideBoxes([Min], [Max])["collided"][0])

    def CollidePoint(self, Point) -> bool:
        """Physics::CollidePoint (Physics.cpp:175-201): the box Point -+ 0.01."""
        p = np.asarray(Point, dtype=np.float32)
        return self.CollideBox(p - np.float32(0.01), p + np.float32(0.01))

    # -- queries: device buffers (raw pointers, e.g. torch tensors' data_ptr()) --------------------
    def intersect_closest_device(self, d_rays: int, n_rays: int, d_hits: int, flags: int = 0, stream: int = 0):
        self._check(self._lib.cndl_intersect_closest_device(self._h, d_rays, n_rays, flags, d_hits, stream or None))

    def intersect_any_device(self, d_rays: int, n_rays: int, d_t: int, stream: int = 0):
        self._check(self._lib.cndl_intersect_any_device(self._h, d_rays, n_rays, d_t, stream or None))

    def generate_bounce_rays_device(self, d_rays: int, d_hits: int, n_rays: int, d_rays_out: int, spp: int = 1, offset: float = 0.05,
                                    tmax: float = 1.0e6, seed: int = 1, d_parent_out: int = 0, stream: int = 0) -> int:
        """Wavefront compaction between bounces; returns the number of rays written."""
        n = C.c_size_t(0)
        self._check(self._lib.cndl_generate_bounce_rays_device(self._h, d_rays, d_hits, n_rays, spp, offset, tmax, seed, d_rays_out,
                                                               d_parent_out or None, C.byref(n), stream or None))
        return int(n.value)

    def generate_rays_device(self, kind: int, d_rays: int, d_hits: int, n_rays: int, d_rays_out: int, spp: int = 1, offset: float = 0.05,
                             tmax: float = 1.0e6, seed: int = 1, roughness: float = 0.0, light_dir=(0.0, 1.0, 0.0), light_cone: float = 0.0,
                             bucket_octants: bool = False, d_parent_out: int = 0, stream: int = 0, d_ids_in: int = 0, d_ids_out: int = 0) -> int:
        """Diffuse / specular / shadow rays from the hits of the previous batch (cndl_generate_rays_device); returns the count."""
        p = RaygenParams(kind, spp, seed, GEN_BUCKET_OCTANTS if bucket_octants else 0, offset, tmax, roughness, (C.c_float * 3)(*light_dir), light_cone,
                         d_ids_in or None, d_ids_out or None)
        n = C.c_size_t(0)
        self._check(self._lib.cndl_generate_rays_device(self._h, C.byref(p), d_rays, d_hits, n_rays, d_rays_out, d_parent_out or None, C.byref(n),
                                                        stream or None))
        return int(n.value)

    def generate_probe_rays_device(self, box_origin, size, res, seed: int, d_rays_out: int, stream: int = 0):
        """Probe-update rays (UpdateRadianceProbes.glsl:408-427) for a res[0] x res[1] x res[2] probe grid."""
        o = np.ascontiguousarray(box_origin, dtype=np.float32)
        sz = np.ascontiguousarray(size, dtype=np.float32)
        r = np.ascontiguousarray(res, dtype=np.int32)
        self._check(self._lib.cndl_generate_probe_rays_device(self._h, _p(o), _p(sz), _p(r), seed, d_rays_out, stream or None))

    def intersect_primary_device(self, inv_view, inv_proj, Width: int, Height: int, d_hits: int, d_rays: int = 0, stream: int = 0):
        iv, ip = _colmajor(inv_view), _colmajor(inv_proj)
        self._check(self._lib.cndl_intersect_primary_device(self._h, _p(iv), _p(ip), Width, Height, d_hits, d_rays or None, stream or None))

    # -- frame-level calls (cndl_trace_frame and friends) --------------------------------------------------------------
    def TraceFrame(self, params: "FrameParams", out=None) -> np.ndarray:
        """One diffuse-GI frame on the device (camera rays -> hits -> diffuse rays -> hits -> records); returns the records of the
        layout / format `params` selects.  `out` (optional): a numpy array over pinned memory (PinnedBuffer.array)."""
        n = self.frame_records(params)
        if out is None:
            out = np.zeros(n, dtype=frame_record_dtype(params.out_format))
        assert out.nbytes >= n * out.dtype.itemsize
        self._check(self._lib.cndl_trace_frame(self._h, C.byref(params), _p(out)))
        return out

    def frame_records(self, params: "FrameParams") -> int:
        f = self._lib.cndl_frame_shard_records if (params.flags & FRAME_LOCAL_LAYOUT) else self._lib.cndl_frame_records
        return int(f(C.byref(params)))

    def frame_submit(self, params: "FrameParams", out: np.ndarray, slot: int = 0):
        self._check(self._lib.cndl_frame_submit(self._h, C.byref(params), _p(out), slot))

    def frame_wait(self, slot: int = 0):
        self._check(self._lib.cndl_frame_wait(self._h, slot))

    def frame_rays_traced(self, slot: int = 0) -> int:
        return int(self._lib.cndl_frame_rays_traced(self._h, slot))

    def trace_frame_device(self, params: "FrameParams", d_out: int, slot: int = 0, stream: int = 0):
        self._check(self._lib.cndl_trace_frame_device(self._h, C.byref(params), d_out, slot, stream or None))

    def frame_untile_device(self, params: "FrameParams", d_shard: int, d_frame: int, stream: int = 0):
        self._check(self._lib.cndl_frame_untile_device(self._h, C.byref(params), d_shard, d_frame, stream or None))

    # -- scene replication ------------------------------------------------------------------------------------------------
    def CloneSceneFrom(self, other: "RayIntersector"):
        """Replicates every object of `other` (any device) into this EMPTY intersector, device to device; call BufferData() afterwards."""
        self._check(self._lib.cndl_clone_scene(self._h, other._h))

    def object_device_view(self, object_id: int) -> dict:
        vs = [C.c_void_p() for _ in range(3)]
        ns = [C.c_size_t() for _ in range(3)]
        self._check(self._lib.cndl_object_device_view(self._h, object_id, C.byref(vs[0]), C.byref(ns[0]), C.byref(vs[1]), C.byref(ns[1]), C.byref(vs[2]), C.byref(ns[2])))
        return dict(d_nodes=vs[0].value, n_nodes=ns[0].value, d_tris=vs[1].value, n_tris=ns[1].value, d_verts=vs[2].value, n_verts=ns[2].value)

    def AddPrebuiltObjectDevice(self, object_id: int, d_nodes: int, n_nodes: int, d_tris: int, n_tris: int, d_verts: int, n_verts: int,
                                vertex_index_base: int = 0, leaf_triangle_offset: int = 0):
        self._check(self._lib.cndl_add_prebuilt_object_device(self._h, object_id, d_nodes, n_nodes, d_tris, n_tris, d_verts, n_verts, vertex_index_base,
                                                              leaf_triangle_offset))

    def get_data_device(self, d_hits: int, n: int, d_out: int, stream: int = 0):
        self._check(self._lib.cndl_get_data_device(self._h, d_hits, n, d_out, stream or None))

    def collide_boxes_device(self, d_boxes: int, n: int, d_out: int, stream: int = 0):
        self._check(self._lib.cndl_collide_boxes_device(self._h, d_boxes, n, d_out, stream or None))


class MultiRayIntersector:
    """Several GPUs behind one handle (cndl_multi_*): the BVH replicated per device (built once, copied device to device), screen
    tiles dealt round-robin, every device's records stored straight into the first device's frame over NVLink peer memory."""

    def __init__(self, node_format: int = STACKLESS, devices=(0,)):
        self._lib = load_library()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._lib.cndl_multi_create(C.byref(h), node_format, devs, len(devices))
        if rc != 0:
            raise CandelaError(rc, "cndl_multi_create failed: a device is missing or not sm_100 (there is no CPU fallback)")
        self._h = h
        self.devices = tuple(devices)
        self.node_format = node_format

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cndl_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise CandelaError(rc, self._lib.cndl_multi_last_error(self._h).decode())

    def AddObject(self, object_id: int, verts, indices, mesh_ids=None, builder: int = BUILDER_SAH_EXACT, swap_policy: int = SWAP_NONE, swap_seed: int = 0):
        verts = np.ascontiguousarray(verts, dtype=VERTEX_DT)
        indices = np.ascontiguousarray(indices, dtype=np.uint32).ravel()
        if mesh_ids is not None:
            mesh_ids = np.ascontiguousarray(mesh_ids, dtype=np.int32)
        opts = BuildOpts(builder, swap_policy, swap_seed)
        self._check(self._lib.cndl_multi_add_object(self._h, object_id, _p(verts), len(verts), _p(indices), len(indices), _p(mesh_ids), C.byref(opts)))

    def BufferData(self, ClearCPUData: bool = True):
        self._check(self._lib.cndl_multi_commit(self._h))

    def PushEntity(self, object_id: int, model=None, emissive: float = 0.0, translucency: float = 0.0):
        m = _colmajor(np.eye(4, dtype=np.float32) if model is None else model)
        self._check(self._lib.cndl_multi_push_entity(self._h, object_id, _p(m), emissive, translucency))

    def BufferEntities(self):
        self._check(self._lib.cndl_multi_buffer_entities(self._h))

    def context(self, i: int) -> "RayIntersector":
        """A borrowed view of device i's context (do not close it)."""
        ri = RayIntersector.__new__(RayIntersector)
        ri._lib = self._lib
        ri._borrowed = True
        ri._h = C.c_void_p(self._lib.cndl_multi_context(self._h, i))
        ri.node_format = self.node_format
        ri.node_dtype = NODE_DT if self.node_format == STACKLESS else STACK_NODE_DT
        return ri

    def TraceFrame(self, params: FrameParams, out=None) -> np.ndarray:
        n = int(self._lib.cndl_frame_records(C.byref(params)))
        if out is None:
            out = np.zeros(n, dtype=frame_record_dtype(params.out_format))
        self._check(self._lib.cndl_multi_trace_frame(self._h, C.byref(params), _p(out)))
        return out

    def frame_submit(self, params: FrameParams, out: np.ndarray, slot: int = 0):
        self._check(self._lib.cndl_multi_frame_submit(self._h, C.byref(params), _p(out), slot))

    def frame_wait(self, slot: int = 0):
        self._check(self._lib.cndl_multi_frame_wait(self._h, slot))

    def frame_rays_traced(self, slot: int = 0) -> int:
        return int(self._lib.cndl_multi_frame_rays_traced(self._h, slot))

    def set_transport(self, transport: int):
        """TRANSPORT_PEER_STORES (default where peer access exists) or TRANSPORT_STAGED_COPY."""
        self._check(self._lib.cndl_multi_set_transport(self._h, transport))

    @property
    def last_replicate_ms(self) -> float:
        return float(self._lib.cndl_multi_last_replicate_ms(self._h))
