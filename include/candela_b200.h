/* candela_b200 — B200-native (sm_100a) BVH build + ray traversal backend.
 *
 * C ABI that replaces the hot path behind Candela's
 *   template<typename T> class Candela::RayIntersector   (Source/Core/BVH/Intersector.h:60-124)
 *   Candela::BVH::BuildBVH                                (Source/Core/BVH/BVHConstructor.h:86-87)
 * and the GLSL traversal it binds to
 *   Source/Core/Shaders/Intersectors/Include/TraverseBVHStackless.glsl
 *   Source/Core/Shaders/Intersectors/Include/TraverseBVHStack.glsl
 *
 * Buffer records keep the reference's layouts byte for byte; every function
 * below cites the reference interface it stands in for.  All functions return
 * CNDL_OK (0) or a negative cndl_status and never throw across the boundary
 * (the reference throws string literals, Intersector.h:147,:204).  There is
 * no CPU fallback: without a usable CUDA device cndl_create fails.
 *
 * Threading: one context per device; calls on one context are serialised by
 * the caller (the reference is single-threaded and its builder keeps file
 * scope statics, BVHConstructor.cpp:58-66).  Pointers passed in are borrowed
 * for the duration of the call only.
 */
#ifndef CANDELA_B200_H
#define CANDELA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNDL_ABI_VERSION 2

typedef enum cndl_status {
    CNDL_OK = 0,
    CNDL_ERR_INVALID = -1,      /* bad argument */
    CNDL_ERR_CUDA = -2,         /* CUDA runtime error; see cndl_last_error */
    CNDL_ERR_NO_DEVICE = -3,    /* no sm_100 device: there is no CPU fallback */
    CNDL_ERR_UNKNOWN_OBJECT = -4, /* entity pushed for an object never added (Intersector.h:203-205) */
    CNDL_ERR_NOT_COMMITTED = -5,  /* traversal before cndl_commit / cndl_buffer_entities */
    CNDL_ERR_OOM = -6
} cndl_status;

/* node_format: the template argument T of RayIntersector<T> (Intersector.h:138-148) */
enum { CNDL_STACKLESS = 0 /* BVH::FlattenedNode */, CNDL_STACK = 1 /* BVH::FlattenedStackNode */ };

/* Candela::Vertex, Source/Core/Utils/Vertex.h:7-12 */
typedef struct cndl_vertex { float position[4]; uint32_t normal_tangent[3]; uint32_t texcoords; } cndl_vertex;
/* BVH::Triangle, BVHConstructor.h:79-84: three global vertex indices + GlobalMeshNumber */
typedef struct cndl_triangle { int32_t v[3]; int32_t mesh; } cndl_triangle;
/* BVH::FlattenedNode, BVHConstructor.h:66-70. min[3] bits: -1 inner, else (first_tri<<4)|count.
 * max[3] bits: miss link (object-local node index, -1 terminates). First child is index+1. */
typedef struct cndl_node { float min[4]; float max[4]; } cndl_node;
/* BVH::FlattenedStackNode, BVHConstructor.h:72-77. Per child: min[3] bits -1 if the child is inner
 * else its leaf pack; max[3] bits = the inner child's slot (object-local). */
typedef struct cndl_stack_node { cndl_node left; cndl_node right; } cndl_stack_node;
/* Candela::BVHEntity, Intersector.h:43-49. Column-major matrices. data[0] = emissive float bits,
 * data[1] = (1 - translucency) float bits. */
typedef struct cndl_entity { float model[16]; float inverse[16]; int32_t node_offset; int32_t node_count; int32_t data[14]; } cndl_entity;

/* Ray and hit records of the new ABI. tmin is reserved (the reference has no
 * per-ray tmin: 0.0001 in the stackless box test, t > 0 for triangles).  tmax is honoured by
 * any-hit queries only (> 0: use it; <= 0: the reference's 1e6); closest-hit always starts at
 * 1e6 like IntersectScene (…Stackless.glsl:284). */
typedef struct cndl_ray { float ox, oy, oz, tmin; float dx, dy, dz, tmax; } cndl_ray;
/* vec4 TUVW + out ints of IntersectScene (…Stackless.glsl:280-319): t, then the barycentric
 * weights of vertices A, B, C; miss = t,u,v,w all -1.  mesh/tri/entity are -1 when nothing was
 * accepted.  A hit on global triangle 0 reports t=u=v=w=-1 with tri = 0 (reference quirk, :300).
 * iters = node iterations of the last entity traversed (the reference's `out int Iters`). */
typedef struct cndl_hit { float t, u, v, w; int32_t mesh, tri, entity, iters; } cndl_hit;

enum { CNDL_BUILDER_SAH_EXACT = 0, /* GPU binned SAH, buffers byte-identical to BVH::BuildBVH */
       CNDL_BUILDER_LBVH = 1       /* GPU Morton/radix-sort LBVH: same layouts, different tree */ };
enum { CNDL_SWAP_NONE = 0, CNDL_SWAP_HASHED = 1 };

/* Build options; zero-initialised == the reference's compile-time constants
 * (BVHConstructor.cpp:41-51: 64 bins, <= 2 triangles per leaf). */
typedef struct cndl_build_opts {
    int32_t builder;      /* CNDL_BUILDER_* */
    int32_t swap_policy;  /* stackless only: child flips of BVHConstructor.cpp:599-609 */
    uint64_t swap_seed;
} cndl_build_opts;

typedef struct cndl_ctx cndl_ctx;

/* RayIntersector<T>::RayIntersector + Initialize (Intersector.h:126-167). */
int cndl_create(cndl_ctx** out, int node_format, int device);
void cndl_destroy(cndl_ctx* ctx);
const char* cndl_last_error(const cndl_ctx* ctx);
int cndl_abi_version(void);

/* RayIntersector<T>::AddObject (Intersector.h:170-198) with the BVH built ON THE GPU.
 * `indices` are object-local and already carry the per-mesh vertex offset of
 * BuildBVH's concatenation (BVHConstructor.cpp:981-1002); one GlobalMeshNumber per triangle. */
int cndl_add_object(cndl_ctx* ctx, uint32_t object_id, const cndl_vertex* verts, size_t V,
                    const uint32_t* indices, size_t I, const int32_t* mesh_id_per_tri,
                    const cndl_build_opts* opts);

/* AddObject for an object whose BVH the engine already built with its own BVH::BuildBVH:
 * `nodes` is FlattenedNode[] or FlattenedStackNode[] per the context format, triangle vertex
 * indices are object-local (they are rebased by the running vertex count exactly like
 * Intersector.h:190-197) and leaf packs must already include the triangle offset
 * (= cndl_triangle_count before the call), as BuildBVH's last argument makes them. */
int cndl_add_prebuilt_object(cndl_ctx* ctx, uint32_t object_id, const void* nodes, size_t N,
                             const cndl_triangle* tris, size_t T, const cndl_vertex* verts, size_t V);

/* Running totals (m_BVHNodes.size() etc.; Intersector.h:179-181). */
size_t cndl_node_count(const cndl_ctx* ctx);
size_t cndl_triangle_count(const cndl_ctx* ctx);
size_t cndl_vertex_count(const cndl_ctx* ctx);
/* _ObjectData of one object (Intersector.h:51-56). Any out pointer may be NULL. */
int cndl_get_object(const cndl_ctx* ctx, uint32_t object_id, int32_t* node_offset, int32_t* node_count,
                    int32_t* triangle_offset, int32_t* vertex_offset);

/* Number of objects added, and their ids in insertion order (ids_out holds `capacity` entries; returns the number written). */
size_t cndl_object_count(const cndl_ctx* ctx);
size_t cndl_object_ids(const cndl_ctx* ctx, uint32_t* ids_out, size_t capacity);

/* BVH::BuildBVH (BVHConstructor.h:86-87, BVHConstructor.cpp:951-1108) as a stand-alone call, for callers that want the
 * flattened buffers themselves: builds on GPU `device` and returns host buffers exactly as the reference's function
 * leaves them in FlattenedNodes / FlattenedTris — leaf packs include tri_offset (its `int t_offset`), triangle records
 * carry OBJECT-LOCAL vertex indices (the rebasing of AddObject, Intersector.h:190-197, is the caller's).  nodes_out has
 * room for nodes_capacity nodes of the format (2 * I / 3 - 1 always suffices); *n_nodes_out = LastNodeIndex + 1.
 * nodes_out / tris_out / build_ms may be NULL. */
int cndl_build_bvh(int node_format, int device, const cndl_vertex* verts, size_t V, const uint32_t* indices, size_t I,
                   const int32_t* mesh_id_per_tri, int32_t tri_offset, const cndl_build_opts* opts, void* nodes_out,
                   size_t nodes_capacity, size_t* n_nodes_out, cndl_triangle* tris_out, float* build_ms);

/* RayIntersector<T>::BufferData(bool ClearCPUData) (Intersector.h:322-351): makes everything
 * added so far visible to traversal and (re)builds the device-side acceleration copies. */
int cndl_commit(cndl_ctx* ctx, int clear_host);

/* Copies the reference-layout buffers (m_BVHNodes / m_BVHTriangles / m_BVHVertices, public in the
 * reference and read by Physics.cpp:100-126) back to the host.  Any pointer may be NULL. */
int cndl_read_buffers(cndl_ctx* ctx, void* nodes, cndl_triangle* tris, cndl_vertex* verts);
/* Device pointers of the same buffers, valid until the next add/commit; for callers that bind
 * them to their own kernels (the analogue of BindEverything's SSBO bindings 16..20, :269-293). */
int cndl_device_buffers(cndl_ctx* ctx, const void** nodes, const cndl_triangle** tris,
                        const cndl_vertex** verts, const cndl_entity** entities);

/* PushEntity (Intersector.h:201-216): stages {Model, inverse(Model), NodeOffset, NodeCount,
 * emissive, 1-translucency}; model is column-major; the inverse follows glm::inverse. */
int cndl_push_entity(cndl_ctx* ctx, uint32_t object_id, const float model[16], float emissive, float translucency);
/* Stages ready-made 192-byte records (what PushEntity would have produced). */
int cndl_push_entity_records(cndl_ctx* ctx, const cndl_entity* records, size_t E);
/* BufferEntities (Intersector.h:227-239): uploads the staged list and clears the staging list. */
int cndl_buffer_entities(cndl_ctx* ctx);
size_t cndl_entity_count(const cndl_ctx* ctx);

/* Query flags */
enum { CNDL_IGNORE_TRANSPARENT = 1 /* IntersectSceneIgnoreTransparent, …Stackless.glsl:321-366 */ };

/* IntersectRay / IntersectRayIgnoreTransparent (closest hit) for a batch of rays in HOST memory.
 * Copies in, traverses, copies out; pinned buffers (cndl_host_alloc) make the copies asynchronous
 * and chunk-pipelined. */
int cndl_intersect_closest(cndl_ctx* ctx, const cndl_ray* rays, size_t R, int flags, cndl_hit* hits);
/* float IntersectRay(o, d) (any hit; …Stackless.glsl:558-579): t_out[i] = first accepted t or -1. */
int cndl_intersect_any(cndl_ctx* ctx, const cndl_ray* rays, size_t R, float* t_out);
/* Same queries on DEVICE buffers (32-byte aligned, like anything cudaMalloc returns), enqueued on `stream` (a cudaStream_t;
 * NULL = default stream) and not synchronised.  Each call takes its own work counter from a ring of 64, so traversal calls in flight on different
 * streams do not interfere; calls that use the context's ordering / generation scratch (cndl_set_traversal_mode sort_rays != 0,
 * cndl_generate_rays_device) must be enqueued on ONE stream at a time (or ordered by events). */
int cndl_intersect_closest_device(cndl_ctx* ctx, const cndl_ray* d_rays, size_t R, int flags, cndl_hit* d_hits, void* stream);
int cndl_intersect_any_device(cndl_ctx* ctx, const cndl_ray* d_rays, size_t R, float* d_t_out, void* stream);

/* RayIntersector<T>::IntersectPrimary (Intersector.h:241-266) with hit records instead of an
 * albedo image: pixel (x,y) -> hits[y*W+x]; ray generation as in
 * Intersectors/TraverseBVHStack.glsl:133-138,:414-431.  Matrices are column-major.
 * hits is a HOST buffer; rays_out (HOST, optional) receives the generated rays. */
int cndl_intersect_primary(cndl_ctx* ctx, const float inv_view[16], const float inv_proj[16], int W, int H,
                           cndl_hit* hits, cndl_ray* rays_out);
int cndl_intersect_primary_device(cndl_ctx* ctx, const float inv_view[16], const float inv_proj[16], int W, int H,
                                  cndl_hit* d_hits, cndl_ray* d_rays_out, void* stream);

/* Ray generators: the step in front of the path.  For every input ray whose hit record has t > 0
 * they write `spp` rays from the hit point P = o + d*t and the geometric normal N of the hit triangle (world space,
 * turned against the incoming ray; where the shaders read N from a G-buffer texture), compacted in input order (rays that
 * missed emit nothing).  kind:
 *   CNDL_GEN_DIFFUSE   origin P + N*offset, direction CosWeightedHemisphere(N, hash2())
 *                      (DiffuseTrace.glsl:445-446 first bounce, offset 0.05; :516-517 later bounces, 0.02; Include/Sampling.glsl:1-12)
 *   CNDL_GEN_SPECULAR  direction StochasticReflectionDirection(Incident, N, roughness * 0.9) (SpecularTrace.glsl:102-135,:513;
 *                      SampleGGXVNDF, Include/Sampling.glsl:63-83); offset < 0 selects the shader's mix(0.05, 0.1, clamp(roughness*1.4)) (:512)
 *   CNDL_GEN_SHADOW    direction normalize(SampleCone(light_dir, hash2(), sqrt(1 - light_cone^2))) (Include/Sampling.glsl:43-61;
 *                      light_cone = sine of the cone's half angle); hits whose normal faces away from the light emit nothing.
 *                      Meant for cndl_intersect_any_device.
 * Arithmetic contract: the rays are bit-identical to the shader functions evaluated in separately rounded IEEE float
 * operations over glm 0.9.8.5, with the two things GLSL leaves to the implementation DEFINED: hash2() is a counter stream —
 * draw n of element e under `seed` is u01(pcg(pcg(seed ^ pcg(e)) + n * 0x9E3779B9)), e = id * spp + sample, id = d_ids_in[i]
 * or i — and sin / cos / acos / pow are the double-precision formulas of csrc/exact_trig.cuh (within 0.5 ulp).  A CPU
 * restatement in the test tree reproduces every ray bit for bit, and is pinned against the shader functions compiled from the reference.
 * flags & CNDL_GEN_BUCKET_OCTANTS: the rays are written octant-major (all rays whose direction signs are ---, then +--, ...),
 * each octant in input order; incoherent batches traverse ~8 % faster in that order.  d_parent_out maps back.
 * d_rays_out must hold R*spp rays.  *count_out = rays written; the call synchronises `stream`. */
enum { CNDL_GEN_DIFFUSE = 0, CNDL_GEN_SPECULAR = 1, CNDL_GEN_SHADOW = 2 };
enum { CNDL_GEN_BUCKET_OCTANTS = 1 };
typedef struct cndl_raygen_params {
    int32_t kind, spp;
    uint32_t seed, flags;
    float offset, tmax, roughness;
    float light_dir[3];
    float light_cone;
    const uint32_t* d_ids_in; /* optional (device): stream id of every input ray; NULL = its index.  Lets a tile-sharded or
                                 multi-bounce pipeline draw the numbers of pixel p whatever slot p's ray occupies */
    uint32_t* d_ids_out;      /* optional (device): receives id * spp + sample of every ray written */
} cndl_raygen_params;
int cndl_generate_rays_device(cndl_ctx* ctx, const cndl_raygen_params* params, const cndl_ray* d_rays, const cndl_hit* d_hits, size_t R,
                              cndl_ray* d_rays_out, uint32_t* d_parent_out, size_t* count_out, void* stream);
/* CNDL_GEN_DIFFUSE with the remaining parameters at their defaults. */
int cndl_generate_bounce_rays_device(cndl_ctx* ctx, const cndl_ray* d_rays, const cndl_hit* d_hits, size_t R, int spp, float offset,
                                     float tmax, uint32_t seed, cndl_ray* d_rays_out, uint32_t* d_parent_out, size_t* count_out, void* stream);
/* Probe-update rays (UpdateRadianceProbes.glsl:408-427; ProbeGI.cpp:215-217 dispatches PROBE_GRID 48 x 24 x 48, Macros.h:22-24):
 * probe (x, y, z) of a res[0] x res[1] x res[2] grid -> d_rays_out[(z * res[1] + y) * res[0] + x] with origin
 * box_origin + (vec3(x, y, z) / res * 2 - 1) * size and direction ImportanceSample() = normalize(LambertBRDF(vec3(hash2(),
 * hash2().x))) (:351-374; the shader's importance branch is switched off), element = the ray's index, tmax 1e6.  Same
 * arithmetic contract as above.  Needs no scene; enqueued on `stream`, not synchronised. */
int cndl_generate_probe_rays_device(cndl_ctx* ctx, const float box_origin[3], const float size[3], const int32_t res[3], uint32_t seed,
                                    cndl_ray* d_rays_out, void* stream);

/* ---- Frame-level calls: one diffuse-GI frame without any ray ever crossing the host bus ---------------------------------
 * In the reference the rays of a frame never exist on the host: DiffuseTrace.glsl:437-446 forms them in the shader from the
 * G-buffer and traces them on the spot.  cndl_trace_frame does the same on the device: camera rays of the frame's pixels
 * (as cndl_intersect_primary) -> closest hit -> CNDL_GEN_DIFFUSE rays (spp per pixel, offset 0.05, stream id = pixel index)
 * -> IntersectRayIgnoreTransparent (:484) -> for bounce b = 1 .. bounces-1: one CNDL_GEN_DIFFUSE ray per surviving path
 * (offset 0.02, seed + b, :516-517) -> IntersectRay (:518) -> one compact record per pixel (sample) written to the output.
 * Only the 168-byte parameter block goes in and the records come out.
 *
 * Sharding: the frame is cut into tile x tile pixel tiles, tile t = ty * tiles_x + tx; a call with
 * (shard_index, shard_count) traces the tiles with t % shard_count == shard_index and writes ONLY their pixels.  The random
 * stream of a pixel depends on its index, not on the shard, so the shards of any shard_count assemble to the same frame, bit
 * for bit.  Output layouts: the whole frame row-major (pixel * spp + sample; default — every shard writes its pixels into the
 * same buffer, which may live on a peer GPU), or CNDL_FRAME_LOCAL_LAYOUT: this shard's slots only, tile-major
 * ((local_tile * tile * tile + y_in_tile * tile + x_in_tile) * spp + sample, local_tile = t / shard_count; slots of edge tiles
 * that fall outside the image hold miss records), for transports that move contiguous shards (cndl_frame_untile_device
 * scatters one back into the row-major frame). */
enum { CNDL_FRAME_OUT_HIT32 = 0,   /* cndl_hit of the diffuse ray of every (pixel, sample); bounces must be 1 */
       CNDL_FRAME_OUT_HIT16 = 1,   /* the same as cndl_hit16 */
       CNDL_FRAME_OUT_PIXEL32 = 2  /* one resolved cndl_pixel per pixel, any number of bounces */ };
enum { CNDL_FRAME_OCTANT_ORDER = 1, /* every generated batch is ordered by direction octant (inside each 1024-ray segment; with
                                      * CNDL_FRAME_COMPACT_RAYS: over the whole batch); results do not depend on it */
       CNDL_FRAME_LOCAL_LAYOUT = 2,
       CNDL_FRAME_COMPACT_RAYS = 4  /* generate through the three-pass compacting generator (cndl_generate_rays_device) instead of the
                                      * one-pass segmented one; results do not depend on it */ };
/* Compact hit: t, tri and the barycentric weights of vertices B and C (cndl_hit.v, .w; u = 1 - v - w in that order of
 * operations); mesh = triangle[tri].mesh.  Scenes with several entities need the 32-byte form (entity is not carried). */
typedef struct cndl_hit16 { float t; int32_t tri; float v, w; } cndl_hit16;
/* Resolved pixel of a multi-bounce frame.  t / tri / v / w: the camera ray's hit (compact form; t = -1: sky).  ao: mean over the
 * pixel's samples of DiffuseTrace.glsl:494's AO term, t1 > 0 ? pow(clamp(t1 / 2.4, 0, 1), 1.23) : 1, t1 = the first diffuse
 * ray's hit distance, summed in sample order (1 when the pixel traced nothing).  t_mean: mean of the t1 that hit (-1 if none).
 * rays: diffuse rays traced for this pixel over all bounces.  escaped: those that reported no hit. */
typedef struct cndl_pixel { float t; int32_t tri; float v, w; float ao, t_mean; int32_t rays, escaped; } cndl_pixel;
#define CNDL_FRAME_SLOTS 4
typedef struct cndl_frame_params {
    float inv_view[16], inv_proj[16]; /* column-major, as cndl_intersect_primary */
    int32_t width, height;
    int32_t spp;                      /* diffuse samples per pixel at the first bounce (>= 1) */
    int32_t bounces;                  /* diffuse bounces per sample (>= 1; the reference's u_SecondaryBounces + 1) */
    uint32_t seed;
    int32_t tile;                     /* tile edge in pixels; 0 = 64 */
    int32_t shard_index, shard_count; /* shard_count 0 or 1: the whole frame */
    int32_t out_format;               /* CNDL_FRAME_OUT_* */
    uint32_t flags;                   /* CNDL_FRAME_* */
} cndl_frame_params;
/* Records in the whole row-major frame / in this shard's local layout; bytes per record of the format. */
size_t cndl_frame_records(const cndl_frame_params* p);
size_t cndl_frame_shard_records(const cndl_frame_params* p);
size_t cndl_frame_record_bytes(int out_format);
/* Enqueues the shard's work on `stream` (not synchronised); d_out is a device pointer (this device's or a peer's with peer
 * access enabled) in the layout the flags select.  `slot` (0 .. CNDL_FRAME_SLOTS-1) names one of four scratch sets, so several frames can be in
 * flight.  cndl_frame_rays_traced (after the stream has been synchronised) = diffuse rays the last frame of that slot traced. */
int cndl_trace_frame_device(cndl_ctx* ctx, const cndl_frame_params* p, void* d_out, int slot, void* stream);
/* The same with a HOST output buffer (pinned memory makes the copy asynchronous): cndl_frame_submit enqueues frame + copy
 * on the context's own stream pair and returns; cndl_frame_wait blocks until that slot's records are in host_out.
 * Each slot has its own stream: with two or three frames submitted, the device->host copy of one frame overlaps the tracing of
 * the next and the kernels of consecutive frames fill each other's tails (1080p, 1 spp: 1.32 -> 1.23 ms per frame).  cndl_trace_frame = submit + wait on slot 0. */
int cndl_frame_submit(cndl_ctx* ctx, const cndl_frame_params* p, void* host_out, int slot);
int cndl_frame_wait(cndl_ctx* ctx, int slot);
int cndl_trace_frame(cndl_ctx* ctx, const cndl_frame_params* p, void* host_out);
uint64_t cndl_frame_rays_traced(const cndl_ctx* ctx, int slot);
/* Scatters one shard in local layout (d_shard, as written with CNDL_FRAME_LOCAL_LAYOUT by the shard p names) into the
 * row-major frame d_frame.  Enqueued on `stream`. */
int cndl_frame_untile_device(cndl_ctx* ctx, const cndl_frame_params* p, const void* d_shard, void* d_frame, void* stream);

/* One process per GPU (torchrun-style hosts): a frame buffer in ONE process's device memory that the other processes of the node
 * map and pass as d_out of cndl_trace_frame_device — every rank's resolve kernel then stores its tiles' records straight into that
 * frame over NVLink peer memory (row-major layout; no gather, no untile).  cndl_ipc_alloc = cudaMalloc + cudaIpcGetMemHandle on the
 * owner; the 64 handle bytes travel by any means (a broadcast); cndl_ipc_open = cudaIpcOpenMemHandle in a peer process (not in the
 * owner's: CUDA refuses that).  The caller orders the ranks (a barrier or a tiny all-reduce after the frame) before rank 0 reads. */
typedef struct cndl_ipc_handle { unsigned char bytes[64]; } cndl_ipc_handle;
int cndl_ipc_alloc(cndl_ctx* ctx, size_t bytes, void** d_ptr, cndl_ipc_handle* handle);
int cndl_ipc_open(cndl_ctx* ctx, const cndl_ipc_handle* handle, void** d_ptr);
int cndl_ipc_close(cndl_ctx* ctx, void* d_ptr);   /* in a peer process */
int cndl_ipc_free(cndl_ctx* ctx, void* d_ptr);    /* in the owner, after every peer has closed */

/* ---- Several GPUs behind one handle: the BVH replicated per device, screen tiles dealt round-robin, hit
 * records gathered for the final frame only.  One process drives all devices; the C++ mirror's RayIntersector holds one of
 * these when it is given more than one device.  cndl_multi_add_object builds on the first device and replicates the
 * reference-layout buffers device to device (cudaMemcpyPeerAsync over NVLink), so the scene is built once.
 * cndl_multi_trace_frame: every device traces its shard and its resolve kernel stores the records straight into the first
 * device's frame buffer through peer memory (NVLink; no index payload, no padding, no collective), then one device->host
 * copy.  Without peer access the shards travel in local layout with cudaMemcpyPeerAsync and are untiled on the first device. */
typedef struct cndl_multi cndl_multi;
int cndl_multi_create(cndl_multi** out, int node_format, const int* devices, int n_devices);
void cndl_multi_destroy(cndl_multi* m);
int cndl_multi_device_count(const cndl_multi* m);
cndl_ctx* cndl_multi_context(cndl_multi* m, int i);   /* borrowed; per-device queries go through the ordinary calls */
const char* cndl_multi_last_error(const cndl_multi* m);
int cndl_multi_add_object(cndl_multi* m, uint32_t object_id, const cndl_vertex* verts, size_t V, const uint32_t* indices, size_t I,
                          const int32_t* mesh_id_per_tri, const cndl_build_opts* opts);
int cndl_multi_commit(cndl_multi* m);
int cndl_multi_push_entity(cndl_multi* m, uint32_t object_id, const float model[16], float emissive, float translucency);
int cndl_multi_buffer_entities(cndl_multi* m);
int cndl_multi_frame_submit(cndl_multi* m, const cndl_frame_params* p, void* host_out, int slot);  /* p's shard fields are ignored */
int cndl_multi_frame_wait(cndl_multi* m, int slot);
int cndl_multi_trace_frame(cndl_multi* m, const cndl_frame_params* p, void* host_out);
uint64_t cndl_multi_frame_rays_traced(const cndl_multi* m, int slot);
/* How the other devices' records reach the first device's frame: stored by their resolve kernels through peer memory (default
 * where peer access can be enabled), or written in local layout, copied with cudaMemcpyPeerAsync and untiled on the first device
 * (the only way without peer access).  Results are identical. */
enum { CNDL_TRANSPORT_PEER_STORES = 0, CNDL_TRANSPORT_STAGED_COPY = 1 };
int cndl_multi_set_transport(cndl_multi* m, int transport);
float cndl_multi_last_replicate_ms(const cndl_multi* m);  /* device-to-device copy time of the last cndl_multi_add_object */
/* Replicates every object of `src` (any device) into the EMPTY context `dst` device to device; call cndl_commit afterwards. */
int cndl_clone_scene(cndl_ctx* dst, cndl_ctx* src);
/* AddObject for reference-layout buffers that already live in device memory (this device's, or a peer's: the copy is a
 * cudaMemcpyDefault) — e.g. another context's slices (cndl_object_device_view) or buffers received over NCCL.  Vertex indices in
 * d_tris are relative to vertex_index_base (0 = object-local as BuildBVH leaves them; a view's indices are relative to minus
 * its vertex_offset, i.e. pass vertex_offset); they are rebased to this context's running vertex count (Intersector.h:190-197).
 * leaf_triangle_offset is the triangle offset the leaf packs embed (BuildBVH's t_offset): it must equal cndl_triangle_count(ctx),
 * else CNDL_ERR_INVALID — so replicating a subset or a different order fails loudly instead of returning wrong triangles. */
int cndl_add_prebuilt_object_device(cndl_ctx* ctx, uint32_t object_id, const void* d_nodes, size_t N, const cndl_triangle* d_tris, size_t T,
                                    const cndl_vertex* d_verts, size_t V, int32_t vertex_index_base, int32_t leaf_triangle_offset);
/* Device pointers and sizes of one object's slices of the reference-layout buffers (triangle vertex indices are GLOBAL there,
 * leaf packs hold global triangle offsets); valid until the next add / commit.  Any out pointer may be NULL. */
int cndl_object_device_view(cndl_ctx* ctx, uint32_t object_id, const void** d_nodes, size_t* N, const cndl_triangle** d_tris, size_t* T,
                            const cndl_vertex** d_verts, size_t* V);

/* GetData (…/Include/TraverseBVHStackless.glsl:370-408) without the texture fetch — the step right after
 * the path: for every hit record, the interpolated half-float vertex normal (normalised) and UV and the
 * entity's emissive / alpha floats.  A miss (t < 0 or mesh < 0) gives normal (-1,-1,-1) and zeros. */
typedef struct cndl_hit_attr { float nx, ny, nz, u, v, emissivity, alpha; int32_t mesh; } cndl_hit_attr;
int cndl_get_data_device(cndl_ctx* ctx, const cndl_hit* d_hits, size_t R, cndl_hit_attr* d_out, void* stream);
int cndl_get_data(cndl_ctx* ctx, const cndl_hit* hits, size_t R, cndl_hit_attr* out);

/* The per-mesh material table GetData indexes by GlobalMeshNumber: BVH::TextureReferences (Source/Core/BVH/Intersector.h:32-37),
 * 32 bytes, uploaded by RayIntersector::GenerateMeshTextureReferences (:404-409) and bound at SSBO slot 4 (:251). */
typedef struct cndl_texture_reference { float model_color[4]; int32_t albedo, normal, pad[2]; } cndl_texture_reference;
/* One entry of FileLoader::GetMeshTexturePaths() (ModelFileLoader.h:21-25) after the texture cache has resolved its two paths
 * (GLClasses::GetTextureCachedDataForPath, Texture.cpp:168-188: a handle and whether the path was found). */
typedef struct cndl_mesh_material { uint64_t albedo_handle, normal_handle; int32_t albedo_valid, normal_valid; float model_color[3]; float pad; } cndl_mesh_material;
/* GenerateMeshTextureReferences (Intersector.h:367-402) as host arithmetic: walks the meshes in order, gives every handle not seen
 * before the next texture-array index — INVALID handles too, they consume an index like in the reference — and writes
 * {vec4(ModelColor, 1), valid ? index : -1, valid ? index : -1} per mesh.  handles (optional, capacity handles_cap) receives the
 * handle bound to Textures[i] (m_TextureHandleReferenceMap inverted, _BindTextures :423-427); *n_handles the number of indices used.
 * Host only; no GPU needed. */
int cndl_generate_texture_references(const cndl_mesh_material* materials, size_t n, cndl_texture_reference* out, uint64_t* handles, size_t handles_cap,
                                     size_t* n_handles);
/* Uploads the table (replacing the previous one, like the reference's glDeleteBuffers + glGenBuffers); n = 0 removes it. */
int cndl_set_texture_references(cndl_ctx* ctx, const cndl_texture_reference* refs, size_t n);
size_t cndl_texture_reference_count(const cndl_ctx* ctx);
/* GetData in full up to the texture unit: cndl_hit_attr plus the Albedo decision (…Stackless.glsl:393-404).  albedo_ref > -1: the
 * caller samples Textures[albedo_ref] at (u, v) and albedo[] is (0,0,0), the shader's initial value; albedo_ref == -1: albedo[] is the
 * mesh's ModelColor.xyz.  A miss gives albedo (0,0,0), albedo_ref -1.  A hit whose mesh number is not in the table (the shader would read
 * past its SSBO) gives albedo (0,0,0), albedo_ref -2; the host-buffer call turns any such record into CNDL_ERR_INVALID. */
typedef struct cndl_hit_material { float nx, ny, nz, u, v, emissivity, alpha; int32_t mesh; float albedo[3]; int32_t albedo_ref; } cndl_hit_material;
int cndl_get_data_material_device(cndl_ctx* ctx, const cndl_hit* d_hits, size_t R, cndl_hit_material* d_out, void* stream);
int cndl_get_data_material(cndl_ctx* ctx, const cndl_hit* hits, size_t R, cndl_hit_material* out);

/* Physics::CollideBox (Source/Core/Physics.cpp:203-228; declared Physics.h:15) for a batch of axis-aligned boxes: does the
 * box touch any triangle of any entity?  collided = 0/1; mesh / tri / entity identify the first overlapping triangle in the
 * reference's walk order (the reference computes them and returns only the bool).  Physics::CollidePoint is the box
 * (P - 0.01, P + 0.01) (:177-179).  Stackless contexts only, like the reference's signature. */
typedef struct cndl_box { float min[3]; float pad0; float max[3]; float pad1; } cndl_box;
typedef struct cndl_collision { int32_t collided, mesh, tri, entity; } cndl_collision;
int cndl_collide_boxes(cndl_ctx* ctx, const cndl_box* boxes, size_t n, cndl_collision* out);
int cndl_collide_boxes_device(cndl_ctx* ctx, const cndl_box* d_boxes, size_t n, cndl_collision* d_out, void* stream);

/* Scene ingest without Assimp (ModelFileLoader.cpp:101-185): reads a Wavefront OBJ into what AddObject consumes — 32-byte Vertex
 * records (normal / UV packed like glm::packHalf2x16, ModelFileLoader.cpp:133-155; tangents zero), object-local indices with
 * the per-mesh vertex offset applied, and one GlobalMeshNumber per triangle (one mesh per usemtl / o / g run, numbered
 * consecutively from first_mesh_number like GlobalMeshCounter, :104-105).  Like the reference's import flags (:243-252): polygons
 * are fan-triangulated, v is flipped (1 - v), corners without a normal get their face's normal, and vertices are joined per mesh
 * when position index, UV index and normal agree.  err (optional) receives a message on failure.  Host only; no GPU needed. */
typedef struct cndl_model cndl_model;
int cndl_model_load_obj(const char* path, int32_t first_mesh_number, cndl_model** out, char* err, size_t err_cap);
/* glTF 2.0 (.gltf with external or base64 buffers, .glb): scene nodes depth first, one mesh per primitive, node transforms NOT
 * applied (ProcessAssimpNode, ModelFileLoader.cpp:187-227, adds meshes as they stand); flat normals when a primitive has none. */
int cndl_model_load_gltf(const char* path, int32_t first_mesh_number, cndl_model** out, char* err, size_t err_cap);
int cndl_model_load(const char* path, int32_t first_mesh_number, cndl_model** out, char* err, size_t err_cap);  /* by extension */
void cndl_model_free(cndl_model* m);
size_t cndl_model_vertex_count(const cndl_model* m);
size_t cndl_model_index_count(const cndl_model* m);
size_t cndl_model_mesh_count(const cndl_model* m);
const cndl_vertex* cndl_model_vertices(const cndl_model* m);
const uint32_t* cndl_model_indices(const cndl_model* m);
const int32_t* cndl_model_mesh_ids(const cndl_model* m);
const char* cndl_model_mesh_name(const cndl_model* m, size_t mesh);
/* The mesh's material as LoadMaterialTextures records it in FileLoader::GetMeshTexturePaths() (ModelFileLoader.cpp:31-99): texture
 * path = directory of the model file + "/" + the name the material gives ("" when it names none; OBJ: map_Kd and norm / map_Kn, the
 * statements Assimp maps to aiTextureType_DIFFUSE / _NORMALS; glTF: baseColorTexture and normalTexture image URIs), ModelColor = the
 * diffuse colour (OBJ Kd, glTF baseColorFactor; Assimp's defaults 0.6 / 1.0 when the file gives none). */
const char* cndl_model_mesh_albedo_path(const cndl_model* m, size_t mesh);
const char* cndl_model_mesh_normal_path(const cndl_model* m, size_t mesh);
int cndl_model_mesh_color(const cndl_model* m, size_t mesh, float rgb[3]);
int cndl_add_model(cndl_ctx* ctx, uint32_t object_id, const cndl_model* m, const cndl_build_opts* opts);  /* = cndl_add_object */
/* glm::packHalf2x16 (glm 0.9.8.5 detail::toFloat16: round to nearest, ties up in magnitude). */
uint32_t cndl_pack_half2x16(float x, float y);

/* Flat-buffer cache of the built scene (the reference rebuilds every BVH at every launch, Pipeline.cpp:1019-1028):
 * cndl_save writes every object's reference-layout nodes / triangles / vertices and the object table; cndl_load restores
 * them into an EMPTY context of the same node format (call cndl_commit afterwards). */
int cndl_save(cndl_ctx* ctx, const char* path);
int cndl_load(cndl_ctx* ctx, const char* path);

/* Pinned host memory for ray / hit batches. */
void* cndl_host_alloc(size_t bytes);
void* cndl_host_alloc_write_combined(size_t bytes);  /* upload-only (rays): the CPU must not read it back */
void cndl_host_free(void* p);

/* Traversal tuning (never changes results): mode 0 = one thread per ray, 1 = persistent warps with
 * ray re-fetch (stackless format only), 2 = persistent while-while with postponed leaf tests (default);
 * sort_rays reorders the rays inside the call (batches of >= 65536 rays): 1 = stable buckets by direction octant; 2 = direction octant, then
 * Morton order of the origin cell (3 bits per axis) by one counting sort, the rays MOVED into that order and the results scattered back;
 * 3 = the same order through an index list (the rays stay where they are); 4 (default) = automatic: 3 when the scene (nodes + triangle
 * records) exceeds 96 MB, i.e. no longer fits the L2, and the batch has at least 2^20 rays (10 M triangles, 12.5 M random rays: 5.36 -> 4.71 ms
 * with the sort counted), else off (L2-resident scenes: ordering inside the call costs what it saves; let the generator emit octant-major). */
int cndl_set_traversal_mode(cndl_ctx* ctx, int mode, int sort_rays);
enum { CNDL_KNOB_BLOCKS_PER_SM = 0,   /* persistent CTAs (128 threads) per SM */
       CNDL_KNOB_LEAF_THRESHOLD = 1,  /* mode 2: parked-at-leaf lanes that trigger the leaf phase */
       CNDL_KNOB_IDLE_THRESHOLD = 2,  /* mode 2: finished lanes that trigger retire/refill */
       CNDL_KNOB_VARIANT = 3,         /* mode 2: 0 = automatic; low 3 bits = node steps per vote round (1..4); 32+ = top of the tree
                                         staged in shared memory, 40+ = two rays per lane (both stackless only) */
       CNDL_KNOB_STACK_LEAF_THRESHOLD = 5, /* mode 2, stack format: parked lanes that trigger the leaf phase */
       CNDL_KNOB_HOST_CHUNKS = 4,     /* host-buffer queries: chunks in the copy/traverse/copy pipeline (0 = default 12) */
       CNDL_KNOB_HOT_NODES = 6,       /* mode 2, stackless: top-of-tree nodes staged in shared memory (<= 7168; takes effect at cndl_commit) */
       CNDL_KNOB_BLOCK_THREADS = 7,   /* mode 2, stackless, staged kernel: threads per CTA (256, 512 or 1024; CTAs per SM = 1024 / threads) */
       CNDL_KNOB_BUILD_SPLIT_NODE = 8, /* SAH builder: ranges longer than this are processed by one CTA per 512 (2048 beyond 2^20 triangles) references instead of one CTA
                                         per node (0 = default 16384; values below 64 are raised to 64); the buffers do not depend on it */
       CNDL_KNOB_BUILD_PACK_MIN = 9   /* SAH builder: levels with at least this many ranges of 3..64 references handle four ranges per warp
                                         (ranges of <= 8 references side by side, eight lanes each); 0 = default 131072, 1 = always; the
                                         buffers do not depend on it */ };
int cndl_set_tuning(cndl_ctx* ctx, int knob, int value);
/* Number of kernels launched by this context so far (bench.py's gpu_launches). */
uint64_t cndl_launch_count(const cndl_ctx* ctx);
/* Milliseconds of the last cndl_add_object build, measured with CUDA events. */
float cndl_last_build_ms(const cndl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CANDELA_B200_H */
