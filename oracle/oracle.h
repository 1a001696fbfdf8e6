/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the Candela BVH build + ray-traversal hot path: a from-scratch
 * restatement, in scalar C++17, of
 *   - the reference's CPU binned-SAH builder and its two flatteners
 *     (/root/reference/Source/Core/BVH/BVHConstructor.cpp), and
 *   - the reference's GLSL traversal routines
 *     (/root/reference/Source/Core/Shaders/Intersectors/Include/TraverseBVHStackless.glsl,
 *      .../TraverseBVHStack.glsl).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker.  The product
 * (libcandela_b200.so) never links, loads or calls it.
 *
 * Pinning status
 *   builder  : PINNED.  tests/test_oracle_vs_reference.py (run where
 *              /root/reference exists) and tests/golden/ (hashes made by
 *              tests/golden/make_golden.py from the compiled, unmodified
 *              reference builder in oracle/_ref/) require byte equality of the
 *              node, triangle and vertex buffers.
 *   traversal: PINNED.  The reference's traversal exists only as GLSL and has no
 *              tests of its own, but its two shader include files compile as C++
 *              against the reference's vendored glm after a purely syntactic
 *              rewrite (oracle/ref_shim/glsl_to_cpp.py + ref_glsl.cpp ->
 *              oracle/_ref/libcandela_ref.so).  tests/test_oracle_traversal.py
 *              requires bit-identical hit records / any-hit distances on every
 *              case (live where /root/reference exists; everywhere through the
 *              committed outputs tests/golden/reference_traversal_golden.npz).
 *   GetData  : PINNED the same way (tests/test_get_data.py).
 *   collide  : PINNED against the compiled Physics.cpp (tests/test_collide.py).
 *   ray generators: PINNED.  CosWeightedHemisphere, SampleGGXVNDF, SampleCone
 *              (Include/Sampling.glsl), StochasticReflectionDirection
 *              (SpecularTrace.glsl:102-135) and LambertBRDF / ImportanceSample
 *              (UpdateRadianceProbes.glsl:351-374) are compiled from the shader
 *              files with hash2() bound to the counter stream and sin / cos /
 *              acos / pow bound to exact_math_ref.h (GLSL leaves both to the
 *              implementation); tests/test_raygen_oracle.py requires identical
 *              bits, and tests/golden/raygen_golden.npz carries the compiled
 *              shaders' outputs to the GPU box.
 *
 * Build flags are part of the definition of "the reference result":
 *   g++ -std=c++17 -O2 -ffp-contract=off   (no -march=native, no -ffast-math)
 */
#ifndef CANDELA_ORACLE_H
#define CANDELA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 32-byte records as include/candela_b200.h (kept separate on purpose:
 * the oracle does not include product headers). */
typedef struct { float ox, oy, oz, tmin; float dx, dy, dz, tmax; } orc_ray;
typedef struct { float t, u, v, w; int32_t mesh, tri, entity, iters; } orc_hit;

enum { ORC_STACKLESS = 0, ORC_STACK = 1 };
enum { ORC_SWAP_NONE = 0, ORC_SWAP_HASHED = 1 };
enum { ORC_CLOSEST = 0, ORC_CLOSEST_IGNORE_TRANSPARENT = 1, ORC_ANY = 2 };

typedef struct orc_bvh orc_bvh; /* one object's build result */

/* BuildBVH restated (BVHConstructor.cpp:951-1108).  `verts` are 32-byte
 * Vertex records (Utils/Vertex.h:7-12), `indices` already carry the per-mesh
 * vertex offset (:981-1002, see orc_concat_meshes), one mesh id per triangle.
 * swap_policy: ORC_SWAP_NONE keeps (left,right) as built; ORC_SWAP_HASHED
 * flips a child pair when a hash of (seed, range start, range length) is odd
 * -- a reproducible stand-in for the reference's random_device coin
 * (:389-391, :599-609).  Stack format never flips (:599).
 * Returns NULL on bad input (I % 3 != 0, I == 0, or an index >= V). */
orc_bvh* orc_build(int format, const void* verts, uint64_t V, const uint32_t* indices, uint64_t I,
                   const int32_t* mesh_id_per_tri, int32_t t_offset, int swap_policy, uint64_t swap_seed);
void orc_bvh_free(orc_bvh*);
uint64_t orc_bvh_node_count(const orc_bvh*);   /* LastNodeIndex + 1 */
uint64_t orc_bvh_tri_count(const orc_bvh*);
/* stats[0..5] = nodes created (LastNodeIndex), leaves, split failures,
 * max build-stack depth, inner slots used by the stack flattener, tree depth */
void orc_bvh_stats(const orc_bvh*, uint64_t stats[6]);
void orc_bvh_fetch(const orc_bvh*, void* nodes, void* tris);
/* refs[i] = original triangle number stored at sorted position i */
void orc_bvh_fetch_order(const orc_bvh*, int32_t* sorted_refs);
/* Reads the child flips out of a stackless node buffer produced by the
 * reference for the same input (which flips at random), applies them to this
 * tree and re-flattens.  Afterwards orc_bvh_fetch must equal that buffer byte
 * for byte.  Returns the number of flipped inner nodes, or -1 if the buffer
 * does not describe the same tree. */
int64_t orc_bvh_adopt_flips(orc_bvh*, const void* ref_stackless_nodes, uint64_t n_nodes);
/* SAH cost of the tree: sum over inner nodes of area(node)/area(root) plus
 * sum over leaves of len*area(leaf)/area(root).  A quality figure, not part
 * of the reference. */
double orc_bvh_sah_cost(const orc_bvh*);

/* Mesh concatenation of BuildBVH (BVHConstructor.cpp:981-1002). Outputs must
 * hold sum(V), sum(I), sum(I)/3 entries. */
void orc_concat_meshes(int n_meshes, const void* verts, const uint64_t* mesh_vertex_counts,
                       const uint32_t* indices, const uint64_t* mesh_index_counts,
                       const int32_t* mesh_global_numbers,
                       void* out_verts, uint32_t* out_indices, int32_t* out_mesh_id_per_tri);

/* RayIntersector<T>::AddObject (Intersector.h:170-198) applied to a build
 * result: rebases the three vertex indices of every triangle by
 * `index_offset` (= vertices already in the intersector). */
void orc_rebase_triangles(void* tris, uint64_t T, uint32_t index_offset);

/* RayIntersector<T>::PushEntity (Intersector.h:201-216): fills one 192-byte
 * BVHEntity from a column-major model matrix; the inverse follows glm 0.9.8.5
 * compute_inverse<mat4> (glm/detail/func_matrix.inl). */
void orc_make_entity(const float model[16], int32_t node_offset, int32_t node_count,
                     float emissive, float translucency, void* out_entity192);

/* Primary camera rays (Intersectors/TraverseBVHStack.glsl:133-138, :414-431;
 * Intersector.h:255-265): pixel (x,y) -> rays[y*W+x]; tmin=0, tmax=1e6. */
void orc_primary_rays(const float inv_view[16], const float inv_proj[16], int W, int H, orc_ray* rays);

/* Scene-level traversal (IntersectScene / IntersectSceneIgnoreTransparent /
 * any-hit IntersectScene; ...Stackless.glsl:280-366,:558-575 and
 * ...Stack.glsl:327-413,:659-676).
 *   kind ORC_CLOSEST / ORC_CLOSEST_IGNORE_TRANSPARENT: writes `hits` (R records).
 *   kind ORC_ANY: writes `any_t` (R floats): first accepted t or -1.  The
 *   reference fixes TMax = 1e6; a ray with tmax > 0 uses that instead
 *   (documented extension).  Closest-hit ignores ray.tmin/tmax like the reference.
 *   counters (may be NULL): [0] node iterations summed over rays and entities,
 *   [1] triangle tests, [2] rays that reached the 1024-iteration cap in some
 *   entity, [3] rays reporting a hit (t > 0 in the output record).
 *   nthreads <= 1 runs on the calling thread; otherwise std::thread over
 *   contiguous ray ranges. */
void orc_trace(int format, int kind,
               const void* nodes, uint64_t total_nodes, const void* tris, const void* verts,
               const void* entities, int32_t n_entities,
               const orc_ray* rays, uint64_t R, orc_hit* hits, float* any_t,
               uint64_t counters[4], int nthreads);

/* Brute force over all triangles of all entities (no BVH): the closest
 * accepted t and its triangle, ties to the lowest triangle index.  Used to
 * cross-check the traversal restatement, not a reference function. */
void orc_brute_force(const void* tris, uint64_t T, const void* verts, const void* entities, int32_t n_entities,
                     const orc_ray* rays, uint64_t R, orc_hit* hits, int nthreads);

/* GetData (...Stackless.glsl:370-408) without the texture fetch: per hit record, the interpolated half-float
 * vertex normal (normalised) and UV, and the entity's emissive / alpha floats.  out: R records of 32 bytes
 * {float nx,ny,nz,u,v,emissivity,alpha; int32 mesh}; a miss (t < 0 or mesh < 0) gives normal (-1,-1,-1), rest 0. */
void orc_get_data(const void* tris, const void* verts, const void* entities, const orc_hit* hits, uint64_t R, void* out);

/* Physics::CollideBox (Source/Core/Physics.cpp:21-228) for n query boxes of 32 bytes {min.xyz, pad, max.xyz, pad} over the
 * stackless buffers.  out: n x {collided, mesh, tri, entity} (the first overlapping triangle in walk order). */
void orc_collide_boxes(const void* nodes, uint64_t n_nodes, const void* tris, const void* verts, const void* entities, int32_t n_entities,
                       const float* boxes, uint64_t n, int32_t* out);

/* ---- ray generators (oracle_raygen.cpp): the step in front of the path.  Bits of sin / cos / acos / pow and the random
 * stream are DEFINED (exact_math_ref.h, pcg-hash counter stream); everything else follows the shader functions, which
 * oracle/ref_shim compiles for the pin (tests/test_raygen_oracle.py). */
typedef struct { int32_t kind, spp; uint32_t seed, flags; float offset, tmax, roughness; float light_dir[3]; float light_cone; } orc_raygen_params;
float orc_xsin(float x);
float orc_xcos(float x);
float orc_xacos(float x);
float orc_xpow(float x, float y);
void orc_xmath_batch(int which, const float* x, const float* y, uint64_t n, float* out);
void orc_sample_directions(int which, const float* normals, const float* incident, const float* xi, const uint32_t* keys, float roughness, uint64_t n,
                           float* out);
void orc_hash2_stream(uint32_t key, uint32_t m, float* out);
uint32_t orc_stream_key(uint32_t seed, uint32_t element);
uint64_t orc_generate_rays(const void* params, const orc_ray* rays, const orc_hit* hits, const uint32_t* ids_in, uint64_t R, const void* tris,
                           const void* verts, const void* entities, orc_ray* out_rays, uint32_t* parent_out, uint32_t* ids_out);
void orc_probe_rays(const float box_origin[3], const float size[3], const int32_t res[3], uint32_t seed, orc_ray* out);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
