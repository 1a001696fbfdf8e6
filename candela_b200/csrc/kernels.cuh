// Kernel entry points shared between the translation units of libcandela_b200.so.
#pragma once
#include "traverse.cuh"

namespace cndl {

enum QueryKind { Q_CLOSEST = 0, Q_CLOSEST_IGNORE_TRANSPARENT = 1, Q_ANY = 2 };

struct LaunchCounter { unsigned long long n = 0; };

// Function attributes (shared-memory carve-out, dynamic shared memory size) are per device: launchers keep one slot per
// device so that contexts on several GPUs in one process each configure their own copy of a kernel.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return d >= 0 && d < kMaxDevices ? d : 0;
}

// Derives tri48 from the reference-layout triangle and vertex buffers (commit time).
// *d_invalid |= 4 when a triangle names a vertex outside [0, V).
void launch_make_tri48(const int4* tris, const float4* verts /* 2 float4 per vertex */, size_t T, size_t V, float4* tri48, int* d_invalid,
                       cudaStream_t stream, LaunchCounter& lc);

// Adds `offset` to the three vertex indices of T triangles (Intersector.h:190-197).
void launch_rebase_triangles(int4* tris, size_t T, int offset, cudaStream_t stream, LaunchCounter& lc);

// One thread per ray. `order` (optional) maps launch slot -> ray index (sorted traversal).
void launch_trace_simple(const SceneView& s, bool stack, int kind, const cndl_ray* rays, size_t R, const unsigned* order,
                         cndl_hit* hits, float* any_t, cudaStream_t stream, LaunchCounter& lc);

// Persistent warps with ray re-fetch from a global counter (zeroed by the launcher).
void launch_trace_persistent(const SceneView& s, bool stack, int kind, const cndl_ray* rays, size_t R, const unsigned* order,
                             cndl_hit* hits, float* any_t, unsigned* work_counter, int sm_count, cudaStream_t stream,
                             LaunchCounter& lc);

// Mode 2: persistent while-while traversal with postponed leaf tests and a batched service phase (kernels_wavefront.cu).
void launch_trace_ww(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits, float* any_t,
                     unsigned* work_counter, int sm_count, int blocks_per_sm, int park_threshold, int idle_threshold, int steps, bool validated,
                     bool helper_lanes, cudaStream_t stream, LaunchCounter& lc);
void launch_trace_ww_stack(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits, float* any_t,
                           unsigned* work_counter, int sm_count, int blocks_per_sm, int park_threshold, int idle_threshold, int steps, bool validated,
                           cudaStream_t stream, LaunchCounter& lc);
// Stable partition of the ray indices by direction octant (kernels_raygen.cu): order[0..R) lists the rays of octant 0
// in their original order, then octant 1, ...  scratch: octant_partition_scratch_ints(R) ints.
size_t octant_partition_scratch_ints(size_t R);
void launch_octant_partition(const cndl_ray* rays, size_t R, unsigned* order, int* scratch, cudaStream_t stream, LaunchCounter& lc);

// Physics::CollideBox for a batch of boxes over the stackless buffers (kernels_traverse.cu).
void launch_collide_boxes(const SceneView& s, const float4* verts, const cndl_box* boxes, size_t n, cndl_collision* out, cudaStream_t stream,
                          LaunchCounter& lc);

// The camera ray of pixel (x, y): main() of Intersectors/TraverseBVHStack.glsl:414-421 — TexCoords = vec2(Pixel) / u_Dims,
// rD = normalize(GetRayDirectionAt(TexCoords)) (:133-138), rO = u_InverseView[3].xyz — in glm's operation order.
struct Mat2 { float iv[16]; float ip[16]; };
__device__ __forceinline__ void primary_ray(const Mat2& m, int x, int y, int W, int H, float4& o, float4& d) {
    const float tx = fdiv((float)x, (float)W), ty = fdiv((float)y, (float)H);  // vec2(Pixel) / u_Dims
    const float cx = fsub(fmul(tx, 2.0f), 1.0f), cy = fsub(fmul(ty, 2.0f), 1.0f);
    const float* ip = m.ip;
    const float ex = fadd(fadd(fmul(ip[0], cx), fmul(ip[4], cy)), fadd(fmul(ip[8], -1.0f), fmul(ip[12], 1.0f)));
    const float ey = fadd(fadd(fmul(ip[1], cx), fmul(ip[5], cy)), fadd(fmul(ip[9], -1.0f), fmul(ip[13], 1.0f)));
    const V3 dir = xform(m.iv, V3{ex, ey, -1.0f}, 0.0f);
    const float inv_len = fdiv(1.0f, __fsqrt_rn(vdot(dir, dir)));  // glm::normalize: v * inversesqrt(dot(v,v))
    o = make_float4(m.iv[12], m.iv[13], m.iv[14], 0.0f);
    d = make_float4(fmul(dir.x, inv_len), fmul(dir.y, inv_len), fmul(dir.z, inv_len), 1000000.0f);
}

// Camera rays of the primary kernel (Intersectors/TraverseBVHStack.glsl:133-138,:414-431).
void launch_primary_rays(const float* inv_view16, const float* inv_proj16, int W, int H, cndl_ray* rays, cudaStream_t stream,
                         LaunchCounter& lc);

// Wavefront ray generation between bounces (kernels_raygen.cu): diffuse / specular / shadow rays from the hits of the
// previous batch, compacted (and optionally octant-bucketed) with a stable partition.
size_t generate_rays_scratch_ints(size_t R, int spp);
// d_R (optional): the input batch length in device memory, R being its upper bound.  *d_count_out: where the number of rays
// written lives on the device; h_count (optional): also copied to the host, which synchronises `stream`.
cudaError_t generate_rays(const SceneView& s, const cndl_raygen_params& prm, const cndl_ray* rays, const cndl_hit* hits, size_t R, const unsigned* d_R,
                          cndl_ray* out, unsigned* parent, int* scratch, const unsigned** d_count_out, size_t* h_count, cudaStream_t stream,
                          LaunchCounter& lc);
// element e = i * spp + s of the last generate_rays call on `scratch`: keys[e] < 8 iff it produced a ray, stored at dest[e]
void generate_rays_maps(int* scratch, size_t R, int spp, const unsigned char** keys, const unsigned** dest);
cudaError_t generate_rays_tiled(const SceneView& s, const cndl_raygen_params& prm, const cndl_ray* rays, const cndl_hit* hits, size_t R,
                                const unsigned* in_seg_counts, cndl_ray* out, int* scratch, unsigned* seg_counts, unsigned* total, unsigned* oct_cursor,
                                unsigned* oct_list, size_t oct_stride, cudaStream_t stream, LaunchCounter& lc);
// Probe-update rays (UpdateRadianceProbes.glsl:408-427), one per probe of a res[0] x res[1] x res[2] grid.
void launch_probe_rays(const float box_origin[3], const float size[3], const int res[3], unsigned seed, cndl_ray* out, cudaStream_t stream, LaunchCounter& lc);

// GetData without textures: interpolated normal / uv + entity emissive / alpha per hit record (kernels_raygen.cu).
void launch_get_data(const SceneView& s, const float4* verts, const cndl_hit* hits, size_t R, cndl_hit_attr* out, cudaStream_t stream, LaunchCounter& lc);
void launch_get_data_material(const SceneView& s, const float4* verts, const cndl_texture_reference* refs, size_t n_refs, const cndl_hit* hits, size_t R,
                              cndl_hit_material* out, unsigned* out_of_table, cudaStream_t stream, LaunchCounter& lc);

// Hot-first derived node layout + shared-memory staged traversal (kernels_hot.cu).
constexpr int kMaxHotNodes = 7168;  // 224 KB of shared memory
struct HotView {
    const float4* nodes2;       // derived node array (hot nodes first)
    const cndl_entity* ents2;   // entity records whose node_offset is the root's index in nodes2
    int n_hot;                  // nodes staged in shared memory
};
size_t hot_scratch_ints(size_t N, int n_objects);
// objects: (node_offset, node_count) per object in insertion order, on the device and on the host.
// h_roots_out[o] = index of object o's root in nodes2; *h_invalid: bit 0 a link / first child leaves its object, bit 1 a leaf range leaves the scene.
cudaError_t derive_hot_layout(const float4* nodes, size_t N, const int2* d_objects, const int2* h_objects, int n_objects, size_t n_tris, int H,
                              float4* nodes2, int* perm, int* scratch, int* h_roots_out, int* h_n_hot, int* h_invalid, cudaStream_t st,
                              LaunchCounter& lc);
// Stack format: *h_invalid bit 0: a child slot leaves its object, bit 1: a leaf range leaves the scene (d_flag: one int of device scratch).
cudaError_t validate_stack_nodes(const float4* nodes, const int2* h_objects, int n_objects, size_t n_tris, int* d_flag, int* h_invalid, cudaStream_t st,
                                 LaunchCounter& lc);
void launch_trace_hot(const SceneView& s, const HotView& hv, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits,
                      float* any_t, unsigned* work_counter, int sm_count, int block_threads, int park_threshold, int idle_threshold, int steps,
                      cudaStream_t stream, LaunchCounter& lc);

// Two rays per lane over the derived layout (kernels_hot.cu).
void launch_trace_pair(const SceneView& s, const HotView& hv, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits,
                       float* any_t, unsigned* work_counter, int sm_count, int blocks_per_sm, int park_threshold, int idle_threshold, int steps,
                       cudaStream_t stream, LaunchCounter& lc);

}  // namespace cndl
