// Kernel entry points shared between the translation units of libcandela_b200.so.
#pragma once
#include "traverse.cuh"

namespace cndl {

enum QueryKind { Q_CLOSEST = 0, Q_CLOSEST_IGNORE_TRANSPARENT = 1, Q_ANY = 2 };

struct LaunchCounter { unsigned long long n = 0; };

// Derives tri48 from the reference-layout triangle and vertex buffers (commit time).
void launch_make_tri48(const int4* tris, const float4* verts /* 2 float4 per vertex */, size_t T, float4* tri48,
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

// Mode 2: persistent while-while traversal with postponed leaf tests and batched retire/refill (stackless only).
void launch_trace_ww(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits, float* any_t,
                     unsigned* work_counter, int sm_count, int blocks_per_sm, int leaf_threshold, int idle_threshold, int variant, cudaStream_t stream,
                     LaunchCounter& lc);

void launch_trace_ww_stack(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits, float* any_t,
                           unsigned* work_counter, int sm_count, int leaf_threshold, int idle_threshold, cudaStream_t stream, LaunchCounter& lc);

// Camera rays of the primary kernel (Intersectors/TraverseBVHStack.glsl:133-138,:414-431).
void launch_primary_rays(const float* inv_view16, const float* inv_proj16, int W, int H, cndl_ray* rays, cudaStream_t stream,
                         LaunchCounter& lc);

// Wavefront compaction between bounces: diffuse rays from the hits of the previous batch (kernels_raygen.cu).
cudaError_t generate_bounce_rays(const SceneView& s, const cndl_ray* rays, const cndl_hit* hits, size_t R, int spp, float offset, float tmax,
                                 unsigned seed, cndl_ray* out, unsigned* parent, int* scratch, size_t* h_count, cudaStream_t stream,
                                 LaunchCounter& lc);

// GetData without textures: interpolated normal / uv + entity emissive / alpha per hit record (kernels_raygen.cu).
void launch_get_data(const SceneView& s, const float4* verts, const cndl_hit* hits, size_t R, cndl_hit_attr* out, cudaStream_t stream, LaunchCounter& lc);

}  // namespace cndl
