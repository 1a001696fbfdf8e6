// GPU BVH construction for cndl_add_object (replaces the CPU BVH::BuildBVH,
// Source/Core/BVH/BVHConstructor.cpp:951-1108).
#pragma once
#include <string>

#include "kernels.cuh"

namespace cndl {

struct BuildRequest {
    int format;                  // CNDL_STACKLESS / CNDL_STACK
    cndl_build_opts opts;
    const float4* d_verts;       // device, 2 float4 per 32-byte vertex, object-local
    size_t V;
    const uint32_t* h_indices;   // host, 3 per triangle, object-local
    const int32_t* h_mesh_ids;   // host, one per triangle, may be null (-> 0)
    size_t T;
    int tri_offset;              // TriangleOffset_: triangles already in the intersector
    void* d_nodes_out;           // device, room for 2T-1 nodes of the context's format
    int4* d_tris_out;            // device, T records, vertex indices object-local
    size_t n_nodes_out;          // result: LastNodeIndex + 1
    void** arena;                // scratch allocation kept by the context between builds
    size_t* arena_cap;
    int** host_counts;           // mapped pinned memory kept by the context: the host copy of the builder's level descriptors (2 MB)
    cudaStream_t side[2] = {nullptr, nullptr};  // SAH builder: streams for the size classes of one level (optional)
    unsigned split_node = 0;     // SAH builder: ranges longer than this are split across CTAs (0 = default)
    unsigned pack_min = 0;       // SAH builder: levels with at least this many tiny ranges take four ranges per warp (0 = default)
};

// Returns CNDL_OK or a negative cndl_status with `err` set. Work is enqueued on `st` and complete on return.
int build_object(BuildRequest& rq, cudaStream_t st, LaunchCounter& lc, float* build_ms, std::string& err);

// sort_rays = 2: order[0..R) = ray indices sorted by (direction octant, Morton code of the origin cell inside lo..hi).
size_t ray_sort_scratch_ints(size_t R);
// order_out[pos] = original index of the ray sorted to pos; sorted_out (optional, R rays): the rays moved to their sorted places
cudaError_t sort_rays_morton(const cndl_ray* rays, size_t R, const float lo[3], const float hi[3], unsigned* order_out, cndl_ray* sorted_out, int* scratch, cudaStream_t st,
                             LaunchCounter& lc);

// Morton helper shared by the LBVH builder and the ray ordering
static __device__ __forceinline__ unsigned expand10(unsigned v) {  // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

}  // namespace cndl
