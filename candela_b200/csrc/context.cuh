// Internal state behind the C ABI (include/candela_b200.h): what RayIntersector<T> keeps
// (Source/Core/BVH/Intersector.h:60-124), held in device memory.  Shared by context.cu (scene + queries)
// and frame.cu (frame-level and multi-device calls).
#pragma once
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <array>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "builder.cuh"
#include "kernels.cuh"

namespace cndl {

struct DeviceBuffer {
    void* p = nullptr;
    size_t bytes = 0, cap = 0;
    ~DeviceBuffer() { if (p) cudaFree(p); }
    // grows keeping the contents
    cudaError_t reserve(size_t want, cudaStream_t st) {
        if (want <= cap) return cudaSuccess;
        size_t ncap = cap ? cap : 4096;
        while (ncap < want) ncap *= 2;
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap);
        if (e != cudaSuccess) return e;
        if (bytes) e = cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (p) cudaFree(p);
        p = q;
        cap = ncap;
        return e;
    }
    // contents not preserved
    cudaError_t ensure_scratch(size_t want) {
        if (want <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
};

constexpr int kCounterSlots = 64;
constexpr int kFrameSlots = 4;   // frames in flight per context (cndl_frame_submit's `slot`)

struct ObjectData { int tri_offset, vert_offset, node_offset, node_count, tri_count, vert_count; };  // _ObjectData, Intersector.h:51-56 (+ counts)

// Scratch of one frame in flight (frame.cu): two of these per context, so that the device->host copy of one frame overlaps
// the tracing of the next.
struct FrameSlot {
    DeviceBuffer prim_rays, prim_hits, pix_ids, rays[2], hits, rids[2], gen_scratch, acc, out, counts, oct_list;
    unsigned* h_counts = nullptr;      // pinned: rays traced per bounce, copied back with the frame
    int n_counts = 0;
    cudaEvent_t traced = nullptr, copied = nullptr;
    bool pending = false;
    unsigned long long rays_traced = 0;
};

}  // namespace cndl

struct cndl_ctx {
    int format = CNDL_STACKLESS;
    int device = 0;
    int sm_count = 148;
    size_t node_size = 32;
    std::string err;

    // m_BVHNodes / m_BVHTriangles / m_BVHVertices (device-resident) and their element counts
    cndl::DeviceBuffer nodes, tris, verts, tri48, ents;
    size_t n_nodes = 0, n_tris = 0, n_verts = 0;
    size_t committed_nodes = 0, committed_tris = 0;  // m_NodeCountBuffered
    bool committed = false;
    std::unordered_map<uint32_t, cndl::ObjectData> objects;

    std::vector<cndl_entity> staged;  // m_Entities
    size_t n_ents = 0;                 // m_EntityPushed
    cndl::DeviceBuffer tex_refs;       // m_BVHTextureReferencesSSBO (Intersector.h:90)
    size_t n_tex_refs = 0;
    bool ents_buffered = false;

    // query scratch
    cndl::DeviceBuffer d_rays, d_hits, d_order, d_keys, d_sort_tmp, d_counter, d_chunk_counters, d_sorted;
    std::vector<cudaEvent_t> events;
    cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};  // H2D, traversal (even chunks), D2H, traversal (odd chunks)
    cudaStream_t main_stream = nullptr;
    int mode = 2, sort_rays = 4;   // sort_rays: 0 off, 1 octant buckets, 2 octant + origin Morton order (rays moved), 3 the same through an index list, 4 automatic
    float world_lo[3] = {0, 0, 0}, world_hi[3] = {0, 0, 0};  // bounds of all entities (for sort_rays = 2)
    int knobs[10] = {8, 14, 10, 0, 0, 12, 4096, 1024, 0, 0};  // CNDL_KNOB_*
    // hot-first derived layout of the stackless nodes (kernels_hot.cu), rebuilt by cndl_commit
    cndl::DeviceBuffer nodes2, perm, ents2, hot_scratch, d_objects;
    std::vector<int2> h_objects;                 // (node_offset, node_count) in insertion order
    std::vector<int> h_roots;                    // root of each object in nodes2
    std::vector<cndl_entity> buffered;           // the entity records last uploaded
    int n_hot = 0;
    bool hot_ready = false, hot_entities_ok = false;
    bool nodes_valid = false, entities_regular = false;  // links / slots / leaf ranges in bounds; every entity names a whole object
    cndl::LaunchCounter launches;
    float last_build_ms = 0.0f;
    void* build_arena = nullptr;
    size_t build_arena_cap = 0;
    int* build_host_counts = nullptr;
    unsigned counter_next = 0;   // next_counter(): ring of work-counter slots for device calls
    cndl::FrameSlot frame[cndl::kFrameSlots];
    cudaStream_t frame_stream[cndl::kFrameSlots] = {}, frame_copy_stream = nullptr;  // one compute stream per slot: the kernels of two frames in flight interleave
    int tri_offset_bias = 0;  // cndl_build_bvh: BuildBVH's t_offset for a stand-alone build

    int fail(int code, const std::string& msg) { err = msg; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? CNDL_ERR_OOM : CNDL_ERR_CUDA;
    }
};

// Nothing may throw across the C boundary (std::vector / std::string / unordered_map allocate): every entry point that
// allocates is a function-try-block ending in CNDL_CATCH.
#define CNDL_CATCH                                        \
    catch (const std::bad_alloc&) { return CNDL_ERR_OOM; } \
    catch (...) { return CNDL_ERR_INVALID; }

#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return ctx->cuda_fail(e__, #call); \
    } while (0)


namespace cndl {
SceneView scene_view(const cndl_ctx* ctx);
size_t order_region_ints(size_t R);
int effective_sort(const cndl_ctx* ctx, size_t R);
int check_ready(cndl_ctx* ctx);
// Enqueues one traversal batch on `st`.  scratch: 16 unsigned ints ([0] work counter); order_region: order_region_ints(R)
// unsigned ints, used when ray ordering is on; d_R (optional): the batch length in device memory, R being its upper bound;
// preset (optional): the caller has already ordered the batch (the frame call's generator writes segmented batches and octant
// lists): the launch takes this RayOrder as it is and d_R / the in-call ordering are ignored.
int enqueue_trace(cndl_ctx* ctx, int kind, const cndl_ray* d_rays, size_t R, const unsigned* d_R, cndl_hit* d_hits, float* d_any, unsigned* scratch,
                  unsigned* order_region, cndl_ray* sorted_region, cudaStream_t st, const cndl::RayOrder* preset = nullptr);
// One 64-byte work-counter slot from the context's ring: device calls in flight on different streams never share one.
unsigned* next_counter(cndl_ctx* ctx);
}  // namespace cndl
