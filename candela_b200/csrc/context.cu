// C ABI of libcandela_b200.so (include/candela_b200.h): the state RayIntersector<T> keeps
// (Source/Core/BVH/Intersector.h:60-124) held in device memory, plus the query entry points.
#include "context.cuh"

#include <map>
#include <mutex>

using namespace cndl;

namespace cndl {

SceneView scene_view(const cndl_ctx* ctx) {
    SceneView s;
    s.nodes = static_cast<const float4*>(ctx->nodes.p);
    s.tri48 = static_cast<const float4*>(ctx->tri48.p);
    s.tris = static_cast<const int4*>(ctx->tris.p);
    s.ents = static_cast<const cndl_entity*>(ctx->ents.p);
    s.total_nodes = (int)ctx->committed_nodes;
    s.n_ents = (int)ctx->n_ents;
    return s;
}

// Entity records for the derived layout: node_offset becomes the root's index in nodes2.  Every
// record must name an object's whole node range, as PushEntity produces (Intersector.h:210-211).
static int upload_hot_entities(cndl_ctx* ctx) {
    ctx->hot_entities_ok = false;
    ctx->entities_regular = false;
    if (!ctx->nodes_valid) return CNDL_OK;
    std::vector<cndl_entity> e2 = ctx->buffered;
    for (auto& e : e2) {
        auto it = std::lower_bound(ctx->h_objects.begin(), ctx->h_objects.end(), e.node_offset, [](const int2& o, int v) { return o.x < v; });
        if (it == ctx->h_objects.end() || it->x != e.node_offset || it->y != e.node_count) return CNDL_OK;  // irregular: reference-layout kernel
        if (ctx->hot_ready) e.node_offset = ctx->h_roots[(size_t)(it - ctx->h_objects.begin())];
    }
    ctx->entities_regular = true;
    if (!ctx->hot_ready) return CNDL_OK;
    CK(ctx->ents2.ensure_scratch((e2.size() ? e2.size() : 1) * sizeof(cndl_entity)));
    if (!e2.empty()) CK(cudaMemcpy(ctx->ents2.p, e2.data(), e2.size() * sizeof(cndl_entity), cudaMemcpyHostToDevice));
    ctx->hot_entities_ok = true;
    return CNDL_OK;
}

size_t order_region_ints(size_t R) { return R + std::max(octant_partition_scratch_ints(R), ray_sort_scratch_ints(R)) + 16; }

int check_ready(cndl_ctx* ctx) {
    if (!ctx->committed) return ctx->fail(CNDL_ERR_NOT_COMMITTED, "cndl_commit has not been called");
    if (!ctx->ents_buffered) return ctx->fail(CNDL_ERR_NOT_COMMITTED, "cndl_buffer_entities has not been called");
    return CNDL_OK;
}

// sort_rays as it applies to a batch of R rays: 4 = automatic
int effective_sort(const cndl_ctx* ctx, size_t R) {
    if (ctx->sort_rays != 4) return ctx->sort_rays;
    const size_t scene_bytes = ctx->committed_nodes * ctx->node_size + ctx->committed_tris * 48;
    return (scene_bytes > ((size_t)96 << 20) && R >= (1u << 20)) ? 3 : 0;
}

// 64-byte work-counter slots handed out round-robin: two device calls in flight on different streams never share one
// (a slot is reused after kCounterSlots further calls on the context).
unsigned* next_counter(cndl_ctx* ctx) {
    unsigned* p = reinterpret_cast<unsigned*>(static_cast<char*>(ctx->d_counter.p) + 256 + 64 * (size_t)(ctx->counter_next % kCounterSlots));
    ctx->counter_next++;
    return p;
}

// Enqueues one traversal batch on `st`.  scratch: 16 unsigned ints ([0] work counter, [8..15] octant counts);
// order_region: order_region_ints(R) unsigned ints, used when ray bucketing is on.
int enqueue_trace(cndl_ctx* ctx, int kind, const cndl_ray* d_rays, size_t R, const unsigned* d_R, cndl_hit* d_hits, float* d_any, unsigned* scratch,
                  unsigned* order_region, cndl_ray* sorted_region, cudaStream_t st, const RayOrder* preset) {
    if (R > 0xFFFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "more than 2^32-16 rays in one call");
    if ((reinterpret_cast<uintptr_t>(d_rays) & 31u) || (reinterpret_cast<uintptr_t>(d_hits) & 31u))
        return ctx->fail(CNDL_ERR_INVALID, "ray and hit buffers must be 32-byte aligned (the kernels move each record with one 256-bit access)");
    if ((d_R || preset) && ctx->mode != 2) return ctx->fail(CNDL_ERR_INVALID, "a device-side batch length needs traversal mode 2");
    const SceneView s = scene_view(ctx);
    const bool stack = ctx->format == CNDL_STACK;
    if (ctx->mode == 0 || (stack && ctx->mode == 1)) launch_trace_simple(s, stack, kind, d_rays, R, nullptr, d_hits, d_any, st, ctx->launches);
    else if (ctx->mode == 1) launch_trace_persistent(s, stack, kind, d_rays, R, nullptr, d_hits, d_any, scratch, ctx->sm_count, st, ctx->launches);
    else {
        RayOrder order{nullptr, nullptr, 0, d_R, nullptr, nullptr};
        if (preset) order = *preset;
        // sort_rays 4 (the default) decides by the scene: once nodes + triangle records no longer fit the L2, ordering a large batch by
        // (octant, origin cell) pays for itself (10 M triangles, 12.5 M random rays: 5.36 -> 4.71 ms with the sort counted)
        const int sort_mode = effective_sort(ctx, R);
        if (sort_mode && order_region && R >= 65536 && !d_R && !preset) {
            if (sort_mode >= 2) {
                // 2: the rays are MOVED into sorted order (sorted_region) and the results scattered back through the index list;
                // 3: the rays stay and are read through the index list
                cndl_ray* sorted = sort_mode == 2 ? sorted_region : nullptr;
                cudaError_t se = sort_rays_morton(d_rays, R, ctx->world_lo, ctx->world_hi, order_region, sorted, reinterpret_cast<int*>(order_region + R), st,
                                                  ctx->launches);
                if (se != cudaSuccess) return ctx->cuda_fail(se, "ray sort");
                if (sorted) {
                    d_rays = sorted;
                    order = RayOrder{nullptr, nullptr, 0, nullptr, order_region};
                } else {
                    order = RayOrder{order_region, nullptr, 0, nullptr, nullptr};
                }
            } else {
                launch_octant_partition(d_rays, R, order_region, reinterpret_cast<int*>(order_region + R), st, ctx->launches);
                order = RayOrder{order_region, nullptr, 0, nullptr, nullptr};
            }
        }
        int variant = ctx->knobs[CNDL_KNOB_VARIANT];
        if (variant == 0) {
            // automatic: a scene that fits the 126 MB L2 is served best by the plain kernel; once the node and triangle
            // records spill to DRAM, staging the top of the tree in shared memory wins (10 M triangles: +8.7 %)
            const size_t working_set = ctx->committed_nodes * ctx->node_size + ctx->committed_tris * 48;
            // ... unless the batch has just been ordered: neighbouring rays then share nodes, L2 hits rise and the plain kernel's
            // higher residency wins (10 M triangles: 4.71 vs 5.03 ms)
            variant = (!stack && working_set > ((size_t)96 << 20) && !(sort_mode >= 2 && order_region && R >= 65536 && !d_R && !preset)) ? 34 : 18;
        }
        int steps = variant & 7;
        if (steps < 1 || steps > 4) steps = 2;
        const int park = ctx->knobs[stack ? CNDL_KNOB_STACK_LEAF_THRESHOLD : CNDL_KNOB_LEAF_THRESHOLD], idle = ctx->knobs[CNDL_KNOB_IDLE_THRESHOLD];
        if (stack) {
            launch_trace_ww_stack(s, kind, d_rays, R, order, d_hits, d_any, scratch, ctx->sm_count, ctx->knobs[CNDL_KNOB_BLOCKS_PER_SM], park, idle, steps,
                                  ctx->nodes_valid && ctx->entities_regular && !(variant & 8), st, ctx->launches);
        } else if (variant >= 32 && ctx->hot_ready && ctx->hot_entities_ok) {
            HotView hv;
            hv.nodes2 = static_cast<const float4*>(ctx->nodes2.p);
            hv.ents2 = static_cast<const cndl_entity*>(ctx->ents2.p);
            hv.n_hot = ctx->n_hot;
            if (variant >= 40)
                launch_trace_pair(s, hv, kind, d_rays, R, order, d_hits, d_any, scratch, ctx->sm_count, ctx->knobs[CNDL_KNOB_BLOCKS_PER_SM], park, idle, steps,
                                  st, ctx->launches);
            else
                launch_trace_hot(s, hv, kind, d_rays, R, order, d_hits, d_any, scratch, ctx->sm_count, ctx->knobs[CNDL_KNOB_BLOCK_THREADS], park, idle, steps,
                                 st, ctx->launches);
        } else {
            launch_trace_ww(s, kind, d_rays, R, order, d_hits, d_any, scratch, ctx->sm_count, ctx->knobs[CNDL_KNOB_BLOCKS_PER_SM], park, idle, steps,
                            ctx->hot_ready && ctx->hot_entities_ok && !(variant & 8), (variant & 16) != 0, st,
                            ctx->launches);  // variant bit 3: keep the range checks; bit 4: helper lanes in the leaf phase
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ctx->cuda_fail(e, "traversal launch");
    return CNDL_OK;
}

// glm 0.9.8.5 compute_inverse<tmat4x4> (glm/detail/func_matrix.inl:297-353), host side, used by
// cndl_push_entity like Intersector.h:209.  Compiled without contraction (-ffp-contract=off).
void glm_inverse(const float* m, float* out) {
    auto M = [&](int c, int r) { return m[4 * c + r]; };
    const float c00 = M(2,2) * M(3,3) - M(3,2) * M(2,3), c02 = M(1,2) * M(3,3) - M(3,2) * M(1,3), c03 = M(1,2) * M(2,3) - M(2,2) * M(1,3);
    const float c04 = M(2,1) * M(3,3) - M(3,1) * M(2,3), c06 = M(1,1) * M(3,3) - M(3,1) * M(1,3), c07 = M(1,1) * M(2,3) - M(2,1) * M(1,3);
    const float c08 = M(2,1) * M(3,2) - M(3,1) * M(2,2), c10 = M(1,1) * M(3,2) - M(3,1) * M(1,2), c11 = M(1,1) * M(2,2) - M(2,1) * M(1,2);
    const float c12 = M(2,0) * M(3,3) - M(3,0) * M(2,3), c14 = M(1,0) * M(3,3) - M(3,0) * M(1,3), c15 = M(1,0) * M(2,3) - M(2,0) * M(1,3);
    const float c16 = M(2,0) * M(3,2) - M(3,0) * M(2,2), c18 = M(1,0) * M(3,2) - M(3,0) * M(1,2), c19 = M(1,0) * M(2,2) - M(2,0) * M(1,2);
    const float c20 = M(2,0) * M(3,1) - M(3,0) * M(2,1), c22 = M(1,0) * M(3,1) - M(3,0) * M(1,1), c23 = M(1,0) * M(2,1) - M(2,0) * M(1,1);
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {M(1,0), M(0,0), M(0,0), M(0,0)}, v1[4] = {M(1,1), M(0,1), M(0,1), M(0,1)};
    const float v2[4] = {M(1,2), M(0,2), M(0,2), M(0,2)}, v3[4] = {M(1,3), M(0,3), M(0,3), M(0,3)};
    const float sa[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sb[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
    float inv[16];
    for (int k = 0; k < 4; ++k) {
        inv[0 + k] = (v1[k] * f0[k] - v2[k] * f1[k] + v3[k] * f2[k]) * sa[k];
        inv[4 + k] = (v0[k] * f0[k] - v2[k] * f3[k] + v3[k] * f4[k]) * sb[k];
        inv[8 + k] = (v0[k] * f1[k] - v1[k] * f3[k] + v3[k] * f5[k]) * sa[k];
        inv[12 + k] = (v0[k] * f2[k] - v1[k] * f4[k] + v2[k] * f5[k]) * sb[k];
    }
    const float d0 = M(0,0) * inv[0], d1 = M(0,1) * inv[4], d2 = M(0,2) * inv[8], d3 = M(0,3) * inv[12];
    const float ood = 1.0f / ((d0 + d1) + (d2 + d3));
    for (int k = 0; k < 16; ++k) out[k] = inv[k] * ood;
}

}  // namespace cndl

extern "C" {

int cndl_abi_version(void) { return CNDL_ABI_VERSION; }

int cndl_create(cndl_ctx** out, int node_format, int device) {
    if (!out) return CNDL_ERR_INVALID;
    *out = nullptr;
    if (node_format != CNDL_STACKLESS && node_format != CNDL_STACK) return CNDL_ERR_INVALID;  // Intersector.h:146-148
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return CNDL_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CNDL_ERR_NO_DEVICE;
    if (prop.major != 10) return CNDL_ERR_NO_DEVICE;  // built for sm_100a only
    cndl_ctx* ctx = new (std::nothrow) cndl_ctx;
    if (!ctx) return CNDL_ERR_OOM;
    ctx->format = node_format;
    ctx->node_size = node_format == CNDL_STACKLESS ? sizeof(cndl_node) : sizeof(cndl_stack_node);
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CNDL_ERR_CUDA; }
    for (auto& s : ctx->streams)
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CNDL_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CNDL_ERR_CUDA; }
    if (ctx->d_counter.ensure_scratch(256 + 64 * kCounterSlots) != cudaSuccess) { delete ctx; return CNDL_ERR_CUDA; }
    *out = ctx;
    return CNDL_OK;
}

void cndl_destroy(cndl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& s : ctx->streams) if (s) cudaStreamDestroy(s);
    if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
    for (auto& fs : ctx->frame_stream) if (fs) cudaStreamDestroy(fs);
    if (ctx->frame_copy_stream) cudaStreamDestroy(ctx->frame_copy_stream);
    for (auto& f : ctx->frame) {
        if (f.traced) cudaEventDestroy(f.traced);
        if (f.copied) cudaEventDestroy(f.copied);
        if (f.h_counts) cudaFreeHost(f.h_counts);
    }
    if (ctx->build_arena) cudaFree(ctx->build_arena);
    if (ctx->build_host_counts) cudaFreeHost(ctx->build_host_counts);
    for (auto& e : ctx->events) cudaEventDestroy(e);
    delete ctx;
}

const char* cndl_last_error(const cndl_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

size_t cndl_node_count(const cndl_ctx* ctx) { return ctx ? ctx->n_nodes : 0; }
size_t cndl_triangle_count(const cndl_ctx* ctx) { return ctx ? ctx->n_tris : 0; }
size_t cndl_vertex_count(const cndl_ctx* ctx) { return ctx ? ctx->n_verts : 0; }
size_t cndl_entity_count(const cndl_ctx* ctx) { return ctx ? ctx->n_ents : 0; }
uint64_t cndl_launch_count(const cndl_ctx* ctx) { return ctx ? ctx->launches.n : 0; }
float cndl_last_build_ms(const cndl_ctx* ctx) { return ctx ? ctx->last_build_ms : 0.0f; }

size_t cndl_object_count(const cndl_ctx* ctx) { return ctx ? ctx->objects.size() : 0; }

size_t cndl_object_ids(const cndl_ctx* ctx, uint32_t* ids_out, size_t capacity) {
    if (!ctx || !ids_out) return 0;
    std::vector<std::pair<int, uint32_t>> order;
    for (const auto& kv : ctx->objects) order.emplace_back(kv.second.node_offset, kv.first);
    std::sort(order.begin(), order.end());
    size_t n = 0;
    for (const auto& o : order)
        if (n < capacity) ids_out[n++] = o.second;
    return n;
}

int cndl_get_object(const cndl_ctx* ctx, uint32_t object_id, int32_t* node_offset, int32_t* node_count, int32_t* triangle_offset,
                    int32_t* vertex_offset) {
    if (!ctx) return CNDL_ERR_INVALID;
    auto it = ctx->objects.find(object_id);
    if (it == ctx->objects.end()) return CNDL_ERR_UNKNOWN_OBJECT;
    if (node_offset) *node_offset = it->second.node_offset;
    if (node_count) *node_count = it->second.node_count;
    if (triangle_offset) *triangle_offset = it->second.tri_offset;
    if (vertex_offset) *vertex_offset = it->second.vert_offset;
    return CNDL_OK;
}

int cndl_add_prebuilt_object(cndl_ctx* ctx, uint32_t object_id, const void* nodes, size_t N, const cndl_triangle* tris, size_t T,
                             const cndl_vertex* verts, size_t V) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!nodes || !tris || !verts || N == 0 || T == 0 || V == 0) return ctx->fail(CNDL_ERR_INVALID, "null or empty buffer");
    if (ctx->n_nodes + N > 0x7FFFFFF0ull || ctx->n_tris + T > (1ull << 27) || ctx->n_verts + V > 0x7FFFFFF0ull)
        return ctx->fail(CNDL_ERR_INVALID, "scene exceeds the leaf-pack limits (2^27 triangles, BVHConstructor.cpp:794)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->main_stream;
    const size_t ns = ctx->node_size;
    // one spare zeroed node: the reference's bounds check admits Pointer == NodeStart+NodeCount (…Stackless.glsl:196)
    CK(ctx->nodes.reserve((ctx->n_nodes + N + 1) * ns, st));
    CK(ctx->tris.reserve((ctx->n_tris + T) * sizeof(cndl_triangle), st));
    CK(ctx->verts.reserve((ctx->n_verts + V) * sizeof(cndl_vertex), st));
    char* dn = static_cast<char*>(ctx->nodes.p) + ctx->n_nodes * ns;
    char* dt = static_cast<char*>(ctx->tris.p) + ctx->n_tris * sizeof(cndl_triangle);
    char* dv = static_cast<char*>(ctx->verts.p) + ctx->n_verts * sizeof(cndl_vertex);
    CK(cudaMemcpyAsync(dn, nodes, N * ns, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dn + N * ns, 0, ns, st));
    CK(cudaMemcpyAsync(dt, tris, T * sizeof(cndl_triangle), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dv, verts, V * sizeof(cndl_vertex), cudaMemcpyHostToDevice, st));
    launch_rebase_triangles(reinterpret_cast<int4*>(dt), T, (int)ctx->n_verts, st, ctx->launches);  // Intersector.h:190-197
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ObjectData od;  // Intersector.h:179-181,:188
    od.node_offset = (int)ctx->n_nodes;
    od.tri_offset = (int)ctx->n_tris;
    od.vert_offset = (int)ctx->n_verts;
    od.node_count = (int)N;
    od.tri_count = (int)T;
    od.vert_count = (int)V;
    ctx->objects[object_id] = od;
    ctx->n_nodes += N;
    ctx->n_tris += T;
    ctx->n_verts += V;
    ctx->nodes.bytes = ctx->n_nodes * ns;
    ctx->tris.bytes = ctx->n_tris * sizeof(cndl_triangle);
    ctx->verts.bytes = ctx->n_verts * sizeof(cndl_vertex);
    ctx->committed = false;
    return CNDL_OK;
} CNDL_CATCH

int cndl_add_object(cndl_ctx* ctx, uint32_t object_id, const cndl_vertex* verts, size_t V, const uint32_t* indices, size_t I,
                    const int32_t* mesh_id_per_tri, const cndl_build_opts* opts) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!verts || !indices || V == 0 || I == 0 || I % 3 != 0) return ctx->fail(CNDL_ERR_INVALID, "null or empty geometry, or index count not a multiple of 3");
    const size_t T = I / 3;
    if (ctx->n_tris + T + (size_t)ctx->tri_offset_bias > (1ull << 27) || ctx->n_verts + V > 0x7FFFFFF0ull)
        return ctx->fail(CNDL_ERR_INVALID, "scene exceeds the leaf-pack limits (2^27 triangles, BVHConstructor.cpp:794)");
    cndl_build_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts) o = *opts;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->main_stream;
    const size_t ns = ctx->node_size;
    const size_t n_max = 2 * T - 1;  // upper bound of LastNodeIndex + 1 (every leaf holds >= 1 triangle)
    CK(ctx->nodes.reserve((ctx->n_nodes + n_max + 1) * ns, st));
    CK(ctx->tris.reserve((ctx->n_tris + T) * sizeof(cndl_triangle), st));
    CK(ctx->verts.reserve((ctx->n_verts + V) * sizeof(cndl_vertex), st));
    char* dn = static_cast<char*>(ctx->nodes.p) + ctx->n_nodes * ns;
    char* dt = static_cast<char*>(ctx->tris.p) + ctx->n_tris * sizeof(cndl_triangle);
    char* dv = static_cast<char*>(ctx->verts.p) + ctx->n_verts * sizeof(cndl_vertex);
    CK(cudaMemcpyAsync(dv, verts, V * sizeof(cndl_vertex), cudaMemcpyHostToDevice, st));
    BuildRequest rq;
    rq.format = ctx->format;
    rq.opts = o;
    rq.d_verts = reinterpret_cast<const float4*>(dv);
    rq.V = V;
    rq.h_indices = indices;
    rq.h_mesh_ids = mesh_id_per_tri;
    rq.T = T;
    rq.tri_offset = (int)ctx->n_tris + ctx->tri_offset_bias;
    rq.d_nodes_out = dn;
    rq.d_tris_out = reinterpret_cast<int4*>(dt);
    rq.n_nodes_out = 0;
    rq.arena = &ctx->build_arena;
    rq.arena_cap = &ctx->build_arena_cap;
    rq.host_counts = &ctx->build_host_counts;
    rq.side[0] = ctx->streams[1];
    rq.side[1] = ctx->streams[3];
    rq.split_node = (unsigned)ctx->knobs[CNDL_KNOB_BUILD_SPLIT_NODE];
    rq.pack_min = (unsigned)ctx->knobs[CNDL_KNOB_BUILD_PACK_MIN];
    std::string berr;
    float ms = 0.0f;
    const int rc = build_object(rq, st, ctx->launches, &ms, berr);
    if (rc != CNDL_OK) return ctx->fail(rc, berr);
    ctx->last_build_ms = ms;
    const size_t N = rq.n_nodes_out;
    CK(cudaMemsetAsync(dn + N * ns, 0, ns, st));
    launch_rebase_triangles(reinterpret_cast<int4*>(dt), T, (int)ctx->n_verts, st, ctx->launches);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ObjectData od;
    od.node_offset = (int)ctx->n_nodes;
    od.tri_offset = (int)ctx->n_tris;
    od.vert_offset = (int)ctx->n_verts;
    od.node_count = (int)N;
    od.tri_count = (int)T;
    od.vert_count = (int)V;
    ctx->objects[object_id] = od;
    ctx->n_nodes += N;
    ctx->n_tris += T;
    ctx->n_verts += V;
    ctx->nodes.bytes = ctx->n_nodes * ns;
    ctx->tris.bytes = ctx->n_tris * sizeof(cndl_triangle);
    ctx->verts.bytes = ctx->n_verts * sizeof(cndl_vertex);
    ctx->committed = false;
    return CNDL_OK;
} CNDL_CATCH

// BVH::BuildBVH as a free function keeps no state in the reference beyond file-scope statics (BVHConstructor.cpp:58-66); here the
// state worth keeping between calls is the CUDA side: streams, the build arena, the mapped level descriptors.  One hidden context
// per (device, node format) is created on first use and reused by later calls (serialised by a mutex; never destroyed: tearing CUDA
// objects down from a static destructor races the runtime's own shutdown).
int cndl_build_bvh(int node_format, int device, const cndl_vertex* verts, size_t V, const uint32_t* indices, size_t I,
                   const int32_t* mesh_id_per_tri, int32_t tri_offset, const cndl_build_opts* opts, void* nodes_out, size_t nodes_capacity,
                   size_t* n_nodes_out, cndl_triangle* tris_out, float* build_ms) try {
    if (!n_nodes_out || tri_offset < 0) return CNDL_ERR_INVALID;
    static std::mutex mu;
    static std::map<std::pair<int, int>, cndl_ctx*> cache;
    std::lock_guard<std::mutex> lock(mu);
    cndl_ctx*& ctx = cache[std::make_pair(device, node_format)];
    if (!ctx) {
        const int rc0 = cndl_create(&ctx, node_format, device);
        if (rc0 != CNDL_OK) { cache.erase(std::make_pair(device, node_format)); return rc0; }
    }
    // an empty scene again (the device buffers keep their capacity)
    ctx->n_nodes = ctx->n_tris = ctx->n_verts = 0;
    ctx->nodes.bytes = ctx->tris.bytes = ctx->verts.bytes = 0;
    ctx->objects.clear();
    ctx->committed = false;
    ctx->tri_offset_bias = tri_offset;
    int rc = cndl_add_object(ctx, 2, verts, V, indices, I, mesh_id_per_tri, opts);
    if (rc == CNDL_OK) {
        *n_nodes_out = ctx->n_nodes;
        if (build_ms) *build_ms = ctx->last_build_ms;
        if (nodes_out && nodes_capacity < ctx->n_nodes) rc = CNDL_ERR_INVALID;  // 2 * T - 1 always suffices
        else rc = cndl_read_buffers(ctx, nodes_out, tris_out, nullptr);
    }
    return rc;
} CNDL_CATCH

int cndl_commit(cndl_ctx* ctx, int clear_host) try {
    (void)clear_host;  // the host never keeps a copy: the device buffers are the only ones
    if (!ctx) return CNDL_ERR_INVALID;
    if (ctx->n_tris == 0) return ctx->fail(CNDL_ERR_INVALID, "nothing to commit");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->main_stream;
    CK(ctx->tri48.ensure_scratch(ctx->n_tris * 48));
    ctx->committed = false;
    ctx->hot_ready = false;
    ctx->nodes_valid = false;
    int* d_tri_flag = static_cast<int*>(ctx->d_counter.p) + 33;
    int tri_flag = 0;
    CK(cudaMemsetAsync(d_tri_flag, 0, sizeof(int), st));
    launch_make_tri48(static_cast<const int4*>(ctx->tris.p), static_cast<const float4*>(ctx->verts.p), ctx->n_tris, ctx->n_verts,
                      static_cast<float4*>(ctx->tri48.p), d_tri_flag, st, ctx->launches);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&tri_flag, d_tri_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (tri_flag) return ctx->fail(CNDL_ERR_INVALID, "a triangle names a vertex outside the vertex buffer (corrupt cache file or prebuilt buffer)");
    const char* kBadLeaf = "a leaf names triangles outside the scene (corrupt cache file or prebuilt buffer)";
    ctx->committed_nodes = ctx->n_nodes;  // m_NodeCountBuffered, Intersector.h:345
    ctx->committed_tris = ctx->n_tris;
    ctx->h_objects.clear();
    for (const auto& kv : ctx->objects) ctx->h_objects.push_back(make_int2(kv.second.node_offset, kv.second.node_count));
    std::sort(ctx->h_objects.begin(), ctx->h_objects.end(), [](const int2& a, const int2& b) { return a.x < b.x; });
    const int n_obj = (int)ctx->h_objects.size();
    if (ctx->format == CNDL_STACK) {
        int invalid = 0;
        CK(validate_stack_nodes(static_cast<const float4*>(ctx->nodes.p), ctx->h_objects.data(), n_obj, ctx->n_tris, static_cast<int*>(ctx->d_counter.p) + 32,
                                &invalid, st, ctx->launches));
        if (invalid & 2) return ctx->fail(CNDL_ERR_INVALID, kBadLeaf);
        ctx->nodes_valid = invalid == 0;
        ctx->committed = true;
        if (ctx->ents_buffered) {
            const int rc = upload_hot_entities(ctx);
            if (rc != CNDL_OK) return rc;
        }
    }
    if (ctx->format == CNDL_STACKLESS) {
        // derived hot-first node layout for the shared-memory staged kernel (kernels_hot.cu)
        const size_t N = ctx->n_nodes;
        CK(ctx->nodes2.ensure_scratch(N * sizeof(cndl_node)));
        CK(ctx->perm.ensure_scratch(N * sizeof(int)));
        CK(ctx->hot_scratch.ensure_scratch(hot_scratch_ints(N, n_obj) * sizeof(int)));
        CK(ctx->d_objects.ensure_scratch((size_t)n_obj * sizeof(int2)));
        CK(cudaMemcpyAsync(ctx->d_objects.p, ctx->h_objects.data(), (size_t)n_obj * sizeof(int2), cudaMemcpyHostToDevice, st));
        ctx->h_roots.assign((size_t)n_obj, -1);
        int invalid = 0;
        CK(derive_hot_layout(static_cast<const float4*>(ctx->nodes.p), N, static_cast<const int2*>(ctx->d_objects.p), ctx->h_objects.data(), n_obj,
                             ctx->n_tris, ctx->knobs[CNDL_KNOB_HOT_NODES], static_cast<float4*>(ctx->nodes2.p), static_cast<int*>(ctx->perm.p),
                             static_cast<int*>(ctx->hot_scratch.p), ctx->h_roots.data(), &ctx->n_hot, &invalid, st, ctx->launches));
        if (invalid & 2) return ctx->fail(CNDL_ERR_INVALID, kBadLeaf);
        ctx->hot_ready = invalid == 0;  // a buffer with out-of-range links keeps the reference-layout kernel and its range checks
        ctx->nodes_valid = ctx->hot_ready;
        ctx->committed = true;
        if (ctx->ents_buffered) {
            const int rc = upload_hot_entities(ctx);
            if (rc != CNDL_OK) return rc;
        }
    }
    return CNDL_OK;
} CNDL_CATCH

int cndl_read_buffers(cndl_ctx* ctx, void* nodes, cndl_triangle* tris, cndl_vertex* verts) try {
    if (!ctx) return CNDL_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (nodes && ctx->n_nodes) CK(cudaMemcpy(nodes, ctx->nodes.p, ctx->n_nodes * ctx->node_size, cudaMemcpyDeviceToHost));
    if (tris && ctx->n_tris) CK(cudaMemcpy(tris, ctx->tris.p, ctx->n_tris * sizeof(cndl_triangle), cudaMemcpyDeviceToHost));
    if (verts && ctx->n_verts) CK(cudaMemcpy(verts, ctx->verts.p, ctx->n_verts * sizeof(cndl_vertex), cudaMemcpyDeviceToHost));
    return CNDL_OK;
} CNDL_CATCH

int cndl_device_buffers(cndl_ctx* ctx, const void** nodes, const cndl_triangle** tris, const cndl_vertex** verts,
                        const cndl_entity** entities) {
    if (!ctx) return CNDL_ERR_INVALID;
    if (nodes) *nodes = ctx->nodes.p;
    if (tris) *tris = static_cast<const cndl_triangle*>(ctx->tris.p);
    if (verts) *verts = static_cast<const cndl_vertex*>(ctx->verts.p);
    if (entities) *entities = static_cast<const cndl_entity*>(ctx->ents.p);
    return CNDL_OK;
}

int cndl_push_entity(cndl_ctx* ctx, uint32_t object_id, const float model[16], float emissive, float translucency) try {
    if (!ctx || !model) return CNDL_ERR_INVALID;
    auto it = ctx->objects.find(object_id);
    if (it == ctx->objects.end())
        return ctx->fail(CNDL_ERR_UNKNOWN_OBJECT, "Trying to push entity whose parent object hasn't been added to global BVH");
    cndl_entity e;
    std::memset(&e, 0, sizeof(e));
    std::memcpy(e.model, model, 64);
    glm_inverse(model, e.inverse);
    e.node_offset = it->second.node_offset;
    e.node_count = it->second.node_count;
    const float alpha = 1.0f - translucency;
    std::memcpy(&e.data[0], &emissive, 4);  // Intersector.h:212
    std::memcpy(&e.data[1], &alpha, 4);     // Intersector.h:213
    ctx->staged.push_back(e);
    return CNDL_OK;
} CNDL_CATCH

int cndl_push_entity_records(cndl_ctx* ctx, const cndl_entity* records, size_t E) try {
    if (!ctx || (!records && E)) return CNDL_ERR_INVALID;
    ctx->staged.insert(ctx->staged.end(), records, records + E);
    return CNDL_OK;
} CNDL_CATCH

int cndl_buffer_entities(cndl_ctx* ctx) try {
    if (!ctx) return CNDL_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const size_t E = ctx->staged.size();
    CK(ctx->ents.ensure_scratch((E ? E : 1) * sizeof(cndl_entity)));
    if (E) CK(cudaMemcpy(ctx->ents.p, ctx->staged.data(), E * sizeof(cndl_entity), cudaMemcpyHostToDevice));
    ctx->n_ents = E;  // m_EntityPushed
    {   // world bounds of the scene: the eight corners of every entity's root box through its model matrix
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        std::unordered_map<int, std::array<float, 6>> root_box;
        for (const auto& e : ctx->staged) {
            if (e.node_offset < 0 || (size_t)e.node_offset >= ctx->n_nodes) continue;
            auto it = root_box.find(e.node_offset);
            if (it == root_box.end()) {
                float b[16];
                CK(cudaMemcpy(b, static_cast<const char*>(ctx->nodes.p) + (size_t)e.node_offset * ctx->node_size, ctx->node_size, cudaMemcpyDeviceToHost));
                std::array<float, 6> r;
                if (ctx->format == CNDL_STACKLESS) r = {b[0], b[1], b[2], b[4], b[5], b[6]};
                else r = {std::min(b[0], b[8]), std::min(b[1], b[9]), std::min(b[2], b[10]), std::max(b[4], b[12]), std::max(b[5], b[13]), std::max(b[6], b[14])};
                it = root_box.emplace(e.node_offset, r).first;
            }
            const auto& r = it->second;
            for (int c = 0; c < 8; ++c) {
                const float x = (c & 1) ? r[3] : r[0], y = (c & 2) ? r[4] : r[1], z = (c & 4) ? r[5] : r[2];
                for (int k = 0; k < 3; ++k) {
                    const float w = e.model[k] * x + e.model[4 + k] * y + e.model[8 + k] * z + e.model[12 + k];
                    if (w == w) { lo[k] = std::min(lo[k], w); hi[k] = std::max(hi[k], w); }
                }
            }
        }
        for (int k = 0; k < 3; ++k) { ctx->world_lo[k] = lo[k] <= hi[k] ? lo[k] : 0.0f; ctx->world_hi[k] = lo[k] <= hi[k] ? hi[k] : 0.0f; }
    }
    ctx->buffered.swap(ctx->staged);
    ctx->staged.clear();
    ctx->ents_buffered = true;
    return upload_hot_entities(ctx);
} CNDL_CATCH

int cndl_set_traversal_mode(cndl_ctx* ctx, int mode, int sort_rays) {
    if (!ctx || mode < 0 || mode > 2 || sort_rays < 0 || sort_rays > 4) return CNDL_ERR_INVALID;
    ctx->mode = mode;
    ctx->sort_rays = sort_rays;
    return CNDL_OK;
}

int cndl_set_tuning(cndl_ctx* ctx, int knob, int value) {
    if (!ctx || knob < 0 || knob >= 10 || value < 0) return CNDL_ERR_INVALID;
    if (knob == CNDL_KNOB_BLOCKS_PER_SM && (value < 1 || value > 16)) return CNDL_ERR_INVALID;
    ctx->knobs[knob] = value;
    return CNDL_OK;
}

int cndl_intersect_closest_device(cndl_ctx* ctx, const cndl_ray* d_rays, size_t R, int flags, cndl_hit* d_hits, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!d_rays || !d_hits)) return ctx->fail(CNDL_ERR_INVALID, "null ray or hit buffer");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const int kind = (flags & CNDL_IGNORE_TRANSPARENT) ? Q_CLOSEST_IGNORE_TRANSPARENT : Q_CLOSEST;
    if (effective_sort(ctx, R)) CK(ctx->d_order.ensure_scratch(order_region_ints(R) * sizeof(unsigned)));
    if (effective_sort(ctx, R) == 2) CK(ctx->d_sorted.ensure_scratch(R * sizeof(cndl_ray)));
    return enqueue_trace(ctx, kind, d_rays, R, nullptr, d_hits, nullptr, next_counter(ctx), static_cast<unsigned*>(ctx->d_order.p),
                         static_cast<cndl_ray*>(ctx->d_sorted.p), static_cast<cudaStream_t>(stream));
} CNDL_CATCH

int cndl_intersect_any_device(cndl_ctx* ctx, const cndl_ray* d_rays, size_t R, float* d_t_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!d_rays || !d_t_out)) return ctx->fail(CNDL_ERR_INVALID, "null ray or output buffer");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    if (effective_sort(ctx, R)) CK(ctx->d_order.ensure_scratch(order_region_ints(R) * sizeof(unsigned)));
    if (effective_sort(ctx, R) == 2) CK(ctx->d_sorted.ensure_scratch(R * sizeof(cndl_ray)));
    return enqueue_trace(ctx, Q_ANY, d_rays, R, nullptr, nullptr, d_t_out, next_counter(ctx), static_cast<unsigned*>(ctx->d_order.p),
                         static_cast<cndl_ray*>(ctx->d_sorted.p), static_cast<cudaStream_t>(stream));
} CNDL_CATCH

// Host-buffer queries: the batch is cut into chunks that rotate over three streams, so the
// host->device copy of chunk k+1, the traversal of chunk k and the device->host copy of chunk
// k-1 overlap (two copy engines + SMs) when the host buffers are pinned.
static int host_query(cndl_ctx* ctx, int kind, const cndl_ray* rays, size_t R, cndl_hit* hits, float* any_t) {
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    if (R == 0) return CNDL_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t out_elt = kind == Q_ANY ? sizeof(float) : sizeof(cndl_hit);
    CK(ctx->d_rays.ensure_scratch(R * sizeof(cndl_ray)));
    CK(ctx->d_hits.ensure_scratch(R * out_elt));
    // Three-stage pipeline over chunks: streams[0] carries every host->device copy back to back,
    // streams[1] / streams[3] the traversal kernels of even / odd chunks (so that the CTAs of the next
    // chunk fill the SMs while the persistent kernel of the previous one drains its last rays),
    // streams[2] every device->host copy; events chain the stages.
    const int n_chunks_want = ctx->knobs[CNDL_KNOB_HOST_CHUNKS] > 0 ? ctx->knobs[CNDL_KNOB_HOST_CHUNKS] : 12;
    size_t chunk = (R + n_chunks_want - 1) / n_chunks_want;
    if (chunk < (1u << 16)) chunk = 1u << 16;
    const size_t n_chunks = (R + chunk - 1) / chunk;
    while (ctx->events.size() < 2 * n_chunks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    CK(ctx->d_chunk_counters.ensure_scratch(n_chunks * 64));
    const int chunk_sort = effective_sort(ctx, chunk);
    if (chunk_sort) CK(ctx->d_order.ensure_scratch((order_region_ints(chunk) * n_chunks) * sizeof(unsigned)));
    if (chunk_sort == 2) CK(ctx->d_sorted.ensure_scratch(R * sizeof(cndl_ray)));
    size_t k = 0;
    for (size_t lo = 0; lo < R; lo += chunk, ++k) {
        const size_t n = R - lo < chunk ? R - lo : chunk;
        cndl_ray* dr = static_cast<cndl_ray*>(ctx->d_rays.p) + lo;
        char* dout = static_cast<char*>(ctx->d_hits.p) + lo * out_elt;
        CK(cudaMemcpyAsync(dr, rays + lo, n * sizeof(cndl_ray), cudaMemcpyHostToDevice, ctx->streams[0]));
        CK(cudaEventRecord(ctx->events[2 * k], ctx->streams[0]));
        cudaStream_t ks = ctx->streams[(k & 1) ? 3 : 1];
        CK(cudaStreamWaitEvent(ks, ctx->events[2 * k], 0));
        unsigned* counter = reinterpret_cast<unsigned*>(static_cast<char*>(ctx->d_chunk_counters.p) + 64 * k);
        rc = enqueue_trace(ctx, kind, dr, n, nullptr, kind == Q_ANY ? nullptr : reinterpret_cast<cndl_hit*>(dout),
                           kind == Q_ANY ? reinterpret_cast<float*>(dout) : nullptr, counter,
                           chunk_sort ? static_cast<unsigned*>(ctx->d_order.p) + order_region_ints(chunk) * k : nullptr,
                           chunk_sort == 2 ? static_cast<cndl_ray*>(ctx->d_sorted.p) + lo : nullptr, ks);
        if (rc != CNDL_OK) return rc;
        CK(cudaEventRecord(ctx->events[2 * k + 1], ks));
        CK(cudaStreamWaitEvent(ctx->streams[2], ctx->events[2 * k + 1], 0));
        char* hout = kind == Q_ANY ? reinterpret_cast<char*>(any_t + lo) : reinterpret_cast<char*>(hits + lo);
        CK(cudaMemcpyAsync(hout, dout, n * out_elt, cudaMemcpyDeviceToHost, ctx->streams[2]));
    }
    for (auto& s : ctx->streams) CK(cudaStreamSynchronize(s));
    return CNDL_OK;
}

int cndl_intersect_closest(cndl_ctx* ctx, const cndl_ray* rays, size_t R, int flags, cndl_hit* hits) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!rays || !hits)) return ctx->fail(CNDL_ERR_INVALID, "null ray or hit buffer");
    return host_query(ctx, (flags & CNDL_IGNORE_TRANSPARENT) ? Q_CLOSEST_IGNORE_TRANSPARENT : Q_CLOSEST, rays, R, hits, nullptr);
} CNDL_CATCH

int cndl_intersect_any(cndl_ctx* ctx, const cndl_ray* rays, size_t R, float* t_out) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!rays || !t_out)) return ctx->fail(CNDL_ERR_INVALID, "null ray or output buffer");
    return host_query(ctx, Q_ANY, rays, R, nullptr, t_out);
} CNDL_CATCH

int cndl_intersect_primary_device(cndl_ctx* ctx, const float inv_view[16], const float inv_proj[16], int W, int H, cndl_hit* d_hits,
                                  cndl_ray* d_rays_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!inv_view || !inv_proj || W <= 0 || H <= 0 || !d_hits) return ctx->fail(CNDL_ERR_INVALID, "bad primary-ray arguments");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t R = (size_t)W * (size_t)H;
    cndl_ray* dr = d_rays_out;
    if (!dr) {
        CK(ctx->d_rays.ensure_scratch(R * sizeof(cndl_ray)));
        dr = static_cast<cndl_ray*>(ctx->d_rays.p);
    }
    launch_primary_rays(inv_view, inv_proj, W, H, dr, st, ctx->launches);
    CK(cudaGetLastError());
    return enqueue_trace(ctx, Q_CLOSEST, dr, R, nullptr, d_hits, nullptr, next_counter(ctx), nullptr, nullptr, st);  // camera rays are coherent already
} CNDL_CATCH

int cndl_intersect_primary(cndl_ctx* ctx, const float inv_view[16], const float inv_proj[16], int W, int H, cndl_hit* hits,
                           cndl_ray* rays_out) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!hits || W <= 0 || H <= 0) return ctx->fail(CNDL_ERR_INVALID, "bad primary-ray arguments");
    const size_t R = (size_t)W * (size_t)H;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_rays.ensure_scratch(R * sizeof(cndl_ray)));
    CK(ctx->d_hits.ensure_scratch(R * sizeof(cndl_hit)));
    cudaStream_t st = ctx->main_stream;
    int rc = cndl_intersect_primary_device(ctx, inv_view, inv_proj, W, H, static_cast<cndl_hit*>(ctx->d_hits.p),
                                           static_cast<cndl_ray*>(ctx->d_rays.p), st);
    if (rc != CNDL_OK) return rc;
    CK(cudaMemcpyAsync(hits, ctx->d_hits.p, R * sizeof(cndl_hit), cudaMemcpyDeviceToHost, st));
    if (rays_out) CK(cudaMemcpyAsync(rays_out, ctx->d_rays.p, R * sizeof(cndl_ray), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CNDL_OK;
} CNDL_CATCH

int cndl_generate_rays_device(cndl_ctx* ctx, const cndl_raygen_params* params, const cndl_ray* d_rays, const cndl_hit* d_hits, size_t R,
                              cndl_ray* d_rays_out, uint32_t* d_parent_out, size_t* count_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!params || !count_out || params->spp < 1 || params->kind < CNDL_GEN_DIFFUSE || params->kind > CNDL_GEN_SHADOW ||
        (R && (!d_rays || !d_hits || !d_rays_out)))
        return ctx->fail(CNDL_ERR_INVALID, "bad ray-generation arguments");
    if (R * (size_t)params->spp > 0x7FFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "too many rays in one call");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_sort_tmp.ensure_scratch(generate_rays_scratch_ints(R, params->spp) * sizeof(int)));
    CK(generate_rays(scene_view(ctx), *params, d_rays, d_hits, R, nullptr, d_rays_out, d_parent_out, static_cast<int*>(ctx->d_sort_tmp.p), nullptr, count_out,
                     static_cast<cudaStream_t>(stream), ctx->launches));
    return CNDL_OK;
} CNDL_CATCH

int cndl_generate_bounce_rays_device(cndl_ctx* ctx, const cndl_ray* d_rays, const cndl_hit* d_hits, size_t R, int spp, float offset,
                                     float tmax, uint32_t seed, cndl_ray* d_rays_out, uint32_t* d_parent_out, size_t* count_out, void* stream) try {
    cndl_raygen_params p;
    std::memset(&p, 0, sizeof(p));
    p.kind = CNDL_GEN_DIFFUSE;
    p.spp = spp;
    p.seed = seed;
    p.offset = offset;
    p.tmax = tmax;
    return cndl_generate_rays_device(ctx, &p, d_rays, d_hits, R, d_rays_out, d_parent_out, count_out, stream);
} CNDL_CATCH

int cndl_generate_probe_rays_device(cndl_ctx* ctx, const float box_origin[3], const float size[3], const int32_t res[3], uint32_t seed,
                                    cndl_ray* d_rays_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!box_origin || !size || !res || !d_rays_out || res[0] <= 0 || res[1] <= 0 || res[2] <= 0 ||
        (size_t)res[0] * (size_t)res[1] * (size_t)res[2] > 0x7FFFFFF0ull)
        return ctx->fail(CNDL_ERR_INVALID, "bad probe-grid arguments");
    CK(cudaSetDevice(ctx->device));
    launch_probe_rays(box_origin, size, res, seed, d_rays_out, static_cast<cudaStream_t>(stream), ctx->launches);
    CK(cudaGetLastError());
    return CNDL_OK;
} CNDL_CATCH

int cndl_get_data_device(cndl_ctx* ctx, const cndl_hit* d_hits, size_t R, cndl_hit_attr* d_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!d_hits || !d_out)) return ctx->fail(CNDL_ERR_INVALID, "null hit or attribute buffer");
    if (R > 0xFFFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "more than 2^32-16 records in one call");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    launch_get_data(scene_view(ctx), static_cast<const float4*>(ctx->verts.p), d_hits, R, d_out, static_cast<cudaStream_t>(stream), ctx->launches);
    CK(cudaGetLastError());
    return CNDL_OK;
} CNDL_CATCH

int cndl_get_data(cndl_ctx* ctx, const cndl_hit* hits, size_t R, cndl_hit_attr* out) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!hits || !out)) return ctx->fail(CNDL_ERR_INVALID, "null hit or attribute buffer");
    if (R == 0) return check_ready(ctx);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_hits.ensure_scratch(R * sizeof(cndl_hit)));
    CK(ctx->d_rays.ensure_scratch(R * sizeof(cndl_hit_attr)));
    cudaStream_t st = ctx->main_stream;
    CK(cudaMemcpyAsync(ctx->d_hits.p, hits, R * sizeof(cndl_hit), cudaMemcpyHostToDevice, st));
    int rc = cndl_get_data_device(ctx, static_cast<const cndl_hit*>(ctx->d_hits.p), R, static_cast<cndl_hit_attr*>(ctx->d_rays.p), st);
    if (rc != CNDL_OK) return rc;
    CK(cudaMemcpyAsync(out, ctx->d_rays.p, R * sizeof(cndl_hit_attr), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CNDL_OK;
} CNDL_CATCH

int cndl_set_texture_references(cndl_ctx* ctx, const cndl_texture_reference* refs, size_t n) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (n && !refs) return ctx->fail(CNDL_ERR_INVALID, "null texture-reference table");
    if (n > 0x7FFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "more than 2^31-16 texture references");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->main_stream));  // a GetData launch may still read the old table
    ctx->n_tex_refs = 0;
    if (n) {
        CK(ctx->tex_refs.ensure_scratch(n * sizeof(cndl_texture_reference)));
        CK(cudaMemcpyAsync(ctx->tex_refs.p, refs, n * sizeof(cndl_texture_reference), cudaMemcpyHostToDevice, ctx->main_stream));
        CK(cudaStreamSynchronize(ctx->main_stream));
        ctx->n_tex_refs = n;
    }
    return CNDL_OK;
} CNDL_CATCH

size_t cndl_texture_reference_count(const cndl_ctx* ctx) { return ctx ? ctx->n_tex_refs : 0; }

int cndl_get_data_material_device(cndl_ctx* ctx, const cndl_hit* d_hits, size_t R, cndl_hit_material* d_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!d_hits || !d_out)) return ctx->fail(CNDL_ERR_INVALID, "null hit or material buffer");
    if (R > 0xFFFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "more than 2^32-16 records in one call");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    if (ctx->n_tex_refs == 0) return ctx->fail(CNDL_ERR_INVALID, "no texture-reference table: call cndl_set_texture_references first (GenerateMeshTextureReferences, Intersector.h:367)");
    CK(cudaSetDevice(ctx->device));
    launch_get_data_material(scene_view(ctx), static_cast<const float4*>(ctx->verts.p), static_cast<const cndl_texture_reference*>(ctx->tex_refs.p), ctx->n_tex_refs,
                             d_hits, R, d_out, static_cast<unsigned*>(ctx->d_counter.p) + 34, static_cast<cudaStream_t>(stream), ctx->launches);
    CK(cudaGetLastError());
    return CNDL_OK;
} CNDL_CATCH

int cndl_get_data_material(cndl_ctx* ctx, const cndl_hit* hits, size_t R, cndl_hit_material* out) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (R && (!hits || !out)) return ctx->fail(CNDL_ERR_INVALID, "null hit or material buffer");
    if (R == 0) return check_ready(ctx);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_hits.ensure_scratch(R * sizeof(cndl_hit)));
    CK(ctx->d_rays.ensure_scratch(R * sizeof(cndl_hit_material)));
    cudaStream_t st = ctx->main_stream;
    unsigned* flag = static_cast<unsigned*>(ctx->d_counter.p) + 34;
    CK(cudaMemsetAsync(flag, 0, sizeof(unsigned), st));
    CK(cudaMemcpyAsync(ctx->d_hits.p, hits, R * sizeof(cndl_hit), cudaMemcpyHostToDevice, st));
    int rc = cndl_get_data_material_device(ctx, static_cast<const cndl_hit*>(ctx->d_hits.p), R, static_cast<cndl_hit_material*>(ctx->d_rays.p), st);
    if (rc != CNDL_OK) return rc;
    unsigned bad = 0;
    CK(cudaMemcpyAsync(out, ctx->d_rays.p, R * sizeof(cndl_hit_material), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&bad, flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad) {
        char msg[160];
        std::snprintf(msg, sizeof(msg), "%u hit records name a mesh outside the texture-reference table of %zu entries", bad, ctx->n_tex_refs);
        return ctx->fail(CNDL_ERR_INVALID, msg);
    }
    return CNDL_OK;
} CNDL_CATCH

int cndl_collide_boxes_device(cndl_ctx* ctx, const cndl_box* d_boxes, size_t n, cndl_collision* d_out, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (ctx->format != CNDL_STACKLESS) return ctx->fail(CNDL_ERR_INVALID, "the collide query walks FlattenedNode buffers (Physics.h:15): stackless contexts only");
    if (n && (!d_boxes || !d_out)) return ctx->fail(CNDL_ERR_INVALID, "null box or result buffer");
    if (n > 0xFFFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "more than 2^32-16 boxes in one call");
    int rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    launch_collide_boxes(scene_view(ctx), static_cast<const float4*>(ctx->verts.p), d_boxes, n, d_out, static_cast<cudaStream_t>(stream), ctx->launches);
    CK(cudaGetLastError());
    return CNDL_OK;
} CNDL_CATCH

int cndl_collide_boxes(cndl_ctx* ctx, const cndl_box* boxes, size_t n, cndl_collision* out) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (n && (!boxes || !out)) return ctx->fail(CNDL_ERR_INVALID, "null box or result buffer");
    if (n == 0) return check_ready(ctx);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_rays.ensure_scratch(n * sizeof(cndl_box)));
    CK(ctx->d_hits.ensure_scratch(n * sizeof(cndl_collision)));
    cudaStream_t st = ctx->main_stream;
    CK(cudaMemcpyAsync(ctx->d_rays.p, boxes, n * sizeof(cndl_box), cudaMemcpyHostToDevice, st));
    int rc = cndl_collide_boxes_device(ctx, static_cast<const cndl_box*>(ctx->d_rays.p), n, static_cast<cndl_collision*>(ctx->d_hits.p), st);
    if (rc != CNDL_OK) return rc;
    CK(cudaMemcpyAsync(out, ctx->d_hits.p, n * sizeof(cndl_collision), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CNDL_OK;
} CNDL_CATCH

// ---- flat-buffer serialisation of the built scene (the reference rebuilds every BVH at every
// launch, Pipeline.cpp:1019-1028).  File: header, object table, then the reference-layout node / triangle / vertex
// buffers exactly as cndl_read_buffers returns them.
namespace {
struct FileHeader { char magic[8]; uint32_t version, format; uint64_t n_objects, n_nodes, n_tris, n_verts; };
struct FileObject { uint32_t id; int32_t node_offset, node_count, tri_offset, tri_count, vert_offset, vert_count; };
const char kMagic[8] = {'C', 'N', 'D', 'L', 'B', 'V', 'H', '1'};
}  // namespace

int cndl_save(cndl_ctx* ctx, const char* path) try {
    if (!ctx || !path) return CNDL_ERR_INVALID;
    if (ctx->n_tris == 0) return ctx->fail(CNDL_ERR_INVALID, "nothing to save");
    CK(cudaSetDevice(ctx->device));
    std::vector<char> nodes(ctx->n_nodes * ctx->node_size);
    std::vector<cndl_triangle> tris(ctx->n_tris);
    std::vector<cndl_vertex> verts(ctx->n_verts);
    int rc = cndl_read_buffers(ctx, nodes.data(), tris.data(), verts.data());
    if (rc != CNDL_OK) return rc;
    FileHeader h;
    std::memcpy(h.magic, kMagic, 8);
    h.version = 1;
    h.format = (uint32_t)ctx->format;
    h.n_objects = ctx->objects.size();
    h.n_nodes = ctx->n_nodes; h.n_tris = ctx->n_tris; h.n_verts = ctx->n_verts;
    std::vector<FileObject> table;
    for (const auto& kv : ctx->objects)
        table.push_back(FileObject{kv.first, kv.second.node_offset, kv.second.node_count, kv.second.tri_offset, kv.second.tri_count, kv.second.vert_offset,
                                   kv.second.vert_count});
    std::sort(table.begin(), table.end(), [](const FileObject& a, const FileObject& b) { return a.node_offset < b.node_offset; });
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return ctx->fail(CNDL_ERR_INVALID, std::string("cannot open ") + path + " for writing");
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(table.data(), sizeof(FileObject), table.size(), f) == table.size() &&
              std::fwrite(nodes.data(), 1, nodes.size(), f) == nodes.size() && std::fwrite(tris.data(), sizeof(cndl_triangle), tris.size(), f) == tris.size() &&
              std::fwrite(verts.data(), sizeof(cndl_vertex), verts.size(), f) == verts.size();
    ok = (std::fclose(f) == 0) && ok;
    return ok ? CNDL_OK : ctx->fail(CNDL_ERR_INVALID, std::string("short write to ") + path);
} CNDL_CATCH

int cndl_load(cndl_ctx* ctx, const char* path) try {
    if (!ctx || !path) return CNDL_ERR_INVALID;
    if (ctx->n_tris != 0) return ctx->fail(CNDL_ERR_INVALID, "cndl_load needs an empty context: leaf packs in the file hold global triangle offsets");
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return ctx->fail(CNDL_ERR_INVALID, std::string("cannot open ") + path);
    FileHeader h;
    auto bad = [&](const char* why) { std::fclose(f); return ctx->fail(CNDL_ERR_INVALID, std::string(path) + ": " + why); };
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) return bad("not a candela_b200 BVH file");
    if ((int)h.format != ctx->format) return bad("node format differs from the context's (stackless vs stack)");
    if (h.n_objects == 0 || h.n_objects > (1u << 24) || h.n_nodes == 0 || h.n_nodes > 0x7FFFFFF0ull || h.n_tris == 0 || h.n_tris > (1ull << 27) ||
        h.n_verts == 0 || h.n_verts > 0x7FFFFFF0ull)
        return bad("implausible sizes");
    std::vector<FileObject> table(h.n_objects);
    std::vector<char> nodes(h.n_nodes * ctx->node_size);
    std::vector<cndl_triangle> tris(h.n_tris);
    std::vector<cndl_vertex> verts(h.n_verts);
    if (std::fread(table.data(), sizeof(FileObject), table.size(), f) != table.size() || std::fread(nodes.data(), 1, nodes.size(), f) != nodes.size() ||
        std::fread(tris.data(), sizeof(cndl_triangle), tris.size(), f) != tris.size() || std::fread(verts.data(), sizeof(cndl_vertex), verts.size(), f) != verts.size())
        return bad("truncated file");
    std::fclose(f);
    for (const auto& o : table)
        if (o.node_offset < 0 || o.node_count <= 0 || (uint64_t)o.node_offset + (uint64_t)o.node_count > h.n_nodes || o.tri_offset < 0 || o.tri_count <= 0 ||
            (uint64_t)o.tri_offset + (uint64_t)o.tri_count > h.n_tris || o.vert_offset < 0 || o.vert_count <= 0 ||
            (uint64_t)o.vert_offset + (uint64_t)o.vert_count > h.n_verts)
            return ctx->fail(CNDL_ERR_INVALID, std::string(path) + ": object table out of range");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->main_stream;
    const size_t ns = ctx->node_size;
    CK(ctx->nodes.reserve((h.n_nodes + 1) * ns, st));
    CK(ctx->tris.reserve(h.n_tris * sizeof(cndl_triangle), st));
    CK(ctx->verts.reserve(h.n_verts * sizeof(cndl_vertex), st));
    CK(cudaMemcpyAsync(ctx->nodes.p, nodes.data(), h.n_nodes * ns, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(static_cast<char*>(ctx->nodes.p) + h.n_nodes * ns, 0, ns, st));
    CK(cudaMemcpyAsync(ctx->tris.p, tris.data(), h.n_tris * sizeof(cndl_triangle), cudaMemcpyHostToDevice, st));  // vertex indices are global already
    CK(cudaMemcpyAsync(ctx->verts.p, verts.data(), h.n_verts * sizeof(cndl_vertex), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    for (const auto& o : table) {
        ObjectData od;
        od.node_offset = o.node_offset; od.node_count = o.node_count; od.tri_offset = o.tri_offset; od.tri_count = o.tri_count;
        od.vert_offset = o.vert_offset; od.vert_count = o.vert_count;
        ctx->objects[o.id] = od;
    }
    ctx->n_nodes = h.n_nodes; ctx->n_tris = h.n_tris; ctx->n_verts = h.n_verts;
    ctx->nodes.bytes = h.n_nodes * ns;
    ctx->tris.bytes = h.n_tris * sizeof(cndl_triangle);
    ctx->verts.bytes = h.n_verts * sizeof(cndl_vertex);
    ctx->committed = false;
    return CNDL_OK;
} CNDL_CATCH

// Peer memory across processes (one process per GPU): a plain cudaMalloc allocation exported with cudaIpcGetMemHandle; another
// process on the node maps it with cudaIpcOpenMemHandle (peer access is enabled lazily) and its kernels store into it over NVLink.
int cndl_ipc_alloc(cndl_ctx* ctx, size_t bytes, void** d_ptr, cndl_ipc_handle* handle) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!d_ptr || !handle || bytes == 0) return ctx->fail(CNDL_ERR_INVALID, "null output or zero size");
    *d_ptr = nullptr;
    CK(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == sizeof(handle->bytes), "cudaIpcMemHandle_t is 64 bytes");
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return ctx->cuda_fail(e, "cudaIpcGetMemHandle");
    }
    std::memcpy(handle->bytes, &h, sizeof(h));
    *d_ptr = p;
    return CNDL_OK;
} CNDL_CATCH

int cndl_ipc_open(cndl_ctx* ctx, const cndl_ipc_handle* handle, void** d_ptr) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!d_ptr || !handle) return ctx->fail(CNDL_ERR_INVALID, "null handle or output");
    *d_ptr = nullptr;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle->bytes, sizeof(h));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return CNDL_OK;
} CNDL_CATCH

int cndl_ipc_close(cndl_ctx* ctx, void* d_ptr) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!d_ptr) return CNDL_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return CNDL_OK;
} CNDL_CATCH

int cndl_ipc_free(cndl_ctx* ctx, void* d_ptr) try {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!d_ptr) return CNDL_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaFree(d_ptr));
    return CNDL_OK;
} CNDL_CATCH

void* cndl_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

// Upload-only buffers (ray batches): write-combined pinned memory is not snooped by the CPU caches, which some hosts
// turn into a higher host-to-device rate (no difference on the B200 boxes measured); reading it back on the CPU is slow,
// so never use it for results.
void* cndl_host_alloc_write_combined(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocWriteCombined) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void cndl_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
