// Mode 2, third generation: the stackless while-while kernel of kernels_wavefront.cu with the TOP
// LEVELS OF THE TREE STAGED IN SHARED MEMORY.
//
// Why: ncu on the second-generation kernel shows the L1 data pipe at 70 % of its peak (one
// wavefront per distinct 32-byte node per lane; the 262k-triangle scene is L2-resident, DRAM is
// idle) and the issue slots at 59 %.  The nodes of depth <= 8 (511 of 315k) receive 53 % of all
// node visits of the diffuse batch, depth <= 10 (2047 nodes) 68 %.  Serving those from shared
// memory takes them off the L1 tag/data pipe and shortens their latency.
//
// How: the reference's pre-order layout scatters the top levels over the whole array (the left
// child is index+1, the right child is wherever the left subtree ends), so a pointer does not tell
// whether its node is hot.  At cndl_commit a second, DERIVED node array is written in which the
// breadth-first top of every object's tree comes first (indices 0 .. n_hot-1) and all other nodes
// follow in their original order.  A derived node is still 32 bytes:
//     lo = {Min.xyz, word}   word >= 0: the leaf pack (first_tri << 4 | count), unchanged
//                            word <  0: ~(index of the first child)        (was implicit: index+1)
//     hi = {Max.xyz, link}   the miss link as an absolute index into the derived array, -1 = stop
// The walk visits exactly the same nodes in exactly the same order with exactly the same box and
// triangle arithmetic, so every hit record (including `iters`) is bit-identical; only addresses
// change.  The reference-layout buffer stays what cndl_read_buffers / cndl_device_buffers expose.
//
// The derivation also validates the node buffer (links and leaf ranges inside the object), which
// is what lets the kernel drop the per-step range checks of …Stackless.glsl:196: on a validated
// buffer they can never fire.  Buffers that fail validation keep using the reference-layout kernel.
#include "kernels.cuh"
#include "scan.cuh"

namespace cndl {

namespace {

enum LaneState : int { EMPTY = 0, WALK = 1, LEAF = 2, DONE = 3 };

__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// ---------------------------------------------------------------------------------------------
// commit-time derivation

// Breadth-first list of the top nodes of every object, at most H entries.  One warp; a frontier of
// 32 nodes per iteration, children appended in order with a warp prefix sum (deterministic).
// perm[] must be -1 everywhere on entry; on exit perm[node] = position for every listed node.
__global__ void hot_bfs_kernel(const float4* __restrict__ nodes, const int2* __restrict__ objects, int n_objects, int H,
                               int* __restrict__ hot_node, int* __restrict__ hot_obj, int* __restrict__ perm, int* __restrict__ n_hot_out) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x;
    int tail = 0;
    for (int base = 0; base < n_objects && tail < H; base += 32) {
        const int o = base + lane;
        int2 ob = make_int2(0, 0);
        if (o < n_objects) ob = objects[o];
        const bool valid = o < n_objects && ob.y > 0;
        const unsigned m = __ballot_sync(FULL, valid);
        const int pos = tail + __popc(m & ((1u << lane) - 1u));
        if (valid && pos < H) {
            hot_node[pos] = ob.x;
            hot_obj[pos] = o;
            perm[ob.x] = pos;
        }
        tail = min(H, tail + __popc(m));
    }
    __syncwarp();
    int head = 0;
    while (head < tail && tail < H) {
        const int end = min(head + 32, tail);
        const int idx = head + lane;
        int c0 = -1, c1 = -1, obj = 0;
        if (idx < end) {
            const int n = hot_node[idx];
            obj = hot_obj[idx];
            const int st = objects[obj].x, cnt = objects[obj].y;
            if (__float_as_int(nodes[2 * (size_t)n].w) == -1) {  // inner: children are n+1 and the miss link of n+1
                const int l = n + 1;
                if (l < st + cnt) {
                    const int ll = __float_as_int(nodes[2 * (size_t)l + 1].w);
                    if (atomicCAS(&perm[l], -1, -2) == -1) c0 = l;
                    const int r = ll >= 0 ? ll + st : -1;
                    if (r >= st && r < st + cnt && atomicCAS(&perm[r], -1, -2) == -1) c1 = r;
                }
            }
        }
        const int nc = (c0 >= 0) + (c1 >= 0);
        int inc = nc;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        const int total = __shfl_sync(FULL, inc, 31);
        int pos = tail + inc - nc;
        if (c0 >= 0) {
            if (pos < H) { hot_node[pos] = c0; hot_obj[pos] = obj; perm[c0] = pos; } else perm[c0] = -1;
            ++pos;
        }
        if (c1 >= 0) {
            if (pos < H) { hot_node[pos] = c1; hot_obj[pos] = obj; perm[c1] = pos; } else perm[c1] = -1;
        }
        tail = min(H, tail + total);
        head = end;
        __syncwarp();
    }
    if (lane == 0) *n_hot_out = tail;
}

__global__ void cold_flags_kernel(const int* __restrict__ perm, int N, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flags[i] = perm[i] < 0 ? 1 : 0;
}

__global__ void cold_perm_kernel(int* __restrict__ perm, int N, const int* __restrict__ cold_rank, const int* __restrict__ n_hot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && perm[i] < 0) perm[i] = *n_hot + cold_rank[i];
}

// One object: writes its nodes in the derived format at their new places and validates them.
__global__ void derive_nodes_kernel(const float4* __restrict__ nodes, int start, int count, int n_tris, const int* __restrict__ perm,
                                    float4* __restrict__ nodes2, int* __restrict__ invalid) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = start + k;
    float4 lo = nodes[2 * (size_t)i], hi = nodes[2 * (size_t)i + 1];
    const int pack = __float_as_int(lo.w), link = __float_as_int(hi.w);
    bool bad = false;
    int word = pack, link2 = -1;
    if (pack == -1) {
        if (k + 1 >= count) { bad = true; word = 0; }  // an inner node needs a first child inside the object
        else word = ~perm[i + 1];
    } else {
        const int first = pack >> 4, len = pack & 0xF;
        if (pack < 0 || first + len > n_tris) { bad = true; word = 0; }
    }
    if (link >= 0) {
        if (link >= count) bad = true;
        else link2 = perm[start + link];
    }
    if (bad) atomicExch(invalid, 1);
    lo.w = __int_as_float(word);
    hi.w = __int_as_float(link2);
    const size_t j = (size_t)perm[i];
    nodes2[2 * j] = lo;
    nodes2[2 * j + 1] = hi;
}

__global__ void gather_kernel(const int* __restrict__ perm, const int2* __restrict__ objects, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = objects[i].y > 0 ? perm[objects[i].x] : -1;
}

// ---------------------------------------------------------------------------------------------
// traversal

__device__ __forceinline__ void store_hit(cndl_hit* __restrict__ hits, size_t i, float t, float u, float v, float w, int mesh, int tri, int ent, int iters) {
    float4* p = reinterpret_cast<float4*>(hits + i);
    p[0] = make_float4(t, u, v, w);
    reinterpret_cast<int4*>(p)[1] = make_int4(mesh, tri, ent, iters);
}

struct HLane {
    RayState r;          // object-space ray of the entity being traversed
    float tmax, closest;
    int ptr, iters, ent;
    int best_tri, best_ent;
    int pend_pack, pend_link;
    unsigned rid;
    int state;
};

// Scene loop bookkeeping (SL:290-301 / :333-337).  `ents` is the derived entity array: node_offset
// holds the root's index in the derived node array (-1: the entity's object is empty).
template <int KIND>
__device__ __forceinline__ void next_entity_hot(const SceneView& s, const cndl_entity* __restrict__ ents2, const cndl_ray* __restrict__ rays, HLane& L,
                                                int from) {
    int e = from;
    while (e < s.n_ents) {
        const cndl_entity* ent = ents2 + e;
        if (KIND == Q_CLOSEST_IGNORE_TRANSPARENT && __int_as_float(__ldg(&ent->data[1])) < 0.99f) { ++e; continue; }
        const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
        const float4 a = __ldg(rp), b = __ldg(rp + 1);
        L.r = to_object_space(ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
        L.ptr = __ldg(&ent->node_offset);
        L.iters = 0;
        L.ent = e;
        L.state = L.ptr >= 0 ? WALK : DONE;  // an entity without nodes is finished at once (SL:196 with iters = 0)
        return;
    }
    L.state = DONE;
    L.ent = 0x3FFFFFFF;
}

// One node visit of a walking lane (SL:192-246), branch-free apart from where the node lives.
__device__ __forceinline__ void node_step_hot(const float4* __restrict__ nodes2, const float4* s_lo, const float4* s_hi, int n_hot, HLane& L, bool warp_exact) {
    if (L.state == WALK) {
        if (L.iters >= 1024) {  // SL:192; the pointer range checks of SL:196 cannot fire on a validated buffer
            L.state = DONE;
        } else {
            ++L.iters;
            float4 mn, mx;
            if (L.ptr < n_hot) {
                mn = s_lo[L.ptr];
                mx = s_hi[L.ptr];
            } else {
                ldg256(nodes2 + 2 * (size_t)L.ptr, mn, mx);
            }
            const int word = __float_as_int(mn.w), link = __float_as_int(mx.w);
            const bool enter = enter_stackless(mn, mx, L.r, L.tmax, warp_exact);
            L.pend_pack = word;
            L.pend_link = link;
            L.ptr = enter ? ~word : link;  // a leaf's pointer is set again after its triangles
            L.state = enter ? (word >= 0 ? LEAF : WALK) : (link < 0 ? DONE : WALK);
        }
    }
}

template <int KIND, int BLOCK, int STEPS>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) trace_hot_stackless_kernel(SceneView s, const float4* __restrict__ nodes2,
                                                                                const cndl_entity* __restrict__ ents2, int n_hot,
                                                                                const cndl_ray* __restrict__ rays, unsigned R,
                                                                                const unsigned* __restrict__ order, cndl_hit* __restrict__ hits,
                                                                                float* __restrict__ any_t, unsigned* __restrict__ work_counter,
                                                                                int park_threshold, int idle_threshold) {
    constexpr bool ANY = KIND == Q_ANY;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    extern __shared__ float4 s_nodes[];
    float4* s_lo = s_nodes;
    float4* s_hi = s_nodes + n_hot;
    for (int i = threadIdx.x; i < n_hot; i += BLOCK) {
        s_lo[i] = __ldg(nodes2 + 2 * (size_t)i);
        s_hi[i] = __ldg(nodes2 + 2 * (size_t)i + 1);
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    HLane L;
    L.state = EMPTY;
    L.rid = 0;
    L.iters = 0;
    L.ent = 0;
    L.ptr = 0;
    L.best_tri = -1;
    L.best_ent = -1;
    L.closest = -1.0f;
    L.pend_pack = 0;
    L.pend_link = -1;
    bool drained = false;
    bool warp_exact = false;  // some lane's ray needs the literal GLSL min/max (warp-uniform: no divergence on it)

    while (true) {
        // ---------------- service: next entity / retire / refill ----------------
        {
            const unsigned b0 = __ballot_sync(FULL, (L.state & 1) != 0), b1 = __ballot_sync(FULL, (L.state & 2) != 0);
            const unsigned done = b0 & b1, empty = ~(b0 | b1), busy = b0 ^ b1;
            if (busy == 0u && done == 0u && drained) break;
            const int serviceable = __popc(done) + (drained ? 0 : __popc(empty));
            if (serviceable >= idle_threshold || busy == 0u) {
                if (L.state == DONE) {
                    const int last_ent = L.ent;
                    next_entity_hot<KIND>(s, ents2, rays, L, L.ent + 1);  // WALK again, or still DONE: the scene loop is over
                    if (L.state == DONE) {
                        if (ANY) {
                            any_t[L.rid] = L.closest;
                        } else {
                            // tail of IntersectScene (SL:300-318)
                            float t = -1.0f, u = -1.0f, v = -1.0f, w = -1.0f;
                            int mesh = -1;
                            if (L.best_tri >= 0) mesh = __ldg(&s.tris[L.best_tri]).w;
                            if (L.closest > 0.0f && L.best_tri > 0) {
                                RayState r = L.r;
                                if (L.best_ent != last_ent) {
                                    const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
                                    const float4 a = __ldg(rp), b = __ldg(rp + 1);
                                    r = to_object_space(ents2 + L.best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
                                }
                                const V3 p = {fadd(r.o.x, fmul(r.d.x, L.closest)), fadd(r.o.y, fmul(r.d.y, L.closest)), fadd(r.o.z, fmul(r.d.z, L.closest))};
                                t = L.closest;
                                barycentrics(s.tri48, L.best_tri, p, u, v, w);
                            }
                            store_hit(hits, L.rid, t, u, v, w, mesh, L.best_tri, L.best_ent, L.iters);
                        }
                        L.state = EMPTY;
                    }
                }
                if (!drained) {
                    const unsigned want = __ballot_sync(FULL, L.state == EMPTY);
                    const int n = __popc(want);
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(work_counter, (unsigned)n);
                    base = __shfl_sync(FULL, base, 0);
                    if (base + (unsigned)n >= R) drained = true;
                    if (L.state == EMPTY) {
                        const unsigned slot = base + (unsigned)__popc(want & ((1u << lane) - 1u));
                        if (slot < R) {
                            L.rid = order ? __ldg(order + slot) : slot;
                            L.closest = -1.0f;
                            L.best_tri = -1;
                            L.best_ent = -1;
                            L.iters = 0;
                            L.ent = 0;
                            if (ANY) {
                                const float rt = __ldg(&rays[L.rid].tmax);
                                L.tmax = rt > 0.0f ? rt : 1000000.0f;
                            } else {
                                L.tmax = 1000000.0f;
                            }
                            next_entity_hot<KIND>(s, ents2, rays, L, 0);
                        }
                    }
                }
                warp_exact = __any_sync(FULL, (L.state == WALK || L.state == LEAF) && L.r.nan_path);
            }
        }

        // ---------------- node phase ----------------
        // runs until `park` of the lanes that were walking at its start have parked (at a leaf or at the
        // end of an entity): a quarter of the walkers, at most park_threshold
        {
            const int walk0 = __popc(__ballot_sync(FULL, L.state == WALK));
            if (walk0 > 0) {
                int park = walk0 >> 2;
                park = park < 1 ? 1 : (park > park_threshold ? park_threshold : park);
                const int min_walk = walk0 - park + 1;  // >= 1
                do {
#pragma unroll
                    for (int step = 0; step < STEPS; ++step) node_step_hot(nodes2, s_lo, s_hi, n_hot, L, warp_exact);
                } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
            }
        }

        // ---------------- leaf phase ----------------
        if (L.state == LEAF) {
            EntityResult er{-1.0f, -1, 0};
            const bool found = leaf_triangles<ANY>(s, L.pend_pack, L.r, L.tmax, er);
            if (er.tri >= 0) { L.closest = er.t; L.best_tri = er.tri; L.best_ent = L.ent; }
            L.ptr = L.pend_link;
            L.state = L.pend_link < 0 ? DONE : WALK;
            if (ANY && found) {  // SL:567-569: the scene loop returns the first T > 0
                L.state = DONE;
                L.ent = 0x3FFFFFFF;
            }
        }
    }
}

template <int KIND, int BLOCK, int STEPS>
void launch_hot_one(unsigned grid, size_t smem, cudaStream_t stream, const SceneView& s, const HotView& hv, const cndl_ray* rays, unsigned R,
                    const unsigned* order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
    auto k = trace_hot_stackless_kernel<KIND, BLOCK, STEPS>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        // everything the staged nodes do not need stays L1: carve out just enough for the resident CTAs
        const size_t per_sm = smem * (1024 / BLOCK) + 1024 * (1024 / BLOCK);
        const size_t pct = (per_sm * 100 + 228 * 1024 - 1) / (228 * 1024);
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)(pct > 100 ? 100 : pct));
        configured = smem;
    }
    k<<<grid, BLOCK, smem, stream>>>(s, hv.nodes2, hv.ents2, (int)(smem / 32), rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold);
}

template <int KIND, int BLOCK>
void launch_hot_steps(int steps, unsigned grid, size_t smem, cudaStream_t stream, const SceneView& s, const HotView& hv, const cndl_ray* rays,
                      unsigned R, const unsigned* order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold,
                      int idle_threshold) {
#define CNDL_HOT_ARGS grid, smem, stream, s, hv, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold
    switch (steps) {
        case 1: launch_hot_one<KIND, BLOCK, 1>(CNDL_HOT_ARGS); break;
        case 3: launch_hot_one<KIND, BLOCK, 3>(CNDL_HOT_ARGS); break;
        default: launch_hot_one<KIND, BLOCK, 2>(CNDL_HOT_ARGS); break;
    }
#undef CNDL_HOT_ARGS
}

template <int BLOCK>
void launch_hot_kind(int kind, int steps, unsigned grid, size_t smem, cudaStream_t stream, const SceneView& s, const HotView& hv,
                     const cndl_ray* rays, unsigned R, const unsigned* order, cndl_hit* hits, float* any_t, unsigned* work_counter,
                     int park_threshold, int idle_threshold) {
#define CNDL_HOT_ARGS steps, grid, smem, stream, s, hv, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold
    switch (kind) {
        case Q_CLOSEST: launch_hot_steps<Q_CLOSEST, BLOCK>(CNDL_HOT_ARGS); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: launch_hot_steps<Q_CLOSEST_IGNORE_TRANSPARENT, BLOCK>(CNDL_HOT_ARGS); break;
        default: launch_hot_steps<Q_ANY, BLOCK>(CNDL_HOT_ARGS); break;
    }
#undef CNDL_HOT_ARGS
}

}  // namespace

cudaError_t derive_hot_layout(const float4* nodes, size_t N, const int2* d_objects, const int2* h_objects, int n_objects, size_t n_tris, int H,
                              float4* nodes2, int* perm, int* scratch, int* h_roots_out, int* h_n_hot, int* h_invalid, cudaStream_t st,
                              LaunchCounter& lc) {
    // scratch: [0] n_hot, [1] invalid, [2] scan total, 8192 hot nodes, 8192 their objects, then N flags, N ranks, block sums, n_objects roots
    int* d_n_hot = scratch;
    int* d_invalid = scratch + 1;
    int* d_total = scratch + 2;
    int* d_hot = scratch + 16;
    int* d_hot_obj = d_hot + 8192;
    int* d_flags = d_hot_obj + 8192;
    int* d_rank = d_flags + N;
    int* d_bsums = d_rank + N;
    int* d_roots = d_bsums + (N / 2048 + 8);
    if (H > kMaxHotNodes) H = kMaxHotNodes;
    if (H < 1) H = 1;
    cudaMemsetAsync(perm, 0xFF, N * sizeof(int), st);
    cudaMemsetAsync(scratch, 0, 16 * sizeof(int), st);
    hot_bfs_kernel<<<1, 32, 0, st>>>(nodes, d_objects, n_objects, H, d_hot, d_hot_obj, perm, d_n_hot);
    const unsigned nb = (unsigned)((N + 255) / 256);
    cold_flags_kernel<<<nb, 256, 0, st>>>(perm, (int)N, d_flags);
    lc.n += 2;
    exclusive_scan(d_flags, (int)N, d_rank, d_bsums, d_total, st, lc);
    cold_perm_kernel<<<nb, 256, 0, st>>>(perm, (int)N, d_rank, d_n_hot);
    lc.n++;
    for (int o = 0; o < n_objects; ++o) {
        if (h_objects[o].y <= 0) continue;
        derive_nodes_kernel<<<(unsigned)((h_objects[o].y + 255) / 256), 256, 0, st>>>(nodes, h_objects[o].x, h_objects[o].y, (int)n_tris, perm, nodes2,
                                                                                     d_invalid);
        lc.n++;
    }
    gather_kernel<<<(unsigned)((n_objects + 255) / 256), 256, 0, st>>>(perm, d_objects, n_objects, d_roots);
    lc.n++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(h_n_hot, d_n_hot, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_invalid, d_invalid, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_roots_out, d_roots, n_objects * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}

size_t hot_scratch_ints(size_t N, int n_objects) { return 16 + 2 * 8192 + 2 * N + (N / 2048 + 8) + (size_t)n_objects + 16; }

void launch_trace_hot(const SceneView& s, const HotView& hv, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits,
                      float* any_t, unsigned* work_counter, int sm_count, int block_threads, int park_threshold, int idle_threshold, int steps,
                      cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    const int block = block_threads >= 1024 ? 1024 : (block_threads >= 512 ? 512 : 256);
    unsigned grid = (unsigned)(sm_count * (1024 / block));
    const unsigned need = (unsigned)((R + block - 1) / block);
    if (grid > need) grid = need;
    // any prefix of the hot list may be staged: keep all resident CTAs of an SM within its shared memory
    size_t n_stage = (size_t)hv.n_hot;
    const size_t per_cta = (size_t)220 * 1024 / (size_t)(1024 / block) / 32;
    if (n_stage > per_cta) n_stage = per_cta;
    const size_t smem = n_stage * 32;
#define CNDL_HOT_ARGS kind, steps, grid, smem, stream, s, hv, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold
    if (block == 1024) launch_hot_kind<1024>(CNDL_HOT_ARGS);
    else if (block == 512) launch_hot_kind<512>(CNDL_HOT_ARGS);
    else launch_hot_kind<256>(CNDL_HOT_ARGS);
#undef CNDL_HOT_ARGS
    lc.n++;
}

}  // namespace cndl
