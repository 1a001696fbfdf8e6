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
    // invalid bit 0: a link or first child leaves the object (the reference-layout kernels with their range checks handle that like
    // the shader does); bit 1: a leaf names triangles outside the scene (nothing can traverse that: cndl_commit fails)
    int bad = 0;
    int word = pack, link2 = -1;
    if (pack == -1) {
        if (k + 1 >= count) { bad |= 1; word = 0; }  // an inner node needs a first child inside the object
        else word = ~perm[i + 1];
    } else {
        const int first = pack >> 4, len = pack & 0xF;
        if (pack < 0 || first + len > n_tris) { bad |= 2; word = 0; }
    }
    if (link >= 0) {
        if (link >= count) bad |= 1;
        else link2 = perm[start + link];
    }
    if (bad) atomicOr(invalid, bad);
    lo.w = __int_as_float(word);
    hi.w = __int_as_float(link2);
    const size_t j = (size_t)perm[i];
    nodes2[2 * j] = lo;
    nodes2[2 * j + 1] = hi;
}

// Stack format: every inner child's slot and every leaf child's triangle range must lie inside the object / the scene.
__global__ void validate_stack_kernel(const float4* __restrict__ nodes, int start, int count, int n_tris, int* __restrict__ invalid) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int bad = 0;  // bit 0: child slot outside the object; bit 1: leaf range outside the scene (see derive_nodes_kernel)
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const float4 mn = nodes[4 * (size_t)(start + k) + 2 * side], mx = nodes[4 * (size_t)(start + k) + 2 * side + 1];
        const int pack = __float_as_int(mn.w), slot = __float_as_int(mx.w);
        if (pack == -1) bad |= (slot < 0 || slot >= count) ? 1 : 0;
        else bad |= (pack < 0 || (pack >> 4) + (pack & 0xF) > n_tris) ? 2 : 0;
    }
    if (bad) atomicOr(invalid, bad);
}

__global__ void gather_kernel(const int* __restrict__ perm, const int2* __restrict__ objects, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = objects[i].y > 0 ? perm[objects[i].x] : -1;
}

// ---------------------------------------------------------------------------------------------
// traversal

__device__ __forceinline__ void store_hit(cndl_hit* __restrict__ hits, size_t i, float t, float u, float v, float w, int mesh, int tri, int ent, int iters) {
    stg256(reinterpret_cast<float4*>(hits + i), make_float4(t, u, v, w),
           make_float4(__int_as_float(mesh), __int_as_float(tri), __int_as_float(ent), __int_as_float(iters)));  // one 256-bit store
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
        float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
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

// One node visit of a walking lane (SL:192-246), branch-free apart from where the node lives.  The pointer is
// always valid on a validated buffer, so a lane at the iteration cap (SL:192) loads its node anyway and the visit is
// discarded by a select; the iteration counter runs on and is clamped to 1024 when the ray retires.
template <bool EXACT>
__device__ __forceinline__ void node_step_hot(const float4* __restrict__ nodes2, const float4* s_lo, const float4* s_hi, int n_hot, HLane& L) {
    if (L.state == WALK) {
        const bool capped = L.iters >= 1024;
        float4 mn, mx;
        if (L.ptr < n_hot) {
            mn = s_lo[L.ptr];
            mx = s_hi[L.ptr];
        } else {
            ldg256(nodes2 + 2 * (size_t)L.ptr, mn, mx);
        }
        const int word = __float_as_int(mn.w), link = __float_as_int(mx.w);
        const bool enter = enter_stackless(mn, mx, L.r, L.tmax, EXACT);
        L.pend_pack = word;
        L.pend_link = link;
        L.ptr = enter ? ~word : link;  // a leaf's pointer is set again after its triangles
        const int next = enter ? (word >= 0 ? LEAF : WALK) : (link < 0 ? DONE : WALK);
        L.state = capped ? DONE : next;
        ++L.iters;
    }
}

template <int KIND, int BLOCK, int STEPS>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) trace_hot_stackless_kernel(SceneView s, const float4* __restrict__ nodes2,
                                                                                const cndl_entity* __restrict__ ents2, int n_hot,
                                                                                const cndl_ray* __restrict__ rays, unsigned R,
                                                                                RayOrder order, cndl_hit* __restrict__ hits,
                                                                                float* __restrict__ any_t, unsigned* __restrict__ work_counter,
                                                                                int park_threshold, int idle_threshold) {
    R = batch_length(order, R);
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
                            any_t[out_slot(order, L.rid)] = L.closest;
                        } else {
                            // tail of IntersectScene (SL:300-318)
                            float t = -1.0f, u = -1.0f, v = -1.0f, w = -1.0f;
                            int mesh = -1;
                            if (L.best_tri >= 0) mesh = __ldg(&s.tris[L.best_tri]).w;
                            if (L.closest > 0.0f && L.best_tri > 0) {
                                RayState r = L.r;
                                if (L.best_ent != last_ent) {
                                    const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
                                    float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
                                    r = to_object_space(ents2 + L.best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
                                }
                                const V3 p = {fadd(r.o.x, fmul(r.d.x, L.closest)), fadd(r.o.y, fmul(r.d.y, L.closest)), fadd(r.o.z, fmul(r.d.z, L.closest))};
                                t = L.closest;
                                barycentrics(s.tri48, L.best_tri, p, u, v, w);
                            }
                            store_hit(hits, out_slot(order, L.rid), t, u, v, w, mesh, L.best_tri, L.best_ent, min(L.iters, 1024));
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
                        if (slot_live(order, slot, R)) {
                            L.rid = ray_of_slot(order, slot);
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
                if (!warp_exact) {
                    do {
#pragma unroll
                        for (int step = 0; step < STEPS; ++step) node_step_hot<false>(nodes2, s_lo, s_hi, n_hot, L);
                    } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
                } else {
                    do {
                        node_step_hot<true>(nodes2, s_lo, s_hi, n_hot, L);
                    } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
                }
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


// ---------------------------------------------------------------------------------------------
// TWO RAYS PER LANE.  The second-generation kernel is latency bound (ncu: 50 % of the warp stall
// samples are long-scoreboard waits on the node load, issue slots 59 % busy, 32 warps per SM at
// 62 registers).  Here every lane walks two independent rays over the derived layout:
//   * a node step issues both rays' node loads back to back and then both box tests, in one
//     straight-line block (idle slots load node 0 and discard it), so twice as many loads are in
//     flight per warp and the two dependency chains fill each other's fixed-latency gaps;
//   * the long phases (triangles, next entity / retire / refill) run once per outer iteration on a
//     MERGED row: each lane contributes whichever of its two slots needs the phase, so sparse
//     per-slot work packs into fewer, fuller passes.
// The per-ray operation sequence is unchanged: results stay bit-identical.
struct PSlot {
    V3 o, d, inv;  // object-space ray of the entity being traversed
    float tmax;
    int ptr, iters, ent, best_tri, best_ent, pend_pack, pend_link;
    unsigned rid;
    int state;
};

__device__ __forceinline__ bool slot_needs_exact(const PSlot& S) {
    return (S.state == WALK || S.state == LEAF) && !(finite3(S.o) && finite3(S.inv));
}

template <int KIND>
__device__ __forceinline__ void pair_next_entity(const SceneView& s, const cndl_entity* __restrict__ ents2, const cndl_ray* __restrict__ rays, PSlot& S,
                                                 int from) {
    int e = from;
    while (e < s.n_ents) {
        const cndl_entity* ent = ents2 + e;
        if (KIND == Q_CLOSEST_IGNORE_TRANSPARENT && __int_as_float(__ldg(&ent->data[1])) < 0.99f) { ++e; continue; }
        const float4* rp = reinterpret_cast<const float4*>(rays + S.rid);
        float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
        const RayState r = to_object_space(ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
        S.o = r.o; S.d = r.d; S.inv = r.inv;
        S.ptr = __ldg(&ent->node_offset);
        S.iters = 0;
        S.ent = e;
        S.state = S.ptr >= 0 ? WALK : DONE;
        return;
    }
    S.state = DONE;
    S.ent = 0x3FFFFFFF;
}

template <bool EXACT>
__device__ __forceinline__ void pair_step(const float4* __restrict__ nodes2, PSlot& A, PSlot& B) {
    const bool actA = A.state == WALK && A.iters < 1024, actB = B.state == WALK && B.iters < 1024;  // SL:192
    float4 amn, amx, bmn, bmx;
    ldg256(nodes2 + 2 * (size_t)(actA ? A.ptr : 0), amn, amx);
    ldg256(nodes2 + 2 * (size_t)(actB ? B.ptr : 0), bmn, bmx);
    {
        const int word = __float_as_int(amn.w), link = __float_as_int(amx.w);
        const RayState r{A.o, A.d, A.inv, false};
        const bool enter = enter_stackless(amn, amx, r, A.tmax, EXACT);
        const int nstate = enter ? (word >= 0 ? LEAF : WALK) : (link < 0 ? DONE : WALK);
        const int nptr = enter ? ~word : link;
        A.state = actA ? nstate : (A.state == WALK ? DONE : A.state);  // a walker that is not active hit the iteration cap
        A.ptr = actA ? nptr : A.ptr;
        A.pend_pack = actA ? word : A.pend_pack;
        A.pend_link = actA ? link : A.pend_link;
        A.iters += actA ? 1 : 0;
    }
    {
        const int word = __float_as_int(bmn.w), link = __float_as_int(bmx.w);
        const RayState r{B.o, B.d, B.inv, false};
        const bool enter = enter_stackless(bmn, bmx, r, B.tmax, EXACT);
        const int nstate = enter ? (word >= 0 ? LEAF : WALK) : (link < 0 ? DONE : WALK);
        const int nptr = enter ? ~word : link;
        B.state = actB ? nstate : (B.state == WALK ? DONE : B.state);
        B.ptr = actB ? nptr : B.ptr;
        B.pend_pack = actB ? word : B.pend_pack;
        B.pend_link = actB ? link : B.pend_link;
        B.iters += actB ? 1 : 0;
    }
}

template <int KIND, int MINB, int STEPS>
__global__ void __launch_bounds__(128, MINB) trace_pair_stackless_kernel(SceneView s, const float4* __restrict__ nodes2,
                                                                      const cndl_entity* __restrict__ ents2, const cndl_ray* __restrict__ rays,
                                                                      unsigned R, RayOrder order, cndl_hit* __restrict__ hits,
                                                                      float* __restrict__ any_t, unsigned* __restrict__ work_counter,
                                                                      int park_threshold, int idle_threshold) {
    R = batch_length(order, R);
    constexpr bool ANY = KIND == Q_ANY;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    PSlot A, B;
    A.state = B.state = EMPTY;
    A.rid = B.rid = 0;
    A.iters = B.iters = 0;
    A.ent = B.ent = 0;
    A.ptr = B.ptr = 0;
    A.best_tri = B.best_tri = -1;
    A.best_ent = B.best_ent = -1;
    A.pend_pack = B.pend_pack = 0;
    A.pend_link = B.pend_link = -1;
    A.tmax = B.tmax = 0.0f;
    A.o = A.d = A.inv = B.o = B.d = B.inv = V3{0.0f, 0.0f, 0.0f};
    bool drained = false;
    bool warp_exact = false;

    while (true) {
        // ---------------- service on the merged row: next entity / retire / refill ----------------
        {
            const unsigned doneA = __ballot_sync(FULL, A.state == DONE), doneB = __ballot_sync(FULL, B.state == DONE);
            const unsigned emptyA = __ballot_sync(FULL, A.state == EMPTY), emptyB = __ballot_sync(FULL, B.state == EMPTY);
            const int busy = 64 - __popc(doneA) - __popc(doneB) - __popc(emptyA) - __popc(emptyB);
            const int n_done = __popc(doneA) + __popc(doneB);
            if (busy == 0 && n_done == 0 && drained) break;
            const int serviceable = n_done + (drained ? 0 : __popc(emptyA) + __popc(emptyB));
            if (serviceable >= idle_threshold || busy == 0) {
                // each lane services one slot: A if it needs it, else B
                const bool needA = A.state == DONE || (A.state == EMPTY && !drained);
                const bool needB = B.state == DONE || (B.state == EMPTY && !drained);
                const int sel = needA ? 0 : (needB ? 1 : -1);
                PSlot T = sel == 1 ? B : A;
                if (sel < 0) T.state = WALK;  // neither: take no part below
                if (T.state == DONE) {
                    const int last_ent = T.ent;
                    pair_next_entity<KIND>(s, ents2, rays, T, T.ent + 1);
                    if (T.state == DONE) {
                        if (ANY) {
                            any_t[out_slot(order, T.rid)] = T.best_tri >= 0 ? T.tmax : -1.0f;
                        } else {
                            // tail of IntersectScene (SL:300-318); TMax == ClosestT once something was accepted
                            float t = -1.0f, u = -1.0f, v = -1.0f, w = -1.0f;
                            int mesh = -1;
                            if (T.best_tri >= 0) mesh = __ldg(&s.tris[T.best_tri]).w;
                            if (T.best_tri > 0) {
                                V3 ro = T.o, rd = T.d;
                                if (T.best_ent != last_ent) {
                                    const float4* rp = reinterpret_cast<const float4*>(rays + T.rid);
                                    float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
                                    const RayState r = to_object_space(ents2 + T.best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
                                    ro = r.o; rd = r.d;
                                }
                                t = T.tmax;
                                const V3 p = {fadd(ro.x, fmul(rd.x, t)), fadd(ro.y, fmul(rd.y, t)), fadd(ro.z, fmul(rd.z, t))};
                                barycentrics(s.tri48, T.best_tri, p, u, v, w);
                            }
                            store_hit(hits, out_slot(order, T.rid), t, u, v, w, mesh, T.best_tri, T.best_ent, T.iters);
                        }
                        T.state = EMPTY;
                    }
                }
                if (!drained) {
                    const unsigned want = __ballot_sync(FULL, sel >= 0 && T.state == EMPTY);
                    const int n = __popc(want);
                    unsigned base = 0;
                    if (lane == 0 && n > 0) base = atomicAdd(work_counter, (unsigned)n);
                    base = __shfl_sync(FULL, base, 0);
                    if (base + (unsigned)n >= R) drained = true;
                    if (sel >= 0 && T.state == EMPTY) {
                        const unsigned slot = base + (unsigned)__popc(want & ((1u << lane) - 1u));
                        if (slot_live(order, slot, R)) {
                            T.rid = ray_of_slot(order, slot);
                            T.best_tri = -1;
                            T.best_ent = -1;
                            T.iters = 0;
                            T.ent = 0;
                            if (ANY) {
                                const float rt = __ldg(&rays[T.rid].tmax);
                                T.tmax = rt > 0.0f ? rt : 1000000.0f;
                            } else {
                                T.tmax = 1000000.0f;
                            }
                            pair_next_entity<KIND>(s, ents2, rays, T, 0);
                        }
                    }
                }
                if (sel == 0) A = T;
                if (sel == 1) B = T;
                warp_exact = __any_sync(FULL, slot_needs_exact(A) || slot_needs_exact(B));
            }
        }

        // ---------------- node phase ----------------
        {
            const int walk0 = __popc(__ballot_sync(FULL, A.state == WALK)) + __popc(__ballot_sync(FULL, B.state == WALK));
            if (walk0 > 0) {
                int park = walk0 >> 2;
                park = park < 1 ? 1 : (park > park_threshold ? park_threshold : park);
                const int min_walk = walk0 - park + 1;  // >= 1
                if (!warp_exact) {
                    do {
#pragma unroll
                        for (int step = 0; step < STEPS; ++step) pair_step<false>(nodes2, A, B);
                    } while (__popc(__ballot_sync(FULL, A.state == WALK)) + __popc(__ballot_sync(FULL, B.state == WALK)) >= min_walk);
                } else {
                    do {
                        pair_step<true>(nodes2, A, B);
                    } while (__popc(__ballot_sync(FULL, A.state == WALK)) + __popc(__ballot_sync(FULL, B.state == WALK)) >= min_walk);
                }
            }
        }

        // ---------------- leaf phase on the merged row ----------------
        {
            const int sel = A.state == LEAF ? 0 : (B.state == LEAF ? 1 : -1);
            if (sel >= 0) {
                const bool b = sel == 1;
                const RayState r{b ? B.o : A.o, b ? B.d : A.d, b ? B.inv : A.inv, false};
                float tmax = b ? B.tmax : A.tmax;
                const int pack = b ? B.pend_pack : A.pend_pack, link = b ? B.pend_link : A.pend_link;
                EntityResult er{-1.0f, -1, 0};
                const bool found = leaf_triangles<ANY>(s, pack, r, tmax, er);
                int nstate = link < 0 ? DONE : WALK;
                int nent = b ? B.ent : A.ent;
                const int cur_ent = nent;
                if (ANY && found) {  // SL:567-569: the scene loop returns the first T > 0
                    nstate = DONE;
                    nent = 0x3FFFFFFF;
                }
                if (!b) {
                    A.tmax = tmax; A.ptr = link; A.state = nstate; A.ent = nent;
                    if (er.tri >= 0) { A.best_tri = er.tri; A.best_ent = cur_ent; }
                } else {
                    B.tmax = tmax; B.ptr = link; B.state = nstate; B.ent = nent;
                    if (er.tri >= 0) { B.best_tri = er.tri; B.best_ent = cur_ent; }
                }
            }
        }
    }
}

template <int KIND, int MINB>
void launch_pair_steps(int steps, unsigned grid, cudaStream_t stream, const SceneView& s, const HotView& hv, const cndl_ray* rays, unsigned R,
                       const RayOrder& order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
#define CNDL_PAIR_LAUNCH(STEPS)                                                                                  \
    {                                                                                                            \
        auto k = trace_pair_stackless_kernel<KIND, MINB, STEPS>;                                                 \
        static bool configured[kMaxDevices] = {};                                                                          \
        const int dev_slot = current_device_slot();                                                                          \
        if (!configured[dev_slot]) {                                                                                       \
            cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 0);                          \
            configured[dev_slot] = true;                                                                                   \
        }                                                                                                        \
        k<<<grid, 128, 0, stream>>>(s, hv.nodes2, hv.ents2, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); \
    }
    switch (steps) {
        case 1: CNDL_PAIR_LAUNCH(1) break;
        case 3: CNDL_PAIR_LAUNCH(3) break;
        default: CNDL_PAIR_LAUNCH(2) break;
    }
#undef CNDL_PAIR_LAUNCH
}

template <int MINB>
void launch_pair_kind(int kind, int steps, unsigned grid, cudaStream_t stream, const SceneView& s, const HotView& hv, const cndl_ray* rays, unsigned R,
                      const RayOrder& order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
    switch (kind) {
        case Q_CLOSEST: launch_pair_steps<Q_CLOSEST, MINB>(steps, grid, stream, s, hv, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: launch_pair_steps<Q_CLOSEST_IGNORE_TRANSPARENT, MINB>(steps, grid, stream, s, hv, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
        default: launch_pair_steps<Q_ANY, MINB>(steps, grid, stream, s, hv, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
    }
}

template <int KIND, int BLOCK, int STEPS>
void launch_hot_one(unsigned grid, size_t smem, cudaStream_t stream, const SceneView& s, const HotView& hv, const cndl_ray* rays, unsigned R,
                    const RayOrder& order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
    auto k = trace_hot_stackless_kernel<KIND, BLOCK, STEPS>;
    static size_t configured[kMaxDevices] = {};
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot] || (smem == 0 && configured[dev_slot] == 0)) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        // everything the staged nodes do not need stays L1: carve out just enough for the resident CTAs
        const size_t per_sm = smem * (1024 / BLOCK) + 1024 * (1024 / BLOCK);
        const size_t pct = (per_sm * 100 + 228 * 1024 - 1) / (228 * 1024);
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)(pct > 100 ? 100 : pct));
        configured[dev_slot] = smem ? smem : 1;
    }
    k<<<grid, BLOCK, smem, stream>>>(s, hv.nodes2, hv.ents2, (int)(smem / 32), rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold);
}

template <int KIND, int BLOCK>
void launch_hot_steps(int steps, unsigned grid, size_t smem, cudaStream_t stream, const SceneView& s, const HotView& hv, const cndl_ray* rays,
                      unsigned R, const RayOrder& order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold,
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
                     const cndl_ray* rays, unsigned R, const RayOrder& order, cndl_hit* hits, float* any_t, unsigned* work_counter,
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

cudaError_t validate_stack_nodes(const float4* nodes, const int2* h_objects, int n_objects, size_t n_tris, int* d_flag, int* h_invalid, cudaStream_t st,
                                 LaunchCounter& lc) {
    cudaMemsetAsync(d_flag, 0, sizeof(int), st);
    for (int o = 0; o < n_objects; ++o) {
        if (h_objects[o].y <= 0) continue;
        validate_stack_kernel<<<(unsigned)((h_objects[o].y + 255) / 256), 256, 0, st>>>(nodes, h_objects[o].x, h_objects[o].y, (int)n_tris, d_flag);
        lc.n++;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_invalid, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}

size_t hot_scratch_ints(size_t N, int n_objects) { return 16 + 2 * 8192 + 2 * N + (N / 2048 + 8) + (size_t)n_objects + 16; }

void launch_trace_hot(const SceneView& s, const HotView& hv, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits,
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

void launch_trace_pair(const SceneView& s, const HotView& hv, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits,
                       float* any_t, unsigned* work_counter, int sm_count, int blocks_per_sm, int park_threshold, int idle_threshold, int steps,
                       cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    unsigned grid = (unsigned)(sm_count * blocks_per_sm);
    const unsigned need = (unsigned)((R + 255) / 256);  // 128 lanes x 2 rays
    if (grid > need) grid = need;
    // register budget follows the requested residency: 4 CTAs/SM -> 128 regs, 5 -> 96, 6 -> 80
    if (blocks_per_sm <= 4) launch_pair_kind<4>(kind, steps, grid, stream, s, hv, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold);
    else if (blocks_per_sm <= 5) launch_pair_kind<5>(kind, steps, grid, stream, s, hv, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold);
    else launch_pair_kind<6>(kind, steps, grid, stream, s, hv, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold);
    lc.n++;
}

}  // namespace cndl
