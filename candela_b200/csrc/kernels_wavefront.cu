// Mode 2: persistent warps, "while-while" stackless traversal with postponed leaf tests and
// batched retire/refill (after Aila & Laine's persistent while-while, adapted to the threaded
// stackless walk of …/Include/TraverseBVHStackless.glsl:175-278).
//
// Per ray the sequence of node visits, box tests, triangle tests and TMax updates is exactly the
// reference's; only the interleaving BETWEEN rays of a warp changes:
//   * node phase   : lanes that are walking take node steps; a lane that enters a leaf parks
//                    (it cannot go on before its triangles are tested: TMax feeds the next box test);
//   * leaf phase   : runs when enough lanes are parked, so the long triangle path executes with many
//                    lanes active instead of one or two per iteration;
//   * retire/refill: finished rays park too; barycentrics + hit store + fetching fresh rays run for
//                    a batch of lanes at once, from a global counter (one atomic per refill).
// Nodes are fetched with one 256-bit load (LDG.E.256) per visit: a FlattenedNode is one 32 B sector.
#include "kernels.cuh"

namespace cndl {

namespace {

enum LaneState : int { EMPTY = 0, WALK = 1, LEAF = 2, DONE = 3 };

__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void store_hit(cndl_hit* __restrict__ hits, size_t i, float t, float u, float v, float w, int mesh, int tri, int ent, int iters) {
    float4* p = reinterpret_cast<float4*>(hits + i);
    p[0] = make_float4(t, u, v, w);
    reinterpret_cast<int4*>(p)[1] = make_int4(mesh, tri, ent, iters);
}

struct WLane {
    RayState r;          // object-space ray of the entity being traversed
    float tmax, closest;
    int ptr, start, lo, hi, iters, ent;  // lo..hi: pointers the loop header admits (SL:196)
    int best_tri, best_ent;
    int pend_pack, pend_link;
    unsigned rid;
    int state;
};

// Scene loop bookkeeping (SL:290-301 / :333-337): move to the next entity to traverse, or finish the ray.
template <int KIND>
__device__ __forceinline__ void next_entity(const SceneView& s, const cndl_ray* __restrict__ rays, WLane& L, int from) {
    int e = from;
    while (e < s.n_ents) {
        const cndl_entity* ent = s.ents + e;
        if (KIND == Q_CLOSEST_IGNORE_TRANSPARENT && __int_as_float(__ldg(&ent->data[1])) < 0.99f) { ++e; continue; }
        const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
        const float4 a = __ldg(rp), b = __ldg(rp + 1);
        L.r = to_object_space(ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
        L.start = __ldg(&ent->node_offset);
        // Pointer >= 0 && Pointer >= NodeStart && Pointer <= NodeStart+NodeCount && Pointer <= u_TotalNodes
        L.lo = L.start > 0 ? L.start : 0;
        L.hi = min(L.start + __ldg(&ent->node_count), s.total_nodes);
        L.ptr = L.start;
        L.iters = 0;
        L.ent = e;
        L.state = WALK;
        return;
    }
    L.state = DONE;
}

template <int KIND, int MINB, int STEPS, bool PREFETCH>
__global__ void __launch_bounds__(128, MINB) trace_ww_stackless_kernel(SceneView s, const cndl_ray* __restrict__ rays, unsigned R,
                                                                    const unsigned* __restrict__ order, cndl_hit* __restrict__ hits,
                                                                    float* __restrict__ any_t, unsigned* __restrict__ work_counter,
                                                                    int leaf_threshold, int idle_threshold) {
    constexpr bool ANY = KIND == Q_ANY;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    WLane L;
    L.state = EMPTY;
    L.rid = 0;
    L.iters = 0;
    L.ent = 0;
    L.best_tri = -1;
    L.best_ent = -1;
    L.closest = -1.0f;
    bool drained = false;

    while (true) {
        // ---------------- retire finished rays, fetch fresh ones ----------------
        {
            const unsigned done = __ballot_sync(FULL, L.state == DONE);
            const unsigned empty = __ballot_sync(FULL, L.state == EMPTY);
            const unsigned busy = ~(done | empty);
            const int serviceable = __popc(done) + (drained ? 0 : __popc(empty));
            if (busy == 0u && done == 0u && drained) break;
            if (serviceable >= idle_threshold || busy == 0u) {
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
                            if (L.best_ent != L.ent) {
                                const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
                                const float4 a = __ldg(rp), b = __ldg(rp + 1);
                                r = to_object_space(s.ents + L.best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
                            }
                            const V3 p = {fadd(r.o.x, fmul(r.d.x, L.closest)), fadd(r.o.y, fmul(r.d.y, L.closest)), fadd(r.o.z, fmul(r.d.z, L.closest))};
                            t = L.closest;
                            barycentrics(s.tri48, L.best_tri, p, u, v, w);
                        }
                        store_hit(hits, L.rid, t, u, v, w, mesh, L.best_tri, L.best_ent, L.iters);
                    }
                    L.state = EMPTY;
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
                            next_entity<KIND>(s, rays, L, 0);
                        }
                    }
                }
            }
        }

        // ---------------- node phase ----------------
        while (true) {
            // lane states are two bits: two ballots give all four masks
            const unsigned b0 = __ballot_sync(FULL, (L.state & 1) != 0), b1 = __ballot_sync(FULL, (L.state & 2) != 0);
            const unsigned walking = b0 & ~b1;
            if (walking == 0u) break;
            if (__popc(b1 & ~b0) >= leaf_threshold) break;                                  // parked at a leaf
            if (__popc(b0 & b1) + (drained ? 0 : __popc(~(b0 | b1))) >= idle_threshold) break;  // finished / empty
#pragma unroll
            for (int step = 0; step < STEPS; ++step) {
                if (L.state == WALK) {
                    // loop header of SL:192-199 (Pointer >= 0, Iterations < 1024, range checks)
                    if (L.iters >= 1024 || L.ptr < L.lo || L.ptr > L.hi) {
                        next_entity<KIND>(s, rays, L, L.ent + 1);
                    } else {
                        ++L.iters;
                        float4 mn, mx;
                        ldg256(s.nodes + 2 * (size_t)L.ptr, mn, mx);
                        const int link = __float_as_int(mx.w);
                        if (enter_stackless(mn, mx, L.r, L.tmax)) {
                            const int pack = __float_as_int(mn.w);
                            if (pack != -1) {
                                L.pend_pack = pack;
                                L.pend_link = link;
                                L.state = LEAF;
                            } else {
                                ++L.ptr;
                            }
                        } else if (link < 0) {
                            next_entity<KIND>(s, rays, L, L.ent + 1);
                        } else {
                            L.ptr = link + L.start;
                        }
                    }
                }
            }
            // the vote round separates this step from the next load: start fetching the next node now
            if (PREFETCH && L.state == WALK) prefetch_l1(s.nodes + 2 * (size_t)L.ptr);
        }

        // ---------------- leaf phase ----------------
        if (L.state == LEAF) {
            EntityResult er{-1.0f, -1, 0};
            const bool found = leaf_triangles<ANY>(s, L.pend_pack, L.r, L.tmax, er);
            if (er.tri >= 0) { L.closest = er.t; L.best_tri = er.tri; L.best_ent = L.ent; }
            if (ANY && found) {
                L.state = DONE;  // SL:567-569: the scene loop returns the first T > 0
            } else if (L.pend_link < 0) {
                next_entity<KIND>(s, rays, L, L.ent + 1);
            } else {
                L.ptr = L.pend_link + L.start;
                L.state = WALK;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Second generation of the stackless while-while kernel.  Same per-ray sequence of operations;
// what changed is what the WARP executes around it (ncu source counters of the first version:
// 36 of ~150 warp instructions per vote round were the three-way threshold test, and the
// enter / miss-link sides of each node step ran one after the other with half the lanes each):
//   * a node step is branch-free: next pointer and next state are selected, never branched on;
//   * DONE means "this entity is finished".  Moving on to the next entity (matrix loads, the
//     object-space ray, three IEEE divisions) happens in the batched service phase together with
//     retiring and refilling, not inline in the node loop (three inlined copies before);
//   * one ballot per vote round: the node phase runs until `park` of the lanes that were walking
//     when it started have parked (at a leaf or at the end of an entity); the count adapts to the
//     number of walkers (a quarter of them, at most `park_threshold`) so the tail of a batch does
//     not serialise behind its last long rays.
// One node visit of a walking lane (SL:192-246), branch-free: the next pointer and the next state are selected.
// DONE = this entity is finished (miss link -1, or the loop header's cap / range checks fail).
__device__ __forceinline__ void node_step(const SceneView& s, WLane& L, bool warp_exact) {
    if (L.state == WALK) {
        if (L.iters >= 1024 || L.ptr < L.lo || L.ptr > L.hi) {  // loop header of SL:192-199
            L.state = DONE;
        } else {
            ++L.iters;
            float4 mn, mx;
            ldg256(s.nodes + 2 * (size_t)L.ptr, mn, mx);
            const int link = __float_as_int(mx.w), pack = __float_as_int(mn.w);
            const bool enter = enter_stackless(mn, mx, L.r, L.tmax, warp_exact);
            L.pend_pack = pack;
            L.pend_link = link;
            L.ptr = enter ? L.ptr + 1 : link + L.start;  // a leaf's pointer is set again after its triangles
            L.state = enter ? (pack != -1 ? LEAF : WALK) : (link < 0 ? DONE : WALK);
        }
    }
}

template <int KIND, int MINB, int STEPS, int POLICY>
__global__ void __launch_bounds__(128, MINB) trace_ww2_stackless_kernel(SceneView s, const cndl_ray* __restrict__ rays, unsigned R,
                                                                     const unsigned* __restrict__ order, cndl_hit* __restrict__ hits,
                                                                     float* __restrict__ any_t, unsigned* __restrict__ work_counter,
                                                                     int park_threshold, int idle_threshold) {
    constexpr bool ANY = KIND == Q_ANY;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int NO_MORE_ENTITIES = 0x3FFFFFFF;
    const unsigned lane = threadIdx.x & 31u;
    WLane L;
    L.state = EMPTY;
    L.rid = 0;
    L.iters = 0;
    L.ent = 0;
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
            int ks = idle_threshold;
            if (POLICY == 1) {  // the node phase's adaptive threshold (see there)
                const int half = (__popc(busy) + 1) >> 1;
                ks = half < idle_threshold ? (half < 1 ? 1 : half) : idle_threshold;
            }
            if (serviceable >= ks || busy == 0u) {
                if (L.state == DONE) {
                    next_entity<KIND>(s, rays, L, L.ent + 1);  // WALK again, or still DONE: the scene loop is over
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
                                if (L.best_ent != L.ent) {
                                    const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
                                    const float4 a = __ldg(rp), b = __ldg(rp + 1);
                                    r = to_object_space(s.ents + L.best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
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
                            next_entity<KIND>(s, rays, L, 0);
                        }
                    }
                }
                warp_exact = __any_sync(FULL, (L.state == WALK || L.state == LEAF) && L.r.nan_path);
            }
        }

        // ---------------- node phase ----------------
        bool run_leaves = true;
        if (POLICY == 0) {
            const int walk0 = __popc(__ballot_sync(FULL, L.state == WALK));
            if (walk0 > 0) {
                int park = walk0 >> 2;
                park = park < 1 ? 1 : (park > park_threshold ? park_threshold : park);
                const int min_walk = walk0 - park + 1;  // >= 1
                do {
#pragma unroll
                    for (int step = 0; step < STEPS; ++step) {
                        node_step(s, L, warp_exact);
                    }
                } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
            }
        } else {
            // Parked lanes wait until enough of them can share the long triangle / service code: the node
            // phase runs until `kl` lanes sit at a leaf, or `ks` lanes can be serviced, or nobody walks.
            // Both thresholds shrink with the number of live rays so that the tail does not serialise.
            const unsigned b0 = __ballot_sync(FULL, (L.state & 1) != 0), b1 = __ballot_sync(FULL, (L.state & 2) != 0);
            int nw = __popc(b0 & ~b1), nl = __popc(b1 & ~b0);
            const int half = (nw + nl + 1) >> 1;
            const int kl = half < park_threshold ? (half < 1 ? 1 : half) : park_threshold;
            const int ks = half < idle_threshold ? (half < 1 ? 1 : half) : idle_threshold;
            const int floor_busy = 32 - (drained ? __popc(~(b0 | b1)) : 0) - ks;  // walking + at-leaf lanes at which ks are serviceable
            while (nw > 0 && nl < kl && nw + nl > floor_busy) {
#pragma unroll
                for (int step = 0; step < STEPS; ++step) {
                        node_step(s, L, warp_exact);
                }
                nw = __popc(__ballot_sync(FULL, L.state == WALK));
                nl = __popc(__ballot_sync(FULL, L.state == LEAF));
            }
            run_leaves = nl >= kl || nw == 0;
        }

        // ---------------- leaf phase ----------------
        if (run_leaves && L.state == LEAF) {
            EntityResult er{-1.0f, -1, 0};
            const bool found = leaf_triangles<ANY>(s, L.pend_pack, L.r, L.tmax, er);
            if (er.tri >= 0) { L.closest = er.t; L.best_tri = er.tri; L.best_ent = L.ent; }
            L.ptr = L.pend_link + L.start;
            L.state = L.pend_link < 0 ? DONE : WALK;
            if (ANY && found) {  // SL:567-569: the scene loop returns the first T > 0
                L.state = DONE;
                L.ent = NO_MORE_ENTITIES;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same scheme for the stack format (…/Include/TraverseBVHStack.glsl:168-324, :509-657).  One step
// loads a 64-byte node (two children).  The choice of the next node (near child first, far child
// pushed, pop when nothing is entered) uses box distances computed with TMax as it was BEFORE this
// node's leaf children are intersected (ST:215-216), so it is taken first; the lane then parks with
// up to two pending leaves, which the leaf phase tests left then right, exactly in the reference order.
struct SLane {
    RayState r;
    float tmax;
    int cur, start, lo, hi, iters, ent, sp;
    int best_tri, best_ent;
    int pend_l, pend_r, after;  // leaf packs to test (-1: none); state to enter afterwards (WALK, or DONE = entity finished)
    unsigned rid;
    int state;
};

template <int KIND>
__device__ __forceinline__ void next_entity_stack(const SceneView& s, const cndl_ray* __restrict__ rays, SLane& L, int from) {
    int e = from;
    while (e < s.n_ents) {
        const cndl_entity* ent = s.ents + e;
        if (KIND == Q_CLOSEST_IGNORE_TRANSPARENT && __int_as_float(__ldg(&ent->data[1])) < 0.99f) { ++e; continue; }
        const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
        const float4 a = __ldg(rp), b = __ldg(rp + 1);
        L.r = to_object_space(ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
        L.start = __ldg(&ent->node_offset);
        L.lo = L.start > 0 ? L.start : 0;  // ST:201-205
        L.hi = min(L.start + __ldg(&ent->node_count), s.total_nodes);
        L.cur = L.start;
        L.sp = 0;
        L.iters = 0;
        L.ent = e;
        L.state = WALK;
        return;
    }
    L.state = DONE;
}

template <int KIND, int STEPS>
__global__ void __launch_bounds__(128, 6) trace_ww_stack_kernel(SceneView s, const cndl_ray* __restrict__ rays, unsigned R,
                                                                const unsigned* __restrict__ order, cndl_hit* __restrict__ hits,
                                                                float* __restrict__ any_t, unsigned* __restrict__ work_counter,
                                                                int leaf_threshold, int idle_threshold) {
    constexpr bool ANY = KIND == Q_ANY;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    int stack[64];
    SLane L;
    L.state = EMPTY;
    L.rid = 0;
    L.iters = 0;
    L.ent = 0;
    L.sp = 0;
    L.best_tri = -1;
    L.best_ent = -1;
    bool drained = false;

    while (true) {
        {   // retire + refill
            const unsigned b0 = __ballot_sync(FULL, (L.state & 1) != 0), b1 = __ballot_sync(FULL, (L.state & 2) != 0);
            const unsigned done = b0 & b1, empty = ~(b0 | b1), busy = b0 ^ b1;
            const int serviceable = __popc(done) + (drained ? 0 : __popc(empty));
            if (busy == 0u && done == 0u && drained) break;
            if (serviceable >= idle_threshold || busy == 0u) {
                if (L.state == DONE) {
                    if (ANY) {
                        any_t[L.rid] = L.best_tri >= 0 ? L.tmax : -1.0f;
                    } else {
                        float t = -1.0f, u = -1.0f, v = -1.0f, w = -1.0f;
                        int mesh = -1;
                        if (L.best_tri >= 0) mesh = __ldg(&s.tris[L.best_tri]).w;
                        if (L.best_tri > 0) {  // ClosestT > 0 && TriangleIdx > 0 (ST:347)
                            RayState r = L.r;
                            if (L.best_ent != L.ent) {
                                const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
                                const float4 a = __ldg(rp), b = __ldg(rp + 1);
                                r = to_object_space(s.ents + L.best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
                            }
                            t = L.tmax;  // TMax == ClosestT once something was accepted
                            const V3 p = {fadd(r.o.x, fmul(r.d.x, t)), fadd(r.o.y, fmul(r.d.y, t)), fadd(r.o.z, fmul(r.d.z, t))};
                            barycentrics(s.tri48, L.best_tri, p, u, v, w);
                        }
                        store_hit(hits, L.rid, t, u, v, w, mesh, L.best_tri, L.best_ent, L.iters);
                    }
                    L.state = EMPTY;
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
                            next_entity_stack<KIND>(s, rays, L, 0);
                        }
                    }
                }
            }
        }
        while (true) {  // node phase
            const unsigned b0 = __ballot_sync(FULL, (L.state & 1) != 0), b1 = __ballot_sync(FULL, (L.state & 2) != 0);
            if ((b0 & ~b1) == 0u) break;
            if (__popc(b1 & ~b0) >= leaf_threshold) break;
            if (__popc(b0 & b1) + (drained ? 0 : __popc(~(b0 | b1))) >= idle_threshold) break;
#pragma unroll
            for (int step = 0; step < STEPS; ++step) {
                if (L.state == WALK) {
                    if (L.iters >= 1024 || L.sp >= 64 || L.sp < 0 || L.cur < L.lo || L.cur > L.hi) {  // ST:198-205
                        next_entity_stack<KIND>(s, rays, L, L.ent + 1);
                    } else {
                        ++L.iters;
                        float4 lmn, lmx, rmn, rmx;
                        ldg256(s.nodes + 4 * (size_t)L.cur, lmn, lmx);
                        ldg256(s.nodes + 4 * (size_t)L.cur + 2, rmn, rmx);
                        const int lpack = __float_as_int(lmn.w), rpack = __float_as_int(rmn.w);
                        const bool lleaf = lpack != -1, rleaf = rpack != -1;
                        const float lt = lleaf ? -1.0f : slab_stack(lmn, lmx, L.r, L.tmax);
                        const float rt = rleaf ? -1.0f : slab_stack(rmn, rmx, L.r, L.tmax);
                        const int lslot = __float_as_int(lmx.w) + L.start, rslot = __float_as_int(rmx.w) + L.start;
                        int after = WALK;
                        if (lt > 0.0f && rt > 0.0f) {  // ST:280-299
                            int postponed = rslot;
                            L.cur = lslot;
                            if (rt < lt) { L.cur = rslot; postponed = lslot; }
                            if (L.sp >= 63) after = DONE;
                            else stack[L.sp++] = postponed;
                        } else if (lt > 0.0f) {
                            L.cur = lslot;
                        } else if (rt > 0.0f) {
                            L.cur = rslot;
                        } else if (L.sp <= 0) {
                            after = DONE;
                        } else {
                            L.cur = stack[--L.sp];
                        }
                        if (lleaf || rleaf) {
                            L.pend_l = lleaf ? lpack : -1;
                            L.pend_r = rleaf ? rpack : -1;
                            L.after = after;
                            L.state = LEAF;
                        } else if (after == DONE) {
                            next_entity_stack<KIND>(s, rays, L, L.ent + 1);
                        }
                    }
                }
            }
        }
        if (L.state == LEAF) {  // leaf phase: left leaf, then right leaf (ST:221-277)
            EntityResult er{-1.0f, -1, 0};
            bool found = false;
            if (L.pend_l != -1) found = leaf_triangles<ANY>(s, L.pend_l, L.r, L.tmax, er);
            if (!found && L.pend_r != -1) found = leaf_triangles<ANY>(s, L.pend_r, L.r, L.tmax, er);
            if (er.tri >= 0) { L.best_tri = er.tri; L.best_ent = L.ent; }
            if (ANY && found) L.state = DONE;
            else if (L.after == DONE) next_entity_stack<KIND>(s, rays, L, L.ent + 1);
            else L.state = WALK;
        }
    }
}

template <int KIND, int MINB, int STEPS, bool PREFETCH>
void launch_one(unsigned grid, unsigned block, cudaStream_t stream, const SceneView& s, const cndl_ray* rays, unsigned R, const unsigned* order,
                cndl_hit* hits, float* any_t, unsigned* work_counter, int leaf_threshold, int idle_threshold) {
    auto k = trace_ww_stackless_kernel<KIND, MINB, STEPS, PREFETCH>;
    static bool configured = false;
    if (!configured) {  // no shared memory is used: give the whole 256 KB array to L1
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        configured = true;
    }
    k<<<grid, block, 0, stream>>>(s, rays, R, order, hits, any_t, work_counter, leaf_threshold, idle_threshold);
}

template <int KIND, int MINB, int STEPS, int POLICY>
void launch_one2(unsigned grid, unsigned block, cudaStream_t stream, const SceneView& s, const cndl_ray* rays, unsigned R, const unsigned* order,
                 cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
    auto k = trace_ww2_stackless_kernel<KIND, MINB, STEPS, POLICY>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        configured = true;
    }
    k<<<grid, block, 0, stream>>>(s, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold);
}

template <int KIND, int MINB>
void launch_variant(int variant, unsigned grid, unsigned block, cudaStream_t stream, const SceneView& s, const cndl_ray* rays, unsigned R,
                    const unsigned* order, cndl_hit* hits, float* any_t, unsigned* work_counter, int leaf_threshold, int idle_threshold) {
    // variant = node steps per vote round (1..4), +8 to prefetch the parked leaf's triangles
#define CNDL_WW_ARGS grid, block, stream, s, rays, R, order, hits, any_t, work_counter, leaf_threshold, idle_threshold
    switch (variant) {
        case 1: launch_one<KIND, MINB, 1, false>(CNDL_WW_ARGS); break;
        case 3: launch_one<KIND, MINB, 3, false>(CNDL_WW_ARGS); break;
        case 4: launch_one<KIND, MINB, 4, false>(CNDL_WW_ARGS); break;
        case 9: launch_one<KIND, MINB, 1, true>(CNDL_WW_ARGS); break;
        case 10: launch_one<KIND, MINB, 2, true>(CNDL_WW_ARGS); break;
        case 17: launch_one2<KIND, MINB, 1, 0>(CNDL_WW_ARGS); break;
        case 18: launch_one2<KIND, MINB, 2, 0>(CNDL_WW_ARGS); break;
        case 19: launch_one2<KIND, MINB, 3, 0>(CNDL_WW_ARGS); break;
        case 20: launch_one2<KIND, MINB, 4, 0>(CNDL_WW_ARGS); break;
        case 25: launch_one2<KIND, MINB, 1, 1>(CNDL_WW_ARGS); break;
        case 26: launch_one2<KIND, MINB, 2, 1>(CNDL_WW_ARGS); break;
        case 27: launch_one2<KIND, MINB, 3, 1>(CNDL_WW_ARGS); break;
        case 28: launch_one2<KIND, MINB, 4, 1>(CNDL_WW_ARGS); break;
        default: launch_one<KIND, MINB, 2, false>(CNDL_WW_ARGS); break;
    }
#undef CNDL_WW_ARGS
}

}  // namespace

void launch_trace_ww_stack(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits, float* any_t,
                           unsigned* work_counter, int sm_count, int leaf_threshold, int idle_threshold, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    const unsigned block = 128;
    unsigned grid = (unsigned)(sm_count * 7);
    const unsigned need = (unsigned)((R + block - 1) / block);
    if (grid > need) grid = need;
    switch (kind) {
        case Q_CLOSEST: trace_ww_stack_kernel<Q_CLOSEST, 1><<<grid, block, 0, stream>>>(s, rays, (unsigned)R, order, hits, any_t, work_counter, leaf_threshold, idle_threshold); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: trace_ww_stack_kernel<Q_CLOSEST_IGNORE_TRANSPARENT, 1><<<grid, block, 0, stream>>>(s, rays, (unsigned)R, order, hits, any_t, work_counter, leaf_threshold, idle_threshold); break;
        default: trace_ww_stack_kernel<Q_ANY, 1><<<grid, block, 0, stream>>>(s, rays, (unsigned)R, order, hits, any_t, work_counter, leaf_threshold, idle_threshold); break;
    }
    lc.n++;
}

void launch_trace_ww(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits, float* any_t,
                     unsigned* work_counter, int sm_count, int blocks_per_sm, int leaf_threshold, int idle_threshold, int variant, cudaStream_t stream,
                     LaunchCounter& lc) {
    if (R == 0) return;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    const unsigned block = 128;
    unsigned grid = (unsigned)(sm_count * blocks_per_sm);
    const unsigned need = (unsigned)((R + block - 1) / block);
    if (grid > need) grid = need;
#define CNDL_WW_LAUNCH(KIND, MINB) \
    launch_variant<KIND, MINB>(variant, grid, block, stream, s, rays, (unsigned)R, order, hits, any_t, work_counter, leaf_threshold, idle_threshold)
#define CNDL_WW_KIND(MINB)                                                                  \
    switch (kind) {                                                                         \
        case Q_CLOSEST: CNDL_WW_LAUNCH(Q_CLOSEST, MINB); break;                             \
        case Q_CLOSEST_IGNORE_TRANSPARENT: CNDL_WW_LAUNCH(Q_CLOSEST_IGNORE_TRANSPARENT, MINB); break; \
        default: CNDL_WW_LAUNCH(Q_ANY, MINB); break;                                        \
    }
    // register budget follows the requested residency: 7 CTAs/SM -> <= 72 regs, 8 -> 64, 10 -> 48, 12 -> 40
    if (blocks_per_sm <= 7) { CNDL_WW_KIND(7) }
    else if (blocks_per_sm <= 8) { CNDL_WW_KIND(8) }
    else if (blocks_per_sm <= 10) { CNDL_WW_KIND(10) }
    else { CNDL_WW_KIND(12) }
    lc.n++;
}

}  // namespace cndl
