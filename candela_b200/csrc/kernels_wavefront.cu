// Mode 2: persistent warps, "while-while" traversal with postponed leaf tests and a batched
// service phase (after Aila & Laine's persistent while-while, adapted to the threaded stackless
// walk of …/Include/TraverseBVHStackless.glsl:175-278 and to the stack walk of
// …/Include/TraverseBVHStack.glsl:168-324).
//
// Per ray the sequence of node visits, box tests, triangle tests and TMax updates is exactly the
// reference's; only the interleaving BETWEEN the rays of a warp changes.  Every lane owns one ray
// as a small state machine {EMPTY, WALK, LEAF, DONE}:
//   * node phase   : walking lanes take node steps.  A step is branch-free (next pointer and next
//                    state are selected, not branched on).  A lane that enters a leaf parks (LEAF):
//                    it cannot go on before its triangles are tested, because TMax feeds the next
//                    box test.  A lane whose walk of the current entity ends parks as DONE.
//                    One ballot per vote round: the phase runs until a quarter (at most
//                    `park_threshold`) of the lanes that were walking at its start have parked, so
//                    the tail of a batch does not serialise behind its last long rays.
//   * leaf phase   : the long triangle path runs for all parked lanes at once.
//   * service phase: runs when `idle_threshold` lanes are DONE or EMPTY: moving on to the next
//                    entity (matrix loads, object-space ray, three IEEE divisions), or retiring the
//                    ray (barycentrics + hit store), then fetching fresh rays from a global counter
//                    (one atomic per warp).
// Nodes are fetched with one 256-bit load (LDG.E.256) per visit: a FlattenedNode is one 32 B sector.
//
// History (profiles/r1_experiments.md): the first version tested three thresholds per vote round
// (36 of ~150 warp instructions per round) and branched on enter / miss-link, which ran both sides
// with half the lanes each; removing both took the diffuse batch from 0.97 to 0.78 ms.
#include "kernels.cuh"

namespace cndl {

namespace {

enum LaneState : int { EMPTY = 0, WALK = 1, LEAF = 2, DONE = 3 };
constexpr int NO_MORE_ENTITIES = 0x3FFFFFFF;
constexpr unsigned FULL = 0xFFFFFFFFu;


__device__ __forceinline__ void store_hit(cndl_hit* __restrict__ hits, size_t i, float t, float u, float v, float w, int mesh, int tri, int ent, int iters) {
    stg256(reinterpret_cast<float4*>(hits + i), make_float4(t, u, v, w),
           make_float4(__int_as_float(mesh), __int_as_float(tri), __int_as_float(ent), __int_as_float(iters)));  // one 256-bit store
}

// tail of IntersectScene (SL:300-318 / ST:347-365): `closest` is TMax after the last acceptance
__device__ __forceinline__ void retire_closest(const SceneView& s, const cndl_ray* __restrict__ rays, cndl_hit* __restrict__ hits, unsigned rid, unsigned out,
                                               const RayState& cur, int cur_ent, float closest, int best_tri, int best_ent, int iters) {
    float t = -1.0f, u = -1.0f, v = -1.0f, w = -1.0f;
    int mesh = -1;
    if (best_tri >= 0) mesh = __ldg(&s.tris[best_tri]).w;
    if (best_tri > 0) {  // ClosestT > 0 && TriangleIdx > 0: global triangle 0 reports as a miss
        RayState r = cur;
        if (best_ent != cur_ent) {
            const float4* rp = reinterpret_cast<const float4*>(rays + rid);
            float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
            r = to_object_space(s.ents + best_ent, V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z});
        }
        t = closest;
        const V3 p = {fadd(r.o.x, fmul(r.d.x, t)), fadd(r.o.y, fmul(r.d.y, t)), fadd(r.o.z, fmul(r.d.z, t))};
        barycentrics(s.tri48, best_tri, p, u, v, w);
    }
    store_hit(hits, out, t, u, v, w, mesh, best_tri, best_ent, iters);
}

// ---------------------------------------------------------------------------------------------
// stackless format

struct WLane {
    RayState r;          // object-space ray of the entity being traversed
    float tmax;          // TMax; equals the closest accepted t once best_tri >= 0
    int ptr, start, lo, hi, iters, ent;  // lo..hi: pointers the loop header admits (SL:196)
    int best_tri, best_ent;
    int pend_pack, pend_link;
    unsigned rid;
    int state;
};

// Scene loop bookkeeping (SL:290-301 / :333-337): start on the next entity to traverse, or stay DONE.
template <int KIND>
__device__ __forceinline__ void next_entity(const SceneView& s, const cndl_ray* __restrict__ rays, WLane& L, int from) {
    int e = from;
    while (e < s.n_ents) {
        const cndl_entity* ent = s.ents + e;
        if (KIND == Q_CLOSEST_IGNORE_TRANSPARENT && __int_as_float(__ldg(&ent->data[1])) < 0.99f) { ++e; continue; }
        const float4* rp = reinterpret_cast<const float4*>(rays + L.rid);
        float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
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

// One node visit of a walking lane (SL:192-246).  DONE = this entity is finished (miss link -1, or the
// loop header's cap / range checks fail).  CHECKED = false: the node buffer and the entity ranges were validated at
// commit (every link and first child inside its object), so the pointer range checks of SL:196 cannot fire.
template <bool CHECKED, bool EXACT>
__device__ __forceinline__ void node_step(const SceneView& s, WLane& L) {
    if (L.state == WALK) {
        if (CHECKED) {
            if (L.iters >= 1024 || L.ptr < L.lo || L.ptr > L.hi) {  // loop header of SL:192-199
                L.state = DONE;
                return;
            }
        }
        // !CHECKED: the pointer is valid, so a lane that has reached the iteration cap may load its node anyway; the
        // visit is then discarded by the selects below instead of being branched around
        const bool capped = !CHECKED && L.iters >= 1024;
        float4 mn, mx;
        ldg256(s.nodes + 2 * (size_t)L.ptr, mn, mx);
        const int link = __float_as_int(mx.w), pack = __float_as_int(mn.w);
        const bool enter = enter_stackless(mn, mx, L.r, L.tmax, EXACT);
        L.pend_pack = pack;
        L.pend_link = link;
        L.ptr = enter ? L.ptr + 1 : link + L.start;  // a leaf's pointer is set again after its triangles
        const int next = enter ? (pack != -1 ? LEAF : WALK) : (link < 0 ? DONE : WALK);
        L.state = capped ? DONE : next;
        ++L.iters;  // a discarded visit counts too: the ray retires with min(iters, 1024)
    }
}

template <int KIND, int MINB, int STEPS, bool CHECKED, bool HELP>
__global__ void __launch_bounds__(128, MINB) trace_ww_stackless_kernel(SceneView s, const cndl_ray* __restrict__ rays, unsigned R, RayOrder order,
                                                                    cndl_hit* __restrict__ hits, float* __restrict__ any_t,
                                                                    unsigned* __restrict__ work_counter, int park_threshold, int idle_threshold) {
    R = batch_length(order, R);
    constexpr bool ANY = KIND == Q_ANY;
    __shared__ float4 s_box[HELP ? 4 : 1][HELP ? 64 : 1];  // per warp: up to 32 posted (ray, triangle) pairs
    __shared__ float s_res[HELP ? 4 : 1][HELP ? 32 : 1];
    const unsigned lane = threadIdx.x & 31u;
    WLane L;
    L.state = EMPTY;
    L.rid = 0;
    L.iters = 0;
    L.ent = 0;
    L.best_tri = -1;
    L.best_ent = -1;
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
                    next_entity<KIND>(s, rays, L, L.ent + 1);  // WALK again, or still DONE: the scene loop is over
                    if (L.state == DONE) {
                        if (ANY) any_t[out_slot(order, L.rid)] = L.best_tri >= 0 ? L.tmax : -1.0f;
                        else retire_closest(s, rays, hits, L.rid, out_slot(order, L.rid), L.r, L.ent, L.tmax, L.best_tri, L.best_ent, min(L.iters, 1024));
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
        {
            const int walk0 = __popc(__ballot_sync(FULL, L.state == WALK));
            if (walk0 > 0) {
                int park = walk0 >> 2;
                park = park < 1 ? 1 : (park > park_threshold ? park_threshold : park);
                const int min_walk = walk0 - park + 1;  // >= 1
                if (!warp_exact) {  // the common case: no lane needs the literal GLSL min/max, and the loop does not test for it
                    do {
#pragma unroll
                        for (int step = 0; step < STEPS; ++step) node_step<CHECKED, false>(s, L);
                    } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
                } else {
                    do {
                        node_step<CHECKED, true>(s, L);
                    } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
                }
            }
        }

        // ---------------- leaf phase ----------------
        if (!HELP) {
            if (L.state == LEAF) {
                EntityResult er{-1.0f, -1, 0};
                const bool found = leaf_triangles<ANY>(s, L.pend_pack, L.r, L.tmax, er);
                if (er.tri >= 0) { L.best_tri = er.tri; L.best_ent = L.ent; }
                L.ptr = L.pend_link + L.start;
                L.state = L.pend_link < 0 ? DONE : WALK;
                if (ANY && found) {  // SL:567-569: the scene loop returns the first T > 0
                    L.state = DONE;
                    L.ent = NO_MORE_ENTITIES;
                }
            }
        } else {
            // Helper lanes: a leaf holds one or two triangles, and only a handful of lanes are parked, so the second
            // triangle of every parked lane is tested in the SAME pass by a lane that is not parked.  The owner posts
            // its ray and the triangle index in a shared-memory mailbox, the helper posts t back, and the owner applies
            // the accept rule to its two results in the reference's order (t itself does not depend on TMax).
            const bool own = L.state == LEAF;
            const unsigned owners = __ballot_sync(FULL, own);
            if (owners != 0u) {
                const int first = L.pend_pack >> 4, len = L.pend_pack & 0xF;
                const bool has2 = own && len >= 2;
                const unsigned extra = __ballot_sync(FULL, has2), helpers = ~owners;
                const unsigned lt = (1u << lane) - 1u;
                const int my_extra = __popc(extra & lt), n_helpers = __popc(helpers);
                float4* box = s_box[threadIdx.x >> 5];
                float* res = s_res[threadIdx.x >> 5];
                const bool posted = has2 && my_extra < n_helpers;
                if (posted) {
                    box[2 * my_extra] = make_float4(L.r.o.x, L.r.o.y, L.r.o.z, __int_as_float(first + 1));
                    box[2 * my_extra + 1] = make_float4(L.r.d.x, L.r.d.y, L.r.d.z, 0.0f);
                }
                __syncwarp();
                const int my_help = __popc(helpers & lt);
                const bool help = !own && my_help < __popc(extra);
                V3 o = L.r.o, d = L.r.d;
                int tri = first;
                if (help) {
                    const float4 a = box[2 * my_help], b = box[2 * my_help + 1];
                    o = {a.x, a.y, a.z};
                    d = {b.x, b.y, b.z};
                    tri = __float_as_int(a.w);
                }
                float t0 = -1.0f;
                if ((own && len > 0) || help) t0 = ray_triangle(s.tri48, tri, o, d);
                if (help) res[my_help] = t0;
                __syncwarp();
                if (own) {
                    bool found = false;
                    if (t0 > 0.0f && t0 < L.tmax) { L.tmax = t0; L.best_tri = first; L.best_ent = L.ent; found = true; }
                    if (len >= 2 && !(ANY && found)) {
                        const float t1 = posted ? res[my_extra] : ray_triangle(s.tri48, first + 1, L.r.o, L.r.d);
                        if (t1 > 0.0f && t1 < L.tmax) { L.tmax = t1; L.best_tri = first + 1; L.best_ent = L.ent; found = true; }
                        for (int idx = first + 2; idx < first + len && !(ANY && found); ++idx) {  // leaves of 3+ triangles: prebuilt buffers only
                            const float t = ray_triangle(s.tri48, idx, L.r.o, L.r.d);
                            if (t > 0.0f && t < L.tmax) { L.tmax = t; L.best_tri = idx; L.best_ent = L.ent; found = true; }
                        }
                    }
                    L.ptr = L.pend_link + L.start;
                    L.state = L.pend_link < 0 ? DONE : WALK;
                    if (ANY && found) {  // SL:567-569: the scene loop returns the first T > 0
                        L.state = DONE;
                        L.ent = NO_MORE_ENTITIES;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The same scheme for the stack format (ST:168-324, :509-657).  One step loads a 64-byte node (two
// children).  The choice of the next node (near child first, far child pushed, pop when nothing is
// entered) uses box distances computed with TMax as it was BEFORE this node's leaf children are
// intersected (ST:215-216), so it is taken first; the lane then parks with up to two pending leaves,
// which the leaf phase tests left then right, exactly in the reference order.
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
        float4 a, b;
        ldg256(rp, a, b);  // rays are 32-byte records, 32-byte aligned
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

// CHECKED = false: slots and leaf ranges were validated at commit and every entity names a whole object, so the pointer
// range checks of ST:201-205 cannot fire; the stack pointer stays in 0..63 by construction (a push at 63 ends the walk).
// The traversal stack (ST:170 `int Stack[64]`): its first kSmemStack entries live in shared memory ([entry][thread]: conflict-free),
// the rest in local memory; pushes and pops then stay off the L1 tag path that the node loads saturate.
constexpr int kSmemStack = 24;
__device__ __forceinline__ void stack_push(int* __restrict__ local_stack, int (*smem_stack)[128], int sp, int value) {
    if (sp < kSmemStack) smem_stack[sp][threadIdx.x] = value;
    else local_stack[sp] = value;
}
__device__ __forceinline__ int stack_pop(const int* __restrict__ local_stack, int (*smem_stack)[128], int sp) {
    return sp < kSmemStack ? smem_stack[sp][threadIdx.x] : local_stack[sp];
}

template <bool EXACT, bool CHECKED>
__device__ __forceinline__ void node_step_stack(const SceneView& s, SLane& L, int* stack, int (*smem_stack)[128]) {
    if (L.state == WALK) {
        if (L.iters >= 1024 || (CHECKED && (L.sp >= 64 || L.sp < 0 || L.cur < L.lo || L.cur > L.hi))) {  // ST:198-205
            L.state = DONE;
        } else {
            ++L.iters;
            float4 lmn, lmx, rmn, rmx;
            ldg256(s.nodes + 4 * (size_t)L.cur, lmn, lmx);
            ldg256(s.nodes + 4 * (size_t)L.cur + 2, rmn, rmx);
            const int lpack = __float_as_int(lmn.w), rpack = __float_as_int(rmn.w);
            const bool lleaf = lpack != -1, rleaf = rpack != -1;
            const float lt = lleaf ? -1.0f : slab_stack(lmn, lmx, L.r, L.tmax, EXACT);
            const float rt = rleaf ? -1.0f : slab_stack(rmn, rmx, L.r, L.tmax, EXACT);
            const int lslot = __float_as_int(lmx.w) + L.start, rslot = __float_as_int(rmx.w) + L.start;
            const bool hl = lt > 0.0f, hr = rt > 0.0f, both = hl && hr, right_first = rt < lt;
            int after = WALK;
            if (both) {  // ST:280-299: near child first (ties go left), far child pushed
                L.cur = right_first ? rslot : lslot;
                if (L.sp >= 63) after = DONE;
                else { stack_push(stack, smem_stack, L.sp, right_first ? lslot : rslot); ++L.sp; }
            } else if (hl || hr) {
                L.cur = hl ? lslot : rslot;
            } else if (L.sp <= 0) {
                after = DONE;
            } else {
                L.cur = stack_pop(stack, smem_stack, --L.sp);
            }
            L.pend_l = lleaf ? lpack : -1;
            L.pend_r = rleaf ? rpack : -1;
            L.after = after;
            L.state = (lleaf || rleaf) ? LEAF : after;
        }
    }
}

template <int KIND, int STEPS, bool CHECKED>
__global__ void __launch_bounds__(128, 8) trace_ww_stack_kernel(SceneView s, const cndl_ray* __restrict__ rays, unsigned R, RayOrder order,
                                                                cndl_hit* __restrict__ hits, float* __restrict__ any_t,
                                                                unsigned* __restrict__ work_counter, int park_threshold, int idle_threshold) {
    R = batch_length(order, R);
    constexpr bool ANY = KIND == Q_ANY;
    __shared__ int s_stack[kSmemStack][128];
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
    L.pend_l = L.pend_r = -1;
    L.after = WALK;
    bool drained = false;
    bool warp_exact = false;

    while (true) {
        {   // service: next entity / retire / refill
            const unsigned b0 = __ballot_sync(FULL, (L.state & 1) != 0), b1 = __ballot_sync(FULL, (L.state & 2) != 0);
            const unsigned done = b0 & b1, empty = ~(b0 | b1), busy = b0 ^ b1;
            if (busy == 0u && done == 0u && drained) break;
            const int serviceable = __popc(done) + (drained ? 0 : __popc(empty));
            if (serviceable >= idle_threshold || busy == 0u) {
                if (L.state == DONE) {
                    next_entity_stack<KIND>(s, rays, L, L.ent + 1);
                    if (L.state == DONE) {
                        if (ANY) any_t[out_slot(order, L.rid)] = L.best_tri >= 0 ? L.tmax : -1.0f;
                        else retire_closest(s, rays, hits, L.rid, out_slot(order, L.rid), L.r, L.ent, L.tmax, L.best_tri, L.best_ent, L.iters);
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
                warp_exact = __any_sync(FULL, (L.state == WALK || L.state == LEAF) && L.r.nan_path);
            }
        }
        {   // node phase
            const int walk0 = __popc(__ballot_sync(FULL, L.state == WALK));
            if (walk0 > 0) {
                int park = walk0 >> 2;
                park = park < 1 ? 1 : (park > park_threshold ? park_threshold : park);
                const int min_walk = walk0 - park + 1;
                if (!warp_exact) {
                    do {
#pragma unroll
                        for (int step = 0; step < STEPS; ++step) node_step_stack<false, CHECKED>(s, L, stack, s_stack);
                    } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
                } else {
                    do {
                        node_step_stack<true, CHECKED>(s, L, stack, s_stack);
                    } while (__popc(__ballot_sync(FULL, L.state == WALK)) >= min_walk);
                }
            }
        }
        if (L.state == LEAF) {  // leaf phase: left leaf, then right leaf (ST:221-277)
            EntityResult er{-1.0f, -1, 0};
            bool found = false;
            if (L.pend_l != -1) found = leaf_triangles<ANY>(s, L.pend_l, L.r, L.tmax, er);
            if (!found && L.pend_r != -1) found = leaf_triangles<ANY>(s, L.pend_r, L.r, L.tmax, er);
            if (er.tri >= 0) { L.best_tri = er.tri; L.best_ent = L.ent; }
            L.state = L.after;
            if (ANY && found) {
                L.state = DONE;
                L.ent = NO_MORE_ENTITIES;
            }
        }
    }
}

template <int KIND, int MINB, bool CHECKED, bool HELP>
void launch_steps(int steps, unsigned grid, cudaStream_t stream, const SceneView& s, const cndl_ray* rays, unsigned R, const RayOrder& order,
                  cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
#define CNDL_WW_LAUNCH(STEPS)                                                                                     \
    {                                                                                                             \
        auto k = trace_ww_stackless_kernel<KIND, MINB, STEPS, CHECKED, HELP>;                                     \
        static bool configured[kMaxDevices] = {};                                                                           \
        const int dev_slot = current_device_slot();                                                                           \
        if (!configured[dev_slot]) { /* no shared memory is used: give the whole 256 KB array to L1 */                      \
            cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, HELP ? 20 : 0);               \
            configured[dev_slot] = true;                                                                                    \
        }                                                                                                         \
        k<<<grid, 128, 0, stream>>>(s, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); \
    }
    switch (steps) {
        case 1: CNDL_WW_LAUNCH(1) break;
        case 3: CNDL_WW_LAUNCH(3) break;
        case 4: CNDL_WW_LAUNCH(4) break;
        default: CNDL_WW_LAUNCH(2) break;
    }
#undef CNDL_WW_LAUNCH
}

template <int MINB, bool CHECKED, bool HELP>
void launch_kind(int kind, int steps, unsigned grid, cudaStream_t stream, const SceneView& s, const cndl_ray* rays, unsigned R, const RayOrder& order,
                 cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
    switch (kind) {
        case Q_CLOSEST: launch_steps<Q_CLOSEST, MINB, CHECKED, HELP>(steps, grid, stream, s, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: launch_steps<Q_CLOSEST_IGNORE_TRANSPARENT, MINB, CHECKED, HELP>(steps, grid, stream, s, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
        default: launch_steps<Q_ANY, MINB, CHECKED, HELP>(steps, grid, stream, s, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
    }
}

template <int KIND>
void launch_stack_steps(int steps, bool validated, unsigned grid, cudaStream_t stream, const SceneView& s, const cndl_ray* rays, unsigned R,
                        const RayOrder& order, cndl_hit* hits, float* any_t, unsigned* work_counter, int park_threshold, int idle_threshold) {
#define CNDL_ST_ARGS s, rays, R, order, hits, any_t, work_counter, park_threshold, idle_threshold
    if (steps == 2 && validated) trace_ww_stack_kernel<KIND, 2, false><<<grid, 128, 0, stream>>>(CNDL_ST_ARGS);
    else if (steps == 2) trace_ww_stack_kernel<KIND, 2, true><<<grid, 128, 0, stream>>>(CNDL_ST_ARGS);
    else if (validated) trace_ww_stack_kernel<KIND, 1, false><<<grid, 128, 0, stream>>>(CNDL_ST_ARGS);
    else trace_ww_stack_kernel<KIND, 1, true><<<grid, 128, 0, stream>>>(CNDL_ST_ARGS);
#undef CNDL_ST_ARGS
}

}  // namespace

void launch_trace_ww_stack(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits, float* any_t,
                           unsigned* work_counter, int sm_count, int blocks_per_sm, int park_threshold, int idle_threshold, int steps, bool validated,
                           cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    unsigned grid = (unsigned)(sm_count * (blocks_per_sm < 1 ? 1 : (blocks_per_sm > 8 ? 8 : blocks_per_sm)));  // 64 registers: 8 CTAs per SM
    const unsigned need = (unsigned)((R + 127) / 128);
    if (grid > need) grid = need;
    switch (kind) {
        case Q_CLOSEST: launch_stack_steps<Q_CLOSEST>(steps, validated, grid, stream, s, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: launch_stack_steps<Q_CLOSEST_IGNORE_TRANSPARENT>(steps, validated, grid, stream, s, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
        default: launch_stack_steps<Q_ANY>(steps, validated, grid, stream, s, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold); break;
    }
    lc.n++;
}

void launch_trace_ww(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const RayOrder& order, cndl_hit* hits, float* any_t,
                     unsigned* work_counter, int sm_count, int blocks_per_sm, int park_threshold, int idle_threshold, int steps, bool validated,
                     bool helper_lanes, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    unsigned grid = (unsigned)(sm_count * blocks_per_sm);
    const unsigned need = (unsigned)((R + 127) / 128);
    if (grid > need) grid = need;
    // register budget follows the requested residency: <= 8 CTAs/SM -> 64 registers, 10 -> 48, 12 -> 40
#define CNDL_WW_ARGS kind, steps, grid, stream, s, rays, (unsigned)R, order, hits, any_t, work_counter, park_threshold, idle_threshold
    if (blocks_per_sm <= 8) {
        if (validated && helper_lanes) launch_kind<8, false, true>(CNDL_WW_ARGS);
        else if (validated) launch_kind<8, false, false>(CNDL_WW_ARGS);
        else launch_kind<8, true, false>(CNDL_WW_ARGS);
    }
    else if (blocks_per_sm == 9 && validated) launch_kind<9, false, false>(CNDL_WW_ARGS);  // 56 registers
    else if (blocks_per_sm <= 10) launch_kind<10, true, false>(CNDL_WW_ARGS);
    else launch_kind<12, true, false>(CNDL_WW_ARGS);
#undef CNDL_WW_ARGS
    lc.n++;
}

}  // namespace cndl
