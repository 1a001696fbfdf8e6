// Device-side traversal of the reference-layout BVH, bit-exact with the
// reference's GLSL (paths relative to Source/Core/Shaders/Intersectors/Include/):
//   SL = TraverseBVHStackless.glsl, ST = TraverseBVHStack.glsl
//
// Data the kernels read
//   nodes  : the reference's node buffer as float4s (2 per FlattenedNode, 4 per
//            FlattenedStackNode), unchanged.
//   tri48  : 48 bytes per triangle, derived at commit time from the reference's
//            triangle + vertex buffers: {v0, e1 = v1 - v0, e2 = v2 - v0,
//            n = cross(e1, e2)}.  These are exactly the ray-independent
//            sub-expressions of RayTriangle (SL:81-86) evaluated with the same
//            IEEE operations, so using them does not change any result bit; it
//            replaces one 16 B + three 32 B-strided dependent loads per test by
//            three adjacent 16 B loads.
#pragma once
#include "exact_math.cuh"
#include "../../include/candela_b200.h"

namespace cndl {

struct SceneView {
    const float4* __restrict__ nodes;
    const float4* __restrict__ tri48;
    const int4* __restrict__ tris;
    const cndl_entity* __restrict__ ents;
    int total_nodes;  // u_TotalNodes = m_NodeCountBuffered (Intersector.h:345)
    int n_ents;       // u_EntityCount
};

// Which ray a launch slot processes.  list == nullptr: slot i is ray i.  counts == nullptr: ray list[i].
// Otherwise the rays are bucketed by direction octant (octant_partition_kernel): bucket o holds counts[o]
// ray indices at list + o * stride, and slots run through bucket 0, then bucket 1, ...
struct RayOrder {
    const unsigned* __restrict__ list;
    const unsigned* __restrict__ counts;
    unsigned stride;
    const unsigned* __restrict__ d_count;  // optional: the batch length in device memory (the launch's R is then its upper bound)
    const unsigned* __restrict__ out_index;  // optional: the batch was physically reordered; the result of ray slot i belongs at out_index[i]
    const unsigned* __restrict__ seg_counts;  // optional: the batch is cut into segments of kRaySegment slots; only the first seg_counts[k] slots of segment k hold rays
};
constexpr int kRaySegmentShift = 10, kRaySegment = 1 << kRaySegmentShift;

// Does launch slot `slot` hold a ray?  (R: the batch length; segmented batches keep their dead slots at the end of every segment.)
__device__ __forceinline__ bool slot_live(const RayOrder& ro, unsigned slot, unsigned R) {
    if (slot >= R) return false;
    return !ro.seg_counts || (slot & (unsigned)(kRaySegment - 1)) < __ldg(ro.seg_counts + (slot >> kRaySegmentShift));
}

__device__ __forceinline__ unsigned out_slot(const RayOrder& ro, unsigned rid) { return ro.out_index ? __ldg(ro.out_index + rid) : rid; }

__device__ __forceinline__ unsigned batch_length(const RayOrder& ro, unsigned R) {
    if (!ro.d_count) return R;
    const unsigned n = __ldg(ro.d_count);
    return n < R ? n : R;
}

__device__ __forceinline__ unsigned ray_of_slot(const RayOrder& ro, unsigned slot) {
    if (!ro.list) return slot;
    if (!ro.counts) return __ldg(ro.list + slot);
    unsigned base = 0, o = 0, next = 0;
#pragma unroll
    for (unsigned k = 0; k < 7; ++k) {
        next += __ldg(ro.counts + k);
        if (slot >= next) { base = next; o = k + 1; }
    }
    return __ldg(ro.list + (size_t)o * ro.stride + (slot - base));
}

struct RayState {
    V3 o, d, inv;
    bool nan_path;  // some operand can make 0*inf: use the literal GLSL min/max
};

// One 256-bit load (LDG.E.256, read-only path): a 32-byte record costs one L1 wavefront instead of two.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float4* p, float4 a, float4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z),
                 "f"(b.w)
                 : "memory");
}

// The per-triangle record: 48 bytes = three float4s {v0, e1.x | e1.yz, e2.xy | e2.z, n}.  (A 64-byte record read with two 256-bit
// loads — two L1 wavefronts instead of three, with the mesh id in the spare lane — was measured: 0.723 vs 0.717 ms on the diffuse
// batch, the larger footprint costs more L1 hits than the saved wavefronts give back.)
constexpr int kTriStride = 3;  // float4s per record
__device__ __forceinline__ void load_tri48(const float4* __restrict__ tri48, int idx, V3& v0, V3& e1, V3& e2, V3& n) {
    const float4 a = __ldg(tri48 + kTriStride * (size_t)idx), b = __ldg(tri48 + kTriStride * (size_t)idx + 1), c = __ldg(tri48 + kTriStride * (size_t)idx + 2);
    v0 = {a.x, a.y, a.z};
    e1 = {a.w, b.x, b.y};
    e2 = {b.z, b.w, c.x};
    n = {c.y, c.z, c.w};
}

// RayTriangle, SL:79-97 / ST:87-105. Returns t, or -1 when outside.
__device__ __forceinline__ float ray_triangle(const float4* __restrict__ tri48, int idx, V3 ro, V3 rd) {
    V3 v0, e1, e2, n;
    load_tri48(tri48, idx, v0, e1, e2, n);
    const V3 rov0 = vsub(ro, v0);
    const V3 q = vcross(rov0, rd);
    const float d = fdiv(1.0f, vdot(rd, n));
    const float u = fmul(d, vdot(vneg(q), e2));
    const float v = fmul(d, vdot(q, e1));
    float t = fmul(d, vdot(vneg(n), rov0));
    if (u < 0.0f || v < 0.0f || fadd(u, v) > 1.0f) t = -1.0f;
    return t;
}

// RayBounds, SL:100-109, followed by the caller's `Box > 0 && Box < TMax` (SL:207).
// Returns whether the node is entered.  `exact` selects the literal GLSL min/max forms, which is
// required whenever a slab product can be NaN (0 * inf) and correct always.
__device__ __forceinline__ bool enter_stackless(float4 mn, float4 mx, const RayState& r, float tmax_cur, bool exact) {
    const float t0x = fmul(fsub(mn.x, r.o.x), r.inv.x), t0y = fmul(fsub(mn.y, r.o.y), r.inv.y), t0z = fmul(fsub(mn.z, r.o.z), r.inv.z);
    const float t1x = fmul(fsub(mx.x, r.o.x), r.inv.x), t1y = fmul(fsub(mx.y, r.o.y), r.inv.y), t1z = fmul(fsub(mx.z, r.o.z), r.inv.z);
    if (!exact) {
        // No NaN can occur (finite origin, finite non-zero inverse direction): fminf/fmaxf equal the GLSL
        // forms up to the sign of a zero, which no comparison below can observe.  With tmin >= 0.0001:
        //   Box = (min(far, TMax) >= tmin) ? tmin : -1;  Box > 0 && Box < TMax   <=>   far >= tmin && tmin < TMax
        const float tmin = fmaxf(fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z)), 0.0001f);
        const float far = fminf(fmaxf(t0x, t1x), fminf(fmaxf(t0y, t1y), fmaxf(t0z, t1z)));
        return far >= tmin && tmin < tmax_cur;
    }
    const float tmin = glsl_max(glsl_max(glsl_max(glsl_min(t0x, t1x), glsl_min(t0y, t1y)), glsl_min(t0z, t1z)), 0.0001f);
    const float tmax = glsl_min(glsl_min(glsl_max(t0x, t1x), glsl_min(glsl_max(t0y, t1y), glsl_max(t0z, t1z))), tmax_cur);
    const float box = (tmax >= tmin) ? tmin : -1.0f;
    return box > 0.0f && box < tmax_cur;
}
__device__ __forceinline__ bool enter_stackless(float4 mn, float4 mx, const RayState& r, float tmax_cur) {
    return enter_stackless(mn, mx, r, tmax_cur, r.nan_path);
}

// RayBounds, ST:107-116.  `exact`: see enter_stackless.
__device__ __forceinline__ float slab_stack(float4 mn, float4 mx, const RayState& r, float maxt, bool exact) {
    const float fx = fmul(fsub(mx.x, r.o.x), r.inv.x), fy = fmul(fsub(mx.y, r.o.y), r.inv.y), fz = fmul(fsub(mx.z, r.o.z), r.inv.z);
    const float nx = fmul(fsub(mn.x, r.o.x), r.inv.x), ny = fmul(fsub(mn.y, r.o.y), r.inv.y), nz = fmul(fsub(mn.z, r.o.z), r.inv.z);
    float t0, t1;
    if (!exact) {
        t1 = fminf(fminf(fmaxf(fx, nx), fminf(fmaxf(fy, ny), fmaxf(fz, nz))), maxt);
        t0 = fmaxf(fmaxf(fminf(fx, nx), fmaxf(fminf(fy, ny), fminf(fz, nz))), 0.0f);
    } else {
        t1 = glsl_min(glsl_min(glsl_max(fx, nx), glsl_min(glsl_max(fy, ny), glsl_max(fz, nz))), maxt);
        t0 = glsl_max(glsl_max(glsl_min(fx, nx), glsl_max(glsl_min(fy, ny), glsl_min(fz, nz))), 0.0f);
    }
    return (t1 >= t0) ? (t0 > 0.0f ? t0 : t1) : -1.0f;
}
__device__ __forceinline__ float slab_stack(float4 mn, float4 mx, const RayState& r, float maxt) { return slab_stack(mn, mx, r, maxt, r.nan_path); }

__device__ __forceinline__ RayState to_object_space(const cndl_entity* __restrict__ e, V3 ro, V3 rd) {
    // SL:177-180: the direction is not renormalised, so t is the same in both spaces
    float m[16];
    const float4* mp = reinterpret_cast<const float4*>(e->inverse);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 c = __ldg(mp + k);
        m[4 * k + 0] = c.x; m[4 * k + 1] = c.y; m[4 * k + 2] = c.z; m[4 * k + 3] = c.w;
    }
    RayState r;
    r.o = xform(m, ro, 1.0f);
    r.d = xform(m, rd, 0.0f);
    r.inv = {fdiv(1.0f, r.d.x), fdiv(1.0f, r.d.y), fdiv(1.0f, r.d.z)};
    r.nan_path = !(finite3(r.o) && finite3(r.inv));
    return r;
}

struct EntityResult { float t; int tri; int iters; };

// Triangle loop of a leaf visit (SL:212-233, ST:221-277). ANY: returns true on first acceptance.
template <bool ANY>
__device__ __forceinline__ bool leaf_triangles(const SceneView& s, int pack, const RayState& r, float& tmax, EntityResult& res) {
    const int len = pack & 0xF, first = pack >> 4;
    for (int idx = first; idx < first + len; ++idx) {
        const float t = ray_triangle(s.tri48, idx, r.o, r.d);
        if (t > 0.0f && t < tmax) {
            tmax = t;
            res.t = t;
            res.tri = idx;
            if (ANY) return true;
        }
    }
    return false;
}

// IntersectBVHStackless SL:175-278 / IntersectBVHStacklessOcclusion SL:463-556
template <bool ANY>
__device__ __forceinline__ EntityResult walk_stackless(const SceneView& s, const RayState& r, int start, int count, float tmax) {
    EntityResult res{-1.0f, -1, 0};
    int ptr = start, iters = 0;
    while (ptr >= 0 && iters < 1024) {
        if (ptr < start || ptr > start + count || ptr > s.total_nodes) break;  // SL:196
        ++iters;
        const float4 mn = __ldg(s.nodes + 2 * (size_t)ptr), mx = __ldg(s.nodes + 2 * (size_t)ptr + 1);
        const int link = __float_as_int(mx.w);
        if (enter_stackless(mn, mx, r, tmax)) {
            const int pack = __float_as_int(mn.w);
            if (pack != -1) {
                if (leaf_triangles<ANY>(s, pack, r, tmax, res)) { res.iters = iters; return res; }
                ptr = link;
                if (ptr < 0) break;
                ptr += start;
            } else {
                ++ptr;
            }
        } else {
            ptr = link;
            if (ptr < 0) break;
            ptr += start;
        }
    }
    res.iters = iters;
    if (ANY) res.t = -1.0f;  // SL:555
    return res;
}

// IntersectBVHStack ST:168-324 / IntersectBVHStackOcclusion ST:509-657
template <bool ANY>
__device__ __forceinline__ EntityResult walk_stack(const SceneView& s, const RayState& r, int start, int count, float tmax) {
    EntityResult res{-1.0f, -1, 0};
    int stack[64];
    int sp = 0, iters = 0, cur = start;
    while (iters < 1024) {
        if (sp >= 64 || sp < 0 || cur < start || cur > start + count || cur < 0 || cur > s.total_nodes) break;  // ST:201-205
        ++iters;
        const float4* np = s.nodes + 4 * (size_t)cur;
        const float4 lmn = __ldg(np), lmx = __ldg(np + 1), rmn = __ldg(np + 2), rmx = __ldg(np + 3);
        const int lpack = __float_as_int(lmn.w), rpack = __float_as_int(rmn.w);
        const bool lleaf = lpack != -1, rleaf = rpack != -1;
        // box tests see TMax as it was before this node's leaves are intersected (ST:215-216)
        const float lt = lleaf ? -1.0f : slab_stack(lmn, lmx, r, tmax);
        const float rt = rleaf ? -1.0f : slab_stack(rmn, rmx, r, tmax);
        bool done = false;
        if (lleaf) done = leaf_triangles<ANY>(s, lpack, r, tmax, res);
        if (!done && rleaf) done = leaf_triangles<ANY>(s, rpack, r, tmax, res);
        if (ANY && done) { res.iters = iters; return res; }
        const int lslot = __float_as_int(lmx.w) + start, rslot = __float_as_int(rmx.w) + start;
        if (lt > 0.0f && rt > 0.0f) {  // ST:280-299
            int postponed = rslot;
            cur = lslot;
            if (rt < lt) { cur = rslot; postponed = lslot; }
            if (sp >= 63) break;
            stack[sp++] = postponed;
            continue;
        } else if (lt > 0.0f) {
            cur = lslot;
            continue;
        } else if (rt > 0.0f) {
            cur = rslot;
            continue;
        }
        if (sp <= 0) break;
        cur = stack[--sp];
    }
    res.iters = iters;
    if (ANY) res.t = -1.0f;
    return res;
}

// ComputeBarycentrics, SL:156-172, with b - a and c - a taken from tri48 (same bits).
__device__ __forceinline__ void barycentrics(const float4* __restrict__ tri48, int idx, V3 p, float& u, float& v, float& w) {
    V3 a, v0, v1, n;
    load_tri48(tri48, idx, a, v0, v1, n);
    const V3 v2 = vsub(p, a);
    const float d00 = vdot(v0, v0), d01 = vdot(v0, v1), d11 = vdot(v1, v1), d20 = vdot(v2, v0), d21 = vdot(v2, v1);
    const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
    v = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
    w = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
    u = fsub(fsub(1.0f, v), w);
}

// IntersectScene / IntersectSceneIgnoreTransparent, SL:280-366, ST:327-413
template <bool STACK>
__device__ __forceinline__ cndl_hit scene_closest(const SceneView& s, V3 ro, V3 rd, bool ignore_transparent) {
    float closest = -1.0f, tmax = 1000000.0f;
    cndl_hit h{-1.0f, -1.0f, -1.0f, -1.0f, -1, -1, -1, 0};
    for (int i = 0; i < s.n_ents; ++i) {
        const cndl_entity* e = s.ents + i;
        if (ignore_transparent && __int_as_float(__ldg(&e->data[1])) < 0.99f) continue;  // SL:333-337
        const RayState r = to_object_space(e, ro, rd);
        const int start = __ldg(&e->node_offset), count = __ldg(&e->node_count);
        const EntityResult er = STACK ? walk_stack<false>(s, r, start, count, tmax) : walk_stackless<false>(s, r, start, count, tmax);
        h.iters = er.iters;
        if (er.t > 0.0f && er.t < tmax) {
            tmax = er.t;
            closest = er.t;
            h.tri = er.tri;
            h.entity = i;
        }
    }
    if (h.tri >= 0) h.mesh = __ldg(&s.tris[h.tri]).w;
    if (closest > 0.0f && h.tri > 0) {  // SL:300 — global triangle 0 reports as a miss
        const RayState r = to_object_space(s.ents + h.entity, ro, rd);
        const V3 p = {fadd(r.o.x, fmul(r.d.x, closest)), fadd(r.o.y, fmul(r.d.y, closest)), fadd(r.o.z, fmul(r.d.z, closest))};
        h.t = closest;
        barycentrics(s.tri48, h.tri, p, h.u, h.v, h.w);
    }
    return h;
}

// any-hit IntersectScene, SL:558-575, ST:659-676 (+ per-ray tmax extension)
template <bool STACK>
__device__ __forceinline__ float scene_any(const SceneView& s, V3 ro, V3 rd, float ray_tmax) {
    const float tmax = ray_tmax > 0.0f ? ray_tmax : 1000000.0f;
    for (int i = 0; i < s.n_ents; ++i) {
        const cndl_entity* e = s.ents + i;
        const RayState r = to_object_space(e, ro, rd);
        const int start = __ldg(&e->node_offset), count = __ldg(&e->node_count);
        const EntityResult er = STACK ? walk_stack<true>(s, r, start, count, tmax) : walk_stackless<true>(s, r, start, count, tmax);
        if (er.t > 0.0f) return er.t;
    }
    return -1.0f;
}

}  // namespace cndl
