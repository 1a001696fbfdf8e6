// Traversal kernels for sm_100a.  See traverse.cuh for the arithmetic contract.
#include "kernels.cuh"

namespace cndl {

namespace {

__device__ __forceinline__ void load_ray(const cndl_ray* __restrict__ rays, size_t i, V3& o, V3& d, float& tmax) {
    const float4* p = reinterpret_cast<const float4*>(rays + i);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    o = {a.x, a.y, a.z};
    d = {b.x, b.y, b.z};
    tmax = b.w;
}

__device__ __forceinline__ void store_hit(cndl_hit* __restrict__ hits, size_t i, const cndl_hit& h) {
    float4* p = reinterpret_cast<float4*>(hits + i);
    p[0] = make_float4(h.t, h.u, h.v, h.w);
    reinterpret_cast<int4*>(p)[1] = make_int4(h.mesh, h.tri, h.entity, h.iters);
}

// ---------------------------------------------------------------------------------------------
// Mode 0: one thread per ray.
template <bool STACK, int KIND>
__global__ void __launch_bounds__(128) trace_simple_kernel(SceneView s, const cndl_ray* __restrict__ rays, size_t R,
                                                           const unsigned* __restrict__ order, cndl_hit* __restrict__ hits,
                                                           float* __restrict__ any_t) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= R) return;
    const size_t i = order ? (size_t)__ldg(order + slot) : slot;
    V3 o, d;
    float ray_tmax;
    load_ray(rays, i, o, d, ray_tmax);
    if (KIND == Q_ANY) {
        any_t[i] = scene_any<STACK>(s, o, d, ray_tmax);
    } else {
        store_hit(hits, i, scene_closest<STACK>(s, o, d, KIND == Q_CLOSEST_IGNORE_TRANSPARENT));
    }
}

// ---------------------------------------------------------------------------------------------
// Mode 1: persistent warps, stackless walk as a per-lane state machine.  A lane whose ray has
// finished goes idle; when enough lanes of the warp are idle the warp claims a fresh run of rays
// from a global counter (one atomic per refill), so divergence in walk length does not leave
// lanes empty until the longest ray of the warp ends.  The per-ray sequence of node visits,
// triangle tests and TMax updates is exactly walk_stackless()'s.
struct Lane {
    RayState r;
    float tmax, closest;
    int ptr, start, count, iters, ent;
    int best_tri, best_ent;
    unsigned rid;  // ray index, 0xFFFFFFFF = idle
};

template <int KIND>
__device__ __forceinline__ bool begin_entity(const SceneView& s, const cndl_ray* __restrict__ rays, Lane& L) {
    // advances L.ent to the next entity to traverse; false when the scene loop is over
    while (L.ent < s.n_ents) {
        const cndl_entity* e = s.ents + L.ent;
        if (KIND == Q_CLOSEST_IGNORE_TRANSPARENT && __int_as_float(__ldg(&e->data[1])) < 0.99f) { ++L.ent; continue; }
        V3 o, d;
        float unused;
        load_ray(rays, L.rid, o, d, unused);
        L.r = to_object_space(e, o, d);
        L.start = __ldg(&e->node_offset);
        L.count = __ldg(&e->node_count);
        L.ptr = L.start;
        L.iters = 0;
        return true;
    }
    return false;
}

template <int KIND>
__global__ void __launch_bounds__(128, 4) trace_persistent_stackless_kernel(SceneView s, const cndl_ray* __restrict__ rays, unsigned R,
                                                                            const unsigned* __restrict__ order,
                                                                            cndl_hit* __restrict__ hits, float* __restrict__ any_t,
                                                                            unsigned* __restrict__ work_counter) {
    constexpr bool ANY = KIND == Q_ANY;
    const unsigned lane = threadIdx.x & 31u;
    Lane L;
    L.rid = 0xFFFFFFFFu;
    L.iters = 0;
    bool drained = false;  // the global counter ran past R

    while (true) {
        // ---- refill idle lanes ----
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, L.rid == 0xFFFFFFFFu);
        if (idle == 0xFFFFFFFFu && drained) break;
        if (!drained && (__popc(idle) >= 8 || idle == 0xFFFFFFFFu)) {
            unsigned base = 0;
            const int n = __popc(idle);
            if (lane == 0) base = atomicAdd(work_counter, (unsigned)n);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + (unsigned)n >= R) drained = true;  // uniform across the warp
            if (L.rid == 0xFFFFFFFFu) {
                const unsigned slot = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (slot < R) {
                    L.rid = order ? __ldg(order + slot) : slot;
                    L.ent = 0;
                    L.closest = -1.0f;
                    L.best_tri = -1;
                    L.best_ent = -1;
                    L.iters = 0;
                    if (ANY) {
                        const float rt = __ldg(&rays[L.rid].tmax);
                        L.tmax = rt > 0.0f ? rt : 1000000.0f;
                    } else {
                        L.tmax = 1000000.0f;
                    }
                    if (!begin_entity<KIND>(s, rays, L)) {
                        // no entity to traverse: miss
                        if (ANY) any_t[L.rid] = -1.0f;
                        else store_hit(hits, L.rid, cndl_hit{-1.0f, -1.0f, -1.0f, -1.0f, -1, -1, -1, 0});
                        L.rid = 0xFFFFFFFFu;
                    }
                }
            }
        }

        // ---- a bounded burst of node steps ----
#pragma unroll 1
        for (int burst = 0; burst < 16; ++burst) {
            if (L.rid != 0xFFFFFFFFu) {
                bool entity_done = false;
                bool any_found = false;
                // loop header of SL:192-199
                if (!(L.ptr >= 0 && L.iters < 1024) || L.ptr < L.start || L.ptr > L.start + L.count || L.ptr > s.total_nodes) {
                    entity_done = true;
                } else {
                    ++L.iters;
                    const float4 mn = __ldg(s.nodes + 2 * (size_t)L.ptr), mx = __ldg(s.nodes + 2 * (size_t)L.ptr + 1);
                    const int link = __float_as_int(mx.w);
                    bool follow_link = true;
                    if (enter_stackless(mn, mx, L.r, L.tmax)) {
                        const int pack = __float_as_int(mn.w);
                        if (pack != -1) {
                            EntityResult er{-1.0f, -1, 0};
                            any_found = leaf_triangles<ANY>(s, pack, L.r, L.tmax, er);
                            if (er.tri >= 0) { L.closest = er.t; L.best_tri = er.tri; L.best_ent = L.ent; }
                        } else {
                            ++L.ptr;
                            follow_link = false;
                        }
                    }
                    if (follow_link) {
                        L.ptr = link;
                        if (link < 0) entity_done = true;
                        else L.ptr += L.start;
                    }
                }
                if (ANY && any_found) {
                    any_t[L.rid] = L.closest;  // SL:567-569: the scene loop returns the first T > 0
                    L.rid = 0xFFFFFFFFu;
                } else if (entity_done) {
                    // scene loop bookkeeping.  In the reference the per-entity result is accepted when
                    // T > 0 && T < TMax(scene); the entity walk started from TMax(scene) and only
                    // accepts t < TMax, so every accepted hit already satisfies that test and L.tmax
                    // is the scene TMax carried into the next entity (SL:293-301).
                    const int last_iters = L.iters;
                    ++L.ent;
                    if (!begin_entity<KIND>(s, rays, L)) {
                        if (ANY) {
                            any_t[L.rid] = -1.0f;
                        } else {
                            cndl_hit h{-1.0f, -1.0f, -1.0f, -1.0f, -1, L.best_tri, L.best_ent, last_iters};
                            if (L.best_tri >= 0) h.mesh = __ldg(&s.tris[L.best_tri]).w;
                            if (L.closest > 0.0f && L.best_tri > 0) {
                                V3 o, d;
                                float unused;
                                load_ray(rays, L.rid, o, d, unused);
                                const RayState r = to_object_space(s.ents + L.best_ent, o, d);
                                const V3 p = {fadd(r.o.x, fmul(r.d.x, L.closest)), fadd(r.o.y, fmul(r.d.y, L.closest)),
                                              fadd(r.o.z, fmul(r.d.z, L.closest))};
                                h.t = L.closest;
                                barycentrics(s.tri48, L.best_tri, p, h.u, h.v, h.w);
                            }
                            store_hit(hits, L.rid, h);
                        }
                        L.rid = 0xFFFFFFFFu;
                    }
                }
            }
        }
    }
}

__global__ void make_tri48_kernel(const int4* __restrict__ tris, const float4* __restrict__ verts, size_t T, unsigned V, float4* __restrict__ tri48,
                                  int* __restrict__ invalid) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int4 t = tris[i];
    if ((unsigned)t.x >= V || (unsigned)t.y >= V || (unsigned)t.z >= V) {  // a vertex index outside the vertex buffer (corrupt cache file / prebuilt buffer)
        atomicOr(invalid, 4);
        tri48[3 * i + 0] = tri48[3 * i + 1] = tri48[3 * i + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return;
    }
    const float4 a = verts[2 * (size_t)t.x], b = verts[2 * (size_t)t.y], c = verts[2 * (size_t)t.z];
    const V3 v0 = {a.x, a.y, a.z};
    const V3 e1 = vsub(V3{b.x, b.y, b.z}, v0);  // v1v0, SL:81
    const V3 e2 = vsub(V3{c.x, c.y, c.z}, v0);  // v2v0, SL:82
    const V3 n = vcross(e1, e2);                 // SL:85
    tri48[3 * i + 0] = make_float4(v0.x, v0.y, v0.z, e1.x);
    tri48[3 * i + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    tri48[3 * i + 2] = make_float4(e2.z, n.x, n.y, n.z);
}

__global__ void rebase_triangles_kernel(int4* __restrict__ tris, size_t T, int offset) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    int4 t = tris[i];
    t.x += offset; t.y += offset; t.z += offset;
    tris[i] = t;
}

// ---------------------------------------------------------------------------------------------
// Physics::CollideBox / CollideBVH (Source/Core/Physics.cpp:21-228): does an axis-aligned box touch the scene?  The
// same threaded walk with a box-overlap predicate; one thread per query box.  Quirks kept: the box is taken to object
// space corner by corner (:84-85), `e` is the full extent (:31), nine cross-product axes only, the sixth repeating
// cross(u2, f2) (:46-54); the first overlapping triangle in walk order wins.
__device__ __forceinline__ bool box_triangle_overlap(V3 v0, V3 v1, V3 v2, V3 bmin, V3 bmax) {
    const V3 c = {fdiv(fadd(bmin.x, bmax.x), 2.0f), fdiv(fadd(bmin.y, bmax.y), 2.0f), fdiv(fadd(bmin.z, bmax.z), 2.0f)};
    const V3 e = vsub(bmax, bmin);
    v0 = vsub(v0, c); v1 = vsub(v1, c); v2 = vsub(v2, c);
    const V3 f0 = vsub(v1, v0), f1 = vsub(v2, v1), f2 = vsub(v0, v2);
    const V3 u0 = {1.0f, 0.0f, 0.0f}, u1 = {0.0f, 1.0f, 0.0f}, u2 = {0.0f, 0.0f, 1.0f};
    const V3 axes[9] = {vcross(u0, f0), vcross(u0, f1), vcross(u0, f2), vcross(u1, f0), vcross(u1, f1), vcross(u2, f2), vcross(u2, f0), vcross(u2, f1), vcross(u2, f2)};
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const V3 a = axes[i];
        const float p0 = vdot(v0, a), p1 = vdot(v1, a), p2 = vdot(v2, a);
        const float r = fadd(fadd(fmul(e.x, fabsf(vdot(u0, a))), fmul(e.y, fabsf(vdot(u1, a)))), fmul(e.z, fabsf(vdot(u2, a))));
        if (glsl_max(-glsl_max(glsl_max(p0, p1), p2), glsl_min(glsl_min(p0, p1), p2)) > r) return false;
    }
    return true;
}

__global__ void collide_boxes_kernel(SceneView s, const float4* __restrict__ verts, const float4* __restrict__ boxes, unsigned n, int4* __restrict__ out) {
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float4 b0 = __ldg(boxes + 2 * (size_t)q), b1 = __ldg(boxes + 2 * (size_t)q + 1);
    int4 res = make_int4(0, -1, -1, -1);
    for (int ei = 0; ei < s.n_ents && !res.x; ++ei) {  // CollideBox :203-228
        const cndl_entity* en = s.ents + ei;
        float m[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m[k] = __ldg(&en->inverse[k]);
        const V3 cmin = xform(m, V3{b0.x, b0.y, b0.z}, 1.0f), cmax = xform(m, V3{b1.x, b1.y, b1.z}, 1.0f);
        const int start = __ldg(&en->node_offset), count = __ldg(&en->node_count);
        int ptr = start, iters = 0;
        while (ptr >= 0 && iters < 1024) {
            // :101 also admits Pointer == m_BVHNodes.size(), an out-of-bounds read; the walk stops there instead
            if (ptr < start || ptr > start + count || ptr >= s.total_nodes) break;
            ++iters;
            const float4 mn = __ldg(s.nodes + 2 * (size_t)ptr), mx = __ldg(s.nodes + 2 * (size_t)ptr + 1);
            const int link = __float_as_int(mx.w);
            const bool overlap = (mn.x <= cmax.x && mx.x >= cmin.x) && (mn.y <= cmax.y && mx.y >= cmin.y) && (mn.z <= cmax.z && mx.z >= cmin.z);  // :23-27
            if (overlap) {
                const int pack = __float_as_int(mn.w);
                if (pack != -1) {
                    const int len = pack & 0xF;
                    for (int idx = pack >> 4; idx < (pack >> 4) + len; ++idx) {
                        const int4 t = __ldg(s.tris + idx);
                        const float4 a = __ldg(verts + 2 * (size_t)t.x), b = __ldg(verts + 2 * (size_t)t.y), c = __ldg(verts + 2 * (size_t)t.z);
                        if (box_triangle_overlap(V3{a.x, a.y, a.z}, V3{b.x, b.y, b.z}, V3{c.x, c.y, c.z}, cmin, cmax)) {
                            res = make_int4(1, t.w, idx, ei);
                            break;
                        }
                    }
                    if (res.x) break;
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
    }
    out[q] = res;
}

__global__ void primary_rays_kernel(Mat2 m, int W, int H, cndl_ray* __restrict__ rays) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    float4 o, d;
    primary_ray(m, x, y, W, H, o, d);
    float4* p = reinterpret_cast<float4*>(rays + ((size_t)y * (size_t)W + (size_t)x));
    p[0] = o;
    p[1] = d;
}

template <bool STACK>
void dispatch_simple(const SceneView& s, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits, float* any_t,
                     cudaStream_t stream) {
    const unsigned block = 128;
    const unsigned grid = (unsigned)((R + block - 1) / block);
    switch (kind) {
        case Q_CLOSEST: trace_simple_kernel<STACK, Q_CLOSEST><<<grid, block, 0, stream>>>(s, rays, R, order, hits, any_t); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: trace_simple_kernel<STACK, Q_CLOSEST_IGNORE_TRANSPARENT><<<grid, block, 0, stream>>>(s, rays, R, order, hits, any_t); break;
        default: trace_simple_kernel<STACK, Q_ANY><<<grid, block, 0, stream>>>(s, rays, R, order, hits, any_t); break;
    }
}

}  // namespace

void launch_trace_simple(const SceneView& s, bool stack, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits,
                         float* any_t, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    if (stack) dispatch_simple<true>(s, kind, rays, R, order, hits, any_t, stream);
    else dispatch_simple<false>(s, kind, rays, R, order, hits, any_t, stream);
    lc.n++;
}

void launch_trace_persistent(const SceneView& s, bool stack, int kind, const cndl_ray* rays, size_t R, const unsigned* order, cndl_hit* hits,
                             float* any_t, unsigned* work_counter, int sm_count, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    if (stack) {  // the stack walk keeps a 64-entry private stack: one ray per thread is the fitting shape
        launch_trace_simple(s, stack, kind, rays, R, order, hits, any_t, stream, lc);
        return;
    }
    cudaMemsetAsync(work_counter, 0, sizeof(unsigned), stream);
    const unsigned block = 128;
    unsigned grid = (unsigned)sm_count * 8u;  // 8 CTAs x 4 warps resident per SM
    const unsigned need = (unsigned)((R + block - 1) / block);
    if (grid > need) grid = need;
    switch (kind) {
        case Q_CLOSEST: trace_persistent_stackless_kernel<Q_CLOSEST><<<grid, block, 0, stream>>>(s, rays, (unsigned)R, order, hits, any_t, work_counter); break;
        case Q_CLOSEST_IGNORE_TRANSPARENT: trace_persistent_stackless_kernel<Q_CLOSEST_IGNORE_TRANSPARENT><<<grid, block, 0, stream>>>(s, rays, (unsigned)R, order, hits, any_t, work_counter); break;
        default: trace_persistent_stackless_kernel<Q_ANY><<<grid, block, 0, stream>>>(s, rays, (unsigned)R, order, hits, any_t, work_counter); break;
    }
    lc.n++;
}

void launch_make_tri48(const int4* tris, const float4* verts, size_t T, size_t V, float4* tri48, int* d_invalid, cudaStream_t stream, LaunchCounter& lc) {
    if (T == 0) return;
    make_tri48_kernel<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(tris, verts, T, (unsigned)V, tri48, d_invalid);
    lc.n++;
}

void launch_rebase_triangles(int4* tris, size_t T, int offset, cudaStream_t stream, LaunchCounter& lc) {
    if (T == 0 || offset == 0) return;
    rebase_triangles_kernel<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(tris, T, offset);
    lc.n++;
}

void launch_collide_boxes(const SceneView& s, const float4* verts, const cndl_box* boxes, size_t n, cndl_collision* out, cudaStream_t stream,
                          LaunchCounter& lc) {
    if (n == 0) return;
    collide_boxes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(s, verts, reinterpret_cast<const float4*>(boxes), (unsigned)n,
                                                                        reinterpret_cast<int4*>(out));
    lc.n++;
}

void launch_primary_rays(const float* iv, const float* ip, int W, int H, cndl_ray* rays, cudaStream_t stream, LaunchCounter& lc) {
    if (W <= 0 || H <= 0) return;
    Mat2 m;
    for (int k = 0; k < 16; ++k) { m.iv[k] = iv[k]; m.ip[k] = ip[k]; }
    const dim3 block(16, 16);  // the reference's local size (Intersectors/TraverseBVHStack.glsl:5)
    const dim3 grid((unsigned)((W + 15) / 16), (unsigned)((H + 15) / 16));
    primary_rays_kernel<<<grid, block, 0, stream>>>(m, W, H, rays);
    lc.n++;
}

}  // namespace cndl
