// TEST INFRASTRUCTURE ONLY — see oracle.h.  PINNED: the reference's traversal exists only as GLSL, but its two
// shader include files compile as C++ against the reference's own glm after a purely syntactic rewrite
// (oracle/ref_shim/glsl_to_cpp.py, ref_glsl.cpp -> oracle/_ref); tests/test_oracle_traversal.py requires this
// restatement to return bit-identical hit records, any-hit distances and GetData outputs.
//
// Scalar C++ restatement of the reference's GLSL traversal.  Paths are
// relative to /root/reference/Source/Core/Shaders/Intersectors/Include/ :
//   SL = TraverseBVHStackless.glsl,  ST = TraverseBVHStack.glsl.
// GLSL leaves FP contraction and min/max-on-NaN to the implementation; the
// oracle fixes them: no contraction (-ffp-contract=off), IEEE division,
// min(x,y) = x<y?x:y and max(x,y) = x>y?x:y (glm's forms; GLSL leaves NaN operands undefined), dot and
// mat*vec summed in glm 0.9.8.5's order.
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct V3 { float x, y, z; };
struct Vertex32 { float pos[4]; uint32_t packed[4]; };
struct Tri16 { int32_t v[4]; };
struct Node32 { float mn[4]; float mx[4]; };
struct Node64 { Node32 l, r; };
struct Entity192 { float model[16]; float inv[16]; int32_t node_offset, node_count; int32_t data[14]; };
static_assert(sizeof(Entity192) == 192, "BVHEntity is 192 bytes (Intersector.h:43-49)");

inline int32_t fbits(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline float ibits(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

// min / max as the reference's vendored glm 0.9.8.5 evaluates them (glm/detail/func_common.inl:15-28), which is what the
// compiled reference shaders (oracle/_ref, ref_glsl.cpp) execute.  GLSL leaves the result undefined when an operand is
// NaN; the two forms differ from `y < x ? y : x` / `x < y ? y : x` only there (and in the sign of a zero).
inline float smin(float x, float y) { return x < y ? x : y; }
inline float smax(float x, float y) { return x > y ? x : y; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 pos3(const Vertex32& v) { return {v.pos[0], v.pos[1], v.pos[2]}; }

// vec3(M * vec4(v, w)), column-major M, glm's summation order
// (glm/detail/type_mat4x4.inl:526-540): (c0*x + c1*y) + (c2*z + c3*w).
inline V3 xform(const float* m, V3 v, float w) {
    V3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
    return r;
}

// RayTriangle, SL:79-97 (identical in ST:87-105). Returns t or -1.
inline float ray_triangle(V3 ro, V3 rd, V3 v0, V3 v1, V3 v2) {
    const V3 v1v0 = sub(v1, v0), v2v0 = sub(v2, v0), rov0 = sub(ro, v0);
    const V3 n = cross(v1v0, v2v0);
    const V3 q = cross(rov0, rd);
    const float d = 1.0f / dot(rd, n);
    const float u = d * dot(neg(q), v2v0);
    const float v = d * dot(q, v1v0);
    float t = d * dot(neg(n), rov0);
    if (u < 0.0f || v < 0.0f || (u + v) > 1.0f) t = -1.0f;
    return t;
}

// RayBounds, SL:100-109
inline float slab_stackless(const float* mn, const float* mx, V3 o, V3 inv, float t_min, float t_max) {
    const float t0x = (mn[0] - o.x) * inv.x, t0y = (mn[1] - o.y) * inv.y, t0z = (mn[2] - o.z) * inv.z;
    const float t1x = (mx[0] - o.x) * inv.x, t1y = (mx[1] - o.y) * inv.y, t1z = (mx[2] - o.z) * inv.z;
    const float lox = smin(t0x, t1x), loy = smin(t0y, t1y), loz = smin(t0z, t1z);
    const float hix = smax(t0x, t1x), hiy = smax(t0y, t1y), hiz = smax(t0z, t1z);
    const float tmin = smax(smax(smax(lox, loy), loz), t_min);   // max3 = max(max(x,y),z), SL:68-71
    const float tmax = smin(smin(hix, smin(hiy, hiz)), t_max);   // min3 = min(x,min(y,z)), SL:73-76
    return (tmax >= tmin) ? tmin : -1.0f;
}

// RayBounds, ST:107-116
inline float slab_stack(V3 ro, V3 inv, const Node32& box, float maxt) {
    const float fx = (box.mx[0] - ro.x) * inv.x, fy = (box.mx[1] - ro.y) * inv.y, fz = (box.mx[2] - ro.z) * inv.z;
    const float nx = (box.mn[0] - ro.x) * inv.x, ny = (box.mn[1] - ro.y) * inv.y, nz = (box.mn[2] - ro.z) * inv.z;
    const float hx = smax(fx, nx), hy = smax(fy, ny), hz = smax(fz, nz);
    const float lx = smin(fx, nx), ly = smin(fy, ny), lz = smin(fz, nz);
    const float t1 = smin(smin(hx, smin(hy, hz)), maxt);
    const float t0 = smax(smax(lx, smax(ly, lz)), 0.0f);
    return (t1 >= t0) ? (t0 > 0.0f ? t0 : t1) : -1.0f;
}

// ComputeBarycentrics, SL:156-172
inline void barycentrics(V3 p, V3 a, V3 b, V3 c, float& u, float& v, float& w) {
    const V3 v0 = sub(b, a), v1 = sub(c, a), v2 = sub(p, a);
    const float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    const float denom = d00 * d11 - d01 * d01;
    v = (d11 * d20 - d01 * d21) / denom;
    w = (d00 * d21 - d01 * d20) / denom;
    u = 1.0f - v - w;
}

struct Scene {
    const Node32* n32;
    const Node64* n64;
    int64_t total_nodes;
    const Tri16* tris;
    const Vertex32* verts;
    const Entity192* ents;
    int32_t n_ents;
};

struct Tally { uint64_t nodes = 0, tri_tests = 0; bool capped = false; };

struct EntityResult { float t; int32_t mesh, tri, iters; };

// Triangle loop shared by every leaf visit (SL:212-233, ST:221-277).  `any`
// returns true on the first accepted triangle (SL:526-531).
inline bool leaf_triangles(const Scene& s, int32_t pack, V3 o, V3 d, float& tmax, EntityResult& r, Tally& tally, bool any) {
    const int32_t len = pack & 0xF, first = pack >> 4;
    for (int32_t idx = first; idx < first + len; ++idx) {
        const Tri16& tri = s.tris[idx];
        tally.tri_tests++;
        const float t = ray_triangle(o, d, pos3(s.verts[tri.v[0]]), pos3(s.verts[tri.v[1]]), pos3(s.verts[tri.v[2]]));
        if (t > 0.0f && t < tmax) {
            tmax = t;
            r.t = t;
            r.mesh = tri.v[3];
            r.tri = idx;
            if (any) return true;
        }
    }
    return false;
}

// IntersectBVHStackless SL:175-278 and IntersectBVHStacklessOcclusion SL:463-556
EntityResult walk_stackless(const Scene& s, V3 ro, V3 rd, const Entity192& e, float tmax, bool any, Tally& tally) {
    const V3 o = xform(e.inv, ro, 1.0f);
    const V3 d = xform(e.inv, rd, 0.0f);
    const V3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    const int32_t start = e.node_offset, count = e.node_count;
    EntityResult r{-1.0f, -1, -1, 0};
    int32_t ptr = start;
    int32_t iters = 0;
    while (ptr >= 0 && iters < 1024) {
        if (ptr < start || ptr > start + count || ptr < 0 || (int64_t)ptr > s.total_nodes) break;  // SL:196
        iters++;
        const Node32& n = s.n32[ptr];
        const float box = slab_stackless(n.mn, n.mx, o, inv, 0.0001f, tmax);
        if (box > 0.0f && box < tmax) {
            const int32_t pack = fbits(n.mn[3]);
            if (pack != -1) {
                if (leaf_triangles(s, pack, o, d, tmax, r, tally, any)) {
                    tally.nodes += (uint64_t)iters;
                    r.iters = iters;
                    return r;  // any-hit early return, SL:530
                }
                ptr = fbits(n.mx[3]);
                if (ptr < 0) break;
                ptr += start;
            } else {
                ptr++;
            }
        } else {
            ptr = fbits(n.mx[3]);
            if (ptr < 0) break;
            ptr += start;
        }
    }
    if (iters >= 1024) tally.capped = true;
    tally.nodes += (uint64_t)iters;
    r.iters = iters;
    if (any) r.t = -1.0f;  // SL:555: the occlusion walk returns -1 unless it returned early
    return r;
}

// IntersectBVHStack ST:168-324 and IntersectBVHStackOcclusion ST:509-657
EntityResult walk_stack(const Scene& s, V3 ro, V3 rd, const Entity192& e, float tmax, bool any, Tally& tally) {
    const V3 o = xform(e.inv, ro, 1.0f);
    const V3 d = xform(e.inv, rd, 0.0f);
    const V3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    const int32_t start = e.node_offset, count = e.node_count;
    EntityResult r{-1.0f, -1, -1, 0};
    int32_t stack[64];
    int32_t sp = 0;
    int32_t iters = 0;
    int32_t cur = start;
    while (iters < 1024) {
        if (sp >= 64 || sp < 0 || cur < start || cur > start + count || cur < 0 || (int64_t)cur > s.total_nodes) break;  // ST:201-205
        iters++;
        const Node64& n = s.n64[cur];
        const int32_t lpack = fbits(n.l.mn[3]), rpack = fbits(n.r.mn[3]);
        const bool lleaf = lpack != -1, rleaf = rpack != -1;
        // Box tests use TMax as it was before this node's leaves are intersected (ST:215-216).
        const float lt = lleaf ? -1.0f : slab_stack(o, inv, n.l, tmax);
        const float rt = rleaf ? -1.0f : slab_stack(o, inv, n.r, tmax);
        bool done = false;
        if (lleaf) done = leaf_triangles(s, lpack, o, d, tmax, r, tally, any);
        if (!done && rleaf) done = leaf_triangles(s, rpack, o, d, tmax, r, tally, any);
        if (done) {
            tally.nodes += (uint64_t)iters;
            r.iters = iters;
            return r;
        }
        if (lt > 0.0f && rt > 0.0f) {  // ST:280-299
            cur = fbits(n.l.mx[3]) + start;
            int32_t postponed = fbits(n.r.mx[3]) + start;
            if (rt < lt) { const int32_t tmp = cur; cur = postponed; postponed = tmp; }
            if (sp >= 63) break;
            stack[sp++] = postponed;
            continue;
        } else if (lt > 0.0f) {
            cur = fbits(n.l.mx[3]) + start;
            continue;
        } else if (rt > 0.0f) {
            cur = fbits(n.r.mx[3]) + start;
            continue;
        }
        if (sp <= 0) break;
        cur = stack[--sp];
    }
    if (iters >= 1024) tally.capped = true;
    tally.nodes += (uint64_t)iters;
    r.iters = iters;
    if (any) r.t = -1.0f;
    return r;
}

// IntersectScene / IntersectSceneIgnoreTransparent, SL:280-366, ST:327-413
orc_hit scene_closest(const Scene& s, bool stack, bool ignore_transparent, const orc_ray& ray, Tally& tally) {
    const V3 ro = {ray.ox, ray.oy, ray.oz}, rd = {ray.dx, ray.dy, ray.dz};
    float closest = -1.0f, tmax = 1000000.0f;
    orc_hit h{-1.0f, -1.0f, -1.0f, -1.0f, -1, -1, -1, 0};
    for (int32_t i = 0; i < s.n_ents; ++i) {
        const Entity192& e = s.ents[i];
        if (ignore_transparent && ibits(e.data[1]) < 0.99f) continue;  // SL:333-337
        const EntityResult r = stack ? walk_stack(s, ro, rd, e, tmax, false, tally) : walk_stackless(s, ro, rd, e, tmax, false, tally);
        h.iters = r.iters;  // `out Iters` is overwritten by every call
        if (r.t > 0.0f && r.t < tmax) {
            tmax = r.t;
            closest = r.t;
            h.mesh = r.mesh;
            h.tri = r.tri;
            h.entity = i;
        }
    }
    if (closest > 0.0f && h.tri > 0) {  // SL:300 — global triangle 0 reports as a miss
        const Entity192& e = s.ents[h.entity];
        const V3 o = xform(e.inv, ro, 1.0f), d = xform(e.inv, rd, 0.0f);
        const Tri16& tri = s.tris[h.tri];
        const V3 p = {o.x + d.x * closest, o.y + d.y * closest, o.z + d.z * closest};
        h.t = closest;
        barycentrics(p, pos3(s.verts[tri.v[0]]), pos3(s.verts[tri.v[1]]), pos3(s.verts[tri.v[2]]), h.u, h.v, h.w);
    }
    return h;
}

// any-hit IntersectScene, SL:558-575, ST:659-676
float scene_any(const Scene& s, bool stack, const orc_ray& ray, Tally& tally) {
    const V3 ro = {ray.ox, ray.oy, ray.oz}, rd = {ray.dx, ray.dy, ray.dz};
    const float tmax = ray.tmax > 0.0f ? ray.tmax : 1000000.0f;  // reference: always 1e6
    for (int32_t i = 0; i < s.n_ents; ++i) {
        const EntityResult r = stack ? walk_stack(s, ro, rd, s.ents[i], tmax, true, tally) : walk_stackless(s, ro, rd, s.ents[i], tmax, true, tally);
        if (r.t > 0.0f) return r.t;
    }
    return -1.0f;
}

template <class F>
void parallel_ranges(uint64_t R, int nthreads, F&& body) {
    if (nthreads <= 1 || R < 1024) { body(0, R, 0); return; }
    std::vector<std::thread> pool;
    const uint64_t chunk = (R + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const uint64_t lo = chunk * (uint64_t)t, hi = lo + chunk < R ? lo + chunk : R;
        if (lo >= hi) break;
        pool.emplace_back([=, &body] { body(lo, hi, t); });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int orc_hardware_threads(void) {
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

void orc_trace(int format, int kind, const void* nodes, uint64_t total_nodes, const void* tris, const void* verts,
               const void* entities, int32_t n_entities, const orc_ray* rays, uint64_t R, orc_hit* hits, float* any_t,
               uint64_t counters[4], int nthreads) {
    Scene s;
    s.n32 = static_cast<const Node32*>(nodes);
    s.n64 = static_cast<const Node64*>(nodes);
    s.total_nodes = (int64_t)total_nodes;
    s.tris = static_cast<const Tri16*>(tris);
    s.verts = static_cast<const Vertex32*>(verts);
    s.ents = static_cast<const Entity192*>(entities);
    s.n_ents = n_entities;
    const bool stack = format == ORC_STACK;
    const int nt = nthreads < 1 ? 1 : nthreads;
    std::vector<uint64_t> part(4 * (size_t)nt, 0);
    parallel_ranges(R, nt, [&](uint64_t lo, uint64_t hi, int tid) {
        uint64_t c_nodes = 0, c_tris = 0, c_capped = 0, c_hits = 0;
        for (uint64_t i = lo; i < hi; ++i) {
            Tally tally;
            if (kind == ORC_ANY) {
                const float t = scene_any(s, stack, rays[i], tally);
                any_t[i] = t;
                if (t > 0.0f) c_hits++;
            } else {
                const orc_hit h = scene_closest(s, stack, kind == ORC_CLOSEST_IGNORE_TRANSPARENT, rays[i], tally);
                hits[i] = h;
                if (h.t > 0.0f) c_hits++;
            }
            c_nodes += tally.nodes;
            c_tris += tally.tri_tests;
            if (tally.capped) c_capped++;
        }
        part[4 * (size_t)tid + 0] = c_nodes;
        part[4 * (size_t)tid + 1] = c_tris;
        part[4 * (size_t)tid + 2] = c_capped;
        part[4 * (size_t)tid + 3] = c_hits;
    });
    if (counters) {
        for (int k = 0; k < 4; ++k) counters[k] = 0;
        for (int t = 0; t < nt; ++t)
            for (int k = 0; k < 4; ++k) counters[k] += part[4 * (size_t)t + k];
    }
}

void orc_brute_force(const void* tris_, uint64_t T, const void* verts_, const void* entities, int32_t n_entities,
                     const orc_ray* rays, uint64_t R, orc_hit* hits, int nthreads) {
    const Tri16* tris = static_cast<const Tri16*>(tris_);
    const Vertex32* verts = static_cast<const Vertex32*>(verts_);
    const Entity192* ents = static_cast<const Entity192*>(entities);
    parallel_ranges(R, nthreads < 1 ? 1 : nthreads, [&](uint64_t lo, uint64_t hi, int) {
        for (uint64_t i = lo; i < hi; ++i) {
            const V3 ro = {rays[i].ox, rays[i].oy, rays[i].oz}, rd = {rays[i].dx, rays[i].dy, rays[i].dz};
            orc_hit h{-1.0f, -1.0f, -1.0f, -1.0f, -1, -1, -1, 0};
            float tmax = 1000000.0f;
            for (int32_t e = 0; e < n_entities; ++e) {
                const V3 o = xform(ents[e].inv, ro, 1.0f), d = xform(ents[e].inv, rd, 0.0f);
                for (uint64_t k = 0; k < T; ++k) {
                    const float t = ray_triangle(o, d, pos3(verts[tris[k].v[0]]), pos3(verts[tris[k].v[1]]), pos3(verts[tris[k].v[2]]));
                    if (t > 0.0f && t < tmax) {
                        tmax = t;
                        h.t = t;
                        h.mesh = tris[k].v[3];
                        h.tri = (int32_t)k;
                        h.entity = e;
                    }
                }
            }
            hits[i] = h;
        }
    });
}

void orc_primary_rays(const float iv[16], const float ip[16], int W, int H, orc_ray* rays) {
    // GetRayDirectionAt + main(), Intersectors/TraverseBVHStack.glsl:133-138,:414-431
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const float tx = (float)x / (float)W, ty = (float)y / (float)H;  // vec2(Pixel) / u_Dims
            const float cx = tx * 2.0f - 1.0f, cy = ty * 2.0f - 1.0f;         // clip = (uv*2-1, -1, 1)
            const float ex = (ip[0] * cx + ip[4] * cy) + (ip[8] * -1.0f + ip[12] * 1.0f);
            const float ey = (ip[1] * cx + ip[5] * cy) + (ip[9] * -1.0f + ip[13] * 1.0f);
            const V3 eye = {ex, ey, -1.0f};                                    // eye = (.., -1, 0)
            const V3 dir = xform(iv, eye, 0.0f);
            const float inv_len = 1.0f / std::sqrt(dot(dir, dir));             // glm::normalize = v * inversesqrt(dot)
            orc_ray& r = rays[(size_t)y * (size_t)W + (size_t)x];
            r.ox = iv[12]; r.oy = iv[13]; r.oz = iv[14]; r.tmin = 0.0f;
            r.dx = dir.x * inv_len; r.dy = dir.y * inv_len; r.dz = dir.z * inv_len; r.tmax = 1000000.0f;
        }
    }
}

void orc_make_entity(const float m[16], int32_t node_offset, int32_t node_count, float emissive, float translucency,
                     void* out) {
    Entity192 e;
    std::memset(&e, 0, sizeof(e));
    std::memcpy(e.model, m, 64);
    // glm 0.9.8.5 compute_inverse<tmat4x4> (glm/detail/func_matrix.inl:297-353); M(c,r) = m[c][r]
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
    const float det = (d0 + d1) + (d2 + d3);
    const float ood = 1.0f / det;
    for (int k = 0; k < 16; ++k) e.inv[k] = inv[k] * ood;
    e.node_offset = node_offset;
    e.node_count = node_count;
    e.data[0] = fbits(emissive);                // Intersector.h:212
    e.data[1] = fbits(1.0f - translucency);     // Intersector.h:213
    std::memcpy(out, &e, sizeof(e));
}

}  // extern "C"

// ---- GetData: hit attribute fetch (SL:370-408), the step right after the path ----
namespace {
// unpackHalf2x16 component: IEEE binary16 -> binary32, exact
inline float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}
struct Attr32 { float nx, ny, nz, u, v, emissivity, alpha; int32_t mesh; };
}  // namespace

extern "C" void orc_get_data(const void* tris_, const void* verts_, const void* entities, const orc_hit* hits, uint64_t R, void* out_) {
    const Tri16* tris = static_cast<const Tri16*>(tris_);
    const Vertex32* verts = static_cast<const Vertex32*>(verts_);
    const Entity192* ents = static_cast<const Entity192*>(entities);
    Attr32* out = static_cast<Attr32*>(out_);
    for (uint64_t i = 0; i < R; ++i) {
        const orc_hit& h = hits[i];
        Attr32 a{-1.0f, -1.0f, -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, h.mesh};
        if (!(h.t < 0.0f || h.mesh < 0)) {  // SL:377-382
            const Tri16& tri = tris[h.tri];
            const Vertex32 &A = verts[tri.v[0]], &B = verts[tri.v[1]], &C = verts[tri.v[2]];
            auto lo = [](uint32_t p) { return half_to_float((uint16_t)(p & 0xFFFFu)); };
            auto hi = [](uint32_t p) { return half_to_float((uint16_t)(p >> 16)); };
            // vec2 UV = uvA * TUVW.y + uvB * TUVW.z + uvC * TUVW.w  (left to right)
            a.u = (lo(A.packed[3]) * h.u + lo(B.packed[3]) * h.v) + lo(C.packed[3]) * h.w;
            a.v = (hi(A.packed[3]) * h.u + hi(B.packed[3]) * h.v) + hi(C.packed[3]) * h.w;
            // UnpackNormal(Packed.xy) = (unpackHalf2x16(x).xy, unpackHalf2x16(y).x)
            const float nx = (lo(A.packed[0]) * h.u + lo(B.packed[0]) * h.v) + lo(C.packed[0]) * h.w;
            const float ny = (hi(A.packed[0]) * h.u + hi(B.packed[0]) * h.v) + hi(C.packed[0]) * h.w;
            const float nz = (lo(A.packed[1]) * h.u + lo(B.packed[1]) * h.v) + lo(C.packed[1]) * h.w;
            const float inv_len = 1.0f / std::sqrt((nx * nx + ny * ny) + nz * nz);  // normalize = v * inversesqrt(dot(v,v))
            a.nx = nx * inv_len; a.ny = ny * inv_len; a.nz = nz * inv_len;
            a.emissivity = ibits(ents[h.entity].data[0]);
            a.alpha = ibits(ents[h.entity].data[1]);
        }
        out[i] = a;
    }
}

// GetData with the BVHTextureReferences table (SL:393-404): the Albedo decision without the texture unit.  Record = Attr32 +
// {albedo[3], albedo_ref}: albedo_ref > -1 means the shader samples Textures[albedo_ref] at (u, v) (albedo stays at its
// initial vec3(0)); otherwise albedo = ModelColor.xyz.  A mesh outside the table is marked albedo_ref = -2 (the shader has no
// defined behaviour there).  PINNED by tests/test_materials.py against the compiled reference shader (ref_glsl_get_data_material).
namespace {
struct TexRef32 { float color[4]; int32_t albedo, normal, pad[2]; };
struct Mat48 { Attr32 a; float albedo[3]; int32_t albedo_ref; };
static_assert(sizeof(TexRef32) == 32 && sizeof(Mat48) == 48, "record layouts");
}  // namespace

extern "C" void orc_get_data_material(const void* tris_, const void* verts_, const void* entities, const void* refs_, uint64_t n_refs, const orc_hit* hits,
                                      uint64_t R, void* out_) {
    const TexRef32* refs = static_cast<const TexRef32*>(refs_);
    Mat48* out = static_cast<Mat48*>(out_);
    std::vector<Attr32> attr(R);
    orc_get_data(tris_, verts_, entities, hits, R, attr.data());
    for (uint64_t i = 0; i < R; ++i) {
        const orc_hit& h = hits[i];
        Mat48 m{attr[i], {0.0f, 0.0f, 0.0f}, -1};
        if (!(h.t < 0.0f || h.mesh < 0)) {
            if ((uint64_t)h.mesh >= n_refs) m.albedo_ref = -2;
            else {
                const int ref = refs[h.mesh].albedo;                       // SL:393
                if (ref > -1 && h.mesh > -1 && h.t > 0.0f) m.albedo_ref = ref;  // SL:397-399: texture(Textures[Ref], UV)
                else { m.albedo[0] = refs[h.mesh].color[0]; m.albedo[1] = refs[h.mesh].color[1]; m.albedo[2] = refs[h.mesh].color[2]; }  // SL:402-404
            }
        }
        out[i] = m;
    }
}

// ---- AABB collide query: Physics::CollideBox / CollideBVH (Source/Core/Physics.cpp:21-228), the other consumer of the
// stackless buffers.  PINNED: the reference's Physics.cpp compiles here (oracle/_ref) and
// tests/test_collide.py requires identical answers.  Quirks kept: the query box is taken to object space corner by
// corner (:84-85), `e` is the full extent (:31), only nine cross-product axes are tested and the sixth repeats
// cross(u2, f2) (:46-54), and the first overlapping triangle in walk order wins.
namespace {
inline bool aabb_overlap(V3 amin, V3 amax, V3 bmin, V3 bmax) {  // :23-27
    return (amin.x <= bmax.x && amax.x >= bmin.x) && (amin.y <= bmax.y && amax.y >= bmin.y) && (amin.z <= bmax.z && amax.z >= bmin.z);
}
inline bool box_triangle_overlap(V3 v0, V3 v1, V3 v2, V3 bmin, V3 bmax) {  // :29-77
    const V3 c = {(bmin.x + bmax.x) / 2.0f, (bmin.y + bmax.y) / 2.0f, (bmin.z + bmax.z) / 2.0f};
    const V3 e = sub(bmax, bmin);
    v0 = sub(v0, c); v1 = sub(v1, c); v2 = sub(v2, c);
    const V3 f0 = sub(v1, v0), f1 = sub(v2, v1), f2 = sub(v0, v2);
    const V3 u0 = {1.0f, 0.0f, 0.0f}, u1 = {0.0f, 1.0f, 0.0f}, u2 = {0.0f, 0.0f, 1.0f};
    const V3 axes[9] = {cross(u0, f0), cross(u0, f1), cross(u0, f2), cross(u1, f0), cross(u1, f1), cross(u2, f2), cross(u2, f0), cross(u2, f1), cross(u2, f2)};
    for (int i = 0; i < 9; ++i) {
        const V3 a = axes[i];
        const float p0 = dot(v0, a), p1 = dot(v1, a), p2 = dot(v2, a);
        const float r = e.x * std::fabs(dot(u0, a)) + e.y * std::fabs(dot(u1, a)) + e.z * std::fabs(dot(u2, a));
        if (smax(-smax(smax(p0, p1), p2), smin(smin(p0, p1), p2)) > r) return false;
    }
    return true;
}
}  // namespace

extern "C" void orc_collide_boxes(const void* nodes_, uint64_t n_nodes, const void* tris_, const void* verts_, const void* entities, int32_t n_entities,
                                  const float* boxes, uint64_t n, int32_t* out) {
    const Node32* nodes = static_cast<const Node32*>(nodes_);
    const Tri16* tris = static_cast<const Tri16*>(tris_);
    const Vertex32* verts = static_cast<const Vertex32*>(verts_);
    const Entity192* ents = static_cast<const Entity192*>(entities);
    for (uint64_t q = 0; q < n; ++q) {
        const V3 wmin = {boxes[8 * q], boxes[8 * q + 1], boxes[8 * q + 2]}, wmax = {boxes[8 * q + 4], boxes[8 * q + 5], boxes[8 * q + 6]};
        int32_t res[4] = {0, -1, -1, -1};
        for (int32_t ei = 0; ei < n_entities && !res[0]; ++ei) {  // CollideBox :203-228
            const Entity192& en = ents[ei];
            const V3 cmin = xform(en.inv, wmin, 1.0f), cmax = xform(en.inv, wmax, 1.0f);  // :84-85
            const int32_t start = en.node_offset, count = en.node_count;
            int32_t ptr = start, iters = 0;
            while (ptr >= 0 && iters < 1024) {
                // :101 also admits Pointer == m_BVHNodes.size(), an out-of-bounds read; both sides stop there instead
                if (ptr < start || ptr > start + count || (uint64_t)ptr >= n_nodes) break;
                ++iters;
                const Node32& nd = nodes[ptr];
                const int32_t link = fbits(nd.mx[3]);
                if (aabb_overlap({nd.mn[0], nd.mn[1], nd.mn[2]}, {nd.mx[0], nd.mx[1], nd.mx[2]}, cmin, cmax)) {
                    const int32_t pack = fbits(nd.mn[3]);
                    if (pack != -1) {
                        const int32_t len = pack & 0xF;
                        for (int32_t idx = pack >> 4; idx < (pack >> 4) + len; ++idx) {
                            const Tri16& t = tris[idx];
                            if (box_triangle_overlap(pos3(verts[t.v[0]]), pos3(verts[t.v[1]]), pos3(verts[t.v[2]]), cmin, cmax)) {
                                res[0] = 1; res[1] = t.v[3]; res[2] = idx; res[3] = ei;
                                break;
                            }
                        }
                        if (res[0]) break;
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
        for (int k = 0; k < 4; ++k) out[4 * q + k] = res[k];
    }
}

