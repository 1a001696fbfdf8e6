// TEST INFRASTRUCTURE ONLY — see oracle.h.
//
// Scalar C++ restatement of the reference's ray GENERATORS, the step in front of the traversal path
// (paths relative to /root/reference/Source/Core/Shaders/):
//   diffuse  : DiffuseTrace.glsl:445-446 (first bounce: origin = P + N*0.05, direction = CosWeightedHemisphere(N, xi)),
//              :516-517 (later bounces, offset 0.02); CosWeightedHemisphere = Include/Sampling.glsl:1-12
//   specular : SpecularTrace.glsl:512-513 (origin = P + N*mix(0.05, 0.1, clamp(rough*1.4, 0, 1)), direction =
//              StochasticReflectionDirection(Incident, N, rough*0.9)), :102-135; SampleGGXVNDF = Include/Sampling.glsl:63-83
//   shadow   : direction = normalize(SampleCone(L, xi, CosThetaMax)), Include/Sampling.glsl:43-61 (the reference's cone
//              sampler; the reference has no shadow-ray pass of its own, so the call site is this repo's)
//   probe    : UpdateRadianceProbes.glsl:408-427 (RayOrigin = u_BoxOrigin + (Pixel/u_Resolution*2-1)*u_Size, direction =
//              ImportanceSample() = normalize(LambertBRDF(vec3(hash2(), hash2().x))), :351-374)
// PINNED: oracle/ref_shim compiles these shader functions themselves (glsl_to_cpp.py --function) against the
// reference's glm; tests/test_raygen_oracle.py requires identical bits.  Two things the shaders leave to the
// implementation are fixed by definition on both sides: the random numbers (the shaders' fract(sin()) hash2() is replaced by
// the counter-based stream below) and the bits of sin / cos / acos / pow (exact_math_ref.h).
// Where the shaders read the surface normal from a G-buffer texture, the generator uses the geometric normal of
// the hit triangle turned against the incoming ray.
#include "oracle.h"
#include "exact_math_ref.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct V3 { float x, y, z; };
struct Vertex32 { float pos[4]; uint32_t packed[4]; };
struct Tri16 { int32_t v[4]; };
struct Entity192 { float model[16]; float inv[16]; int32_t node_offset, node_count; int32_t data[14]; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                                           // glm compute_dot<tvec3>
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }        // glm compute_cross
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }                                               // v * inversesqrt(dot(v, v))
inline V3 reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }                                                    // glm compute_reflect
inline float smin(float x, float y) { return x < y ? x : y; }   // glm::min (glm/detail/func_common.inl:14-18)
inline float smax(float x, float y) { return x > y ? x : y; }   // glm::max (:22-27)

inline uint32_t pcg_hash(uint32_t v) {
    const uint32_t s = v * 747796405u + 2891336453u;
    const uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
inline float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

// The stream that stands in for the shaders' hash2(): call n of stream k returns (xi(k, 2n), xi(k, 2n+1)).
struct Hash2 {
    uint32_t k, n;
    float next() { return u01(pcg_hash(k + (n++) * 0x9E3779B9u)); }
};
inline uint32_t stream_key(uint32_t seed, uint32_t element) { return pcg_hash(seed ^ pcg_hash(element)); }

const float PI2 = 2.0f * 3.14159265359f;

// Include/Sampling.glsl:1-12
V3 cos_weighted_hemisphere(V3 n, float rx_, float ry_) {
    const V3 uu = normalize(cross(n, V3{0.0f, 1.0f, 1.0f}));
    const V3 vv = cross(uu, n);
    const float ra = std::sqrt(ry_);
    float sn, cs;
    xm::xsincos(PI2 * rx_, sn, cs);
    const float rx = ra * cs, ry = ra * sn, rz = std::sqrt(1.0f - ry_);
    const V3 rr = (rx * uu + ry * vv) + rz * n;
    return normalize(rr);
}

// Include/Sampling.glsl:63-83
V3 sample_ggx_vndf(V3 N, float roughness, float xi_x, float xi_y) {
    const float alpha = roughness * roughness, alpha2 = alpha * alpha;
    const float phi = PI2 * xi_x;
    const float cos_theta = std::sqrt((1.0f - xi_y) / (1.0f + (alpha2 - 1.0f) * xi_y));
    const float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    float sn, cs;
    xm::xsincos(phi, sn, cs);
    const V3 H = {cs * sin_theta, sn * sin_theta, cos_theta};
    const V3 up = std::fabs(N.z) < 0.999f ? V3{0.0f, 0.0f, 1.0f} : V3{1.0f, 0.0f, 0.0f};
    const V3 tangent = normalize(cross(up, N));
    const V3 bitangent = cross(N, tangent);
    const V3 sample_vec = (tangent * H.x + bitangent * H.y) + N * H.z;
    return normalize(sample_vec);
}

// SpecularTrace.glsl:102-135
V3 stochastic_reflection_direction(V3 incident, V3 normal, float roughness, Hash2& h) {
    if (roughness < 0.01f) return reflect(incident, normal);
    V3 microfacet = normal;
    for (int i = 0; i < 12; ++i) {
        const float a = h.next() * 0.8f, b = h.next() * 0.7f;  // hash2() * TailControl
        const V3 s = sample_ggx_vndf(normal, roughness, a, b);
        if (dot(s, normal) > 0.001f) { microfacet = s; break; }
    }
    return reflect(incident, microfacet);
}

// Include/Sampling.glsl:43-61: SampleCone(Direction, Xi, CosTheta) = mat3(T, B, L) * SampleCone(Xi, CosTheta)
V3 sample_cone(V3 L, float xi_x, float xi_y, float cos_theta_max) {
    const float cos_theta = (1.0f - xi_x) + xi_x * cos_theta_max;
    const float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    const float phi = xi_y * 3.14159265359f * 2.0f;
    float sn, cs;
    xm::xsincos(phi, sn, cs);
    const V3 l = {sin_theta * cs, sin_theta * sn, cos_theta};
    const V3 T = normalize(cross(L, V3{0.0f, 1.0f, 1.0f}));
    const V3 B = cross(T, L);
    // glm mat3 * vec3: m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z
    return {T.x * l.x + B.x * l.y + L.x * l.z, T.y * l.x + B.y * l.y + L.y * l.z, T.z * l.x + B.z * l.y + L.z * l.z};
}

// UpdateRadianceProbes.glsl:351-362 (PI is 3.1415926535 in that file, :4)
V3 lambert_brdf(float hx, float hy, float hz) {
    const float phi = 2.0f * 3.1415926535f * hx;
    const float cos_theta = 2.0f * hy - 1.0f;
    const float theta = xm::xacos(cos_theta);
    const float r = xm::xpow(hz, 1.0f / 3.0f);
    float sp, cp, st, ct;
    xm::xsincos(phi, sp, cp);
    xm::xsincos(theta, st, ct);
    return {r * st * cp, r * st * sp, r * ct};
}

struct GenParams { int32_t kind, spp; uint32_t seed, flags; float offset, tmax, roughness; float light_dir[3]; float light_cone; };

struct HitFrame { V3 p, n, in; bool valid; };

HitFrame hit_frame(const orc_ray& r, const orc_hit& h, const Tri16* tris, const Vertex32* verts, const Entity192* ents) {
    HitFrame f{};
    f.valid = h.t > 0.0f;
    if (!f.valid) return f;
    const V3 o = {r.ox, r.oy, r.oz}, d = {r.dx, r.dy, r.dz};
    f.p = o + d * h.t;
    f.in = d;
    const Tri16& t = tris[h.tri];
    const V3 v0 = {verts[t.v[0]].pos[0], verts[t.v[0]].pos[1], verts[t.v[0]].pos[2]};
    const V3 v1 = {verts[t.v[1]].pos[0], verts[t.v[1]].pos[1], verts[t.v[1]].pos[2]};
    const V3 v2 = {verts[t.v[2]].pos[0], verts[t.v[2]].pos[1], verts[t.v[2]].pos[2]};
    const V3 c = cross(v1 - v0, v2 - v0);  // RayTriangle's n (…Stackless.glsl:85), object space
    const float* m = ents[h.entity].model;
    const V3 nw = {m[0] * c.x + m[4] * c.y + m[8] * c.z, m[1] * c.x + m[5] * c.y + m[9] * c.z, m[2] * c.x + m[6] * c.y + m[10] * c.z};
    const float dd = dot(nw, nw);
    if (!(dd > 0.0f) || std::isinf(dd)) { f.valid = false; return f; }  // degenerate triangle: no ray
    V3 n = nw * (1.0f / std::sqrt(dd));
    if (dot(n, d) > 0.0f) n = neg(n);
    f.n = n;
    return f;
}

// Sample s of element e (stream key from (seed, e * spp + s)).  false: no ray.
bool gen_ray(const GenParams& g, const HitFrame& f, uint32_t element, float* out8) {
    Hash2 h{stream_key(g.seed, element), 0};
    float off = g.offset;
    V3 dir;
    if (g.kind == 0) {
        const float a = h.next(), b = h.next();
        dir = cos_weighted_hemisphere(f.n, a, b);
    } else if (g.kind == 1) {
        dir = stochastic_reflection_direction(f.in, f.n, g.roughness * 0.9f, h);
        if (g.offset < 0.0f) off = 0.05f + smin(smax(g.roughness * 1.4f, 0.0f), 1.0f) * (0.1f - 0.05f);  // mix(0.05, 0.1, clamp(PBR.x*1.4, 0, 1))
    } else {
        const V3 L = {g.light_dir[0], g.light_dir[1], g.light_dir[2]};
        if (!(dot(f.n, L) > 0.0f)) return false;
        const float a = h.next(), b = h.next();
        const float cos_max = std::sqrt(1.0f - g.light_cone * g.light_cone);
        dir = normalize(sample_cone(L, a, b, cos_max));
    }
    const V3 o = f.p + f.n * off;
    out8[0] = o.x; out8[1] = o.y; out8[2] = o.z; out8[3] = 0.0f;
    out8[4] = dir.x; out8[5] = dir.y; out8[6] = dir.z; out8[7] = g.tmax;
    return true;
}

inline unsigned octant_of(const float* r8) { return (r8[4] > 0.0f ? 1u : 0u) | (r8[5] > 0.0f ? 2u : 0u) | (r8[6] > 0.0f ? 4u : 0u); }

}  // namespace

extern "C" {

float orc_xsin(float x) { return xm::xsin(x); }
float orc_xcos(float x) { return xm::xcos(x); }
float orc_xacos(float x) { return xm::xacos(x); }
float orc_xpow(float x, float y) { return xm::xpow(x, y); }
void orc_xmath_batch(int which, const float* x, const float* y, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = which == 0 ? xm::xsin(x[i]) : which == 1 ? xm::xcos(x[i]) : which == 2 ? xm::xacos(x[i]) : xm::xpow(x[i], y[i]);
}

// Direction samplers on n inputs (for the pin against the compiled shader functions):
//   which 0: CosWeightedHemisphere(N, xi.xy)            in: normals, xi (2 per input)
//         1: SampleGGXVNDF(N, roughness, xi.xy)
//         2: StochasticReflectionDirection(I, N, roughness) with hash2() = stream keys[i]     in: normals, incident, keys
//         3: SampleCone(N as Direction, xi.xy, cos_theta = roughness argument)
//         4: normalize(LambertBRDF(vec3(hash2(), hash2().x))) with hash2() = stream keys[i]   (ImportanceSample, probe update)
void orc_sample_directions(int which, const float* normals, const float* incident, const float* xi, const uint32_t* keys, float roughness, uint64_t n,
                           float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        V3 N{0, 0, 0}, I{0, 0, 0}, d;
        if (normals) N = {normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
        if (incident) I = {incident[3 * i], incident[3 * i + 1], incident[3 * i + 2]};
        if (which == 0) d = cos_weighted_hemisphere(N, xi[2 * i], xi[2 * i + 1]);
        else if (which == 1) d = sample_ggx_vndf(N, roughness, xi[2 * i], xi[2 * i + 1]);
        else if (which == 2) { Hash2 h{keys[i], 0}; d = stochastic_reflection_direction(I, N, roughness, h); }
        else if (which == 3) d = sample_cone(N, xi[2 * i], xi[2 * i + 1], roughness);
        else { Hash2 h{keys[i], 0}; const float a = h.next(), b = h.next(), c = h.next(); h.next(); d = normalize(lambert_brdf(a, b, c)); }
        out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
    }
}
// xi(k, j), j = 0..m-1 of stream `key`; and the stream key of (seed, element)
void orc_hash2_stream(uint32_t key, uint32_t m, float* out) {
    Hash2 h{key, 0};
    for (uint32_t j = 0; j < m; ++j) out[j] = h.next();
}
uint32_t orc_stream_key(uint32_t seed, uint32_t element) { return stream_key(seed, element); }

// The wavefront generator of the product (cndl_generate_rays_device): for every input ray i whose hit record has t > 0,
// `spp` rays, stream element = id(i) * spp + s with id(i) = ids_in ? ids_in[i] : i.  Output order: input order
// (flags & 1 == 0) or octant-major (bucket = direction signs, each bucket in input order).  parent_out[k] = i,
// ids_out[k] = id(i) * spp + s (either may be NULL).  Returns the number of rays written (out_rays holds R * spp).
uint64_t orc_generate_rays(const void* params_, const orc_ray* rays, const orc_hit* hits, const uint32_t* ids_in, uint64_t R, const void* tris_,
                           const void* verts_, const void* ents_, orc_ray* out_rays, uint32_t* parent_out, uint32_t* ids_out) {
    const GenParams& g = *static_cast<const GenParams*>(params_);
    const Tri16* tris = static_cast<const Tri16*>(tris_);
    const Vertex32* verts = static_cast<const Vertex32*>(verts_);
    const Entity192* ents = static_cast<const Entity192*>(ents_);
    struct Item { float r[8]; uint32_t parent, id; unsigned oct; };
    std::vector<Item> items;
    items.reserve((size_t)R * (size_t)g.spp);
    for (uint64_t i = 0; i < R; ++i) {
        const HitFrame f = hit_frame(rays[i], hits[i], tris, verts, ents);
        if (!f.valid) continue;
        const uint32_t id = ids_in ? ids_in[i] : (uint32_t)i;
        for (int s = 0; s < g.spp; ++s) {
            Item it;
            const uint32_t element = id * (uint32_t)g.spp + (uint32_t)s;
            if (!gen_ray(g, f, element, it.r)) continue;
            it.parent = (uint32_t)i;
            it.id = element;
            it.oct = (g.flags & 1u) ? octant_of(it.r) : 0u;
            items.push_back(it);
        }
    }
    uint64_t k = 0;
    for (unsigned o = 0; o < 8; ++o)
        for (const Item& it : items) {
            if (it.oct != o) continue;
            std::memcpy(&out_rays[k], it.r, 32);
            if (parent_out) parent_out[k] = it.parent;
            if (ids_out) ids_out[k] = it.id;
            ++k;
        }
    return k;
}

// Probe-update rays (UpdateRadianceProbes.glsl:408-427): one ray per probe of a res[0] x res[1] x res[2] grid, probe
// (x, y, z) -> index (z * res[1] + y) * res[0] + x; stream element = that index; tmax = 1e6.
void orc_probe_rays(const float box_origin[3], const float size[3], const int32_t res[3], uint32_t seed, orc_ray* out) {
    const V3 org = {box_origin[0], box_origin[1], box_origin[2]}, sz = {size[0], size[1], size[2]};
    const V3 rs = {(float)res[0], (float)res[1], (float)res[2]};
    uint32_t idx = 0;
    for (int z = 0; z < res[2]; ++z)
        for (int y = 0; y < res[1]; ++y)
            for (int x = 0; x < res[0]; ++x, ++idx) {
                const V3 tex = {(float)x / rs.x, (float)y / rs.y, (float)z / rs.z};                 // vec3(Pixel) / u_Resolution
                const V3 clip = {tex.x * 2.0f - 1.0f, tex.y * 2.0f - 1.0f, tex.z * 2.0f - 1.0f};     // TexCoords * 2.0f - 1.0f
                const V3 o = {org.x + clip.x * sz.x, org.y + clip.y * sz.y, org.z + clip.z * sz.z};  // u_BoxOrigin + Clip * u_Size
                Hash2 h{stream_key(seed, idx), 0};
                const float a = h.next(), b = h.next(), c = h.next();
                const V3 d = normalize(lambert_brdf(a, b, c));
                out[idx] = orc_ray{o.x, o.y, o.z, 0.0f, d.x, d.y, d.z, 1000000.0f};
            }
}

}  // extern "C"
