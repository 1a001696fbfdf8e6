// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// The reference's GLSL traversal (Source/Core/Shaders/Intersectors/Include/TraverseBVHStackless.glsl and
// TraverseBVHStack.glsl) executed on the CPU: oracle/ref_shim/glsl_to_cpp.py rewrites the two include files
// syntactically (qualifiers, literals, swizzles, SSBO blocks -> pointers) into oracle/_ref/gen/*.inc, and this
// file compiles them against the reference's own vendored glm 0.9.8.5, whose vector functions follow the GLSL
// specification except that glm evaluates min(x,y) as x<y?x:y and max(x,y) as x>y?x:y (func_common.inl:14-27; the two forms
// differ only for NaN operands, which GLSL leaves undefined).  Compiled with
// -O2 -ffp-contract=off like the oracle, so every float operation is a separately rounded IEEE operation.
// This is what pins the oracle's traversal restatement: tests require identical hit records.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#undef INFINITY

#include <glm/glm.hpp>
#include <glm/gtc/packing.hpp>

namespace glsl_prelude {
// No texture unit here.  A sampler knows its index in Textures[], and texture() records (index, uv) of the call instead of
// sampling: that is exactly what the product reports for a textured hit (albedo_ref, u, v), so the Albedo branch is pinned too.
struct sampler2D { int id = -1; };
static int g_sampled_id = -1;
static glm::vec2 g_sampled_uv(0.0f);
inline glm::vec4 texture(const sampler2D& s, const glm::vec2& uv) { g_sampled_id = s.id; g_sampled_uv = uv; return glm::vec4(0.0f); }
}  // namespace glsl_prelude

namespace ref_stackless {
using namespace glm;
using namespace glsl_prelude;
#include "gen/stackless.inc"
}  // namespace ref_stackless

namespace ref_stack {
using namespace glm;
using namespace glsl_prelude;
#include "gen/stack.inc"
}  // namespace ref_stack

namespace ref_primary {  // GetRayDirectionAt of the primary-ray kernel (Intersectors/TraverseBVHStack.glsl:133-138)
using namespace glm;
static mat4 u_InverseProjection, u_InverseView;
#include "gen/primary.inc"
}  // namespace ref_primary

namespace ref_sampling {  // CosWeightedHemisphere, SampleGGXVNDF (Shaders/Include/Sampling.glsl:1-12, :63-83)
using namespace glm;
#include "gen/cos_hemisphere.inc"
#include "gen/ggx_vndf.inc"
}  // namespace ref_sampling

// The ray generators' direction samplers, compiled from the shader files with the two things GLSL leaves to the
// implementation bound to ONE definition: hash2() (the shaders' fract(sin()) hash) draws from the counter stream of
// oracle_raygen.cpp, and sin / cos / acos / pow are the functions of exact_math_ref.h.  Everything else — the order of
// operations, normalize, cross, reflect, sqrt, mat3 * vec3 — is the shader's own text over the reference's glm.
//   CosWeightedHemisphere, SampleGGXVNDF, SampleCone x2   Shaders/Include/Sampling.glsl:1-12, :63-83, :43-61
//   StochasticReflectionDirection                          Shaders/SpecularTrace.glsl:102-135
//   LambertBRDF, ImportanceSample                          Shaders/UpdateRadianceProbes.glsl:351-362, :365-406
#include "exact_math_ref.h"
namespace ref_raygen {
using namespace glm;
inline float sin(float x) { return xm::xsin(x); }
inline float cos(float x) { return xm::xcos(x); }
inline float acos(float x) { return xm::xacos(x); }
inline float pow(float x, float y) { return xm::xpow(x, y); }
static uint32_t g_key = 0, g_calls = 0;
inline uint32_t pcg_hash(uint32_t v) {
    const uint32_t s = v * 747796405u + 2891336453u;
    const uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
inline float next_xi() { return (float)(pcg_hash(g_key + (g_calls++) * 0x9E3779B9u) >> 8) * (1.0f / 16777216.0f); }
inline vec2 hash2() { const float a = next_xi(), b = next_xi(); return vec2(a, b); }
#define PI 3.14159265359f  /* SpecularTrace.glsl:6 */
#include "gen/cos_hemisphere.inc"
#include "gen/ggx_vndf.inc"
#include "gen/sample_cone_xi.inc"
#include "gen/sample_cone_dir.inc"
#include "gen/stochastic_reflection.inc"
#undef PI
#define PI 3.1415926535f   /* UpdateRadianceProbes.glsl:4 */
// ImportanceSample's importance-sampling branch is switched off in the shader (`ShouldImportanceSample = false`); what it
// names must exist for the text to compile, and is never reached
struct ProbeMapPixel { vec2 Packed; };
static const ProbeMapPixel* MapData = nullptr;
inline int Get1DIdx(ivec2, ivec2) { std::abort(); }
inline vec3 OctahedronToUnitVector(vec2) { std::abort(); }
inline vec3 SampleDirectionCone(vec3) { std::abort(); }
#include "gen/lambert_brdf.inc"
#include "gen/importance_sample.inc"
#undef PI
}  // namespace ref_raygen

namespace {
struct Hit32 { float t, u, v, w; int32_t mesh, tri, entity, iters; };
struct Attr32 { float nx, ny, nz, u, v, emissivity, alpha; int32_t mesh; };
}  // namespace

#define BIND(NS)                                                                  \
    NS::BVHNodes = static_cast<const NS::Node*>(nodes);                           \
    NS::BVHTris = static_cast<const NS::Triangle*>(tris);                         \
    NS::BVHVertices = static_cast<const NS::Vertex*>(verts);                      \
    NS::BVHEntities = static_cast<const NS::BVHEntity*>(ents);                    \
    NS::u_EntityCount = n_ents;                                                   \
    NS::u_TotalNodes = (int)n_nodes;

extern "C" {

// kind: 0 IntersectScene, 1 IntersectSceneIgnoreTransparent -> 32-byte hit records; 2 any-hit -> one float per ray.
// Any-hit with ray tmax <= 0 calls the shader's own `float IntersectScene(o, d)`; with tmax > 0 (the ABI's extension)
// it runs that function's six-line entity loop around the shader's Intersect*Occlusion with the ray's TMax.
int ref_glsl_trace(int format, int kind, const void* nodes, uint64_t n_nodes, const void* tris, const void* verts, const void* ents, int32_t n_ents,
                   const float* rays, uint64_t R, void* out) {
    if (format == 0) { BIND(ref_stackless) } else { BIND(ref_stack) }
    for (uint64_t i = 0; i < R; ++i) {
        const glm::vec3 o(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]), d(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]);
        const float ray_tmax = rays[8 * i + 7];
        if (kind == 2) {
            float t;
            if (!(ray_tmax > 0.0f)) {
                t = format == 0 ? ref_stackless::IntersectRay(o, d) : ref_stack::IntersectRay(o, d);  // float IntersectRay(o, d): the any-hit entry point of both files
            } else {
                t = -1.0f;
                for (int e = 0; e < n_ents; ++e) {
                    const float tr = format == 0
                        ? ref_stackless::IntersectBVHStacklessOcclusion(o, d, ref_stackless::BVHEntities[e].NodeOffset, ref_stackless::BVHEntities[e].NodeCount,
                                                                        ref_stackless::BVHEntities[e].InverseMatrix, ray_tmax)
                        : ref_stack::IntersectBVHStackOcclusion(o, d, ref_stack::BVHEntities[e].NodeOffset, ref_stack::BVHEntities[e].NodeCount,
                                                                ref_stack::BVHEntities[e].InverseMatrix, ray_tmax);
                    if (tr > 0.0f) { t = tr; break; }
                }
            }
            static_cast<float*>(out)[i] = t;
        } else {
            int mesh = -1, tri = -1, entity = -1, iters = 0;  // `out` parameters the shader leaves unwritten on a miss
            glm::vec4 tuvw;
            if (format == 0) tuvw = kind == 0 ? ref_stackless::IntersectScene(o, d, mesh, tri, entity, iters) : ref_stackless::IntersectSceneIgnoreTransparent(o, d, mesh, tri, entity, iters);
            else tuvw = kind == 0 ? ref_stack::IntersectScene(o, d, mesh, tri, entity, iters) : ref_stack::IntersectSceneIgnoreTransparent(o, d, mesh, tri, entity, iters);
            static_cast<Hit32*>(out)[i] = Hit32{tuvw.x, tuvw.y, tuvw.z, tuvw.w, mesh, tri, entity, iters};
        }
    }
    return 0;
}

// The same over `nthreads` host threads (contiguous ray ranges).  The shader globals are bound once, before the threads
// start, and only read afterwards.  This is the reference arm of bench.py (`--impl reference`, cpu_baseline kind "reference").
int ref_glsl_trace_mt(int format, int kind, const void* nodes, uint64_t n_nodes, const void* tris, const void* verts, const void* ents, int32_t n_ents,
                      const float* rays, uint64_t R, void* out, int nthreads) {
    if (nthreads <= 1 || R < 4096) return ref_glsl_trace(format, kind, nodes, n_nodes, tris, verts, ents, n_ents, rays, R, out);
    ref_glsl_trace(format, kind, nodes, n_nodes, tris, verts, ents, n_ents, rays, 0, out);  // binds the globals
    const uint64_t chunk = (R + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
    const size_t out_elt = kind == 2 ? sizeof(float) : sizeof(Hit32);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        const uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < R ? lo + chunk : R;
        if (lo >= hi) break;
        pool.emplace_back([=] {
            ref_glsl_trace(format, kind, nodes, n_nodes, tris, verts, ents, n_ents, rays + 8 * lo, hi - lo, static_cast<char*>(out) + lo * out_elt);
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}

// Primary rays exactly as main() of Intersectors/TraverseBVHStack.glsl:414-421 forms them: TexCoords = vec2(Pixel) / u_Dims,
// rD = normalize(GetRayDirectionAt(TexCoords)), rO = u_InverseView[3].xyz.  Matrices column-major; rays: 8 floats each.
int ref_glsl_primary_rays(const float* inv_view16, const float* inv_proj16, int W, int H, float* rays) {
    std::memcpy(&ref_primary::u_InverseView[0][0], inv_view16, 64);
    std::memcpy(&ref_primary::u_InverseProjection[0][0], inv_proj16, 64);
    const glm::vec2 dims((float)W, (float)H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const glm::vec2 tex = glm::vec2(glm::ivec2(x, y)) / dims;
            const glm::vec3 rd = glm::normalize(ref_primary::GetRayDirectionAt(tex));
            const glm::vec3 ro = glm::vec3(ref_primary::u_InverseView[3]);
            float* r = rays + 8 * ((size_t)y * W + x);
            r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = 0.0f;
            r[4] = rd.x; r[5] = rd.y; r[6] = rd.z; r[7] = 1000000.0f;
        }
    return 0;
}

// The reference's direction samplers on n inputs: which = 0 CosWeightedHemisphere(N, xi), 1 SampleGGXVNDF(N, roughness, xi).
void ref_glsl_sample(int which, const float* normals, const float* xi, float roughness, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const glm::vec3 N(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        const glm::vec2 X(xi[2 * i], xi[2 * i + 1]);
        const glm::vec3 d = which == 0 ? ref_sampling::CosWeightedHemisphere(N, X) : ref_sampling::SampleGGXVNDF(N, roughness, X);
        out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
    }
}

// The shader functions of namespace ref_raygen on n inputs (same `which` numbering as orc_sample_directions):
//   0 CosWeightedHemisphere(N, xi)   1 SampleGGXVNDF(N, roughness, xi)   2 StochasticReflectionDirection(I, N, roughness), hash2() = stream keys[i]
//   3 SampleCone(N as Direction, xi, CosTheta = roughness argument)      4 ImportanceSample(0) (probe update), hash2() = stream keys[i]
void ref_glsl_sample_directions(int which, const float* normals, const float* incident, const float* xi, const uint32_t* keys, float roughness, uint64_t n,
                                float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        glm::vec3 N(0.0f), I(0.0f), d;
        glm::vec2 X(0.0f);
        if (normals) N = glm::vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        if (incident) I = glm::vec3(incident[3 * i], incident[3 * i + 1], incident[3 * i + 2]);
        if (xi) X = glm::vec2(xi[2 * i], xi[2 * i + 1]);
        if (keys) { ref_raygen::g_key = keys[i]; ref_raygen::g_calls = 0; }
        if (which == 0) d = ref_raygen::CosWeightedHemisphere(N, X);
        else if (which == 1) d = ref_raygen::SampleGGXVNDF(N, roughness, X);
        else if (which == 2) d = ref_raygen::StochasticReflectionDirection(I, N, roughness);
        else if (which == 3) d = ref_raygen::SampleCone(N, X, roughness);
        else d = ref_raygen::ImportanceSample(0);
        out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
    }
}

// glm::inverse as RayIntersector::PushEntity applies it (Intersector.h:209).
void ref_glm_inverse(const float* m16, float* out16) {
    glm::mat4 m;
    std::memcpy(&m[0][0], m16, 64);
    const glm::mat4 inv = glm::inverse(m);
    std::memcpy(out16, &inv[0][0], 64);
}

// GetData (TraverseBVHStackless.glsl:375-408) on hit records; texture references all carry Albedo = -1, so Albedo is the
// ModelColor branch and is not reported.
int ref_glsl_get_data(const void* tris, const void* verts, const void* ents, int32_t n_ents, const void* hits_, uint64_t R, void* out_) {
    const void* nodes = nullptr;
    const uint64_t n_nodes = 0;
    BIND(ref_stackless)
    const Hit32* hits = static_cast<const Hit32*>(hits_);
    int max_mesh = 0;
    for (uint64_t i = 0; i < R; ++i) max_mesh = hits[i].mesh > max_mesh ? hits[i].mesh : max_mesh;
    std::vector<ref_stackless::TextureReferences> refs((size_t)max_mesh + 1);
    for (auto& r : refs) { r.ModelColor = glm::vec4(1.0f); r.Albedo = -1; r.Normal = -1; r.Pad[0] = r.Pad[1] = 0; }
    ref_stackless::BVHTextureReferences = refs.data();
    Attr32* out = static_cast<Attr32*>(out_);
    for (uint64_t i = 0; i < R; ++i) {
        const Hit32& h = hits[i];
        glm::vec3 normal(0.0f), albedo(0.0f);
        float emissivity = 0.0f, alpha = 0.0f;
        // on a miss the shader writes Normal / Albedo / Emissivity and leaves Alpha alone
        ref_stackless::GetData(glm::vec4(h.t, h.u, h.v, h.w), h.mesh, h.tri, h.entity, normal, albedo, emissivity, alpha);
        const bool miss = h.t < 0.0f || h.mesh < 0;
        glm::vec2 uv(0.0f);
        if (!miss) {  // UV is a local of GetData: the same expression, evaluated by the same glm code
            const auto& T = ref_stackless::BVHTris[h.tri];
            const auto &A = ref_stackless::BVHVertices[T.PackedData[0]], &B = ref_stackless::BVHVertices[T.PackedData[1]], &C = ref_stackless::BVHVertices[T.PackedData[2]];
            uv = (glm::unpackHalf2x16(A.PackedData.w) * h.u) + (glm::unpackHalf2x16(B.PackedData.w) * h.v) + (glm::unpackHalf2x16(C.PackedData.w) * h.w);
        }
        out[i] = Attr32{normal.x, normal.y, normal.z, uv.x, uv.y, emissivity, alpha, h.mesh};
    }
    return 0;
}

// GetData with a caller-supplied BVHTextureReferences table (Intersector.h:32-37).  Output per record: the cndl_hit_material
// layout (48 bytes) — Normal, UV, Emissivity, Alpha, Mesh, Albedo, and the index of the sampler texture() was called with
// (-1: not called); tex_uv receives the UV texture() was called with.  Every Mesh must be inside the table (the caller checks).
int ref_glsl_get_data_material(const void* tris, const void* verts, const void* ents, int32_t n_ents, const void* refs, const void* hits_, uint64_t R,
                               void* out_, float* tex_uv) {
    const void* nodes = nullptr;
    const uint64_t n_nodes = 0;
    BIND(ref_stackless)
    ref_stackless::BVHTextureReferences = static_cast<const ref_stackless::TextureReferences*>(refs);
    for (int i = 0; i < 512; ++i) ref_stackless::Textures[i].id = i;
    const Hit32* hits = static_cast<const Hit32*>(hits_);
    struct Mat48 { float nx, ny, nz, u, v, emissivity, alpha; int32_t mesh; float albedo[3]; int32_t albedo_ref; };
    static_assert(sizeof(Mat48) == 48, "record layout");
    Mat48* out = static_cast<Mat48*>(out_);
    for (uint64_t i = 0; i < R; ++i) {
        const Hit32& h = hits[i];
        glm::vec3 normal(0.0f), albedo(0.0f);
        float emissivity = 0.0f, alpha = 0.0f;
        glsl_prelude::g_sampled_id = -1;
        glsl_prelude::g_sampled_uv = glm::vec2(0.0f);
        ref_stackless::GetData(glm::vec4(h.t, h.u, h.v, h.w), h.mesh, h.tri, h.entity, normal, albedo, emissivity, alpha);
        const bool miss = h.t < 0.0f || h.mesh < 0;
        glm::vec2 uv(0.0f);
        if (!miss) {
            const auto& T = ref_stackless::BVHTris[h.tri];
            const auto &A = ref_stackless::BVHVertices[T.PackedData[0]], &B = ref_stackless::BVHVertices[T.PackedData[1]], &C = ref_stackless::BVHVertices[T.PackedData[2]];
            uv = (glm::unpackHalf2x16(A.PackedData.w) * h.u) + (glm::unpackHalf2x16(B.PackedData.w) * h.v) + (glm::unpackHalf2x16(C.PackedData.w) * h.w);
        }
        out[i] = Mat48{normal.x, normal.y, normal.z, uv.x, uv.y, emissivity, alpha, h.mesh, {albedo.x, albedo.y, albedo.z}, glsl_prelude::g_sampled_id};
        if (tex_uv) { tex_uv[2 * i] = glsl_prelude::g_sampled_uv.x; tex_uv[2 * i + 1] = glsl_prelude::g_sampled_uv.y; }
    }
    return 0;
}

}  // extern "C"
