// The kernels either side of traversal in a wavefront pipeline:
//   * ray generation from the hit records of the previous batch — diffuse bounce rays
//     (DiffuseTrace.glsl:445-446,:516-517 + CosWeightedHemisphere, Include/Sampling.glsl:1-12), specular
//     rays (SpecularTrace.glsl:102-135,:512-513 + SampleGGXVNDF, Include/Sampling.glsl:63-83) and shadow
//     rays — compacted so that rays which missed produce nothing, optionally emitted octant-major;
//   * the stable 8-way partition both the generator and cndl_set_traversal_mode(sort_rays) use;
//   * GetData, the hit attribute fetch that follows traversal.
// The shaders' fract(sin()) hash is replaced by a counter-based generator: the rays written
// here are INPUTS to traversal.  Compaction and bucketing are scans, so the output order is deterministic.
#include "kernels.cuh"
#include "scan.cuh"
#include "exact_trig.cuh"

#include <cuda_fp16.h>

namespace cndl {

namespace {

constexpr unsigned FULLMASK = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// Stable 8-way partition.  Element e has a key in 0..7, or 8 = dropped.  Tiles of 2048 elements per
// block; within a tile the order (item, warp, lane) is the element order.  Pass COUNT writes the
// tile's count per key (key-major: block_counts[key * nblocks + block]); an exclusive scan over that
// array gives every (key, tile) its base; pass RANK adds the element's rank inside its tile.
constexpr int kP8Block = 256, kP8Items = 8, kP8Tile = kP8Block * kP8Items;

struct KeyFromBytes {
    const unsigned char* __restrict__ k;
    __device__ __forceinline__ unsigned operator()(unsigned e) const { return k[e]; }
};
struct KeyFromRayOctant {
    const cndl_ray* __restrict__ rays;
    __device__ __forceinline__ unsigned operator()(unsigned e) const {
        const float4 d = __ldg(reinterpret_cast<const float4*>(rays + e) + 1);
        return (d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u);
    }
};

__device__ __forceinline__ unsigned same_key_mask(unsigned key, unsigned b0, unsigned b1, unsigned b2, unsigned valid) {
    return valid & ((key & 1u) ? b0 : ~b0) & ((key & 2u) ? b1 : ~b1) & ((key & 4u) ? b2 : ~b2);
}

// out_is_order: out[position] = element (a gather list); else out[element] = position (a scatter map).
template <bool RANK, class KeyFn>
__global__ void __launch_bounds__(kP8Block) partition8_kernel(KeyFn keyfn, unsigned n, unsigned nblocks, int* __restrict__ block_counts,
                                                              const int* __restrict__ block_bases, unsigned* __restrict__ out, int out_is_order) {
    __shared__ int s_cnt[kP8Items][kP8Block / 32][8];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned first = blockIdx.x * kP8Tile;
    unsigned key[kP8Items];
#pragma unroll
    for (int j = 0; j < kP8Items; ++j) {
        const unsigned e = first + j * kP8Block + threadIdx.x;
        key[j] = e < n ? keyfn(e) : 8u;
        const unsigned b0 = __ballot_sync(FULLMASK, key[j] & 1u), b1 = __ballot_sync(FULLMASK, key[j] & 2u), b2 = __ballot_sync(FULLMASK, key[j] & 4u);
        const unsigned valid = __ballot_sync(FULLMASK, key[j] < 8u);
        if (lane < 8) s_cnt[j][warp][lane] = __popc(same_key_mask(lane, b0, b1, b2, valid));
    }
    __syncthreads();
    if (threadIdx.x < 8) {  // exclusive prefix over (item, warp) for key = threadIdx.x
        int run = 0;
        for (int j = 0; j < kP8Items; ++j)
            for (int w = 0; w < kP8Block / 32; ++w) {
                const int c = s_cnt[j][w][threadIdx.x];
                s_cnt[j][w][threadIdx.x] = run;
                run += c;
            }
        if (!RANK) block_counts[threadIdx.x * nblocks + blockIdx.x] = run;
    }
    if (!RANK) return;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kP8Items; ++j) {
        const unsigned e = first + j * kP8Block + threadIdx.x;
        const unsigned b0 = __ballot_sync(FULLMASK, key[j] & 1u), b1 = __ballot_sync(FULLMASK, key[j] & 2u), b2 = __ballot_sync(FULLMASK, key[j] & 4u);
        const unsigned valid = __ballot_sync(FULLMASK, key[j] < 8u);
        if (key[j] < 8u) {
            const unsigned pos = (unsigned)block_bases[key[j] * nblocks + blockIdx.x] + (unsigned)s_cnt[j][warp][key[j]] +
                                 (unsigned)__popc(same_key_mask(key[j], b0, b1, b2, valid) & ((1u << lane) - 1u));
            if (out_is_order) out[pos] = e;
            else out[e] = pos;
        }
    }
}

struct P8Scratch { int *counts, *bases, *block_sums, *total; };
__host__ size_t p8_scratch_ints(size_t n) {
    const size_t nb = (n + kP8Tile - 1) / kP8Tile;
    return 16 * nb + (8 * nb) / kScanTile + 16;
}
__host__ P8Scratch p8_carve(int* scratch, size_t n) {
    const size_t nb = (n + kP8Tile - 1) / kP8Tile;
    P8Scratch p;
    p.counts = scratch;
    p.bases = scratch + 8 * nb;
    p.block_sums = scratch + 16 * nb;
    p.total = p.block_sums + (8 * nb) / kScanTile + 8;
    return p;
}

template <class KeyFn>
void partition8(KeyFn keyfn, size_t n, unsigned* out, int out_is_order, const P8Scratch& p, cudaStream_t st, LaunchCounter& lc) {
    const unsigned nb = (unsigned)((n + kP8Tile - 1) / kP8Tile);
    partition8_kernel<false, KeyFn><<<nb, kP8Block, 0, st>>>(keyfn, (unsigned)n, nb, p.counts, nullptr, nullptr, 0);
    lc.n++;
    exclusive_scan(p.counts, (int)(8 * nb), p.bases, p.block_sums, p.total, st, lc);
    partition8_kernel<true, KeyFn><<<nb, kP8Block, 0, st>>>(keyfn, (unsigned)n, nb, nullptr, p.bases, out, out_is_order);
    lc.n++;
}

// ---------------------------------------------------------------------------------------------
// ray generation
//
// Bit-exact contract (the test-side CPU checker states the same arithmetic and is pinned against the shader
// functions compiled from the reference): every float operation is a separately rounded IEEE operation in the order the
// shader text (over glm 0.9.8.5) evaluates it — normalize(v) = v * (1 / sqrt(dot(v, v))), dot = (x*x + y*y) + z*z,
// cross and reflect in glm's forms — the random numbers come from the counter stream below, and sin / cos / acos / pow are
// the defined functions of exact_trig.cuh.

__device__ __forceinline__ unsigned pcg_hash(unsigned v) {
    unsigned s = v * 747796405u + 2891336453u;
    unsigned w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
__device__ __forceinline__ float u01(unsigned h) { return fmul((float)(h >> 8), 1.0f / 16777216.0f); }

// The stream that stands in for the shaders' hash2(): draw n of stream k is u01(pcg_hash(k + n * 0x9E3779B9)); a hash2()
// call takes two consecutive draws.  The stream key of element e under `seed` is pcg_hash(seed ^ pcg_hash(e)).
struct Hash2 {
    unsigned k, n;
    __device__ __forceinline__ float next() { return u01(pcg_hash(k + (n++) * 0x9E3779B9u)); }
};
__device__ __forceinline__ unsigned stream_key(unsigned seed, unsigned element) { return pcg_hash(seed ^ pcg_hash(element)); }

__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return {fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
__device__ __forceinline__ V3 vscale(V3 a, float s) { return {fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)}; }
__device__ __forceinline__ V3 vnormalize(V3 v) { return vscale(v, fdiv(1.0f, __fsqrt_rn(vdot(v, v)))); }      // glm compute_normalize
__device__ __forceinline__ V3 vreflect(V3 I, V3 N) { return vsub(I, vscale(vscale(N, vdot(N, I)), 2.0f)); }    // glm compute_reflect: I - N * dot(N, I) * 2

#define CNDL_PI2_F (2.0f * 3.14159265359f)

// CosWeightedHemisphere, Include/Sampling.glsl:1-12
__device__ __forceinline__ V3 cos_weighted_hemisphere(V3 n, float r_x, float r_y) {
    const V3 uu = vnormalize(vcross(n, V3{0.0f, 1.0f, 1.0f}));
    const V3 vv = vcross(uu, n);
    const float ra = __fsqrt_rn(r_y);
    float sn, cs;
    xm::xsincos(fmul(CNDL_PI2_F, r_x), sn, cs);
    const float rx = fmul(ra, cs), ry = fmul(ra, sn), rz = __fsqrt_rn(fsub(1.0f, r_y));
    const V3 rr = vadd(vadd(vscale(uu, rx), vscale(vv, ry)), vscale(n, rz));
    return vnormalize(rr);
}

// SampleGGXVNDF, Include/Sampling.glsl:63-83
__device__ __forceinline__ V3 sample_ggx_vndf(V3 N, float roughness, float xi_x, float xi_y) {
    const float alpha = fmul(roughness, roughness), alpha2 = fmul(alpha, alpha);
    const float phi = fmul(CNDL_PI2_F, xi_x);
    const float cos_theta = __fsqrt_rn(fdiv(fsub(1.0f, xi_y), fadd(1.0f, fmul(fsub(alpha2, 1.0f), xi_y))));
    const float sin_theta = __fsqrt_rn(fsub(1.0f, fmul(cos_theta, cos_theta)));
    float sn, cs;
    xm::xsincos(phi, sn, cs);
    const V3 H = {fmul(cs, sin_theta), fmul(sn, sin_theta), cos_theta};
    const V3 up = fabsf(N.z) < 0.999f ? V3{0.0f, 0.0f, 1.0f} : V3{1.0f, 0.0f, 0.0f};
    const V3 tangent = vnormalize(vcross(up, N));
    const V3 bitangent = vcross(N, tangent);
    const V3 sv = vadd(vadd(vscale(tangent, H.x), vscale(bitangent, H.y)), vscale(N, H.z));
    return vnormalize(sv);
}

// StochasticReflectionDirection, SpecularTrace.glsl:102-135
__device__ __forceinline__ V3 stochastic_reflection_direction(V3 incident, V3 normal, float roughness, Hash2& h) {
    if (roughness < 0.01f) return vreflect(incident, normal);
    V3 microfacet = normal;
    for (int i = 0; i < 12; ++i) {
        const float a = fmul(h.next(), 0.8f), b = fmul(h.next(), 0.7f);  // hash2() * TailControl
        const V3 s = sample_ggx_vndf(normal, roughness, a, b);
        if (vdot(s, normal) > 0.001f) { microfacet = s; break; }
    }
    return vreflect(incident, microfacet);
}

// SampleCone(Direction, Xi, CosTheta) = mat3(T, B, L) * SampleCone(Xi, CosTheta), Include/Sampling.glsl:43-61
__device__ __forceinline__ V3 sample_cone(V3 L, float xi_x, float xi_y, float cos_theta_max) {
    const float cos_theta = fadd(fsub(1.0f, xi_x), fmul(xi_x, cos_theta_max));
    const float sin_theta = __fsqrt_rn(fsub(1.0f, fmul(cos_theta, cos_theta)));
    const float phi = fmul(fmul(xi_y, 3.14159265359f), 2.0f);
    float sn, cs;
    xm::xsincos(phi, sn, cs);
    const V3 l = {fmul(sin_theta, cs), fmul(sin_theta, sn), cos_theta};
    const V3 T = vnormalize(vcross(L, V3{0.0f, 1.0f, 1.0f}));
    const V3 B = vcross(T, L);
    return {fadd(fadd(fmul(T.x, l.x), fmul(B.x, l.y)), fmul(L.x, l.z)), fadd(fadd(fmul(T.y, l.x), fmul(B.y, l.y)), fmul(L.y, l.z)),
            fadd(fadd(fmul(T.z, l.x), fmul(B.z, l.y)), fmul(L.z, l.z))};
}

// LambertBRDF, UpdateRadianceProbes.glsl:351-362 (PI = 3.1415926535 in that file)
__device__ __forceinline__ V3 lambert_brdf(float hx, float hy, float hz) {
    const float phi = fmul(2.0f * 3.1415926535f, hx);
    const float cos_theta = fsub(fmul(2.0f, hy), 1.0f);
    const float theta = xm::xacos(cos_theta);
    const float r = xm::xpow(hz, 1.0f / 3.0f);
    float sp, cp, st, ct;
    xm::xsincos(phi, sp, cp);
    xm::xsincos(theta, st, ct);
    return {fmul(fmul(r, st), cp), fmul(fmul(r, st), sp), fmul(r, ct)};
}

struct GenParams {
    int kind, spp, bucket;
    unsigned seed;
    float offset, tmax, roughness, lx, ly, lz, cone;
};

struct HitFrame {  // what every sample of one hit shares
    V3 p;   // hit point: RayOrigin + RayDirection * TUVW.x (DiffuseTrace.glsl:511)
    V3 n;   // geometric normal, world space, turned against the incoming ray
    V3 in;  // incoming direction
    bool valid;
};

__device__ __forceinline__ HitFrame hit_frame(const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits, const float4* __restrict__ tri48,
                                              const cndl_entity* __restrict__ ents, unsigned i) {
    HitFrame f;
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(hits + i));
    f.valid = h0.x > 0.0f;
    if (!f.valid) return f;
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(hits + i) + 1);
    const float4 ra = __ldg(reinterpret_cast<const float4*>(rays + i)), rb = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
    const V3 o = {ra.x, ra.y, ra.z}, d = {rb.x, rb.y, rb.z};
    f.p = vadd(o, vscale(d, h0.x));
    f.in = d;
    const float4 c = __ldg(tri48 + kTriStride * (size_t)h1.y + 2);  // .yzw = cross(v1 - v0, v2 - v0), object space
    const float* m = ents[h1.z].model;
    const V3 nw = {fadd(fadd(fmul(__ldg(m + 0), c.y), fmul(__ldg(m + 4), c.z)), fmul(__ldg(m + 8), c.w)),
                   fadd(fadd(fmul(__ldg(m + 1), c.y), fmul(__ldg(m + 5), c.z)), fmul(__ldg(m + 9), c.w)),
                   fadd(fadd(fmul(__ldg(m + 2), c.y), fmul(__ldg(m + 6), c.z)), fmul(__ldg(m + 10), c.w))};
    const float dd = vdot(nw, nw);
    if (!(dd > 0.0f) || isinf(dd)) { f.valid = false; return f; }  // degenerate triangle: no ray
    V3 n = vscale(nw, fdiv(1.0f, __fsqrt_rn(dd)));
    if (vdot(n, d) > 0.0f) n = vneg(n);
    f.n = n;
    return f;
}

// The ray of stream element `element` for hit frame f.  false: this sample emits no ray.
__device__ __forceinline__ bool gen_ray(const GenParams& g, const HitFrame& f, unsigned element, float4& o, float4& d) {
    Hash2 h{stream_key(g.seed, element), 0u};
    float off = g.offset;
    V3 dir;
    if (g.kind == CNDL_GEN_DIFFUSE) {
        const float a = h.next(), b = h.next();
        dir = cos_weighted_hemisphere(f.n, a, b);
    } else if (g.kind == CNDL_GEN_SPECULAR) {
        dir = stochastic_reflection_direction(f.in, f.n, fmul(g.roughness, 0.9f), h);  // SpecularTrace.glsl:513
        if (g.offset < 0.0f) {  // mix(0.05f, 0.1f, clamp(PBR.x * 1.4f, 0, 1)) (:512); glm: x + a * (y - x), clamp = min(max(x, lo), hi)
            const float a = glsl_min(glsl_max(fmul(g.roughness, 1.4f), 0.0f), 1.0f);
            off = fadd(0.05f, fmul(a, fsub(0.1f, 0.05f)));
        }
    } else {
        // shadow ray towards the light, jittered inside a cone with the reference's SampleCone; surfaces facing away need no ray
        const V3 L = {g.lx, g.ly, g.lz};
        if (!(vdot(f.n, L) > 0.0f)) return false;
        const float a = h.next(), b = h.next();
        const float cos_max = __fsqrt_rn(fsub(1.0f, fmul(g.cone, g.cone)));
        dir = vnormalize(sample_cone(L, a, b, cos_max));
    }
    const V3 org = vadd(f.p, vscale(f.n, off));
    o = make_float4(org.x, org.y, org.z, 0.0f);
    d = make_float4(dir.x, dir.y, dir.z, g.tmax);
    return true;
}

// Pass 1: the partition key of every potential output ray (8 = none).
__global__ void gen_keys_kernel(GenParams g, const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits, const unsigned* __restrict__ ids,
                                const float4* __restrict__ tri48, const cndl_entity* __restrict__ ents, const unsigned* __restrict__ d_R, unsigned R,
                                unsigned char* __restrict__ keys) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const bool live = !d_R || i < __ldg(d_R);  // d_R: the batch length when it is only known on the device (slots beyond it emit nothing)
    HitFrame f;
    f.valid = false;
    if (live) f = hit_frame(rays, hits, tri48, ents, i);
    const unsigned id = ids && live ? __ldg(ids + i) : i;
    for (int s = 0; s < g.spp; ++s) {
        unsigned key = 8;
        float4 o, d;
        if (f.valid && gen_ray(g, f, id * (unsigned)g.spp + (unsigned)s, o, d))
            key = g.bucket ? ((d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u)) : 0u;
        keys[(size_t)i * g.spp + s] = (unsigned char)key;
    }
}

// Pass 3: the rays, written where the partition put them.
__global__ void gen_emit_kernel(GenParams g, const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits, const unsigned* __restrict__ ids,
                                const float4* __restrict__ tri48, const cndl_entity* __restrict__ ents, unsigned R, const unsigned char* __restrict__ keys,
                                const unsigned* __restrict__ dest, cndl_ray* __restrict__ out, unsigned* __restrict__ parent, unsigned* __restrict__ ids_out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    bool any = false;
    for (int s = 0; s < g.spp; ++s) any = any || keys[(size_t)i * g.spp + s] < 8;
    if (!any) return;
    const HitFrame f = hit_frame(rays, hits, tri48, ents, i);
    const unsigned id = ids ? __ldg(ids + i) : i;
    for (int s = 0; s < g.spp; ++s) {
        const size_t e = (size_t)i * g.spp + s;
        if (keys[e] >= 8) continue;
        float4 o, d;
        const unsigned element = id * (unsigned)g.spp + (unsigned)s;
        gen_ray(g, f, element, o, d);
        const unsigned pos = dest[e];
        float4* q = reinterpret_cast<float4*>(out + pos);
        q[0] = o;
        q[1] = d;
        if (parent) parent[pos] = i;
        if (ids_out) ids_out[pos] = element;
    }
}

// The generator of the frame-level call: one CTA per kRaySegment potential rays (element e = input slot * spp + sample).  Every
// thread forms its four rays ONCE, the CTA orders them by direction octant (a stable counting sort over nine keys — ballots per
// warp, one scan of the 9 x 32 partial counts) and writes them into ITS segment of the output, live rays first.  No global scan,
// no second evaluation of the sample: the traversal learns which slots of a segment are live from seg_counts[segment]
// (RayOrder::seg_counts), and consecutive slots of a segment hold rays of one octant from neighbouring pixels.
// in_seg_counts (optional): the input batch is itself such a segmented batch (a further bounce).
template <int TPB>
__global__ void __launch_bounds__(TPB) gen_tile_kernel(GenParams g, const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits,
                                                       const unsigned* __restrict__ ids, const float4* __restrict__ tri48, const cndl_entity* __restrict__ ents,
                                                       const unsigned* __restrict__ in_seg_counts, unsigned n_elems, unsigned char* __restrict__ keys,
                                                       unsigned* __restrict__ dest, cndl_ray* __restrict__ out, unsigned* __restrict__ ids_out,
                                                       unsigned* __restrict__ seg_counts, unsigned* __restrict__ total, unsigned* __restrict__ oct_cursor,
                                                       unsigned* __restrict__ oct_list, unsigned oct_stride) {
    constexpr int ROUNDS = kRaySegment / TPB, WARPS = TPB / 32, PER_KEY = ROUNDS * WARPS;  // PER_KEY = kRaySegment / 32
    __shared__ int s_cnt[9 * PER_KEY];  // [key][round * WARPS + warp]
    __shared__ unsigned s_oct_base[8];
    const unsigned tid = threadIdx.x, lane = tid & 31u, wp = tid >> 5;
    const unsigned cta_base = blockIdx.x * (unsigned)kRaySegment;
    for (unsigned k = tid; k < 9u * PER_KEY; k += TPB) s_cnt[k] = 0;
    __syncthreads();
    float4 o[ROUNDS], d[ROUNDS];
    unsigned key[ROUNDS], rank[ROUNDS], element[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const unsigned e = cta_base + (unsigned)r * (unsigned)TPB + tid;
        key[r] = 8;
        element[r] = 0;
        if (e < n_elems) {
            const unsigned i = e / (unsigned)g.spp, smp = e - i * (unsigned)g.spp;
            const bool live = !in_seg_counts || (i & (unsigned)(kRaySegment - 1)) < __ldg(in_seg_counts + (i >> kRaySegmentShift));
            if (live) {
                const HitFrame f = hit_frame(rays, hits, tri48, ents, i);
                const unsigned id = ids ? __ldg(ids + i) : i;
                element[r] = id * (unsigned)g.spp + smp;
                if (f.valid && gen_ray(g, f, element[r], o[r], d[r]))
                    key[r] = g.bucket ? ((d[r].x > 0.0f ? 1u : 0u) | (d[r].y > 0.0f ? 2u : 0u) | (d[r].z > 0.0f ? 4u : 0u)) : 0u;
            }
        }
        const unsigned same = __match_any_sync(0xFFFFFFFFu, key[r]);
        rank[r] = (unsigned)__popc(same & ((1u << lane) - 1u));
        if (rank[r] == 0) s_cnt[key[r] * PER_KEY + r * WARPS + wp] = __popc(same);
    }
    __syncthreads();
    if (wp == 0) {  // exclusive scan of the 9 * PER_KEY counts in (key, round, warp) order
        constexpr int PER = 9 * PER_KEY / 32;
        int v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = s_cnt[lane * PER + k]; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if ((int)lane >= off) incl += up;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k) { s_cnt[lane * PER + k] = run; run += v[k]; }
    }
    __syncthreads();
    const unsigned n_live = (unsigned)s_cnt[8 * PER_KEY];  // where key 8 (no ray) begins
    if (oct_list) {
        // octant-major over the WHOLE batch without a global scan: the segment's run of octant o is appended to the batch's list of
        // octant o (one atomic per run; the lists are RayOrder's bucketed form).  Which segment comes first inside an octant depends
        // on scheduling; results do not.
        if (tid < 8) {
            const unsigned n = (unsigned)(s_cnt[(tid + 1) * PER_KEY] - s_cnt[tid * PER_KEY]);
            s_oct_base[tid] = n ? atomicAdd(oct_cursor + tid, n) : 0u;
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const unsigned e = cta_base + (unsigned)r * (unsigned)TPB + tid;
        if (e >= n_elems) continue;
        keys[e] = (unsigned char)key[r];
        if (key[r] >= 8) continue;
        const unsigned in_seg = (unsigned)s_cnt[key[r] * PER_KEY + r * WARPS + wp] + rank[r];
        const unsigned pos = cta_base + in_seg;
        dest[e] = pos;
        if (oct_list) oct_list[(size_t)key[r] * oct_stride + s_oct_base[key[r]] + (in_seg - (unsigned)s_cnt[key[r] * PER_KEY])] = pos;
        stg256(reinterpret_cast<float4*>(out + pos), o[r], d[r]);
        if (ids_out) ids_out[pos] = element[r];
    }
    if (tid == 0) {
        seg_counts[blockIdx.x] = n_live;
        if (n_live) atomicAdd(total, n_live);
    }
}

// Probe-update rays (UpdateRadianceProbes.glsl:408-427): probe (x, y, z) of a res.x * res.y * res.z grid -> ray index
// (z * res.y + y) * res.x + x; RayOrigin = u_BoxOrigin + (vec3(Pixel) / u_Resolution * 2 - 1) * u_Size; direction =
// ImportanceSample() with its importance branch off = normalize(LambertBRDF(vec3(hash2(), hash2().x))) (:365-374).
__global__ void probe_rays_kernel(float3 org, float3 size, int3 res, unsigned seed, cndl_ray* __restrict__ out) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = (unsigned)res.x * (unsigned)res.y * (unsigned)res.z;
    if (idx >= n) return;
    const int x = (int)(idx % (unsigned)res.x), y = (int)((idx / (unsigned)res.x) % (unsigned)res.y), z = (int)(idx / ((unsigned)res.x * (unsigned)res.y));
    const V3 tex = {fdiv((float)x, (float)res.x), fdiv((float)y, (float)res.y), fdiv((float)z, (float)res.z)};
    const V3 clip = {fsub(fmul(tex.x, 2.0f), 1.0f), fsub(fmul(tex.y, 2.0f), 1.0f), fsub(fmul(tex.z, 2.0f), 1.0f)};
    const V3 o = {fadd(org.x, fmul(clip.x, size.x)), fadd(org.y, fmul(clip.y, size.y)), fadd(org.z, fmul(clip.z, size.z))};
    Hash2 h{stream_key(seed, idx), 0u};
    const float a = h.next(), b = h.next(), c = h.next();
    const V3 d = vnormalize(lambert_brdf(a, b, c));
    float4* q = reinterpret_cast<float4*>(out + idx);
    q[0] = make_float4(o.x, o.y, o.z, 0.0f);
    q[1] = make_float4(d.x, d.y, d.z, 1000000.0f);
}

// GetData (…/Include/TraverseBVHStackless.glsl:370-408) without the texture fetch.  MATERIAL adds the Albedo decision (:393-404)
// from the per-mesh BVHTextureReferences table.
template <bool MATERIAL>
__global__ void get_data_kernel(const int4* __restrict__ tris, const float4* __restrict__ verts, const cndl_entity* __restrict__ ents,
                                const cndl_texture_reference* __restrict__ refs, unsigned n_refs, const cndl_hit* __restrict__ hits, unsigned R,
                                void* __restrict__ out_, unsigned* __restrict__ out_of_table) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(hits + i));
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(hits + i) + 1);
    float4 o0 = make_float4(-1.0f, -1.0f, -1.0f, 0.0f);  // Normal = vec3(-1) on a miss (SL:377-382)
    float4 o1 = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(h1.x));
    float4 o2 = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));  // Albedo = vec3(0), no texture
    if (!(h0.x < 0.0f || h1.x < 0)) {
        const int4 t = __ldg(tris + h1.y);
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(verts + 2 * (size_t)t.x + 1));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(verts + 2 * (size_t)t.y + 1));
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(verts + 2 * (size_t)t.z + 1));
        auto lo = [](unsigned p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))); };
        auto hi = [](unsigned p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); };
        auto mix3 = [&](float fa, float fb, float fc) { return fadd(fadd(fmul(fa, h0.y), fmul(fb, h0.z)), fmul(fc, h0.w)); };
        const float u = mix3(lo(a.w), lo(b.w), lo(c.w)), v = mix3(hi(a.w), hi(b.w), hi(c.w));
        const float nx = mix3(lo(a.x), lo(b.x), lo(c.x)), ny = mix3(hi(a.x), hi(b.x), hi(c.x)), nz = mix3(lo(a.y), lo(b.y), lo(c.y));
        const float inv_len = fdiv(1.0f, __fsqrt_rn(fadd(fadd(fmul(nx, nx), fmul(ny, ny)), fmul(nz, nz))));
        o0 = make_float4(fmul(nx, inv_len), fmul(ny, inv_len), fmul(nz, inv_len), u);
        o1 = make_float4(v, __int_as_float(__ldg(&ents[h1.z].data[0])), __int_as_float(__ldg(&ents[h1.z].data[1])), __int_as_float(h1.x));
        if (MATERIAL) {
            if ((unsigned)h1.x >= n_refs) {  // the shader would read past its SSBO: flagged, never guessed
                o2.w = __int_as_float(-2);
                atomicAdd(out_of_table, 1u);
            } else {
                const float4 color = __ldg(reinterpret_cast<const float4*>(refs + h1.x));
                const int ref = __ldg(&refs[h1.x].albedo);
                // `Ref > -1 && Mesh > -1 && TUVW.x > 0.` (SL:397): t == 0 takes the ModelColor branch
                if (ref > -1 && h0.x > 0.0f) o2.w = __int_as_float(ref);
                else o2 = make_float4(color.x, color.y, color.z, __int_as_float(-1));
            }
        }
    }
    float4* p = reinterpret_cast<float4*>(out_) + (size_t)i * (MATERIAL ? 3 : 2);
    p[0] = o0;
    p[1] = o1;
    if (MATERIAL) p[2] = o2;
}

}  // namespace

void launch_get_data(const SceneView& s, const float4* verts, const cndl_hit* hits, size_t R, cndl_hit_attr* out, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    get_data_kernel<false><<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(s.tris, verts, s.ents, nullptr, 0u, hits, (unsigned)R, out, nullptr);
    lc.n++;
}

void launch_get_data_material(const SceneView& s, const float4* verts, const cndl_texture_reference* refs, size_t n_refs, const cndl_hit* hits, size_t R,
                              cndl_hit_material* out, unsigned* out_of_table, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    get_data_kernel<true><<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(s.tris, verts, s.ents, refs, (unsigned)n_refs, hits, (unsigned)R, out, out_of_table);
    lc.n++;
}

size_t octant_partition_scratch_ints(size_t R) { return p8_scratch_ints(R); }

void launch_octant_partition(const cndl_ray* rays, size_t R, unsigned* order, int* scratch, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    partition8(KeyFromRayOctant{rays}, R, order, 1, p8_carve(scratch, R), stream, lc);
}

size_t generate_rays_scratch_ints(size_t R, int spp) {
    const size_t n = R * (size_t)spp;
    return (n + 3) / 4 + n + p8_scratch_ints(n) + 16;
}

// Writes the rays and leaves their number in device memory (returned through *d_count_out, valid until the scratch is
// reused); when h_count is given, also copies it to the host and synchronises `stream`.  d_R (optional): the input batch
// length on the device, R being its upper bound.
cudaError_t generate_rays(const SceneView& s, const cndl_raygen_params& prm, const cndl_ray* rays, const cndl_hit* hits, size_t R, const unsigned* d_R,
                          cndl_ray* out, unsigned* parent, int* scratch, const unsigned** d_count_out, size_t* h_count, cudaStream_t stream,
                          LaunchCounter& lc) {
    if (h_count) *h_count = 0;
    if (d_count_out) *d_count_out = nullptr;
    if (R == 0) return cudaSuccess;
    GenParams g;
    g.kind = prm.kind;
    g.spp = prm.spp;
    g.bucket = (prm.flags & CNDL_GEN_BUCKET_OCTANTS) ? 1 : 0;
    g.seed = prm.seed;
    g.offset = prm.offset;
    g.tmax = prm.tmax;
    g.roughness = prm.roughness;
    g.lx = prm.light_dir[0]; g.ly = prm.light_dir[1]; g.lz = prm.light_dir[2];
    g.cone = prm.light_cone;
    const size_t n = R * (size_t)prm.spp;
    unsigned char* keys = reinterpret_cast<unsigned char*>(scratch);
    unsigned* dest = reinterpret_cast<unsigned*>(scratch + (n + 3) / 4);
    const P8Scratch p = p8_carve(scratch + (n + 3) / 4 + n, n);
    const unsigned grid = (unsigned)((R + 255) / 256);
    gen_keys_kernel<<<grid, 256, 0, stream>>>(g, rays, hits, prm.d_ids_in, s.tri48, s.ents, d_R, (unsigned)R, keys);
    lc.n++;
    partition8(KeyFromBytes{keys}, n, dest, 0, p, stream, lc);
    gen_emit_kernel<<<grid, 256, 0, stream>>>(g, rays, hits, prm.d_ids_in, s.tri48, s.ents, (unsigned)R, keys, dest, out, parent, prm.d_ids_out);
    lc.n++;
    if (d_count_out) *d_count_out = reinterpret_cast<const unsigned*>(p.total);
    if (h_count) {
        int h_total = 0;
        cudaError_t e = cudaMemcpyAsync(&h_total, p.total, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        *h_count = (size_t)h_total;
    }
    return cudaGetLastError();
}

// The segmented variant (gen_tile_kernel): element e's ray lands inside segment e / kRaySegment of `out`; seg_counts[segment] = live
// rays of the segment, *total += all of them (the caller zeroes it).  keys / dest are left in `scratch` like generate_rays does.
// oct_list (optional, with CNDL_GEN_BUCKET_OCTANTS): 8 lists of oct_stride entries each; list o receives the positions of the rays of
// octant o, oct_cursor[o] (zeroed by the caller) their number.
cudaError_t generate_rays_tiled(const SceneView& s, const cndl_raygen_params& prm, const cndl_ray* rays, const cndl_hit* hits, size_t R,
                                const unsigned* in_seg_counts, cndl_ray* out, int* scratch, unsigned* seg_counts, unsigned* total, unsigned* oct_cursor,
                                unsigned* oct_list, size_t oct_stride, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return cudaSuccess;
    GenParams g;
    g.kind = prm.kind;
    g.spp = prm.spp;
    g.bucket = (prm.flags & CNDL_GEN_BUCKET_OCTANTS) ? 1 : 0;
    g.seed = prm.seed;
    g.offset = prm.offset;
    g.tmax = prm.tmax;
    g.roughness = prm.roughness;
    g.lx = prm.light_dir[0]; g.ly = prm.light_dir[1]; g.lz = prm.light_dir[2];
    g.cone = prm.light_cone;
    const size_t n = R * (size_t)prm.spp;
    unsigned char* keys = reinterpret_cast<unsigned char*>(scratch);
    unsigned* dest = reinterpret_cast<unsigned*>(scratch + (n + 3) / 4);
    const unsigned grid = (unsigned)((n + kRaySegment - 1) / kRaySegment);
    // 512 threads x 2 rays: 54 registers, 32 resident warps per SM (256 x 4: 68 registers, 24 warps; 1024 x 1: 38 registers, one CTA per
    // SM, slower) — 1.222 / 1.226 / 1.244 ms per 1080p frame
    gen_tile_kernel<512><<<grid, 512, 0, stream>>>(g, rays, hits, prm.d_ids_in, s.tri48, s.ents, in_seg_counts, (unsigned)n, keys, dest, out, prm.d_ids_out, seg_counts,
                                                   total, oct_cursor, oct_list, (unsigned)oct_stride);
    lc.n++;
    return cudaGetLastError();
}

// keys / dest of the last generate_rays call on `scratch` (element e = i * spp + s: keys[e] < 8 iff it produced a ray, at dest[e])
void generate_rays_maps(int* scratch, size_t R, int spp, const unsigned char** keys, const unsigned** dest) {
    const size_t n = R * (size_t)spp;
    *keys = reinterpret_cast<const unsigned char*>(scratch);
    *dest = reinterpret_cast<const unsigned*>(scratch + (n + 3) / 4);
}

void launch_probe_rays(const float box_origin[3], const float size[3], const int res[3], unsigned seed, cndl_ray* out, cudaStream_t stream, LaunchCounter& lc) {
    const size_t n = (size_t)res[0] * (size_t)res[1] * (size_t)res[2];
    if (n == 0) return;
    probe_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(make_float3(box_origin[0], box_origin[1], box_origin[2]), make_float3(size[0], size[1], size[2]),
                                                                      make_int3(res[0], res[1], res[2]), seed, out);
    lc.n++;
}

}  // namespace cndl
