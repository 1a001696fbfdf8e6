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

__device__ __forceinline__ unsigned pcg_hash(unsigned v) {
    unsigned s = v * 747796405u + 2891336453u;
    unsigned w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
__device__ __forceinline__ float u01(unsigned h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

struct GenParams {
    int kind, spp, bucket;
    unsigned seed;
    float offset, tmax, roughness, lx, ly, lz, cone;
};

struct HitFrame {  // what every sample of one hit shares
    float px, py, pz;     // hit point
    float nx, ny, nz;     // geometric normal, world space, turned against the incoming ray
    float ix, iy, iz;     // incoming direction
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
    const float t = h0.x;
    f.px = ra.x + rb.x * t; f.py = ra.y + rb.y * t; f.pz = ra.z + rb.z * t;
    f.ix = rb.x; f.iy = rb.y; f.iz = rb.z;
    const float4 c = __ldg(tri48 + 3 * (size_t)h1.y + 2);
    const float* m = ents[h1.z].model;
    float nx = m[0] * c.y + m[4] * c.z + m[8] * c.w, ny = m[1] * c.y + m[5] * c.z + m[9] * c.w, nz = m[2] * c.y + m[6] * c.z + m[10] * c.w;
    const float inv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
    nx *= inv; ny *= inv; nz *= inv;
    if (nx * rb.x + ny * rb.y + nz * rb.z > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
    f.nx = nx; f.ny = ny; f.nz = nz;
    return f;
}

// Sample s of hit i: origin + direction.  false: this sample emits no ray.
__device__ __forceinline__ bool gen_ray(const GenParams& g, const HitFrame& f, unsigned i, int s, float4& o, float4& d) {
    const unsigned k = pcg_hash(g.seed ^ pcg_hash(i * (unsigned)g.spp + (unsigned)s));
    float off = g.offset, dx, dy, dz;
    if (g.kind == CNDL_GEN_DIFFUSE) {
        // CosWeightedHemisphere (Sampling.glsl:1-12): uu = normalize(cross(n, (0,1,1))), vv = cross(uu, n)
        float ux = f.ny - f.nz, uy = -f.nx, uz = f.nx;
        float inv = rsqrtf(fmaxf(ux * ux + uy * uy + uz * uz, 1e-30f));
        ux *= inv; uy *= inv; uz *= inv;
        const float vx = uy * f.nz - uz * f.ny, vy = uz * f.nx - ux * f.nz, vz = ux * f.ny - uy * f.nx;
        const float r1 = u01(k), r2 = u01(pcg_hash(k + 0x9E3779B9u));
        const float rad = sqrtf(r2), ang = 6.28318530718f * r1;
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float rx = rad * cs, ry = rad * sn, rz = sqrtf(1.0f - r2);
        dx = rx * ux + ry * vx + rz * f.nx; dy = rx * uy + ry * vy + rz * f.ny; dz = rx * uz + ry * vz + rz * f.nz;
    } else if (g.kind == CNDL_GEN_SPECULAR) {
        // StochasticReflectionDirection(Incident, Normal, PBR.x * 0.9) (SpecularTrace.glsl:102-135,:513)
        const float rough = g.roughness * 0.9f;
        float mx = f.nx, my = f.ny, mz = f.nz;  // Microfacet = Normal
        if (rough >= 0.01f) {
            const float alpha = rough * rough, alpha2 = alpha * alpha;
            // tangent frame of SampleGGXVNDF (Sampling.glsl:77-79)
            const bool upz = fabsf(f.nz) < 0.999f;
            const float ax = upz ? 0.0f : 1.0f, az = upz ? 1.0f : 0.0f;  // up
            float tx = -az * f.ny, ty = az * f.nx - ax * f.nz, tz = ax * f.ny;  // cross(up, N) with up.y = 0
            float inv = rsqrtf(fmaxf(tx * tx + ty * ty + tz * tz, 1e-30f));
            tx *= inv; ty *= inv; tz *= inv;
            const float bx = f.ny * tz - f.nz * ty, by = f.nz * tx - f.nx * tz, bz = f.nx * ty - f.ny * tx;  // cross(N, tangent)
            for (int t = 0; t < 12; ++t) {
                const float x1 = u01(pcg_hash(k + (unsigned)(2 * t + 1) * 0x9E3779B9u)) * 0.8f;  // TailControl (:118)
                const float x2 = u01(pcg_hash(k + (unsigned)(2 * t + 2) * 0x9E3779B9u)) * 0.7f;
                const float phi = 6.28318530718f * x1;
                const float ct = sqrtf((1.0f - x2) / (1.0f + (alpha2 - 1.0f) * x2)), st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
                float sp, cp;
                sincosf(phi, &sp, &cp);
                const float hx = cp * st, hy = sp * st, hz = ct;
                float sx = tx * hx + bx * hy + f.nx * hz, sy = ty * hx + by * hy + f.ny * hz, sz = tz * hx + bz * hy + f.nz * hz;
                inv = rsqrtf(fmaxf(sx * sx + sy * sy + sz * sz, 1e-30f));
                sx *= inv; sy *= inv; sz *= inv;
                if (sx * f.nx + sy * f.ny + sz * f.nz > 0.001f) { mx = sx; my = sy; mz = sz; break; }
            }
        }
        const float di = 2.0f * (mx * f.ix + my * f.iy + mz * f.iz);  // reflect(I, M) = I - 2 dot(M, I) M
        dx = f.ix - di * mx; dy = f.iy - di * my; dz = f.iz - di * mz;
        if (g.offset < 0.0f) off = 0.05f + 0.05f * fminf(fmaxf(g.roughness * 1.4f, 0.0f), 1.0f);  // mix(0.05, 0.1, clamp(PBR.x*1.4)) (:512)
    } else {
        // shadow ray towards the light, jittered inside a cone; surfaces facing away from the light need no ray
        if (f.nx * g.lx + f.ny * g.ly + f.nz * g.lz <= 0.0f) return false;
        const bool upz = fabsf(g.lz) < 0.999f;
        const float ax = upz ? 0.0f : 1.0f, az = upz ? 1.0f : 0.0f;
        float tx = -az * g.ly, ty = az * g.lx - ax * g.lz, tz = ax * g.ly;
        const float inv = rsqrtf(fmaxf(tx * tx + ty * ty + tz * tz, 1e-30f));
        tx *= inv; ty *= inv; tz *= inv;
        const float bx = g.ly * tz - g.lz * ty, by = g.lz * tx - g.lx * tz, bz = g.lx * ty - g.ly * tx;
        const float r1 = u01(k), r2 = u01(pcg_hash(k + 0x9E3779B9u));
        const float rad = sqrtf(r2) * g.cone, ang = 6.28318530718f * r1;
        float sn, cs;
        sincosf(ang, &sn, &cs);
        dx = g.lx + rad * (cs * tx + sn * bx); dy = g.ly + rad * (cs * ty + sn * by); dz = g.lz + rad * (cs * tz + sn * bz);
    }
    const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
    o = make_float4(f.px + f.nx * off, f.py + f.ny * off, f.pz + f.nz * off, 0.0f);
    d = make_float4(dx * inv, dy * inv, dz * inv, g.tmax);
    return true;
}

// Pass 1: the partition key of every potential output ray (8 = none).
__global__ void gen_keys_kernel(GenParams g, const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits, const float4* __restrict__ tri48,
                                const cndl_entity* __restrict__ ents, unsigned R, unsigned char* __restrict__ keys) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const HitFrame f = hit_frame(rays, hits, tri48, ents, i);
    for (int s = 0; s < g.spp; ++s) {
        unsigned key = 8;
        float4 o, d;
        if (f.valid && gen_ray(g, f, i, s, o, d)) key = g.bucket ? ((d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u)) : 0u;
        keys[(size_t)i * g.spp + s] = (unsigned char)key;
    }
}

// Pass 3: the rays, written where the partition put them.
__global__ void gen_emit_kernel(GenParams g, const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits, const float4* __restrict__ tri48,
                                const cndl_entity* __restrict__ ents, unsigned R, const unsigned char* __restrict__ keys,
                                const unsigned* __restrict__ dest, cndl_ray* __restrict__ out, unsigned* __restrict__ parent) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const HitFrame f = hit_frame(rays, hits, tri48, ents, i);
    if (!f.valid) return;
    for (int s = 0; s < g.spp; ++s) {
        const size_t e = (size_t)i * g.spp + s;
        if (keys[e] >= 8) continue;
        float4 o, d;
        gen_ray(g, f, i, s, o, d);
        const unsigned pos = dest[e];
        float4* q = reinterpret_cast<float4*>(out + pos);
        q[0] = o;
        q[1] = d;
        if (parent) parent[pos] = i;
    }
}

// GetData (…/Include/TraverseBVHStackless.glsl:370-408) without the texture fetch.
__global__ void get_data_kernel(const int4* __restrict__ tris, const float4* __restrict__ verts, const cndl_entity* __restrict__ ents,
                                const cndl_hit* __restrict__ hits, unsigned R, cndl_hit_attr* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(hits + i));
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(hits + i) + 1);
    float4 o0 = make_float4(-1.0f, -1.0f, -1.0f, 0.0f);  // Normal = vec3(-1) on a miss (SL:377-382)
    float4 o1 = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(h1.x));
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
    }
    float4* p = reinterpret_cast<float4*>(out + i);
    p[0] = o0;
    p[1] = o1;
}


}  // namespace

void launch_get_data(const SceneView& s, const float4* verts, const cndl_hit* hits, size_t R, cndl_hit_attr* out, cudaStream_t stream, LaunchCounter& lc) {
    if (R == 0) return;
    get_data_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(s.tris, verts, s.ents, hits, (unsigned)R, out);
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

// Returns the number of rays written through *h_count (synchronises `stream`).
cudaError_t generate_rays(const SceneView& s, const cndl_raygen_params& prm, const cndl_ray* rays, const cndl_hit* hits, size_t R, cndl_ray* out,
                          unsigned* parent, int* scratch, size_t* h_count, cudaStream_t stream, LaunchCounter& lc) {
    *h_count = 0;
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
    gen_keys_kernel<<<grid, 256, 0, stream>>>(g, rays, hits, s.tri48, s.ents, (unsigned)R, keys);
    lc.n++;
    partition8(KeyFromBytes{keys}, n, dest, 0, p, stream, lc);
    gen_emit_kernel<<<grid, 256, 0, stream>>>(g, rays, hits, s.tri48, s.ents, (unsigned)R, keys, dest, out, parent);
    lc.n++;
    int h_total = 0;
    cudaError_t e = cudaMemcpyAsync(&h_total, p.total, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    *h_count = (size_t)h_total;
    return cudaGetLastError();
}

}  // namespace cndl
