// Wavefront ray generation between bounces: from the hit records of one batch, emit the next batch of
// diffuse rays, compacted so that rays which missed produce nothing (DiffuseTrace.glsl:445-446 for the
// first bounce, :516-517 for later ones; CosWeightedHemisphere of Shaders/Include/Sampling.glsl:1-12).
// The shader's fract(sin()) hash is replaced by a counter-based generator: the rays
// this kernel writes are INPUTS to traversal, so they carry no parity requirement of their own; the
// compaction is a scan, so their order is deterministic.
#include "kernels.cuh"
#include "scan.cuh"

#include <cuda_fp16.h>

namespace cndl {

namespace {

__device__ __forceinline__ unsigned pcg_hash(unsigned v) {
    unsigned s = v * 747796405u + 2891336453u;
    unsigned w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
    return (w >> 22u) ^ w;
}
__device__ __forceinline__ float u01(unsigned h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

__global__ void bounce_count_kernel(const cndl_hit* __restrict__ hits, unsigned R, int spp, int* __restrict__ counts) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R) counts[i] = __ldg(&hits[i].t) > 0.0f ? spp : 0;
}

__global__ void bounce_emit_kernel(const cndl_ray* __restrict__ rays, const cndl_hit* __restrict__ hits, const float4* __restrict__ tri48,
                                   const cndl_entity* __restrict__ ents, const int* __restrict__ offsets, unsigned R, int spp, float offset,
                                   float tmax, unsigned seed, cndl_ray* __restrict__ out, unsigned* __restrict__ parent) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(hits + i));
    if (!(h0.x > 0.0f)) return;
    const int4 h1 = __ldg(reinterpret_cast<const int4*>(hits + i) + 1);
    const float4 ra = __ldg(reinterpret_cast<const float4*>(rays + i)), rb = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
    const float t = h0.x;
    const float px = ra.x + rb.x * t, py = ra.y + rb.y * t, pz = ra.z + rb.z * t;
    // geometric normal of the hit triangle, taken to world space with the entity's model matrix and turned
    // against the incoming ray
    const float4 c = __ldg(tri48 + 3 * (size_t)h1.y + 2);
    const float* m = ents[h1.z].model;
    float nx = m[0] * c.y + m[4] * c.z + m[8] * c.w, ny = m[1] * c.y + m[5] * c.z + m[9] * c.w, nz = m[2] * c.y + m[6] * c.z + m[10] * c.w;
    float inv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
    nx *= inv; ny *= inv; nz *= inv;
    if (nx * rb.x + ny * rb.y + nz * rb.z > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
    // uu = normalize(cross(n, (0,1,1))), vv = cross(uu, n)
    float ux = ny - nz, uy = -nx, uz = nx;
    inv = rsqrtf(fmaxf(ux * ux + uy * uy + uz * uz, 1e-30f));
    ux *= inv; uy *= inv; uz *= inv;
    const float vx = uy * nz - uz * ny, vy = uz * nx - ux * nz, vz = ux * ny - uy * nx;
    const int base = offsets[i];
    for (int s = 0; s < spp; ++s) {
        const unsigned k = pcg_hash(seed ^ pcg_hash(i * (unsigned)spp + (unsigned)s));
        const float r1 = u01(k), r2 = u01(pcg_hash(k + 0x9E3779B9u));
        const float rad = sqrtf(r2), ang = 6.28318530718f * r1;
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float rx = rad * cs, ry = rad * sn, rz = sqrtf(1.0f - r2);
        float dx = rx * ux + ry * vx + rz * nx, dy = rx * uy + ry * vy + rz * ny, dz = rx * uz + ry * vz + rz * nz;
        inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
        float4* o = reinterpret_cast<float4*>(out + base + s);
        o[0] = make_float4(px + nx * offset, py + ny * offset, pz + nz * offset, 0.0f);
        o[1] = make_float4(dx * inv, dy * inv, dz * inv, tmax);
        if (parent) parent[base + s] = i;
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

// Returns the number of rays written through *h_count (synchronises `stream`). `scratch` must hold
// 2*R + R/2048 + 8 ints.
cudaError_t generate_bounce_rays(const SceneView& s, const cndl_ray* rays, const cndl_hit* hits, size_t R, int spp, float offset, float tmax,
                                 unsigned seed, cndl_ray* out, unsigned* parent, int* scratch, size_t* h_count, cudaStream_t stream,
                                 LaunchCounter& lc) {
    *h_count = 0;
    if (R == 0) return cudaSuccess;
    int* counts = scratch;
    int* offsets = scratch + R;
    int* block_sums = scratch + 2 * R;
    int* total = block_sums + (R / kScanTile + 2);
    const unsigned grid = (unsigned)((R + 255) / 256);
    bounce_count_kernel<<<grid, 256, 0, stream>>>(hits, (unsigned)R, spp, counts);
    lc.n++;
    exclusive_scan(counts, (int)R, offsets, block_sums, total, stream, lc);
    bounce_emit_kernel<<<grid, 256, 0, stream>>>(rays, hits, s.tri48, s.ents, offsets, (unsigned)R, spp, offset, tmax, seed, out, parent);
    lc.n++;
    int h_total = 0;
    cudaError_t e = cudaMemcpyAsync(&h_total, total, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    *h_count = (size_t)h_total;
    return cudaGetLastError();
}

}  // namespace cndl
