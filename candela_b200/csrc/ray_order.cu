// sort_rays = 2 of cndl_set_traversal_mode: one-pass counting sort of a ray batch (declared in builder.cuh).
#include "builder.cuh"
#include "scan.cuh"

namespace cndl {

// Ray ordering for scenes that do not fit the L2 (sort_rays = 2): rays grouped by direction octant (major) and the
// Morton code of their origin cell (3 bits per axis inside `lo`..`hi`, the world bounds of the scene) — a 12-bit key,
// so ONE counting sort does it: a histogram pass, a 4096-entry scan and a scatter pass, both passes with a block-private
// histogram in shared memory.  Rays that follow each other then walk neighbouring subtrees, so node and triangle records
// fetched from DRAM by one warp are found in L2 by the next.  10 M-triangle scene, 12.5 M random rays: traversal 5.27 ->
// 4.41 ms (5 bits per axis would give 4.32, which a single pass cannot hold).  The order inside a bucket follows the
// arrival of the blocks; a ray's result does not depend on its slot.
namespace {
constexpr int kRoBins = 4096, kRoBlock = 256, kRoItems = 16, kRoTile = kRoBlock * kRoItems;

__device__ __forceinline__ unsigned ray_order_key(const cndl_ray* __restrict__ rays, unsigned i, float3 lo, float3 scale) {
    const float4 o = __ldg(reinterpret_cast<const float4*>(rays + i)), d = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
    const unsigned qx = (unsigned)fminf(fmaxf((o.x - lo.x) * scale.x, 0.0f), 7.0f), qy = (unsigned)fminf(fmaxf((o.y - lo.y) * scale.y, 0.0f), 7.0f),
                   qz = (unsigned)fminf(fmaxf((o.z - lo.z) * scale.z, 0.0f), 7.0f);
    const unsigned octant = (d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u);
    return (octant << 9) | ((expand10(qx) << 2) | (expand10(qy) << 1) | expand10(qz));
}

__global__ void __launch_bounds__(kRoBlock) ray_order_hist_kernel(const cndl_ray* __restrict__ rays, unsigned R, float3 lo, float3 scale,
                                                                 unsigned* __restrict__ hist) {
    __shared__ unsigned h[kRoBins];
    for (int b = threadIdx.x; b < kRoBins; b += kRoBlock) h[b] = 0;
    __syncthreads();
    const unsigned first = blockIdx.x * kRoTile;
#pragma unroll 4
    for (int j = 0; j < kRoItems; ++j) {
        const unsigned i = first + j * kRoBlock + threadIdx.x;
        if (i < R) atomicAdd(&h[ray_order_key(rays, i, lo, scale)], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kRoBins; b += kRoBlock)
        if (h[b]) atomicAdd(&hist[b], h[b]);
}

// exclusive scan of the 4096 bucket sizes, in place (one block of 1024 threads, 4 bins each)
__global__ void __launch_bounds__(1024) ray_order_scan_kernel(unsigned* __restrict__ hist) {
    __shared__ int s_warp[1024 / 32 + 1];
    unsigned v[4];
    int sum = 0;
    for (int k = 0; k < 4; ++k) { v[k] = hist[4 * threadIdx.x + k]; sum += (int)v[k]; }
    int total;
    int ex = block_exclusive_scan<1024>(sum, s_warp, total);
    for (int k = 0; k < 4; ++k) { hist[4 * threadIdx.x + k] = (unsigned)ex; ex += (int)v[k]; }
}

// order[pos] = original index of the ray sorted to pos; sorted (optional): the rays themselves, moved to their sorted places
__global__ void __launch_bounds__(kRoBlock) ray_order_scatter_kernel(const cndl_ray* __restrict__ rays, unsigned R, float3 lo, float3 scale,
                                                                    unsigned* __restrict__ cursor, unsigned* __restrict__ order, cndl_ray* __restrict__ sorted) {
    __shared__ unsigned h[kRoBins];
    for (int b = threadIdx.x; b < kRoBins; b += kRoBlock) h[b] = 0;
    __syncthreads();
    const unsigned first = blockIdx.x * kRoTile;
    unsigned key[kRoItems], rank[kRoItems];
#pragma unroll
    for (int j = 0; j < kRoItems; ++j) {
        const unsigned i = first + j * kRoBlock + threadIdx.x;
        key[j] = kRoBins;
        rank[j] = 0;
        if (i < R) {
            key[j] = ray_order_key(rays, i, lo, scale);
            rank[j] = atomicAdd(&h[key[j]], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kRoBins; b += kRoBlock)
        if (h[b]) h[b] = atomicAdd(&cursor[b], h[b]);  // this block's run inside bucket b
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRoItems; ++j) {
        const unsigned i = first + j * kRoBlock + threadIdx.x;
        if (key[j] < (unsigned)kRoBins) {
            const unsigned pos = h[key[j]] + rank[j];
            order[pos] = i;
            if (sorted) {
                const float4* src = reinterpret_cast<const float4*>(rays + i);
                float4* dst = reinterpret_cast<float4*>(sorted + pos);
                dst[0] = __ldg(src);
                dst[1] = __ldg(src + 1);
            }
        }
    }
}
}  // namespace

size_t ray_sort_scratch_ints(size_t) { return kRoBins + 64; }

cudaError_t sort_rays_morton(const cndl_ray* rays, size_t R, const float lo[3], const float hi[3], unsigned* order_out, cndl_ray* sorted_out, int* scratch,
                             cudaStream_t st, LaunchCounter& lc) {
    if (R == 0) return cudaSuccess;
    unsigned* hist = reinterpret_cast<unsigned*>(scratch);
    float3 l = make_float3(lo[0], lo[1], lo[2]), sc;
    sc.x = hi[0] > lo[0] ? 8.0f / (hi[0] - lo[0]) : 0.0f;
    sc.y = hi[1] > lo[1] ? 8.0f / (hi[1] - lo[1]) : 0.0f;
    sc.z = hi[2] > lo[2] ? 8.0f / (hi[2] - lo[2]) : 0.0f;
    const unsigned blocks = (unsigned)((R + kRoTile - 1) / kRoTile);
    cudaMemsetAsync(hist, 0, kRoBins * sizeof(unsigned), st);
    ray_order_hist_kernel<<<blocks, kRoBlock, 0, st>>>(rays, (unsigned)R, l, sc, hist);
    ray_order_scan_kernel<<<1, 1024, 0, st>>>(hist);
    ray_order_scatter_kernel<<<blocks, kRoBlock, 0, st>>>(rays, (unsigned)R, l, sc, hist, order_out, sorted_out);
    lc.n += 3;
    return cudaGetLastError();
}

}  // namespace cndl
