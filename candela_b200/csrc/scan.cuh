// Block-wide and device-wide exclusive scans of 32-bit integers (reduce-then-scan, three launches).
// Header-only: each translation unit that includes it gets its own copy of the kernels.
#pragma once
#include "kernels.cuh"

namespace cndl {
namespace {

// ---------------------------------------------------------------------------------------------
// block-wide helpers
template <int BLOCK>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums /* BLOCK/32 + 1 */, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < BLOCK / 32 ? warp_sums[lane] : 0;
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += n;
        }
        __syncwarp();  // every lane's read of warp_sums above is ordered before the writes below
        if (lane < BLOCK / 32) warp_sums[lane] = winc - w;
        if (lane == 31) warp_sums[BLOCK / 32] = winc;
    }
    __syncthreads();
    total = warp_sums[BLOCK / 32];
    const int r = warp_sums[warp] + inc - v;
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of 32-bit values (reduce-then-scan, three launches)
constexpr int kScanBlock = 512, kScanItems = 4, kScanTile = kScanBlock * kScanItems;

__global__ void scan_reduce_kernel(const int* in, int n, int* block_sums) {
    __shared__ int s_warp[kScanBlock / 32 + 1];
    const int base = blockIdx.x * kScanTile;
    int v = 0;
    for (int j = 0; j < kScanItems; ++j) {
        const int i = base + j * kScanBlock + threadIdx.x;
        if (i < n) v += in[i];
    }
    int total;
    block_exclusive_scan<kScanBlock>(v, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void scan_spine_kernel(int* block_sums, int n_blocks, int* total_out) {
    // single block, sequential over tiles of kScanBlock
    __shared__ int s_warp[kScanBlock / 32 + 1];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += kScanBlock) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan<kScanBlock>(v, s_warp, total);
        if (i < n_blocks) block_sums[i] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

__global__ void scan_apply_kernel(const int* in, int n, const int* block_sums, int* out) {
    __shared__ int s_warp[kScanBlock / 32 + 1];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems], sum = 0;
    for (int j = 0; j < kScanItems; ++j) { v[j] = base + j < n ? in[base + j] : 0; sum += v[j]; }
    int total;
    int ex = block_exclusive_scan<kScanBlock>(sum, s_warp, total) + block_sums[blockIdx.x];
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < n) out[base + j] = ex;
        ex += v[j];
    }
}

void exclusive_scan(const int* d_in, int n, int* d_out, int* d_block_sums, int* d_total, cudaStream_t st, LaunchCounter& lc) {
    const int blocks = (n + kScanTile - 1) / kScanTile;
    scan_reduce_kernel<<<blocks, kScanBlock, 0, st>>>(d_in, n, d_block_sums);
    scan_spine_kernel<<<1, kScanBlock, 0, st>>>(d_block_sums, blocks, d_total);
    scan_apply_kernel<<<blocks, kScanBlock, 0, st>>>(d_in, n, d_block_sums, d_out);
    lc.n += 3;
}

}  // namespace
}  // namespace cndl
