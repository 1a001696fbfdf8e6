// Shared by the two BVH builders (builder.cu: exact binned SAH; builder_lbvh.cu: LBVH): per-triangle boxes, the root box,
// build-node arrays, the scratch arena.  Header-only: each translation unit gets its own copy of the kernels.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "builder.cuh"
#include "scan.cuh"

namespace cndl {
namespace {

constexpr float kSentinelMax = 10000000.0f;   // BVHConstructor.h:20
constexpr float kSentinelMin = -10000000.0f;  // BVHConstructor.h:21
constexpr int kBins = 64;                      // :46
constexpr unsigned kMaxLeaf = 2;               // :50
constexpr float kInfCost = 1e29f;              // :56
constexpr unsigned kBigNode = 512;             // ranges longer than this get a 1024-thread block (2048: +2 % build time)
constexpr unsigned kTinyNode = 64;             // ranges up to this get one warp
constexpr unsigned kSplitNodeDefault = 16384;  // ranges longer than this are split across CTAs (split_* kernels)
constexpr int kSplitBlock = 256;  // a CTA of a split node covers 256 x ITEMS references: 512 up to 2^20 triangles (more CTAs: the passes are
                                   // latency-bound there, 2.39 -> 2.12 ms at 262k), 2048 beyond (fewer merges into the global bins: 22.2 -> 20.9 ms at 10 M)
constexpr int kBinInts = 3 * kBins + 18 * kBins;  // count[3][64], mn[3][3][64], mx[3][3][64] as ordered keys

// glm 0.9.8.5 min/max (func_common.inl:15-28); argument order matters for +0/-0 ties
__device__ __forceinline__ float gmin(float x, float y) { return x < y ? x : y; }
__device__ __forceinline__ float gmax(float x, float y) { return x > y ? x : y; }

// order-preserving float <-> int key (involution), for integer atomics on floats
__device__ __forceinline__ int f2key(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7FFFFFFF; }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

__device__ __forceinline__ float box_area(float mnx, float mny, float mnz, float mxx, float mxy, float mxz) {
    const float ex = fsub(mxx, mnx), ey = fsub(mxy, mny), ez = fsub(mxz, mnz);  // Bounds::GetArea, BVHConstructor.h:41-44
    return fadd(fadd(fmul(ex, ey), fmul(ey, ez)), fmul(ez, ex));
}

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // SplitMix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct BuildArrays {
    // per triangle
    const float4* verts;     // 2 float4 per vertex
    const uint32_t* indices; // 3 per triangle
    const int32_t* mesh_ids; // may be null
    float4* tmin;            // xyz = box min, w = centroid.x
    float4* tmax;            // xyz = box max, w = centroid.y
    float* tcz;              // centroid.z
    int* refs;               // TriangleReferences, partitioned in place
    unsigned T;
    // per build node (ids in level order)
    float4* nmin;            // xyz
    float4* nmax;
    unsigned* nstart;        // build range start (leaf: unchanged; pack derives from it)
    unsigned* nlen;          // build range length
    int* nchild;             // id of the left child (right = +1), -1 for a leaf
    unsigned* nsize;         // subtree size in nodes
    int* npre;               // pre-order index (stackless)
    int* nlink;              // miss link (stackless)
    // outputs
    int4* tris_out;
    int tri_offset;
    // root box scratch: 6 ordered keys + 6 zero tie-break positions
    int* root_scratch;
};

// ---------------------------------------------------------------------------------------------
// per-triangle boxes and centroids (:411-427) + root box
__global__ void tri_precompute_kernel(BuildArrays a) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {kSentinelMax, kSentinelMax, kSentinelMax}, mx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
    const bool valid = t < a.T;
    if (valid) {
        for (int c = 0; c < 3; ++c) {
            const float4 p = a.verts[2 * (size_t)a.indices[3 * (size_t)t + c]];
            const float pv[3] = {p.x, p.y, p.z};
            for (int k = 0; k < 3; ++k) {
                mn[k] = gmin(mn[k], pv[k]);
                mx[k] = gmax(mx[k], pv[k]);
            }
        }
        const float cx = fdiv(fadd(mn[0], mx[0]), 2.0f), cy = fdiv(fadd(mn[1], mx[1]), 2.0f), cz = fdiv(fadd(mn[2], mx[2]), 2.0f);
        a.tmin[t] = make_float4(mn[0], mn[1], mn[2], cx);
        a.tmax[t] = make_float4(mx[0], mx[1], mx[2], cy);
        a.tcz[t] = cz;
        a.refs[t] = (int)t;
    }
    // root box: numeric min/max by ordered-int atomics, reduced per warp and per block first (one set of global atomics per block:
    // at 10 M triangles the per-warp version spent 1.3 ms on twelve contended addresses)
    __shared__ int s_root[12];
    if (threadIdx.x < 12) s_root[threadIdx.x] = threadIdx.x < 3 ? 0x7FFFFFFF : (threadIdx.x < 6 ? (int)0x80000000 : -1);
    __syncthreads();
    for (int k = 0; k < 3; ++k) {
        int kmn = f2key(mn[k] == 0.0f ? 0.0f : mn[k]), kmx = f2key(mx[k] == 0.0f ? 0.0f : mx[k]);
        // MinInitial = glm::min(MinInitial, cur.Min) (:421): ties take the later triangle, so the sign of a
        // zero result is that of the LAST triangle whose component is zero.
        int zmn = valid && mn[k] == 0.0f ? (int)t : -1, zmx = valid && mx[k] == 0.0f ? (int)t : -1;
        kmn = __reduce_min_sync(0xFFFFFFFFu, kmn);
        kmx = __reduce_max_sync(0xFFFFFFFFu, kmx);
        zmn = __reduce_max_sync(0xFFFFFFFFu, zmn);
        zmx = __reduce_max_sync(0xFFFFFFFFu, zmx);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&s_root[k], kmn);
            atomicMax(&s_root[3 + k], kmx);
            if (zmn >= 0) atomicMax(&s_root[6 + k], zmn);
            if (zmx >= 0) atomicMax(&s_root[9 + k], zmx);
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(a.root_scratch + threadIdx.x, s_root[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(a.root_scratch + threadIdx.x, s_root[threadIdx.x]);
    else if (threadIdx.x < 12 && s_root[threadIdx.x] >= 0) atomicMax(a.root_scratch + threadIdx.x, s_root[threadIdx.x]);
}

__global__ void root_finalize_kernel(BuildArrays a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) {
        mn[k] = key2f(a.root_scratch[k]);
        mx[k] = key2f(a.root_scratch[3 + k]);
        if (mn[k] == 0.0f && a.root_scratch[6 + k] >= 0) { const float4 v = a.tmin[a.root_scratch[6 + k]]; mn[k] = k == 0 ? v.x : (k == 1 ? v.y : v.z); }
        if (mx[k] == 0.0f && a.root_scratch[9 + k] >= 0) { const float4 v = a.tmax[a.root_scratch[9 + k]]; mx[k] = k == 0 ? v.x : (k == 1 ? v.y : v.z); }
    }
    a.nmin[0] = make_float4(mn[0], mn[1], mn[2], 0.0f);
    a.nmax[0] = make_float4(mx[0], mx[1], mx[2], 0.0f);
    a.nstart[0] = 0;
    a.nlen[0] = a.T;
    a.nchild[0] = -1;
}

__device__ __forceinline__ float centroid_of(const BuildArrays& a, int r, int axis) {
    return axis == 0 ? a.tmin[r].w : (axis == 1 ? a.tmax[r].w : a.tcz[r]);
}

__device__ __forceinline__ int leaf_pack(const BuildArrays& a, int id) {
    const unsigned s = a.nstart[id], len = a.nlen[id];
    const unsigned at = a.T - (s + len) + (unsigned)a.tri_offset;  // :469
    return (int)((at << 4) | (len & 0xF));                         // :794
}

// a root that is itself a leaf (T <= 2)
__global__ void single_leaf_kernel(BuildArrays a, int stackless, float4* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (unsigned j = 0; j < a.T; ++j)
        a.tris_out[j] = make_int4((int)a.indices[3 * j], (int)a.indices[3 * j + 1], (int)a.indices[3 * j + 2], a.mesh_ids ? a.mesh_ids[j] : 0);
    const float4 mn = a.nmin[0], mx = a.nmax[0];
    const int pack = leaf_pack(a, 0);
    if (stackless) {
        out[0] = make_float4(mn.x, mn.y, mn.z, __int_as_float(pack));
        out[1] = make_float4(mx.x, mx.y, mx.z, __int_as_float(-1));
    } else {
        // the reference dereferences null here; defined as: left = the leaf, right = an empty leaf
        out[0] = make_float4(mn.x, mn.y, mn.z, __int_as_float(pack));
        out[1] = make_float4(mx.x, mx.y, mx.z, 0.0f);
        out[2] = make_float4(kSentinelMax, kSentinelMax, kSentinelMax, __int_as_float(0));
        out[3] = make_float4(kSentinelMin, kSentinelMin, kSentinelMin, 0.0f);
    }
}

// Bump allocator over one device allocation that the context keeps between builds (cudaMalloc / cudaFree
// of two dozen arrays cost far more than the build itself).  First pass measures, second pass assigns.
struct Scratch {
    char* base = nullptr;
    size_t used = 0;
    template <class T>
    void alloc(T** p, size_t count) {
        const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~size_t(255);
        if (base) *p = reinterpret_cast<T*>(base + used);
        used += bytes;
    }
};

}  // namespace

#define BK(call)                                                          \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) {                                         \
            err = std::string(#call) + ": " + cudaGetErrorString(e__);    \
            return e__ == cudaErrorMemoryAllocation ? CNDL_ERR_OOM : CNDL_ERR_CUDA; \
        }                                                                 \
    } while (0)

}  // namespace cndl
