// LBVH builder (CNDL_BUILDER_LBVH) for cndl_add_object: Morton codes -> LSD radix sort -> Karras radix tree over leaves of
// <= 2 triangles -> bottom-up boxes -> the reference's two node layouts.  Same buffers, different tree:
// parity level P2 — traversal returns the same hit triangle, not the same buffer index.
#include "builder_common.cuh"

namespace cndl {

namespace {

constexpr int kRsBlock = 256, kRsItems = 16, kRsTile = kRsBlock * kRsItems, kRsBins = 256;

__global__ void morton_kernel(BuildArrays a, unsigned* keys, unsigned* vals) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.T) return;
    const float4 rmn = a.nmin[0], rmx = a.nmax[0];
    const float c[3] = {a.tmin[t].w, a.tmax[t].w, a.tcz[t]};
    const float lo[3] = {rmn.x, rmn.y, rmn.z}, hi[3] = {rmx.x, rmx.y, rmx.z};
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        const float e = hi[k] - lo[k];
        float u = e > 0.0f ? (c[k] - lo[k]) / e : 0.0f;
        u = fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f);
        q[k] = (unsigned)u;
    }
    keys[t] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[t] = t;
}

// counts[bin * n_tiles + tile]
__global__ void __launch_bounds__(kRsBlock) radix_hist_kernel(const unsigned* __restrict__ keys, unsigned n, int shift, int n_tiles, int* __restrict__ counts) {
    __shared__ int h[kRsBins];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * kRsTile;
    for (int j = 0; j < kRsItems; ++j) {
        const unsigned i = base + j * kRsBlock + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFFu], 1);
    }
    __syncthreads();
    counts[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one tile.  Each warp owns a contiguous run of 512 items (16 rounds of 32) so that tile
// order == (warp, round, lane) order; ranks come from per-warp digit counts + __match_any_sync.
__global__ void __launch_bounds__(kRsBlock) radix_scatter_kernel(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals, unsigned n,
                                                                 int shift, int n_tiles, const int* __restrict__ offsets,
                                                                 unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out) {
    __shared__ int wcount[kRsBlock / 32][kRsBins];  // per warp: digit counts, then running offsets
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = threadIdx.x; b < (kRsBlock / 32) * kRsBins; b += kRsBlock) (&wcount[0][0])[b] = 0;
    __syncthreads();
    const unsigned wbase = blockIdx.x * kRsTile + warp * (kRsItems * 32);
    unsigned k[kRsItems], v[kRsItems];
    for (int j = 0; j < kRsItems; ++j) {
        const unsigned i = wbase + j * 32 + lane;
        const bool valid = i < n;
        k[j] = valid ? keys[i] : 0xFFFFFFFFu;
        v[j] = valid ? vals[i] : 0u;
        const unsigned vm = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const unsigned d = (k[j] >> shift) & 0xFFu;
            const unsigned peers = __match_any_sync(vm, d);
            if (lane == __ffs(peers) - 1) wcount[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over the warps of each digit, plus the tile's global offset for that digit
        const int d = threadIdx.x;
        int run = offsets[d * n_tiles + blockIdx.x];
        for (int w = 0; w < kRsBlock / 32; ++w) { const int c = wcount[w][d]; wcount[w][d] = run; run += c; }
    }
    __syncthreads();
    for (int j = 0; j < kRsItems; ++j) {
        const unsigned i = wbase + j * 32 + lane;
        const bool valid = i < n;
        const unsigned vm = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const unsigned d = (k[j] >> shift) & 0xFFu;
            const unsigned peers = __match_any_sync(vm, d);
            const int pos = wcount[warp][d] + __popc(peers & ((1u << lane) - 1u));
            keys_out[pos] = k[j];
            vals_out[pos] = v[j];
            __syncwarp(peers);
            if (lane == __ffs(peers) - 1) wcount[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
}

struct LbvhArrays {
    const unsigned* keys;  // sorted Morton codes, one per triangle
    const unsigned* order; // sorted position -> original triangle
    unsigned L;            // leaves (clusters of <= 2 consecutive sorted triangles)
    int* parent;           // [2L-1]; ids: internal i in [0, L-1), leaf j at L-1+j
    int* left;             // [L-1] child ids
    int* right;
    int* first;            // [L-1] range of leaves
    int* last;
    float4* bmin;          // [2L-1]
    float4* bmax;
    int* visits;           // [L-1]
};

__device__ __forceinline__ int lbvh_delta(const LbvhArrays& b, int i, int j) {
    if (j < 0 || j >= (int)b.L) return -1;
    const unsigned ki = b.keys[2 * (size_t)i], kj = b.keys[2 * (size_t)j];
    return ki == kj ? 32 + __clz((unsigned)i ^ (unsigned)j) : __clz(ki ^ kj);
}

// Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees, and k-d trees", algorithm of fig. 4
__global__ void lbvh_tree_kernel(LbvhArrays b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)b.L - 1) return;
    const int d = lbvh_delta(b, i, i + 1) - lbvh_delta(b, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(b, i, i - d);
    int lmax = 2;
    while (lbvh_delta(b, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(b, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(b, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(b, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = lo == gamma ? (int)b.L - 1 + gamma : gamma;
    const int rc = hi == gamma + 1 ? (int)b.L - 1 + gamma + 1 : gamma + 1;
    b.left[i] = lc;
    b.right[i] = rc;
    b.first[i] = lo;
    b.last[i] = hi;
    b.parent[lc] = i;
    b.parent[rc] = i;
    if (i == 0) b.parent[0] = -1;
}

__device__ __forceinline__ float canon0(float v) { return v == 0.0f ? 0.0f : v; }  // -0 -> +0: order-independent bits

__global__ void lbvh_refit_kernel(LbvhArrays b, BuildArrays a) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= b.L) return;
    float mn[3] = {kSentinelMax, kSentinelMax, kSentinelMax}, mx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
    for (unsigned p = 2 * j; p < min(2 * j + 2, a.T); ++p) {
        const unsigned r = b.order[p];
        const float4 tm = a.tmin[r], tx = a.tmax[r];
        mn[0] = fminf(mn[0], tm.x); mn[1] = fminf(mn[1], tm.y); mn[2] = fminf(mn[2], tm.z);
        mx[0] = fmaxf(mx[0], tx.x); mx[1] = fmaxf(mx[1], tx.y); mx[2] = fmaxf(mx[2], tx.z);
        a.tris_out[p] = make_int4((int)a.indices[3 * (size_t)r], (int)a.indices[3 * (size_t)r + 1], (int)a.indices[3 * (size_t)r + 2], a.mesh_ids ? a.mesh_ids[r] : 0);
    }
    int id = (int)b.L - 1 + (int)j;
    b.bmin[id] = make_float4(canon0(mn[0]), canon0(mn[1]), canon0(mn[2]), 0.0f);
    b.bmax[id] = make_float4(canon0(mx[0]), canon0(mx[1]), canon0(mx[2]), 0.0f);
    int p = b.parent[id];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(&b.visits[p], 1) == 0) return;  // the second child to arrive carries on
        const float4 lmn = b.bmin[b.left[p]], lmx = b.bmax[b.left[p]], rmn = b.bmin[b.right[p]], rmx = b.bmax[b.right[p]];
        b.bmin[p] = make_float4(fminf(lmn.x, rmn.x), fminf(lmn.y, rmn.y), fminf(lmn.z, rmn.z), 0.0f);
        b.bmax[p] = make_float4(fmaxf(lmx.x, rmx.x), fmaxf(lmx.y, rmx.y), fmaxf(lmx.z, rmx.z), 0.0f);
        p = b.parent[p];
    }
}

// Pre-order index of a node = ancestors + nodes of the subtrees hanging to the left of the root path
//                           = depth + 2*first_leaf - right_turns   (k subtrees holding l leaves have 2l-k nodes)
__global__ void lbvh_number_kernel(LbvhArrays b, int* depth_out, int* pre_out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 2 * (int)b.L - 1;
    if (id >= n) return;
    int depth = 0, right_turns = 0, x = id, p = b.parent[id];
    while (p >= 0) {
        ++depth;
        if (b.right[p] == x) ++right_turns;
        x = p;
        p = b.parent[p];
    }
    const int first = id < (int)b.L - 1 ? b.first[id] : id - ((int)b.L - 1);
    depth_out[id] = depth;
    pre_out[id] = depth + 2 * first - right_turns;
}

__global__ void lbvh_emit_stackless_kernel(LbvhArrays b, BuildArrays a, const int* pre, float4* out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 2 * (int)b.L - 1;
    if (id >= n) return;
    const bool inner = id < (int)b.L - 1;
    const int leaves = inner ? b.last[id] - b.first[id] + 1 : 1;
    const int p = pre[id];
    const int next = p + 2 * leaves - 1;  // first node after this subtree in pre-order == the miss link
    int minw = -1;
    if (!inner) {
        const unsigned j = (unsigned)(id - ((int)b.L - 1));
        const unsigned len = min(2u, a.T - 2 * j);
        minw = (int)(((2 * j + (unsigned)a.tri_offset) << 4) | len);
    }
    const float4 mn = b.bmin[id], mx = b.bmax[id];
    out[2 * (size_t)p] = make_float4(mn.x, mn.y, mn.z, __int_as_float(minw));
    out[2 * (size_t)p + 1] = make_float4(mx.x, mx.y, mx.z, __int_as_float(next >= n ? -1 : next));
}

__global__ void lbvh_depth_keys_kernel(const int* depth, unsigned n_inner, unsigned* keys, unsigned* vals) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner) return;
    keys[i] = (unsigned)depth[i];
    vals[i] = i;
}

__global__ void lbvh_slot_kernel(const unsigned* sorted_nodes, unsigned n_inner, int* slot) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inner) slot[sorted_nodes[i]] = (int)i;
}

__global__ void lbvh_emit_stack_kernel(LbvhArrays b, BuildArrays a, const int* slot, float4* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)b.L - 1) return;
    float4* o = out + 4 * (size_t)slot[i];
    for (int side = 0; side < 2; ++side) {
        const int ch = side == 0 ? b.left[i] : b.right[i];
        const bool leaf = ch >= (int)b.L - 1;
        int minw = -1;
        if (leaf) {
            const unsigned j = (unsigned)(ch - ((int)b.L - 1));
            minw = (int)(((2 * j + (unsigned)a.tri_offset) << 4) | min(2u, a.T - 2 * j));
        }
        const float4 mn = b.bmin[ch], mx = b.bmax[ch];
        o[2 * side] = make_float4(mn.x, mn.y, mn.z, __int_as_float(minw));
        o[2 * side + 1] = make_float4(mx.x, mx.y, mx.z, leaf ? 0.0f : __int_as_float(slot[ch]));
    }
}

// LSD radix sort of (key, value) pairs; `bits` key bits are significant. Result ends in (*keys, *vals).
cudaError_t radix_sort_pairs(unsigned** keys, unsigned** vals, unsigned** keys_alt, unsigned** vals_alt, unsigned n, int bits, int* counts, int* offsets,
                             int* block_sums, int* total, cudaStream_t st, LaunchCounter& lc) {
    const int n_tiles = (int)((n + kRsTile - 1) / kRsTile);
    for (int shift = 0; shift < bits; shift += 8) {
        radix_hist_kernel<<<n_tiles, kRsBlock, 0, st>>>(*keys, n, shift, n_tiles, counts);
        lc.n++;
        exclusive_scan(counts, kRsBins * n_tiles, offsets, block_sums, total, st, lc);
        radix_scatter_kernel<<<n_tiles, kRsBlock, 0, st>>>(*keys, *vals, n, shift, n_tiles, offsets, *keys_alt, *vals_alt);
        lc.n++;
        std::swap(*keys, *keys_alt);
        std::swap(*vals, *vals_alt);
    }
    return cudaGetLastError();
}

}  // namespace

int build_lbvh_object(BuildRequest& rq, cudaStream_t st, LaunchCounter& lc, float* build_ms, std::string& err) {
    const size_t T = rq.T;
    if (T == 0 || T > (1ull << 27)) { err = "triangle count out of range"; return CNDL_ERR_INVALID; }
    for (size_t i = 0; i < 3 * T; ++i)
        if (rq.h_indices[i] >= rq.V) { err = "vertex index out of range"; return CNDL_ERR_INVALID; }
    const bool stackless = rq.format == CNDL_STACKLESS;
    const size_t L = (T + 1) / 2, N = 2 * L - 1;
    const int n_tiles = (int)((std::max(T, N) + kRsTile - 1) / kRsTile);

    BuildArrays a{};
    LbvhArrays b{};
    uint32_t* d_idx = nullptr;
    int32_t* d_mesh = nullptr;
    unsigned *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr;
    int *d_counts = nullptr, *d_offsets = nullptr, *d_block_sums = nullptr, *d_total = nullptr, *d_depth = nullptr, *d_pre = nullptr, *d_slot = nullptr;
    const size_t scan_n = (size_t)kRsBins * n_tiles + 16;
    auto layout = [&](Scratch& sc) {
        sc.alloc(&d_idx, 3 * T);
        if (rq.h_mesh_ids) sc.alloc(&d_mesh, T);
        sc.alloc(&a.tmin, T); sc.alloc(&a.tmax, T); sc.alloc(&a.tcz, T); sc.alloc(&a.refs, T);
        sc.alloc(&a.nmin, 1); sc.alloc(&a.nmax, 1); sc.alloc(&a.nstart, 1); sc.alloc(&a.nlen, 1); sc.alloc(&a.nchild, 1);
        sc.alloc(&a.root_scratch, 16);
        sc.alloc(&k0, std::max(T, N)); sc.alloc(&k1, std::max(T, N)); sc.alloc(&v0, std::max(T, N)); sc.alloc(&v1, std::max(T, N));
        sc.alloc(&d_counts, scan_n); sc.alloc(&d_offsets, scan_n); sc.alloc(&d_block_sums, scan_n / kScanTile + 2); sc.alloc(&d_total, 4);
        sc.alloc(&b.parent, N); sc.alloc(&b.left, L); sc.alloc(&b.right, L); sc.alloc(&b.first, L); sc.alloc(&b.last, L);
        sc.alloc(&b.bmin, N); sc.alloc(&b.bmax, N); sc.alloc(&b.visits, L);
        sc.alloc(&d_depth, N); sc.alloc(&d_pre, N); sc.alloc(&d_slot, N);
    };
    Scratch measure;
    layout(measure);
    if (*rq.arena_cap < measure.used) {
        if (*rq.arena) cudaFree(*rq.arena);
        *rq.arena = nullptr;
        *rq.arena_cap = 0;
        BK(cudaMalloc(rq.arena, measure.used));
        *rq.arena_cap = measure.used;
    }
    Scratch sc;
    sc.base = static_cast<char*>(*rq.arena);
    layout(sc);

    cudaEvent_t ev0, ev1;
    BK(cudaEventCreate(&ev0));
    BK(cudaEventCreate(&ev1));
    BK(cudaMemcpyAsync(d_idx, rq.h_indices, 3 * T * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (rq.h_mesh_ids) BK(cudaMemcpyAsync(d_mesh, rq.h_mesh_ids, T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    BK(cudaEventRecord(ev0, st));
    a.verts = rq.d_verts;
    a.indices = d_idx;
    a.mesh_ids = d_mesh;
    a.T = (unsigned)T;
    a.tris_out = rq.d_tris_out;
    a.tri_offset = rq.tri_offset;
    const int h_root_init[12] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000, -1, -1, -1, -1, -1, -1};
    BK(cudaMemcpyAsync(a.root_scratch, h_root_init, sizeof(h_root_init), cudaMemcpyHostToDevice, st));
    const unsigned gT = (unsigned)((T + 255) / 256);
    tri_precompute_kernel<<<gT, 256, 0, st>>>(a);
    root_finalize_kernel<<<1, 32, 0, st>>>(a);
    morton_kernel<<<gT, 256, 0, st>>>(a, k0, v0);
    lc.n += 3;
    BK(radix_sort_pairs(&k0, &v0, &k1, &v1, (unsigned)T, 30, d_counts, d_offsets, d_block_sums, d_total, st, lc));

    float4* out = static_cast<float4*>(rq.d_nodes_out);
    b.keys = k0;
    b.order = v0;
    b.L = (unsigned)L;
    const unsigned gL = (unsigned)((L + 255) / 256), gN = (unsigned)((N + 255) / 256);
    BK(cudaMemsetAsync(b.visits, 0, L * sizeof(int), st));
    BK(cudaMemsetAsync(b.parent, 0xFF, N * sizeof(int), st));
    if (L > 1) { lbvh_tree_kernel<<<gL, 256, 0, st>>>(b); lc.n++; }
    lbvh_refit_kernel<<<gL, 256, 0, st>>>(b, a);
    lbvh_number_kernel<<<gN, 256, 0, st>>>(b, d_depth, d_pre);
    lc.n += 2;
    if (stackless || L == 1) {
        if (!stackless) {
            // a root that is itself a leaf: one slot, left = the leaf, right = an empty leaf (as in the SAH path)
            BK(cudaMemsetAsync(out, 0, sizeof(cndl_stack_node), st));
            a.nmin = b.bmin; a.nmax = b.bmax;  // single_leaf_kernel reads node 0's box
            unsigned zero_start = 0, len = (unsigned)T;
            BK(cudaMemcpyAsync(a.nstart, &zero_start, 4, cudaMemcpyHostToDevice, st));
            BK(cudaMemcpyAsync(a.nlen, &len, 4, cudaMemcpyHostToDevice, st));
            single_leaf_kernel<<<1, 32, 0, st>>>(a, 0, out);
        } else {
            lbvh_emit_stackless_kernel<<<gN, 256, 0, st>>>(b, a, d_pre, out);
        }
        lc.n++;
    } else {
        // FlattenStackBVH order: inner nodes by (depth, left-to-right) == stable sort of node ids by depth
        BK(cudaMemsetAsync(out, 0, N * sizeof(cndl_stack_node), st));
        // the sorted Morton keys / order are dead after the refit: their buffers serve as the sort's alternates
        unsigned *sk = k1, *sv = v1, *sk2 = k0, *sv2 = v0;
        lbvh_depth_keys_kernel<<<gL, 256, 0, st>>>(d_depth, (unsigned)(L - 1), sk, sv);
        lc.n++;
        BK(radix_sort_pairs(&sk, &sv, &sk2, &sv2, (unsigned)(L - 1), 16, d_counts, d_offsets, d_block_sums, d_total, st, lc));
        lbvh_slot_kernel<<<gL, 256, 0, st>>>(sv, (unsigned)(L - 1), d_slot);
        lbvh_emit_stack_kernel<<<gL, 256, 0, st>>>(b, a, d_slot, out);
        lc.n += 2;
    }
    BK(cudaGetLastError());
    BK(cudaEventRecord(ev1, st));
    BK(cudaStreamSynchronize(st));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    if (build_ms) *build_ms = ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    rq.n_nodes_out = N;
    return CNDL_OK;
}

// ---------------------------------------------------------------------------------------------
}  // namespace cndl
