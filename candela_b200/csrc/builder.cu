// GPU binned-SAH BVH builder whose output is byte-identical to the reference's CPU builder
// (Source/Core/BVH/BVHConstructor.cpp; line numbers below refer to it).
//
// The reference builds depth-first on one thread.  Its result, however, is a pure function of the
// input that can be evaluated level by level:
//   * bin counts / bin boxes, child boxes: min/max/+ reductions, order-independent (the sign of a
//     zero is the only order-dependent bit; see zero-sign notes below);
//   * the split search (:276-363): strict `<` over (axis, bin) in order == first minimum;
//   * the Lomuto partition (:532-549) fixes the order of references inside each range; it is
//     reproduced exactly by a chunk-parallel emulation (partition_range below);
//   * leaves are emitted right-range-first (:624-625), so a leaf with build range [s, s+len) lands
//     at sorted position T-(s+len);
//   * FlattenBVH (:783-845) numbers nodes in pre-order, FlattenStackBVH (:847-930) numbers inner
//     nodes in breadth-first order: both follow from subtree sizes / level order.
// Every float operation uses the non-contracting wrappers of exact_math.cuh.

#include "builder_common.cuh"

namespace cndl {

namespace {

// Split search over the 64 bins of one axis (:319-358), evaluated by one warp.  Bin boxes and counts
// combine with min/max/+, so prefix (from the left) and suffix (from the right) scans give exactly the
// values the reference accumulates sequentially (the sign of a zero aside, which areas cannot see);
// the costs use the reference's expression; strict `<` over ascending splits == first minimum.
struct BinAgg { int n; float mn[3], mx[3]; };
__device__ __forceinline__ BinAgg agg_identity() { return {0, {kSentinelMax, kSentinelMax, kSentinelMax}, {kSentinelMin, kSentinelMin, kSentinelMin}}; }
__device__ __forceinline__ BinAgg agg_join(const BinAgg& a, const BinAgg& b) {
    BinAgg r;
    r.n = a.n + b.n;
    for (int c = 0; c < 3; ++c) { r.mn[c] = fminf(a.mn[c], b.mn[c]); r.mx[c] = fmaxf(a.mx[c], b.mx[c]); }
    return r;
}
__device__ __forceinline__ BinAgg agg_shfl(const BinAgg& a, int src_lane) {
    BinAgg r;
    r.n = __shfl_sync(0xFFFFFFFFu, a.n, src_lane);
    for (int c = 0; c < 3; ++c) { r.mn[c] = __shfl_sync(0xFFFFFFFFu, a.mn[c], src_lane); r.mx[c] = __shfl_sync(0xFFFFFFFFu, a.mx[c], src_lane); }
    return r;
}
__device__ __forceinline__ float agg_cost(const BinAgg& l, const BinAgg& r) {
    const float la = box_area(l.mn[0], l.mn[1], l.mn[2], l.mx[0], l.mx[1], l.mx[2]);
    const float ra = box_area(r.mn[0], r.mn[1], r.mn[2], r.mx[0], r.mx[1], r.mx[2]);
    return fadd(fmul(__int2float_rn(l.n), la), fmul(__int2float_rn(r.n), ra));  // :351
}

__device__ __forceinline__ void warp_sah_eval(const int* s_count, const int (*s_mn)[kBins], const int (*s_mx)[kBins], int axis, float lo,
                                              float extent, float* s_best_cost, int* s_axis, float* s_border) {
    const int lane = threadIdx.x & 31;
    BinAgg b0, b1;
    b0.n = s_count[2 * lane];
    b1.n = s_count[2 * lane + 1];
    for (int c = 0; c < 3; ++c) {
        b0.mn[c] = key2f(s_mn[c][2 * lane]); b0.mx[c] = key2f(s_mx[c][2 * lane]);
        b1.mn[c] = key2f(s_mn[c][2 * lane + 1]); b1.mx[c] = key2f(s_mx[c][2 * lane + 1]);
    }
    const BinAgg pair = agg_join(b0, b1);
    BinAgg incl = pair, incr = pair;  // inclusive scans from the left / from the right
    for (int o = 1; o < 32; o <<= 1) {
        const BinAgg up = agg_shfl(incl, lane - o < 0 ? lane : lane - o);
        if (lane >= o) incl = agg_join(up, incl);
        const BinAgg dn = agg_shfl(incr, lane + o > 31 ? lane : lane + o);
        if (lane + o <= 31) incr = agg_join(incr, dn);
    }
    BinAgg excl = agg_shfl(incl, lane == 0 ? 0 : lane - 1);
    if (lane == 0) excl = agg_identity();
    BinAgg excr = agg_shfl(incr, lane == 31 ? 31 : lane + 1);  // bins 2(lane+1) .. 63
    if (lane == 31) excr = agg_identity();
    // split i = 2*lane: left = bins 0..2lane, right = bins 2lane+1..63; split i = 2*lane+1: left = 0..2lane+1, right = 2lane+2..63
    const float c0 = agg_cost(agg_join(excl, b0), agg_join(b1, excr));
    float c1 = agg_cost(incl, excr);
    float cost = c0;
    int idx = 2 * lane;
    if (lane < 31 && c1 < cost) { cost = c1; idx = 2 * lane + 1; }  // i = 63 is not a split
    if (!(cost == cost)) cost = __int_as_float(0x7F800000);        // a NaN cost is never selected by `<`
    for (int o = 16; o > 0; o >>= 1) {
        const float pc = __shfl_xor_sync(0xFFFFFFFFu, cost, o);
        const int pi = __shfl_xor_sync(0xFFFFFFFFu, idx, o);
        if (pc < cost || (pc == cost && pi < idx)) { cost = pc; idx = pi; }
    }
    const float best_so_far = *s_best_cost;  // every lane reads (the compiler hoists the load anyway); order it before lane 0's write
    __syncwarp();
    if (lane == 0 && cost < best_so_far) {
        *s_best_cost = cost;
        *s_axis = axis;
        *s_border = fadd(lo, fmul(fdiv(extent, (float)kBins), __int2float_rn(idx + 1)));  // :347,:356
    }
}

// One level of the tree as the DEVICE knows it: how many of its active nodes fall into each size class and where its nodes
// sit in the node arrays.  The host launches a level with grid UPPER BOUNDS and never waits for these numbers (it reads them
// back every few levels to tighten the bounds and to notice the end): every kernel takes its real extent from here.
struct LevelDesc {
    int n_class[4];  // split, big, small, tiny
    int base;        // id of the level's first node
    int n;           // nodes in the level (2 x the active nodes of the level above)
    int pad[2];
};
enum { CLS_SPLIT = 0, CLS_BIG = 1, CLS_SMALL = 2, CLS_TINY = 3 };

struct LevelArgs {
    BuildArrays a;
    const int* active;     // node ids to split at this level, in level order
    const int* klist;      // positions in `active` handled by this launch (one size class)
    const LevelDesc* lv;   // this level (device memory)
    int stackless;
    int swap_policy;
    unsigned long long swap_seed;
    unsigned char* nflip;  // per node: children exchanged at flatten time
};

// Child `side` of the node at position k of the active list: box (zero signs resolved), range, and — for a range of
// <= 2 references — the leaf's triangles (:559-572,:612-625).  box_keys / zero_first: 6 entries (min xyz, max xyz).
__device__ __forceinline__ void create_child(const LevelArgs& g, int k, int side, unsigned start, unsigned len, unsigned mid, const int* box_keys,
                                             const int* zero_first, const int* refs) {
    const BuildArrays& a = g.a;
    const int child = g.lv->base + g.lv->n + 2 * k;  // the next level starts right after this one
    float bx[6];
    for (int c = 0; c < 6; ++c) {
        bx[c] = key2f(box_keys[c]);
        if (bx[c] == 0.0f && zero_first[c] != 0x7FFFFFFF) {
            const int r = refs[zero_first[c]];
            const float4 v = c < 3 ? a.tmin[r] : a.tmax[r];
            const int cc = c % 3;
            bx[c] = cc == 0 ? v.x : (cc == 1 ? v.y : v.z);
        }
    }
    const unsigned cstart = side == 0 ? start : mid;
    const unsigned clen = side == 0 ? mid - start : start + len - mid;
    a.nmin[child + side] = make_float4(bx[0], bx[1], bx[2], 0.0f);
    a.nmax[child + side] = make_float4(bx[3], bx[4], bx[5], 0.0f);
    a.nstart[child + side] = cstart;
    a.nlen[child + side] = clen;
    a.nchild[child + side] = -1;
    g.nflip[child + side] = 0;
    if (clen <= kMaxLeaf) {
        // leaf (:456-472): SortedTriangleReferences receives ranges right-first, so this range lands at T-(s+len)
        const unsigned at = a.T - (cstart + clen);
        for (unsigned j = 0; j < clen; ++j) {
            const int r = refs[cstart + j];
            a.tris_out[at + j] = make_int4((int)a.indices[3 * (size_t)r], (int)a.indices[3 * (size_t)r + 1], (int)a.indices[3 * (size_t)r + 2],
                                           a.mesh_ids ? a.mesh_ids[r] : 0);  // GenerateTriangles, :630-650
        }
    }
}

__device__ __forceinline__ void finish_parent(const LevelArgs& g, int k, int id, unsigned start, unsigned len) {
    g.a.nchild[id] = g.lv->base + g.lv->n + 2 * k;
    unsigned char flip = 0;
    if (g.stackless && g.swap_policy == CNDL_SWAP_HASHED)
        flip = (unsigned char)(mix64(g.swap_seed ^ mix64(((unsigned long long)start << 32) | len)) & 1ull);
    g.nflip[id] = flip;
}

// One block owns one node for one level: split search, partition, child boxes, leaf emission.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) level_step_kernel(LevelArgs g, int cls) {
    if ((int)blockIdx.x >= g.lv->n_class[cls]) return;  // the grid is an upper bound
    const BuildArrays& a = g.a;
    __shared__ int s_count[3][kBins];
    __shared__ int s_mn[3][3][kBins], s_mx[3][3][kBins];  // [axis][component][bin]
    constexpr bool AGG = BLOCK >= 128;
    __shared__ int s_warp[BLOCK / 32 + 1];
    __shared__ float s_best_cost, s_border;
    __shared__ int s_axis;
    __shared__ int s_box[2][6];        // child boxes as ordered keys
    __shared__ int s_zero_first[2][6]; // first position holding a zero component (sign tie-break)
    __shared__ int s_elem[BLOCK];
    __shared__ int s_rank[BLOCK];
    __shared__ unsigned char s_flag[BLOCK];

    const int k = g.klist[blockIdx.x];
    const int id = g.active[k];
    const unsigned start = a.nstart[id], len = a.nlen[id];
    const float4 bmn = a.nmin[id], bmx = a.nmax[id];
    const float nmn[3] = {bmn.x, bmn.y, bmn.z}, nmx[3] = {bmx.x, bmx.y, bmx.z};
    const int tid = threadIdx.x;

    // ---- SearchSAHPlaneBinned (:276-363): one pass over the range bins all three axes ----
    if (tid == 0) { s_best_cost = kInfCost; s_axis = 0; s_border = nmn[0]; }
    for (int b = tid; b < 3 * kBins; b += BLOCK) (&s_count[0][0])[b] = 0;
    for (int b = tid; b < 9 * kBins; b += BLOCK) { (&s_mn[0][0][0])[b] = f2key(kSentinelMax); (&s_mx[0][0][0])[b] = f2key(kSentinelMin); }
    __syncthreads();
    float scale[3], extent[3];
    bool axis_on[3];
    for (int ax = 0; ax < 3; ++ax) {
        axis_on[ax] = !(nmn[ax] == nmx[ax]);                 // :285
        extent[ax] = fsub(nmx[ax], nmn[ax]);
        scale[ax] = fdiv((float)kBins, extent[ax]);          // :295
    }
    for (unsigned base = 0; base < len; base += BLOCK) {
        const unsigned i = base + tid;
        const bool valid = i < len;
        const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const int r = a.refs[start + i];
            const float4 tm = a.tmin[r], tx = a.tmax[r];
            const float cz = a.tcz[r];
            // the sign of a zero in a bin box never reaches a decision (only areas use bin boxes)
            const int kmn[3] = {f2key(tm.x), f2key(tm.y), f2key(tm.z)}, kmx[3] = {f2key(tx.x), f2key(tx.y), f2key(tx.z)};
            const float cen[3] = {tm.w, tx.w, cz};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                if (!axis_on[ax]) continue;  // uniform for the block
                int b = __float2int_rz(fmul(fsub(cen[ax], nmn[ax]), scale[ax]));  // :302
                b = b > kBins - 1 ? kBins - 1 : (b < 0 ? 0 : b);
                // warp-level bin reduction: when every lane of the warp falls into the same bin (coherent input:
                // grids, sorted meshes) the warp combines with REDUX and one lane updates shared memory; 32-way
                // same-address atomics would serialise otherwise
                bool uniform = false;
                if (AGG) uniform = __all_sync(vmask, b == __shfl_sync(vmask, b, __ffs(vmask) - 1));
                if (AGG && uniform) {
                    int rmn[3], rmx[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) { rmn[c] = __reduce_min_sync(vmask, kmn[c]); rmx[c] = __reduce_max_sync(vmask, kmx[c]); }
                    if ((int)(threadIdx.x & 31) == __ffs(vmask) - 1) {
                        atomicAdd(&s_count[ax][b], __popc(vmask));
#pragma unroll
                        for (int c = 0; c < 3; ++c) { atomicMin(&s_mn[ax][c][b], rmn[c]); atomicMax(&s_mx[ax][c][b], rmx[c]); }
                    }
                } else {
                    atomicAdd(&s_count[ax][b], 1);
#pragma unroll
                    for (int c = 0; c < 3; ++c) { atomicMin(&s_mn[ax][c][b], kmn[c]); atomicMax(&s_mx[ax][c][b], kmx[c]); }
                }
            }
        }
    }
    __syncthreads();
    if (tid < 32) {
        for (int ax = 0; ax < 3; ++ax) {
            if (!axis_on[ax]) continue;
            warp_sah_eval(s_count[ax], s_mn[ax], s_mx[ax], ax, nmn[ax], extent[ax], &s_best_cost, &s_axis, &s_border);
            __syncwarp();
        }
    }
    __syncthreads();
    const int axis = s_axis;
    const float border = s_border;

    // ---- Lomuto partition (:532-549), emulated chunk by chunk.  Position i is visited holding its
    // original element; an element with flag L at chunk rank j lands at mid+j, and whatever sat there
    // at that moment moves to the L element's old place.  "At that moment" is resolved by following
    // the chain of earlier swaps inside the chunk (builder). ----
    unsigned mid = start;
    // the next chunk's references and centroids are fetched while the current chunk is processed: positions
    // past the current chunk are never written before they are visited
    int r_next = (unsigned)tid < len ? a.refs[start + tid] : -1;
    float c_next = r_next >= 0 ? centroid_of(a, r_next, axis) : 0.0f;
    for (unsigned i0 = start; i0 < start + len; i0 += BLOCK) {
        const unsigned n = min((unsigned)BLOCK, start + len - i0);
        const bool valid = (unsigned)tid < n;
        const int r = r_next;
        const bool f = valid && c_next < border;
        {
            const unsigned nxt = i0 + BLOCK + tid;
            r_next = nxt < start + len ? a.refs[nxt] : -1;
            c_next = r_next >= 0 ? centroid_of(a, r_next, axis) : 0.0f;
        }
        int nL;
        const int rank = block_exclusive_scan<BLOCK>(f ? 1 : 0, s_warp, nL);
        s_elem[tid] = r;
        s_rank[tid] = rank;
        s_flag[tid] = f ? 1 : 0;
        __syncthreads();
        const unsigned P = i0 + tid;
        const unsigned s = mid + (unsigned)rank;
        int displaced = -1;
        const bool writes_back = f && s != P && !(P >= mid && P < mid + (unsigned)nL);
        if (writes_back) {
            long long q = (long long)s - (long long)i0;
            while (q >= 0 && s_flag[q]) q = (long long)mid + s_rank[q] - (long long)i0;
            displaced = q < 0 ? a.refs[(long long)i0 + q] : s_elem[q];
        }
        __syncthreads();
        if (f) a.refs[s] = r;
        if (writes_back) a.refs[P] = displaced;
        mid += (unsigned)nL;
        __syncthreads();
    }
    // ---- split failure (:553-556) ----
    if (mid == start || mid == start + len) mid = start + len / 2;

    // ---- child boxes (:574-597): glm::min(tri, acc) keeps acc on ties => the FIRST zero decides a zero's sign ----
    if (tid < 12) {
        const int c = tid % 6;
        s_box[tid / 6][c] = c < 3 ? f2key(kSentinelMax) : f2key(kSentinelMin);
        s_zero_first[tid / 6][c] = 0x7FFFFFFF;
    }
    __syncthreads();
    {
        // every thread accumulates privately per side, then one set of atomics per thread that saw data
        float lmn[3] = {kSentinelMax, kSentinelMax, kSentinelMax}, lmx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
        float rmn[3] = {kSentinelMax, kSentinelMax, kSentinelMax}, rmx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
        bool sawl = false, sawr = false;
        for (unsigned i = tid; i < len; i += BLOCK) {
            const unsigned pos = start + i;
            const int r = a.refs[pos];
            const float4 tm = a.tmin[r], tx = a.tmax[r];
            const float vmn[3] = {tm.x, tm.y, tm.z}, vmx[3] = {tx.x, tx.y, tx.z};
            const int side = pos < mid ? 0 : 1;
            for (int c = 0; c < 3; ++c) {
                if (side == 0) { lmn[c] = fminf(lmn[c], vmn[c]); lmx[c] = fmaxf(lmx[c], vmx[c]); }
                else { rmn[c] = fminf(rmn[c], vmn[c]); rmx[c] = fmaxf(rmx[c], vmx[c]); }
                if (vmn[c] == 0.0f) atomicMin(&s_zero_first[side][c], (int)pos);
                if (vmx[c] == 0.0f) atomicMin(&s_zero_first[side][3 + c], (int)pos);
            }
            if (side == 0) sawl = true; else sawr = true;
        }
        for (int c = 0; c < 3; ++c) {
            if (sawl) { atomicMin(&s_box[0][c], f2key(lmn[c] == 0.0f ? 0.0f : lmn[c])); atomicMax(&s_box[0][3 + c], f2key(lmx[c] == 0.0f ? 0.0f : lmx[c])); }
            if (sawr) { atomicMin(&s_box[1][c], f2key(rmn[c] == 0.0f ? 0.0f : rmn[c])); atomicMax(&s_box[1][3 + c], f2key(rmx[c] == 0.0f ? 0.0f : rmx[c])); }
        }
    }
    __syncthreads();

    // ---- create the two children (:559-572,:612-625) ----
    if (tid < 2) create_child(g, k, tid, start, len, mid, s_box[tid], s_zero_first[tid], a.refs);
    if (tid == 0) finish_parent(g, k, id, start, len);
}

// The level step for ranges of 3..64 references, one WARP per node, kTinyWarps nodes per CTA.  The bottom levels of a
// tree hold most of its nodes (10 M triangles: 1.6 M ranges in one level); with one-warp CTAs the 32-CTA limit left
// the SMs half empty, and the three-axis bins made shared memory the next limit.  Here a lane keeps its (at most two)
// references in registers, the 64 bins of ONE axis at a time live in 1.8 KB of shared memory per warp, ranks come
// from ballots and the child boxes from warp reductions.  Same arithmetic, same order of decisions as level_step_kernel.
// 64 registers / 4 CTAs per SM: more resident warps (48 or 40 registers, with spills) measured 3-6 % slower.
constexpr int kTinyWarps = 8;
constexpr unsigned kDirectLen = 9;  // ranges up to this length search their split without bins (3 x 10 candidates <= 32 lanes)
constexpr unsigned kPackLen = 8;    // ranges up to this length are handled four to a warp, eight lanes each (level_step_tiny_kernel)
constexpr long long kPackMinRanges = 131072;  // levels with fewer tiny ranges keep one range per warp (262k triangles: 2.06 ms either way below it, 2.25 packed)
constexpr int kPackStride = 12;     // staged words per reference: box min, box max, (unused x3), bins of the three axes
struct TinyShared {
    int count[kTinyWarps][kBins];
    int mn[kTinyWarps][3][kBins], mx[kTinyWarps][3][kBins];
    float best_cost[kTinyWarps], border[kTinyWarps];
    int axis[kTinyWarps];
    float stage[kTinyWarps][4 * kPackLen * kPackStride];
};

// One range of 9..64 references (or any tiny range), the whole warp on it.
__device__ __forceinline__ void tiny_node_full(const LevelArgs& g, TinyShared& sh, int node) {
    const BuildArrays& a = g.a;
    int (&s_count)[kTinyWarps][kBins] = sh.count;
    int (&s_mn)[kTinyWarps][3][kBins] = sh.mn;
    int (&s_mx)[kTinyWarps][3][kBins] = sh.mx;
    float (&s_best_cost)[kTinyWarps] = sh.best_cost;
    float (&s_border)[kTinyWarps] = sh.border;
    int (&s_axis)[kTinyWarps] = sh.axis;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int k = g.klist[node];
    const int id = g.active[k];
    const unsigned start = a.nstart[id], len = a.nlen[id];
    const float4 bmn = a.nmin[id], bmx = a.nmax[id];
    const unsigned lt = (1u << lane) - 1u;

    // this lane's references: positions lane and lane + 32 of the range
    const bool v0 = (unsigned)lane < len, v1 = (unsigned)lane + 32 < len;
    const int r0 = v0 ? a.refs[start + lane] : -1, r1 = v1 ? a.refs[start + 32 + lane] : -1;
    // ---- SearchSAHPlaneBinned (:276-363) ----
    if (len <= kDirectLen) {
        // Up to 9 references (three quarters of all tiny ranges): no bins.  The cost of split i only changes where a bin is
        // occupied, and strict `<` over ascending i keeps the first index of every constant stretch, so the candidates per axis
        // are i = 0 and the bins of the references themselves: at most 10 x 3 (axis, candidate) pairs, one per lane, each
        // accumulating its two sides over the references staged in shared memory.  The minimum over (cost, axis, i) in that
        // order is the pair the sequential search ends with.
        float* stage = reinterpret_cast<float*>(&s_mn[wp][0][0]);  // 9 words per reference: box min, box max, centroid
        if (v0) {
            const float4 tm = a.tmin[r0], tx = a.tmax[r0];
            float* e = stage + 9 * lane;
            e[0] = tm.x; e[1] = tm.y; e[2] = tm.z; e[3] = tx.x; e[4] = tx.y; e[5] = tx.z; e[6] = tm.w; e[7] = tx.w; e[8] = a.tcz[r0];
        }
        __syncwarp();
        const int ax = lane / (kDirectLen + 1), j = lane % (kDirectLen + 1);
        const float lo = ax == 0 ? bmn.x : (ax == 1 ? bmn.y : bmn.z), hi = ax == 0 ? bmx.x : (ax == 1 ? bmx.y : bmx.z);
        const float extent = fsub(hi, lo);
        const float scale = fdiv((float)kBins, extent);         // :295
        int ci = 0;
        if ((unsigned)j < len) {
            ci = __float2int_rz(fmul(fsub(stage[9 * j + 6 + (ax < 3 ? ax : 0)], lo), scale));  // :302
            ci = ci > kBins - 1 ? kBins - 1 : (ci < 0 ? 0 : ci);
        }
        const bool active = ax < 3 && (unsigned)j <= len && !(lo == hi) && ci < kBins - 1;  // :285; i = 63 is not a split
        BinAgg L = agg_identity(), R = agg_identity();
        for (unsigned e = 0; e < len; ++e) {
            const float* el = stage + 9 * e;
            int b = __float2int_rz(fmul(fsub(el[6 + (ax < 3 ? ax : 0)], lo), scale));
            b = b > kBins - 1 ? kBins - 1 : (b < 0 ? 0 : b);
            BinAgg one;
            one.n = 1;
            one.mn[0] = el[0]; one.mn[1] = el[1]; one.mn[2] = el[2]; one.mx[0] = el[3]; one.mx[1] = el[4]; one.mx[2] = el[5];
            if (b <= ci) L = agg_join(L, one); else R = agg_join(R, one);
        }
        float cost = agg_cost(L, R);
        if (!active || !(cost == cost)) cost = __int_as_float(0x7F800000);  // a NaN cost is never selected by `<`
        int key = ax * kBins + ci, who = lane;
        for (int o = 16; o > 0; o >>= 1) {
            const float pc = __shfl_xor_sync(0xFFFFFFFFu, cost, o);
            const int pk = __shfl_xor_sync(0xFFFFFFFFu, key, o), pw = __shfl_xor_sync(0xFFFFFFFFu, who, o);
            if (pc < cost || (pc == cost && pk < key)) { cost = pc; key = pk; who = pw; }
        }
        const float my_border = fadd(lo, fmul(fdiv(extent, (float)kBins), __int2float_rn(ci + 1)));  // :347,:356
        const float win_border = __shfl_sync(0xFFFFFFFFu, my_border, who);
        __syncwarp();
        if (lane == 0) {
            const bool found = cost < kInfCost;
            s_axis[wp] = found ? key / kBins : 0;
            s_border[wp] = found ? win_border : bmn.x;
        }
        __syncwarp();
    } else {
        // one axis at a time through 64 bins
        if (lane == 0) { s_best_cost[wp] = kInfCost; s_axis[wp] = 0; s_border[wp] = bmn.x; }
        for (int ax = 0; ax < 3; ++ax) {
            const float lo = ax == 0 ? bmn.x : (ax == 1 ? bmn.y : bmn.z), hi = ax == 0 ? bmx.x : (ax == 1 ? bmx.y : bmx.z);
            if (lo == hi) continue;                                 // :285
            const float extent = fsub(hi, lo);
            const float scale = fdiv((float)kBins, extent);         // :295
            for (int b = lane; b < kBins; b += 32) s_count[wp][b] = 0;
            for (int b = lane; b < 3 * kBins; b += 32) { (&s_mn[wp][0][0])[b] = f2key(kSentinelMax); (&s_mx[wp][0][0])[b] = f2key(kSentinelMin); }
            __syncwarp();
            // the boxes are re-read per axis (L1 hits) instead of being held in registers across the split search
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = half ? r1 : r0;
                if (r < 0) continue;
                const float4 tm = a.tmin[r], tx = a.tmax[r];
                const float cen = ax == 0 ? tm.w : (ax == 1 ? tx.w : a.tcz[r]);
                int b = __float2int_rz(fmul(fsub(cen, lo), scale));  // :302
                b = b > kBins - 1 ? kBins - 1 : (b < 0 ? 0 : b);
                atomicAdd(&s_count[wp][b], 1);
                atomicMin(&s_mn[wp][0][b], f2key(tm.x)); atomicMin(&s_mn[wp][1][b], f2key(tm.y)); atomicMin(&s_mn[wp][2][b], f2key(tm.z));
                atomicMax(&s_mx[wp][0][b], f2key(tx.x)); atomicMax(&s_mx[wp][1][b], f2key(tx.y)); atomicMax(&s_mx[wp][2][b], f2key(tx.z));
            }
            __syncwarp();
            warp_sah_eval(s_count[wp], s_mn[wp], s_mx[wp], ax, lo, extent, &s_best_cost[wp], &s_axis[wp], &s_border[wp]);
            __syncwarp();
        }
    }
    const int axis = s_axis[wp];
    const float border = s_border[wp];

    // ---- Lomuto partition (:532-549), 32 positions at a time (see level_step_kernel) ----
    unsigned mid = start;
    const float c0 = v0 ? centroid_of(a, r0, axis) : 0.0f, c1 = v1 ? centroid_of(a, r1, axis) : 0.0f;
    for (int half = 0; half < 2; ++half) {
        const unsigned i0 = start + 32u * half;
        if (i0 >= start + len) break;
        const int r = half ? r1 : r0;
        const bool f = (half ? v1 : v0) && (half ? c1 : c0) < border;
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, f);
        const int rank = __popc(mask & lt), nL = __popc(mask);
        const unsigned P = i0 + lane, s = mid + (unsigned)rank;
        const bool writes_back = f && s != P && !(P >= mid && P < mid + (unsigned)nL);
        long long q = (long long)s - (long long)i0;
        if (writes_back)
            while (q >= 0 && ((mask >> q) & 1u)) q = (long long)mid + __popc(mask & ((1u << q) - 1u)) - (long long)i0;
        const int from_chunk = __shfl_sync(0xFFFFFFFFu, r, q < 0 ? 0 : (int)q);
        int displaced = -1;
        if (writes_back) displaced = q < 0 ? a.refs[(long long)i0 + q] : from_chunk;
        __syncwarp();
        if (f) a.refs[s] = r;
        if (writes_back) a.refs[P] = displaced;
        mid += (unsigned)nL;
        __syncwarp();
    }
    if (mid == start || mid == start + len) mid = start + len / 2;  // split failure (:553-556)

    // ---- child boxes (:574-597): the FIRST zero decides a zero's sign ----
    int box[2][6], zf[2][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        box[0][c] = box[1][c] = c < 3 ? f2key(kSentinelMax) : f2key(kSentinelMin);
        zf[0][c] = zf[1][c] = 0x7FFFFFFF;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const unsigned pos = start + 32u * half + lane;
        if (pos < start + len) {
            const int r = a.refs[pos];
            const float4 tm = a.tmin[r], tx = a.tmax[r];
            const float v[6] = {tm.x, tm.y, tm.z, tx.x, tx.y, tx.z};
            const int side = pos < mid ? 0 : 1;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int key = f2key(v[c] == 0.0f ? 0.0f : v[c]);
                if (side == 0) { box[0][c] = c < 3 ? min(box[0][c], key) : max(box[0][c], key); if (v[c] == 0.0f) zf[0][c] = min(zf[0][c], (int)pos); }
                else { box[1][c] = c < 3 ? min(box[1][c], key) : max(box[1][c], key); if (v[c] == 0.0f) zf[1][c] = min(zf[1][c], (int)pos); }
            }
        }
    }
#pragma unroll
    for (int sd = 0; sd < 2; ++sd)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            box[sd][c] = c < 3 ? __reduce_min_sync(0xFFFFFFFFu, box[sd][c]) : __reduce_max_sync(0xFFFFFFFFu, box[sd][c]);
            zf[sd][c] = __reduce_min_sync(0xFFFFFFFFu, zf[sd][c]);
        }
    __syncwarp();  // the partition's writes to refs are visible to the two lanes below

    // ---- create the two children (:559-572,:612-625) ----
    if (lane == 0) create_child(g, k, 0, start, len, mid, box[0], zf[0], a.refs);
    if (lane == 1) create_child(g, k, 1, start, len, mid, box[1], zf[1], a.refs);
    if (lane == 0) finish_parent(g, k, id, start, len);
    __syncwarp();
}

// The level step for the tiny class.  A warp takes FOUR consecutive ranges of the class.  Those of at most kPackLen references —
// three quarters of all tiny ranges; at 10 M triangles 1.6 M of them in one level — are split side by side, eight lanes each: the
// split search is the direct one of tiny_node_full (candidates = the references' own bins and i = 0, four rounds of eight
// candidates), the Lomuto emulation fits one 8-position chunk, the child boxes come from three shuffle steps inside the group, and
// lanes 0 / 1 of every group create the children — so the per-range cost that does not depend on the range's length (partition,
// boxes, children: ~900 warp instructions before) is paid once per four ranges.  The longer ranges of the four follow one at a
// time on the whole warp.  Same arithmetic, same order of decisions.
// PACK = 1: one range per warp (levels with few tiny ranges are latency-bound: four ranges in a row on one warp would only
// lengthen the level); PACK = 4: the packed form for levels with many.
template <int PACK>
__global__ void __launch_bounds__(32 * kTinyWarps, 4) level_step_tiny_kernel(LevelArgs g) {
    const BuildArrays& a = g.a;
    const int n_nodes = g.lv->n_class[CLS_TINY];  // the grid is an upper bound
    __shared__ TinyShared sh;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (PACK == 1) {
        const int node = blockIdx.x * kTinyWarps + wp;
        if (node < n_nodes) tiny_node_full(g, sh, node);
        return;
    }
    const int first = (blockIdx.x * kTinyWarps + wp) * 4;
    if (first >= n_nodes) return;  // whole warps leave; nothing below synchronises across warps
    const int grp = lane >> 3, gl = lane & 7;
    const unsigned gm = 0xFFu << (8 * grp);
    const int node = first + grp;
    const bool have = node < n_nodes;
    int k = 0, id = 0;
    unsigned start = 0, len = 0;
    if (have) {
        k = g.klist[node];
        id = g.active[k];
        start = a.nstart[id];
        len = a.nlen[id];
    }
    if (have && len <= kPackLen) {
        const float4 bmn = a.nmin[id], bmx = a.nmax[id];
        float* stg = sh.stage[wp] + grp * (kPackLen * kPackStride);
        int* sbin = reinterpret_cast<int*>(stg);
        const bool v = (unsigned)gl < len;
        const int r = v ? a.refs[start + gl] : -1;
        float cen[3] = {0.0f, 0.0f, 0.0f};
        if (v) {
            const float4 tm = a.tmin[r], tx = a.tmax[r];
            cen[0] = tm.w; cen[1] = tx.w; cen[2] = a.tcz[r];
            float* e = stg + kPackStride * gl;
            e[0] = tm.x; e[1] = tm.y; e[2] = tm.z; e[3] = tx.x; e[4] = tx.y; e[5] = tx.z;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const float lo = ax == 0 ? bmn.x : (ax == 1 ? bmn.y : bmn.z), hi = ax == 0 ? bmx.x : (ax == 1 ? bmx.y : bmx.z);
                const float scale = fdiv((float)kBins, fsub(hi, lo));       // :295
                int b = __float2int_rz(fmul(fsub(cen[ax], lo), scale));       // :302
                b = b > kBins - 1 ? kBins - 1 : (b < 0 ? 0 : b);
                sbin[kPackStride * gl + 9 + ax] = b;
            }
        }
        __syncwarp(gm);
        // ---- SearchSAHPlaneBinned (:276-363), direct form: 27 (axis, candidate) pairs, eight per round ----
        float best_cost = __int_as_float(0x7F800000), best_border = bmn.x;
        int best_key = 0x7FFFFFFF;
        for (int rr = 0; rr < 4; ++rr) {
            const int c = rr * 8 + gl;
            const int ax = c / 9 > 2 ? 2 : c / 9, j = c - (c / 9) * 9;
            const float lo = ax == 0 ? bmn.x : (ax == 1 ? bmn.y : bmn.z), hi = ax == 0 ? bmx.x : (ax == 1 ? bmx.y : bmx.z);
            int ci = 0;
            if ((unsigned)j < len) ci = sbin[kPackStride * j + 9 + ax];
            const bool active = c < 27 && (unsigned)j <= len && !(lo == hi) && ci < kBins - 1;  // :285; i = 63 is not a split
            float cost = __int_as_float(0x7F800000);
            if (active) {
                BinAgg L = agg_identity(), R = agg_identity();
                for (unsigned e = 0; e < len; ++e) {
                    const float* el = stg + kPackStride * e;
                    BinAgg one;
                    one.n = 1;
                    one.mn[0] = el[0]; one.mn[1] = el[1]; one.mn[2] = el[2]; one.mx[0] = el[3]; one.mx[1] = el[4]; one.mx[2] = el[5];
                    if (sbin[kPackStride * e + 9 + ax] <= ci) L = agg_join(L, one); else R = agg_join(R, one);
                }
                cost = agg_cost(L, R);
                if (!(cost == cost)) cost = __int_as_float(0x7F800000);  // a NaN cost is never selected by `<`
            }
            const int key = ax * kBins + ci;
            if (cost < best_cost || (cost == best_cost && key < best_key)) {
                best_cost = cost;
                best_key = key;
                best_border = fadd(lo, fmul(fdiv(fsub(hi, lo), (float)kBins), __int2float_rn(ci + 1)));  // :347,:356
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float pc = __shfl_xor_sync(gm, best_cost, o), pb = __shfl_xor_sync(gm, best_border, o);
            const int pk = __shfl_xor_sync(gm, best_key, o);
            if (pc < best_cost || (pc == best_cost && pk < best_key)) { best_cost = pc; best_key = pk; best_border = pb; }
        }
        const bool found = best_cost < kInfCost;
        const int axis = found ? best_key / kBins : 0;
        const float border = found ? best_border : bmn.x;

        // ---- Lomuto partition (:532-549): one chunk of at most eight positions (see level_step_kernel) ----
        const bool f = v && (axis == 0 ? cen[0] : (axis == 1 ? cen[1] : cen[2])) < border;
        const unsigned m8 = (__ballot_sync(gm, f) >> (8 * grp)) & 0xFFu;
        const int rank = __popc(m8 & ((1u << gl) - 1u)), nL = __popc(m8);
        const unsigned P = start + gl, s_pos = start + (unsigned)rank;
        const bool writes_back = f && s_pos != P && !(P < start + (unsigned)nL);
        int q = rank;
        if (writes_back)
            while ((m8 >> q) & 1u) q = __popc(m8 & ((1u << q) - 1u));
        const int displaced = __shfl_sync(gm, r, 8 * grp + q);
        __syncwarp(gm);
        if (f) a.refs[s_pos] = r;
        if (writes_back) a.refs[P] = displaced;
        unsigned mid = start + (unsigned)nL;
        __syncwarp(gm);
        if (mid == start || mid == start + len) mid = start + len / 2;  // split failure (:553-556)

        // ---- child boxes (:574-597): the FIRST zero decides a zero's sign ----
        int box[2][6], zf[2][6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            box[0][c] = box[1][c] = c < 3 ? f2key(kSentinelMax) : f2key(kSentinelMin);
            zf[0][c] = zf[1][c] = 0x7FFFFFFF;
        }
        bool zero_seen = false;
        if (v) {
            const int r2 = a.refs[P];
            const float4 tm = a.tmin[r2], tx = a.tmax[r2];
            const float val[6] = {tm.x, tm.y, tm.z, tx.x, tx.y, tx.z};
            const int side = P < mid ? 0 : 1;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int key = f2key(val[c] == 0.0f ? 0.0f : val[c]);
                if (side == 0) box[0][c] = key; else box[1][c] = key;
                if (val[c] == 0.0f) {
                    zero_seen = true;
                    if (side == 0) zf[0][c] = (int)P; else zf[1][c] = (int)P;
                }
            }
        }
        const bool any_zero = (__ballot_sync(gm, zero_seen) & gm) != 0u;
#pragma unroll
        for (int sd = 0; sd < 2; ++sd)
#pragma unroll
            for (int c = 0; c < 6; ++c)
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    const int other = __shfl_xor_sync(gm, box[sd][c], o);
                    box[sd][c] = c < 3 ? min(box[sd][c], other) : max(box[sd][c], other);
                }
        if (any_zero) {
#pragma unroll
            for (int sd = 0; sd < 2; ++sd)
#pragma unroll
                for (int c = 0; c < 6; ++c)
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) zf[sd][c] = min(zf[sd][c], __shfl_xor_sync(gm, zf[sd][c], o));
        }
        // ---- create the two children (:559-572,:612-625): lanes 0 and 1 of the group, side = lane ----
        if (gl < 2) {
            int bk[6], z[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) { bk[c] = gl ? box[1][c] : box[0][c]; z[c] = gl ? zf[1][c] : zf[0][c]; }
            create_child(g, k, gl, start, len, mid, bk, z, a.refs);
        }
        if (gl == 0) finish_parent(g, k, id, start, len);
    }
    __syncwarp();
    // ---- the longer ranges of the four, one at a time on the whole warp ----
#pragma unroll 1
    for (int qn = 0; qn < 4; ++qn) {
        const unsigned lq = __shfl_sync(0xFFFFFFFFu, len, 8 * qn);
        if (lq > kPackLen) tiny_node_full(g, sh, first + qn);
    }
}

// ---------------------------------------------------------------------------------------------
// The same level step for ranges longer than the split threshold, spread over one CTA per 512 or 2048 references.
// One CTA would walk a 262k-reference root three times alone (~1 ms); every result below is a pure function of
// the range, so the passes can be cut anywhere:
//   split_bin     bins its chunk in shared memory and merges into the node's global bins (min / max / + only);
//                 the last CTA of a node runs the split search on the merged bins;
//   split_count   writes each reference's side and counts the left ones per chunk; the last CTA scans the counts;
//   split_rank    gives every left reference its rank r in the range: it lands at start + r (Lomuto keeps the left
//                 references in order);
//   split_gather  fills the positions from start + nL on.  The sequential loop leaves position P alone when its
//                 reference is a right one; when it is the k-th left one, P receives whatever position start + k held
//                 when the loop got to P, which is again either an untouched right reference or the filling of an
//                 earlier left one's place: follow q -> start + rank(q) until a right reference turns up.  The chains of
//                 different P are disjoint (rank is injective), but one chain can be long: inside a run of left references
//                 that follows the c-th right one every hop is q -> q - c, so the walk takes a whole run per iteration
//                 (rpos gives the run's first position) and a chain costs at most O(sqrt(len)) dependent loads;
//   split_finish  copies the new order back, reduces the two child boxes and the last CTA creates the children.
struct SplitArgs {
    LevelArgs g;         // klist = positions of this level's split nodes in the active list; g.lv->n_class[CLS_SPLIT] of them
    int* chunk_base;     // n_split + 1: first chunk of every split node, total
    int* bins;           // n_split x kBinInts
    int* done;           // n_split x 3 arrival counters
    int* split;          // n_split x 4: axis, border bits, nL
    int* boxes;          // n_split x 24: box keys [2][6], first zero position [2][6]
    int* chunk_l;        // per chunk: left count, replaced by its exclusive prefix inside the node
    int* rankflag;       // per position: rank of a left reference, -1 for a right one
    int* rpos;           // start + j: position (relative to start) of the range's j-th right reference
    int* alt;            // per position: the partitioned order
};

struct SplitWhere { int h, chunk, n_chunks, id, k; unsigned start, len; };

__device__ __forceinline__ bool split_locate(const SplitArgs& s, SplitWhere& w) {
    const int n_split = s.g.lv->n_class[CLS_SPLIT];
    if (n_split <= 0) return false;
    const int total = s.chunk_base[n_split];
    if ((int)blockIdx.x >= total) return false;
    int lo = 0, hi = n_split - 1;  // last h with chunk_base[h] <= blockIdx.x
    while (lo < hi) {
        const int m = (lo + hi + 1) >> 1;
        if (s.chunk_base[m] <= (int)blockIdx.x) lo = m; else hi = m - 1;
    }
    w.h = lo;
    w.chunk = (int)blockIdx.x - s.chunk_base[lo];
    w.n_chunks = s.chunk_base[lo + 1] - s.chunk_base[lo];
    w.k = s.g.klist[lo];
    w.id = s.g.active[w.k];
    w.start = s.g.a.nstart[w.id];
    w.len = s.g.a.nlen[w.id];
    return true;
}

// true in the CTA that arrives last at counter `which` of node h; everything written before by the others is visible to it
__device__ __forceinline__ bool split_arrive_last(const SplitArgs& s, const SplitWhere& w, int which, int* s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_flag = atomicAdd(&s.done[3 * w.h + which], 1) == w.n_chunks - 1;
    __syncthreads();
    const bool last = *s_flag != 0;
    if (last) __threadfence();
    return last;
}

__global__ void __launch_bounds__(1024) split_prep_kernel(SplitArgs s, int chunk_len) {
    __shared__ int s_warp[1024 / 32 + 1];
    const int tid = threadIdx.x;
    const int n_split = s.g.lv->n_class[CLS_SPLIT];
    int carry = 0;
    for (int base = 0; base < n_split; base += 1024) {
        const int h = base + tid;
        int nch = 0;
        if (h < n_split) nch = (int)((s.g.a.nlen[s.g.active[s.g.klist[h]]] + chunk_len - 1) / chunk_len);
        int total;
        const int ex = block_exclusive_scan<1024>(nch, s_warp, total);
        if (h < n_split) s.chunk_base[h] = carry + ex;
        carry += total;
    }
    if (tid == 0) s.chunk_base[n_split] = carry;
    for (int i = tid; i < n_split * kBinInts; i += 1024) {
        const int j = i % kBinInts;
        s.bins[i] = j < 3 * kBins ? 0 : (j < 12 * kBins ? f2key(kSentinelMax) : f2key(kSentinelMin));
    }
    for (int i = tid; i < n_split * 3; i += 1024) s.done[i] = 0;
    for (int i = tid; i < n_split * 24; i += 1024) {
        const int j = i % 24;
        s.boxes[i] = j >= 12 ? 0x7FFFFFFF : (j % 6 < 3 ? f2key(kSentinelMax) : f2key(kSentinelMin));
    }
}

template <int ITEMS>
__global__ void __launch_bounds__(kSplitBlock) split_bin_kernel(SplitArgs s) {
    constexpr int kSplitItems = ITEMS, kSplitChunk = kSplitBlock * ITEMS;
    __shared__ int s_count[3][kBins];
    __shared__ int s_mn[3][3][kBins], s_mx[3][3][kBins];
    __shared__ float s_best_cost, s_border;
    __shared__ int s_axis, s_last;
    SplitWhere w;
    if (!split_locate(s, w)) return;
    const BuildArrays& a = s.g.a;
    const int tid = threadIdx.x;
    const float4 bmn = a.nmin[w.id], bmx = a.nmax[w.id];
    const float nmn[3] = {bmn.x, bmn.y, bmn.z}, nmx[3] = {bmx.x, bmx.y, bmx.z};
    for (int b = tid; b < 3 * kBins; b += kSplitBlock) (&s_count[0][0])[b] = 0;
    for (int b = tid; b < 9 * kBins; b += kSplitBlock) { (&s_mn[0][0][0])[b] = f2key(kSentinelMax); (&s_mx[0][0][0])[b] = f2key(kSentinelMin); }
    __syncthreads();
    float scale[3], extent[3];
    bool axis_on[3];
    for (int ax = 0; ax < 3; ++ax) {
        axis_on[ax] = !(nmn[ax] == nmx[ax]);                 // :285
        extent[ax] = fsub(nmx[ax], nmn[ax]);
        scale[ax] = fdiv((float)kBins, extent[ax]);          // :295
    }
    for (int j = 0; j < kSplitItems; ++j) {
        const unsigned i = (unsigned)w.chunk * kSplitChunk + j * kSplitBlock + tid;
        const bool valid = i < w.len;
        const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {
            const int r = a.refs[w.start + i];
            const float4 tm = a.tmin[r], tx = a.tmax[r];
            const float cz = a.tcz[r];
            const int kmn[3] = {f2key(tm.x), f2key(tm.y), f2key(tm.z)}, kmx[3] = {f2key(tx.x), f2key(tx.y), f2key(tx.z)};
            const float cen[3] = {tm.w, tx.w, cz};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                if (!axis_on[ax]) continue;
                int b = __float2int_rz(fmul(fsub(cen[ax], nmn[ax]), scale[ax]));  // :302
                b = b > kBins - 1 ? kBins - 1 : (b < 0 ? 0 : b);
                if (__all_sync(vmask, b == __shfl_sync(vmask, b, __ffs(vmask) - 1))) {
                    int rmn[3], rmx[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) { rmn[c] = __reduce_min_sync(vmask, kmn[c]); rmx[c] = __reduce_max_sync(vmask, kmx[c]); }
                    if ((int)(threadIdx.x & 31) == __ffs(vmask) - 1) {
                        atomicAdd(&s_count[ax][b], __popc(vmask));
#pragma unroll
                        for (int c = 0; c < 3; ++c) { atomicMin(&s_mn[ax][c][b], rmn[c]); atomicMax(&s_mx[ax][c][b], rmx[c]); }
                    }
                } else {
                    atomicAdd(&s_count[ax][b], 1);
#pragma unroll
                    for (int c = 0; c < 3; ++c) { atomicMin(&s_mn[ax][c][b], kmn[c]); atomicMax(&s_mx[ax][c][b], kmx[c]); }
                }
            }
        }
    }
    __syncthreads();
    int* gb = s.bins + (size_t)w.h * kBinInts;
    for (int b = tid; b < 3 * kBins; b += kSplitBlock) {
        const int n = (&s_count[0][0])[b];
        if (n) atomicAdd(gb + b, n);
    }
    for (int b = tid; b < 9 * kBins; b += kSplitBlock) {
        const int mn = (&s_mn[0][0][0])[b], mx = (&s_mx[0][0][0])[b];
        if (mn != f2key(kSentinelMax)) atomicMin(gb + 3 * kBins + b, mn);
        if (mx != f2key(kSentinelMin)) atomicMax(gb + 12 * kBins + b, mx);
    }
    if (!split_arrive_last(s, w, 0, &s_last)) return;
    for (int b = tid; b < 3 * kBins; b += kSplitBlock) (&s_count[0][0])[b] = __ldcg(gb + b);
    for (int b = tid; b < 9 * kBins; b += kSplitBlock) {
        (&s_mn[0][0][0])[b] = __ldcg(gb + 3 * kBins + b);
        (&s_mx[0][0][0])[b] = __ldcg(gb + 12 * kBins + b);
    }
    if (tid == 0) { s_best_cost = kInfCost; s_axis = 0; s_border = nmn[0]; }
    __syncthreads();
    if (tid < 32) {
        for (int ax = 0; ax < 3; ++ax) {
            if (!axis_on[ax]) continue;
            warp_sah_eval(s_count[ax], s_mn[ax], s_mx[ax], ax, nmn[ax], extent[ax], &s_best_cost, &s_axis, &s_border);
            __syncwarp();
        }
        if (tid == 0) {
            s.split[4 * w.h] = s_axis;
            s.split[4 * w.h + 1] = __float_as_int(s_border);
        }
    }
}

template <int ITEMS>
__global__ void __launch_bounds__(kSplitBlock) split_count_kernel(SplitArgs s) {
    constexpr int kSplitItems = ITEMS, kSplitChunk = kSplitBlock * ITEMS;
    __shared__ int s_warp[kSplitBlock / 32 + 1];
    __shared__ int s_last;
    SplitWhere w;
    if (!split_locate(s, w)) return;
    const BuildArrays& a = s.g.a;
    const int tid = threadIdx.x;
    const int axis = s.split[4 * w.h];
    const float border = __int_as_float(s.split[4 * w.h + 1]);
    int cnt = 0;
    for (int j = 0; j < kSplitItems; ++j) {
        const unsigned i = (unsigned)w.chunk * kSplitChunk + j * kSplitBlock + tid;
        if (i < w.len) {
            const bool f = centroid_of(a, a.refs[w.start + i], axis) < border;  // :534
            s.rankflag[w.start + i] = f ? 0 : -1;
            cnt += f ? 1 : 0;
        }
    }
    int total;
    block_exclusive_scan<kSplitBlock>(cnt, s_warp, total);
    int* cl = s.chunk_l + s.chunk_base[w.h];
    if (tid == 0) cl[w.chunk] = total;
    if (!split_arrive_last(s, w, 1, &s_last)) return;
    int carry = 0;
    for (int base = 0; base < w.n_chunks; base += kSplitBlock) {
        const int c = base + tid;
        const int v = c < w.n_chunks ? __ldcg(cl + c) : 0;
        int tot;
        const int ex = block_exclusive_scan<kSplitBlock>(v, s_warp, tot);
        if (c < w.n_chunks) cl[c] = carry + ex;
        carry += tot;
    }
    if (tid == 0) s.split[4 * w.h + 2] = carry;
}

template <int ITEMS>
__global__ void __launch_bounds__(kSplitBlock) split_rank_kernel(SplitArgs s) {
    constexpr int kSplitItems = ITEMS, kSplitChunk = kSplitBlock * ITEMS;
    __shared__ int s_warp[kSplitBlock / 32 + 1];
    SplitWhere w;
    if (!split_locate(s, w)) return;
    const BuildArrays& a = s.g.a;
    const int tid = threadIdx.x;
    const unsigned i0 = (unsigned)w.chunk * kSplitChunk + (unsigned)tid * kSplitItems;  // this thread's kSplitItems consecutive positions
    bool fl[kSplitItems];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kSplitItems; ++j) {
        fl[j] = i0 + j < w.len && s.rankflag[w.start + i0 + j] >= 0;
        cnt += fl[j] ? 1 : 0;
    }
    int total;
    int rank = s.chunk_l[s.chunk_base[w.h] + w.chunk] + block_exclusive_scan<kSplitBlock>(cnt, s_warp, total);
#pragma unroll
    for (int j = 0; j < kSplitItems; ++j) {
        if (fl[j]) {
            s.rankflag[w.start + i0 + j] = rank;
            s.alt[w.start + rank] = a.refs[w.start + i0 + j];
            ++rank;
        } else if (i0 + j < w.len) {
            s.rpos[w.start + (i0 + j - (unsigned)rank)] = (int)(i0 + j);  // i0 + j - rank right references precede this one
        }
    }
}

template <int ITEMS>
__global__ void __launch_bounds__(kSplitBlock) split_gather_kernel(SplitArgs s) {
    constexpr int kSplitItems = ITEMS, kSplitChunk = kSplitBlock * ITEMS;
    SplitWhere w;
    if (!split_locate(s, w)) return;
    const BuildArrays& a = s.g.a;
    const unsigned nL = (unsigned)s.split[4 * w.h + 2];
    for (int j = 0; j < kSplitItems; ++j) {
        const unsigned i = (unsigned)w.chunk * kSplitChunk + j * kSplitBlock + threadIdx.x;
        if (i < w.len && i >= nL) {
            unsigned p = i;  // relative to start
            for (int rf = s.rankflag[w.start + p]; rf >= 0; rf = s.rankflag[w.start + p]) {
                const unsigned c = p - (unsigned)rf;  // right references before p: >= 1 on a chain that starts at or after nL
                if (c == 0) break;
                const unsigned run0 = (unsigned)s.rpos[w.start + c - 1] + 1;  // first position of the run of left references around p
                p -= (p - run0) / c * c;  // the hops that stay inside the run (every position of the run has the same c) ...
                p -= c;                   // ... and the one that leaves it
            }
            s.alt[w.start + i] = a.refs[w.start + p];
        }
    }
}

template <int ITEMS>
__global__ void __launch_bounds__(kSplitBlock) split_finish_kernel(SplitArgs s) {
    constexpr int kSplitItems = ITEMS, kSplitChunk = kSplitBlock * ITEMS;
    __shared__ int s_box[2][6], s_zero_first[2][6];
    __shared__ int s_last;
    SplitWhere w;
    if (!split_locate(s, w)) return;
    const BuildArrays& a = s.g.a;
    const int tid = threadIdx.x;
    const unsigned start = w.start, len = w.len;
    unsigned mid = start + (unsigned)s.split[4 * w.h + 2];
    if (mid == start || mid == start + len) mid = start + len / 2;  // split failure (:553-556)
    if (tid < 12) {
        const int c = tid % 6;
        s_box[tid / 6][c] = c < 3 ? f2key(kSentinelMax) : f2key(kSentinelMin);
        s_zero_first[tid / 6][c] = 0x7FFFFFFF;
    }
    __syncthreads();
    {
        float lmn[3] = {kSentinelMax, kSentinelMax, kSentinelMax}, lmx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
        float rmn[3] = {kSentinelMax, kSentinelMax, kSentinelMax}, rmx[3] = {kSentinelMin, kSentinelMin, kSentinelMin};
        bool sawl = false, sawr = false;
        for (int j = 0; j < kSplitItems; ++j) {
            const unsigned i = (unsigned)w.chunk * kSplitChunk + j * kSplitBlock + tid;
            if (i >= len) continue;
            const unsigned pos = start + i;
            const int r = s.alt[pos];
            a.refs[pos] = r;
            const float4 tm = a.tmin[r], tx = a.tmax[r];
            const float vmn[3] = {tm.x, tm.y, tm.z}, vmx[3] = {tx.x, tx.y, tx.z};
            const int side = pos < mid ? 0 : 1;
            for (int c = 0; c < 3; ++c) {
                if (side == 0) { lmn[c] = fminf(lmn[c], vmn[c]); lmx[c] = fmaxf(lmx[c], vmx[c]); }
                else { rmn[c] = fminf(rmn[c], vmn[c]); rmx[c] = fmaxf(rmx[c], vmx[c]); }
                if (vmn[c] == 0.0f) atomicMin(&s_zero_first[side][c], (int)pos);
                if (vmx[c] == 0.0f) atomicMin(&s_zero_first[side][3 + c], (int)pos);
            }
            if (side == 0) sawl = true; else sawr = true;
        }
        for (int c = 0; c < 3; ++c) {
            if (sawl) { atomicMin(&s_box[0][c], f2key(lmn[c] == 0.0f ? 0.0f : lmn[c])); atomicMax(&s_box[0][3 + c], f2key(lmx[c] == 0.0f ? 0.0f : lmx[c])); }
            if (sawr) { atomicMin(&s_box[1][c], f2key(rmn[c] == 0.0f ? 0.0f : rmn[c])); atomicMax(&s_box[1][3 + c], f2key(rmx[c] == 0.0f ? 0.0f : rmx[c])); }
        }
    }
    __syncthreads();
    int* gb = s.boxes + 24 * (size_t)w.h;
    if (tid < 12) {
        const int side = tid / 6, c = tid % 6;
        if (c < 3) atomicMin(gb + tid, s_box[side][c]); else atomicMax(gb + tid, s_box[side][c]);
        if (s_zero_first[side][c] != 0x7FFFFFFF) atomicMin(gb + 12 + tid, s_zero_first[side][c]);
    }
    if (!split_arrive_last(s, w, 2, &s_last)) return;
    if (tid < 12) {
        s_box[tid / 6][tid % 6] = __ldcg(gb + tid);
        s_zero_first[tid / 6][tid % 6] = __ldcg(gb + 12 + tid);
    }
    __syncthreads();
    // the children read the partitioned order from `alt`: complete since the previous launch, whereas other CTAs' copies
    // into refs are ordered only by the arrival counter
    if (tid < 2) create_child(s.g, w.k, tid, start, len, mid, s_box[tid], s_zero_first[tid], s.alt);
    if (tid == 0) finish_parent(s.g, w.k, w.id, start, len);
}

// Next level's active list.  Both kernels read where the new level sits from the descriptor of the level above
// (base' = base + n, n' = 2 x its active nodes), classify + scan + compact its nodes, and write the new level's descriptor to
// device memory and to its mapped host copy; the host does not wait for it (see build_object).
//
// A level of at most kSmallLevel nodes: one CTA (one launch instead of five).
constexpr int kSmallLevel = 4096, kMaxRunLevels = 2048, kMaxLevels = 1 << 16;
__device__ __forceinline__ void next_level_extent(const LevelDesc& up, int& base, int& n) {
    base = up.base + up.n;
    n = 2 * (up.n_class[0] + up.n_class[1] + up.n_class[2] + up.n_class[3]);
}
__device__ __forceinline__ void publish_level(LevelDesc* lv, LevelDesc* host_lv, int level, const int counts[4], int base, int n) {
    for (int c = 0; c < 4; ++c) { lv[level].n_class[c] = counts[c]; host_lv[level].n_class[c] = counts[c]; }
    lv[level].base = base; lv[level].n = n;
    host_lv[level].base = base; host_lv[level].n = n;
}
__global__ void __launch_bounds__(1024) level_compact_small_kernel(const unsigned* nlen, LevelDesc* lv, LevelDesc* host_lv /* mapped host memory */, int level,
                                                                   unsigned split_node, int* active, int* klist_split, int* klist_big, int* klist_small,
                                                                   int* klist_tiny) {
    __shared__ int s_warp[1024 / 32 + 1];
    const int tid = threadIdx.x;
    int base, n;
    next_level_extent(lv[level - 1], base, n);
    int cls[4], c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tid * 4 + j;
        cls[j] = -1;
        if (i < n) {
            const unsigned len = nlen[base + i];
            cls[j] = len > split_node ? 0 : (len > kBigNode ? 1 : (len > kTinyNode ? 2 : (len > kMaxLeaf ? 3 : -1)));
        }
        c0 += cls[j] == 0; c1 += cls[j] == 1; c2 += cls[j] == 2; c3 += cls[j] == 3;
    }
    int t[4];
    int o0 = block_exclusive_scan<1024>(c0, s_warp, t[0]);
    int o1 = block_exclusive_scan<1024>(c1, s_warp, t[1]);
    int o2 = block_exclusive_scan<1024>(c2, s_warp, t[2]);
    int o3 = block_exclusive_scan<1024>(c3, s_warp, t[3]);
    int k = o0 + o1 + o2 + o3;  // active nodes before this thread's first one, in level order
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (cls[j] < 0) continue;
        active[k] = base + tid * 4 + j;
        if (cls[j] == 0) klist_split[o0++] = k;
        else if (cls[j] == 1) klist_big[o1++] = k;
        else if (cls[j] == 2) klist_small[o2++] = k;
        else klist_tiny[o3++] = k;
        ++k;
    }
    if (tid == 0) publish_level(lv, host_lv, level, t, base, n);
}

// The same for larger levels in ONE pass: tiles of 1024 nodes, chained by decoupled look-back over the four class counts
// (replaces classify + three scan launches + compact, ~25 us of serial launches per level).  Tiles take their index from
// a ticket (the new level's own: lv[level].pad[0], zero since the build started), so a tile only ever waits for tiles that are
// already running; the grid is an upper bound, and CTAs whose ticket lies beyond the level leave at once.  Tile states carry the
// level number (epoch) instead of being cleared between levels; the last tile publishes the descriptor.
struct TileState { int status; int agg[4]; int incl[4]; };
constexpr int kCompactTile = 1024;
__global__ void __launch_bounds__(256) level_compact_chained_kernel(const unsigned* nlen, LevelDesc* lv, LevelDesc* host_lv /* mapped host memory */, int level,
                                                                    unsigned split_node, int* active, int* klist_split, int* klist_big, int* klist_small,
                                                                    int* klist_tiny, TileState* tiles) {
    __shared__ int s_warp[256 / 32 + 1];
    __shared__ int s_tile, s_prefix[4];
    const int tid = threadIdx.x;
    const int epoch = level;
    int base, n;
    next_level_extent(lv[level - 1], base, n);
    if (tid == 0) s_tile = atomicAdd(&lv[level].pad[0], 1);
    __syncthreads();
    const int tile = s_tile, n_tiles = (n + kCompactTile - 1) / kCompactTile;
    if (n_tiles == 0) {  // nothing left to split: the (empty) level is published by whoever came first
        const int zero[4] = {0, 0, 0, 0};
        if (tile == 0 && tid == 0) publish_level(lv, host_lv, level, zero, base, 0);
        return;
    }
    if (tile >= n_tiles) return;
    int cls[4], c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tile * kCompactTile + tid * 4 + j;
        cls[j] = -1;
        if (i < n) {
            const unsigned len = nlen[base + i];
            cls[j] = len > split_node ? 0 : (len > kBigNode ? 1 : (len > kTinyNode ? 2 : (len > kMaxLeaf ? 3 : -1)));
        }
        c0 += cls[j] == 0; c1 += cls[j] == 1; c2 += cls[j] == 2; c3 += cls[j] == 3;
    }
    int t[4];
    int o0 = block_exclusive_scan<256>(c0, s_warp, t[0]);
    int o1 = block_exclusive_scan<256>(c1, s_warp, t[1]);
    int o2 = block_exclusive_scan<256>(c2, s_warp, t[2]);
    int o3 = block_exclusive_scan<256>(c3, s_warp, t[3]);
    if (tid == 0) {
        const int have_agg = 2 * epoch + 1, have_incl = 2 * epoch + 2;
        int pre[4] = {0, 0, 0, 0};
        if (tile > 0) {
            for (int c = 0; c < 4; ++c) tiles[tile].agg[c] = t[c];
            __threadfence();
            *reinterpret_cast<volatile int*>(&tiles[tile].status) = have_agg;
            for (int p = tile - 1; p >= 0; --p) {
                int st;
                do { st = *reinterpret_cast<volatile int*>(&tiles[p].status); } while (st < have_agg);
                __threadfence();
                if (st == have_incl) {
                    for (int c = 0; c < 4; ++c) pre[c] += __ldcg(&tiles[p].incl[c]);
                    break;
                }
                for (int c = 0; c < 4; ++c) pre[c] += __ldcg(&tiles[p].agg[c]);
            }
        }
        for (int c = 0; c < 4; ++c) { tiles[tile].incl[c] = pre[c] + t[c]; s_prefix[c] = pre[c]; }
        __threadfence();
        *reinterpret_cast<volatile int*>(&tiles[tile].status) = have_incl;
        if (tile == n_tiles - 1) {
            const int counts[4] = {pre[0] + t[0], pre[1] + t[1], pre[2] + t[2], pre[3] + t[3]};
            publish_level(lv, host_lv, level, counts, base, n);
        }
    }
    __syncthreads();
    o0 += s_prefix[0]; o1 += s_prefix[1]; o2 += s_prefix[2]; o3 += s_prefix[3];
    int k = o0 + o1 + o2 + o3;  // active nodes before this thread's first one, in level order
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (cls[j] < 0) continue;
        active[k] = base + tile * kCompactTile + tid * 4 + j;
        if (cls[j] == 0) klist_split[o0++] = k;
        else if (cls[j] == 1) klist_big[o1++] = k;
        else if (cls[j] == 2) klist_small[o2++] = k;
        else klist_tiny[o3++] = k;
        ++k;
    }
}

// ---------------------------------------------------------------------------------------------
// flatten
__global__ void subtree_size_kernel(BuildArrays a, int base, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = base + i, c = a.nchild[id];
    a.nsize[id] = c < 0 ? 1u : 1u + a.nsize[c] + a.nsize[c + 1];
}

// the same for a run of consecutive small levels d_hi, d_hi - 1, ..., d_lo in one CTA (levels[2d] = base, levels[2d+1] = count)
__global__ void __launch_bounds__(1024) subtree_size_run_kernel(BuildArrays a, const int* levels, int d_hi, int d_lo) {
    for (int d = d_hi; d >= d_lo; --d) {
        const int base = levels[2 * d], n = levels[2 * d + 1];
        for (int i = threadIdx.x; i < n; i += 1024) {
            const int id = base + i, c = a.nchild[id];
            a.nsize[id] = c < 0 ? 1u : 1u + a.nsize[c] + a.nsize[c + 1];
        }
        __syncthreads();
    }
}

// FlattenBVH (:783-845), one level per launch, top-down
__device__ __forceinline__ void flatten_stackless_node(const BuildArrays& a, const unsigned char* nflip, int id, float4* out) {
    if (id == 0) { a.npre[0] = 0; a.nlink[0] = -1; }
    const int pre = a.npre[id], link = a.nlink[id], c = a.nchild[id];
    const float4 mn = a.nmin[id], mx = a.nmax[id];
    int minw;
    if (c < 0) {
        minw = leaf_pack(a, id);
    } else {
        minw = -1;
        const int first = nflip[id] ? c + 1 : c, second = nflip[id] ? c : c + 1;
        a.npre[first] = pre + 1;
        a.npre[second] = pre + 1 + (int)a.nsize[first];
        a.nlink[first] = pre + 1 + (int)a.nsize[first];
        a.nlink[second] = link;
    }
    out[2 * (size_t)pre] = make_float4(mn.x, mn.y, mn.z, __int_as_float(minw));
    out[2 * (size_t)pre + 1] = make_float4(mx.x, mx.y, mx.z, __int_as_float(link));
}

__global__ void flatten_stackless_kernel(BuildArrays a, const unsigned char* nflip, int base, int n, float4* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flatten_stackless_node(a, nflip, base + i, out);
}

// a run of consecutive small levels d_lo .. d_hi, top-down, in one CTA
__global__ void __launch_bounds__(1024) flatten_stackless_run_kernel(BuildArrays a, const unsigned char* nflip, const int* levels, int d_lo, int d_hi,
                                                                     float4* out) {
    for (int d = d_lo; d <= d_hi; ++d) {
        const int base = levels[2 * d], n = levels[2 * d + 1];
        for (int i = threadIdx.x; i < n; i += 1024) flatten_stackless_node(a, nflip, base + i, out);
        __syncthreads();
    }
}

__global__ void inner_flags_kernel(BuildArrays a, int n, int* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = a.nchild[i] >= 0 ? 1 : 0;
}

// FlattenStackBVH (:847-930): inner nodes in breadth-first order == level order of ids
__global__ void flatten_stack_kernel(BuildArrays a, int n, const int* slot, float4* out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const int c = a.nchild[id];
    if (c < 0) return;
    float4* o = out + 4 * (size_t)slot[id];
    for (int side = 0; side < 2; ++side) {
        const int ch = c + side;
        const float4 mn = a.nmin[ch], mx = a.nmax[ch];
        const bool leaf = a.nchild[ch] < 0;
        o[2 * side] = make_float4(mn.x, mn.y, mn.z, __int_as_float(leaf ? leaf_pack(a, ch) : -1));
        o[2 * side + 1] = make_float4(mx.x, mx.y, mx.z, leaf ? 0.0f : __int_as_float(slot[ch]));
    }
}

}  // namespace

int build_lbvh_object(BuildRequest& rq, cudaStream_t st, LaunchCounter& lc, float* build_ms, std::string& err);

int build_object(BuildRequest& rq, cudaStream_t st, LaunchCounter& lc, float* build_ms, std::string& err) {
    if (rq.opts.builder == CNDL_BUILDER_LBVH) return build_lbvh_object(rq, st, lc, build_ms, err);
    if (rq.opts.builder != CNDL_BUILDER_SAH_EXACT) { err = "unknown builder"; return CNDL_ERR_INVALID; }
    const size_t T = rq.T;
    if (T == 0 || T > (1ull << 27)) { err = "triangle count out of range"; return CNDL_ERR_INVALID; }
    for (size_t i = 0; i < 3 * T; ++i)
        if (rq.h_indices[i] >= rq.V) { err = "vertex index out of range"; return CNDL_ERR_INVALID; }
    const size_t n_max = 2 * T - 1;
    const bool stackless = rq.format == CNDL_STACKLESS;

    BuildArrays a{};
    uint32_t* d_idx = nullptr;
    int32_t* d_mesh = nullptr;
    unsigned char* d_flip = nullptr;
    TileState* d_tiles = nullptr;
    LevelDesc* d_lv = nullptr;
    int *d_levels = nullptr, *d_flags = nullptr, *d_offsets = nullptr, *d_block_sums = nullptr, *d_totals = nullptr, *d_active = nullptr, *d_active_next = nullptr, *d_kl_big = nullptr, *d_kl_small = nullptr, *d_kl_tiny = nullptr, *d_kl_split = nullptr;
    const size_t scan_n = std::max<size_t>(n_max, 16);  // inner-node flags of the stack flatten (the level compaction scans in its own kernels)
    const unsigned split_node = rq.split_node ? std::max(rq.split_node, kTinyNode) : kSplitNodeDefault;
    const long long pack_min = rq.pack_min ? (long long)rq.pack_min : kPackMinRanges;
    const int split_items = T <= (1u << 20) ? 2 : 8, split_chunk = kSplitBlock * split_items;
    const size_t n_split_max = T / split_node + 2, split_chunks_max = T / split_chunk + n_split_max + 1;
    SplitArgs sp{};
    auto layout = [&](Scratch& sc) {
        sc.alloc(&d_idx, 3 * T);
        if (rq.h_mesh_ids) sc.alloc(&d_mesh, T);
        sc.alloc(&a.tmin, T); sc.alloc(&a.tmax, T); sc.alloc(&a.tcz, T); sc.alloc(&a.refs, T);
        sc.alloc(&a.nmin, n_max); sc.alloc(&a.nmax, n_max); sc.alloc(&a.nstart, n_max); sc.alloc(&a.nlen, n_max);
        sc.alloc(&a.nchild, n_max); sc.alloc(&a.nsize, n_max); sc.alloc(&a.npre, n_max); sc.alloc(&a.nlink, n_max);
        sc.alloc(&d_flip, n_max);
        sc.alloc(&a.root_scratch, 16);
        sc.alloc(&d_flags, scan_n); sc.alloc(&d_offsets, scan_n);
        sc.alloc(&d_block_sums, scan_n / kScanTile + 2); sc.alloc(&d_totals, 4);
        sc.alloc(&d_active, n_max); sc.alloc(&d_active_next, n_max); sc.alloc(&d_kl_big, n_max); sc.alloc(&d_kl_small, n_max); sc.alloc(&d_kl_tiny, n_max);
        sc.alloc(&d_kl_split, n_split_max); sc.alloc(&d_levels, 2 * (size_t)kMaxRunLevels);
        sc.alloc(&d_tiles, n_max / kCompactTile + 2); sc.alloc(&d_lv, (size_t)kMaxLevels);
        sc.alloc(&sp.chunk_base, n_split_max + 1); sc.alloc(&sp.bins, n_split_max * kBinInts); sc.alloc(&sp.done, n_split_max * 3);
        sc.alloc(&sp.split, n_split_max * 4); sc.alloc(&sp.boxes, n_split_max * 24); sc.alloc(&sp.chunk_l, split_chunks_max);
        sc.alloc(&sp.rankflag, T); sc.alloc(&sp.rpos, T); sc.alloc(&sp.alt, T);
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

    // the level descriptors' host copy: mapped pinned memory kept by the context, written by the compaction kernels
    if (!*rq.host_counts) BK(cudaHostAlloc(reinterpret_cast<void**>(rq.host_counts), (size_t)kMaxLevels * sizeof(LevelDesc), cudaHostAllocMapped | cudaHostAllocPortable));
    volatile LevelDesc* h_lv = reinterpret_cast<volatile LevelDesc*>(*rq.host_counts);
    LevelDesc* d_host_lv = nullptr;
    BK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_host_lv), *rq.host_counts, 0));

    cudaEvent_t ev0, ev1, ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    const bool side_ok = rq.side[0] && rq.side[1];
    BK(cudaEventCreate(&ev0));
    BK(cudaEventCreate(&ev1));
    if (side_ok) {
        BK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        BK(cudaEventCreateWithFlags(&ev_join[0], cudaEventDisableTiming));
        BK(cudaEventCreateWithFlags(&ev_join[1], cudaEventDisableTiming));
    }
    BK(cudaMemcpyAsync(d_idx, rq.h_indices, 3 * T * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (rq.h_mesh_ids) BK(cudaMemcpyAsync(d_mesh, rq.h_mesh_ids, T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    BK(cudaEventRecord(ev0, st));  // build time: geometry resident, like the reference's timer around BuildBVH (:953)

    a.verts = rq.d_verts;
    a.indices = d_idx;
    a.mesh_ids = d_mesh;
    a.T = (unsigned)T;
    a.tris_out = rq.d_tris_out;
    a.tri_offset = rq.tri_offset;

    const int h_root_init[12] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000, -1, -1, -1, -1, -1, -1};
    BK(cudaMemcpyAsync(a.root_scratch, h_root_init, sizeof(h_root_init), cudaMemcpyHostToDevice, st));
    BK(cudaMemsetAsync(d_flip, 0, n_max, st));
    BK(cudaMemsetAsync(d_tiles, 0, (n_max / kCompactTile + 2) * sizeof(TileState), st));
    BK(cudaMemsetAsync(d_lv, 0, (size_t)kMaxLevels * sizeof(LevelDesc), st));  // also the per-level tickets of the chained compaction
    tri_precompute_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(a);
    root_finalize_kernel<<<1, 32, 0, st>>>(a);
    lc.n += 2;

    std::vector<int> level_base{0}, level_count{1};
    size_t n_nodes = 1;
    float4* out = static_cast<float4*>(rq.d_nodes_out);

    if (T <= kMaxLeaf) {
        single_leaf_kernel<<<1, 32, 0, st>>>(a, stackless ? 1 : 0, out);
        lc.n++;
    } else {
        // Level-synchronous build WITHOUT a host round trip per level.  The device keeps a descriptor per level (class sizes,
        // node range); the host launches each level's kernels with grid UPPER BOUNDS derived from the last counts it has seen
        // (a child is never in a larger size class than its parent, so a class at most doubles per level and is capped by
        // T / its minimum length) and reads the descriptors back only every kSyncEvery levels, to tighten the bounds and to
        // notice that nothing is left to split.  Levels enqueued past the end find empty classes and return at once.
        constexpr int kSyncEvery = 4;
        LevelDesc l0{};
        l0.n_class[CLS_SPLIT] = T > split_node ? 1 : 0;
        l0.n_class[CLS_BIG] = !l0.n_class[CLS_SPLIT] && T > kBigNode ? 1 : 0;
        l0.n_class[CLS_TINY] = T <= kTinyNode ? 1 : 0;
        l0.n_class[CLS_SMALL] = 1 - l0.n_class[CLS_SPLIT] - l0.n_class[CLS_BIG] - l0.n_class[CLS_TINY];
        l0.base = 0;
        l0.n = 1;
        BK(cudaMemcpyAsync(d_lv, &l0, sizeof(l0), cudaMemcpyHostToDevice, st));
        const_cast<LevelDesc&>(h_lv[0]) = l0;
        BK(cudaMemsetAsync(d_active, 0, sizeof(int), st));
        BK(cudaMemsetAsync(d_kl_split, 0, sizeof(int), st));
        BK(cudaMemsetAsync(d_kl_big, 0, sizeof(int), st));
        BK(cudaMemsetAsync(d_kl_small, 0, sizeof(int), st));
        BK(cudaMemsetAsync(d_kl_tiny, 0, sizeof(int), st));
        const long long cap[4] = {(long long)(T / split_node + 1), (long long)(T / kBigNode + 1), (long long)(T / kTinyNode + 1), (long long)(T / (kMaxLeaf + 1) + 1)};
        long long ub[4] = {l0.n_class[0], l0.n_class[1], l0.n_class[2], l0.n_class[3]};  // upper bounds of the current level's class sizes
        int level = 0, known = 0;
        while (ub[0] + ub[1] + ub[2] + ub[3] > 0) {
            if (level + 2 >= kMaxLevels) { err = "tree deeper than 65536 levels"; return CNDL_ERR_INVALID; }
            const int n_split = (int)ub[CLS_SPLIT], n_big = (int)ub[CLS_BIG], n_small = (int)ub[CLS_SMALL], n_tiny = (int)ub[CLS_TINY];
            LevelArgs g;
            g.a = a;
            g.active = d_active;
            g.lv = d_lv + level;
            g.stackless = stackless ? 1 : 0;
            g.swap_policy = rq.opts.swap_policy;
            g.swap_seed = rq.opts.swap_seed;
            g.nflip = d_flip;
            if (n_split) {
                g.klist = d_kl_split;
                sp.g = g;
                // every split node is longer than split_node, so its chunks number at most T / chunk + one partial chunk per node
                const unsigned grid = (unsigned)(T / split_chunk + (size_t)n_split);
                split_prep_kernel<<<1, 1024, 0, st>>>(sp, split_chunk);
                if (split_items == 2) {
                    split_bin_kernel<2><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_count_kernel<2><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_rank_kernel<2><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_gather_kernel<2><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_finish_kernel<2><<<grid, kSplitBlock, 0, st>>>(sp);
                } else {
                    split_bin_kernel<8><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_count_kernel<8><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_rank_kernel<8><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_gather_kernel<8><<<grid, kSplitBlock, 0, st>>>(sp);
                    split_finish_kernel<8><<<grid, kSplitBlock, 0, st>>>(sp);
                }
                lc.n += 6;
            }
            // the size classes of one level touch disjoint nodes and ranges: they run side by side on up to three streams
            const bool st_busy = n_split || n_tiny;  // those two stay on st
            cudaStream_t s_big = side_ok && n_big && st_busy ? rq.side[0] : st;
            cudaStream_t s_small = side_ok && n_small && (st_busy || n_big) ? rq.side[1] : st;
            const bool fork = s_big != st || s_small != st;
            if (fork) {
                BK(cudaEventRecord(ev_fork, st));
                if (s_big != st) BK(cudaStreamWaitEvent(s_big, ev_fork, 0));
                if (s_small != st) BK(cudaStreamWaitEvent(s_small, ev_fork, 0));
            }
            if (n_big) { g.klist = d_kl_big; level_step_kernel<1024><<<n_big, 1024, 0, s_big>>>(g, CLS_BIG); lc.n++; }
            if (n_small) { g.klist = d_kl_small; level_step_kernel<128><<<n_small, 128, 0, s_small>>>(g, CLS_SMALL); lc.n++; }
            if (n_tiny) {
                g.klist = d_kl_tiny;
                if (n_tiny >= pack_min) {  // four ranges per warp
                    level_step_tiny_kernel<4><<<(unsigned)((n_tiny + 4 * kTinyWarps - 1) / (4 * kTinyWarps)), 32 * kTinyWarps, 0, st>>>(g);
                } else {
                    level_step_tiny_kernel<1><<<(unsigned)((n_tiny + kTinyWarps - 1) / kTinyWarps), 32 * kTinyWarps, 0, st>>>(g);
                }
                lc.n++;
            }
            BK(cudaGetLastError());
            if (fork) {
                if (s_big != st) { BK(cudaEventRecord(ev_join[0], s_big)); BK(cudaStreamWaitEvent(st, ev_join[0], 0)); }
                if (s_small != st) { BK(cudaEventRecord(ev_join[1], s_small)); BK(cudaStreamWaitEvent(st, ev_join[1], 0)); }
            }
            // next level's active list and descriptor
            const long long n_next_ub = 2 * (ub[0] + ub[1] + ub[2] + ub[3]);
            ++level;
            if (n_next_ub <= kSmallLevel) {
                level_compact_small_kernel<<<1, 1024, 0, st>>>(a.nlen, d_lv, d_host_lv, level, split_node, d_active_next, d_kl_split, d_kl_big, d_kl_small,
                                                               d_kl_tiny);
            } else {
                level_compact_chained_kernel<<<(unsigned)((n_next_ub + kCompactTile - 1) / kCompactTile), 256, 0, st>>>(
                    a.nlen, d_lv, d_host_lv, level, split_node, d_active_next, d_kl_split, d_kl_big, d_kl_small, d_kl_tiny, d_tiles);
            }
            lc.n++;
            std::swap(d_active, d_active_next);
            const long long all = ub[0] + ub[1] + ub[2] + ub[3];
            const long long nb[4] = {2 * ub[0], 2 * (ub[0] + ub[1]), 2 * (ub[0] + ub[1] + ub[2]), 2 * all};
            for (int c = 0; c < 4; ++c) ub[c] = std::min(nb[c], cap[c]);
            // Bounds that have drifted far above anything plausible would launch grids of mostly empty CTAs (10 M triangles: the build
            // doubled): once the bounded grids of the next level exceed a couple of thousand CTAs, read the real counts every level —
            // a level that large hides the round trip anyway.
            const long long bounded_ctas = ub[CLS_BIG] + ub[CLS_SMALL] + (ub[CLS_TINY] + kTinyWarps - 1) / kTinyWarps + (ub[CLS_SPLIT] ? (long long)(T / split_chunk) : 0);
            if (level - known >= kSyncEvery || bounded_ctas > 2048) {
                BK(cudaStreamSynchronize(st));
                for (int c = 0; c < 4; ++c) ub[c] = h_lv[level].n_class[c];  // exact
                known = level;
            }
        }
        BK(cudaStreamSynchronize(st));
        for (int l = 0; l <= level && l < kMaxLevels; ++l) {
            const int n_l = h_lv[l].n;
            if (l > 0 && n_l == 0) break;
            const int base_l = h_lv[l].base;
            if (l > 0) { level_base.push_back(base_l); level_count.push_back(n_l); }
            n_nodes = (size_t)base_l + (size_t)n_l;
        }
        // flatten.  Runs of consecutive levels of at most kSmallLevel nodes (the top and the bottom of the tree) take one
        // single-CTA launch each instead of one launch per level.
        const int D = (int)level_base.size();
        const bool runs = D <= kMaxRunLevels;
        if (runs) {
            std::vector<int> h_levels(2 * (size_t)D);
            for (int d = 0; d < D; ++d) { h_levels[2 * d] = level_base[d]; h_levels[2 * d + 1] = level_count[d]; }
            BK(cudaMemcpyAsync(d_levels, h_levels.data(), h_levels.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            BK(cudaStreamSynchronize(st));  // h_levels goes out of scope
        }
        auto small = [&](int d) { return runs && level_count[d] <= kSmallLevel; };
        for (int d = D - 1; d >= 0;) {
            int e = d;
            while (small(d) && e > 0 && small(e - 1)) --e;
            if (e < d) subtree_size_run_kernel<<<1, 1024, 0, st>>>(a, d_levels, d, e);
            else subtree_size_kernel<<<(level_count[d] + 255) / 256, 256, 0, st>>>(a, level_base[d], level_count[d]);
            lc.n++;
            d = e - 1;
        }
        if (stackless) {
            for (int d = 0; d < D;) {
                int e = d;
                while (small(d) && e + 1 < D && small(e + 1)) ++e;
                if (e > d) flatten_stackless_run_kernel<<<1, 1024, 0, st>>>(a, d_flip, d_levels, d, e, out);
                else flatten_stackless_kernel<<<(level_count[d] + 255) / 256, 256, 0, st>>>(a, d_flip, level_base[d], level_count[d], out);
                lc.n++;
                d = e + 1;
            }
        } else {
            const int n = (int)n_nodes;
            BK(cudaMemsetAsync(out, 0, n_nodes * sizeof(cndl_stack_node), st));  // unused slots stay zero (:762-765)
            inner_flags_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, n, d_flags);
            exclusive_scan(d_flags, n, d_offsets, d_block_sums, d_totals, st, lc);
            flatten_stack_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, n, d_offsets, out);
            lc.n += 2;
        }
        BK(cudaGetLastError());
    }
    BK(cudaEventRecord(ev1, st));
    BK(cudaStreamSynchronize(st));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    if (build_ms) *build_ms = ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (side_ok) { cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join[0]); cudaEventDestroy(ev_join[1]); }
    rq.n_nodes_out = n_nodes;
    return CNDL_OK;
}

}  // namespace cndl
