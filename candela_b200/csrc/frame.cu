// Frame-level and multi-device calls of the C ABI (include/candela_b200.h).
//
// In the reference the rays of a frame never exist on the host: DiffuseTrace.glsl:437-446 forms them in the shader from the
// G-buffer and traces them on the spot (:484, :516-518).  cndl_trace_frame_device is that frame as one enqueued sequence of
// kernels — camera rays of this shard's tiles, closest hit, diffuse generation (kernels_raygen.cu), traversal with the batch
// length read on the device, resolve — so nothing but a parameter block goes in and compact records come out; with several
// GPUs each traces the tiles it was dealt and its resolve kernel stores the records straight into the first
// device's frame through NVLink peer memory.
#include "context.cuh"
#include "exact_trig.cuh"

#include <chrono>

using namespace cndl;

namespace {

constexpr int kMaxBounces = 8;

// Tiles of T x T pixels, t = ty * tiles_x + tx, dealt round-robin: shard s owns the tiles with t % shards == s, in
// ascending order (local tile l is tile l * shards + s).  Slots: local_tile * T * T + y_in_tile * T + x_in_tile; slots of
// edge tiles outside the image are padding.
struct TileMap {
    int W, H, T, tiles_x, tiles_y, n_tiles, shard, shards, local_tiles;
};

__host__ __device__ inline TileMap make_tile_map(int W, int H, int T, int shard, int shards) {
    TileMap m;
    m.W = W; m.H = H; m.T = T;
    m.tiles_x = (W + T - 1) / T;
    m.tiles_y = (H + T - 1) / T;
    m.n_tiles = m.tiles_x * m.tiles_y;
    m.shard = shard; m.shards = shards;
    m.local_tiles = m.n_tiles > shard ? (m.n_tiles - shard + shards - 1) / shards : 0;
    return m;
}
__device__ __forceinline__ bool slot_to_pixel(const TileMap& m, unsigned slot, int& x, int& y) {
    const unsigned tt = (unsigned)(m.T * m.T);
    const unsigned l = slot / tt, r = slot - l * tt;
    const int tile = (int)l * m.shards + m.shard;
    x = (tile % m.tiles_x) * m.T + (int)(r % (unsigned)m.T);
    y = (tile / m.tiles_x) * m.T + (int)(r / (unsigned)m.T);
    return x < m.W && y < m.H;
}
__device__ __forceinline__ unsigned pixel_to_slot(const TileMap& m, unsigned pixel) {
    const int x = (int)(pixel % (unsigned)m.W), y = (int)(pixel / (unsigned)m.W);
    const int tile = (y / m.T) * m.tiles_x + x / m.T;
    return (unsigned)(tile / m.shards) * (unsigned)(m.T * m.T) + (unsigned)((y % m.T) * m.T + x % m.T);
}

// Camera rays of the shard's slots (the arithmetic of cndl_intersect_primary, kernels.cuh primary_ray).  Padding slots get
// a ray that starts far outside every scene and points away, so it leaves the root box test at once and reports a miss.
__global__ void frame_primary_kernel(Mat2 m, TileMap tm, unsigned P, cndl_ray* __restrict__ rays, unsigned* __restrict__ pix_ids) {
    const unsigned slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P) return;
    int x, y;
    float4 o, d;
    unsigned id = 0xFFFFFFFFu;
    if (slot_to_pixel(tm, slot, x, y)) {
        primary_ray(m, x, y, tm.W, tm.H, o, d);
        id = (unsigned)y * (unsigned)tm.W + (unsigned)x;
    } else {
        o = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.0f);
        d = make_float4(0.57735026f, 0.57735026f, 0.57735026f, 1000000.0f);
    }
    float4* p = reinterpret_cast<float4*>(rays + slot);
    p[0] = o;
    p[1] = d;
    pix_ids[slot] = id;
}

// Output of the hit formats (bounces == 1): element e = slot * spp + sample holds the record of the diffuse ray the
// generator wrote at dest[e] (keys[e] < 8), or a miss when that sample emitted no ray.  One thread per 16 bytes of output, so
// that a warp's stores are contiguous (they may cross NVLink into a peer's frame).
template <bool COMPACT>
__global__ void frame_resolve_hits_kernel(unsigned n_elems, int spp, const unsigned* __restrict__ pix_ids, const unsigned char* __restrict__ keys,
                                          const unsigned* __restrict__ dest, const cndl_hit* __restrict__ hits, float4* __restrict__ out, int local_layout) {
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned e = COMPACT ? q : q >> 1, half = COMPACT ? 0u : q & 1u;
    if (e >= n_elems) return;
    const unsigned slot = e / (unsigned)spp, s = e - slot * (unsigned)spp;
    const unsigned pixel = __ldg(pix_ids + slot);
    if (pixel == 0xFFFFFFFFu && !local_layout) return;
    const bool has = keys[e] < 8;
    const unsigned k = has ? __ldg(dest + e) : 0u;
    const size_t idx = local_layout ? (size_t)e : (size_t)pixel * (size_t)spp + s;
    if (COMPACT) {
        float4 r = make_float4(-1.0f, __int_as_float(-1), -1.0f, -1.0f);
        if (has) {
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(hits + k));
            r = make_float4(h0.x, __int_as_float(__ldg(&hits[k].tri)), h0.z, h0.w);
        }
        out[idx] = r;
    } else {
        float4 r = half ? make_float4(__int_as_float(-1), __int_as_float(-1), __int_as_float(-1), __int_as_float(0)) : make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
        if (has) r = __ldg(reinterpret_cast<const float4*>(hits + k) + half);
        out[2 * idx + half] = r;
    }
}

// PIXEL32, first bounce: the camera hit, the AO term of DiffuseTrace.glsl:494 averaged over the pixel's samples in sample
// order, and the ray / escape counts of bounce 0.
__global__ void frame_first_bounce_kernel(unsigned P, int spp, const unsigned char* __restrict__ keys, const unsigned* __restrict__ dest,
                                          const cndl_hit* __restrict__ prim_hits, const cndl_hit* __restrict__ hits, cndl_pixel* __restrict__ acc) {
    const unsigned slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P) return;
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(prim_hits + slot));
    const int4 p1 = __ldg(reinterpret_cast<const int4*>(prim_hits + slot) + 1);
    float ao_sum = 0.0f, t_sum = 0.0f;
    int n = 0, n_hit = 0;
    for (int s = 0; s < spp; ++s) {
        const size_t e = (size_t)slot * spp + s;
        if (keys[e] >= 8) continue;
        ++n;
        const float t1 = __ldg(&hits[__ldg(dest + e)].t);
        if (t1 > 0.0f) {
            const float c = glsl_min(glsl_max(fdiv(t1, 2.4f), 0.0f), 1.0f);  // clamp(TUVW.x / 2.4f, 0.0f, 1.0f)
            ao_sum = fadd(ao_sum, xm::xpow(c, 1.23f));
            t_sum = fadd(t_sum, t1);
            ++n_hit;
        } else {
            ao_sum = fadd(ao_sum, 1.0f);
        }
    }
    cndl_pixel px;
    px.t = p0.x; px.tri = p1.y; px.v = p0.z; px.w = p0.w;
    px.ao = n ? fdiv(ao_sum, (float)n) : 1.0f;
    px.t_mean = n_hit ? fdiv(t_sum, (float)n_hit) : -1.0f;
    px.rays = n;
    px.escaped = n - n_hit;
    float4* q = reinterpret_cast<float4*>(acc + slot);
    q[0] = make_float4(px.t, __int_as_float(px.tri), px.v, px.w);
    q[1] = make_float4(px.ao, px.t_mean, __int_as_float(px.rays), __int_as_float(px.escaped));
}

// PIXEL32, bounce b >= 1: integer counts only (order-independent, so the result does not depend on scheduling).  The batch is
// either compact (d_count rays) or segmented (seg_counts).
__global__ void frame_accumulate_kernel(TileMap tm, int spp, unsigned cap, const unsigned* __restrict__ d_count, const unsigned* __restrict__ seg_counts,
                                        const unsigned* __restrict__ rids, const cndl_hit* __restrict__ hits, cndl_pixel* __restrict__ acc) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg_counts) {
        if (k >= cap || (k & (unsigned)(kRaySegment - 1)) >= __ldg(seg_counts + (k >> kRaySegmentShift))) return;
    } else if (k >= min(cap, __ldg(d_count))) {
        return;
    }
    const unsigned slot = pixel_to_slot(tm, __ldg(rids + k) / (unsigned)spp);
    atomicAdd(&acc[slot].rays, 1);
    if (!(__ldg(&hits[k].t) > 0.0f)) atomicAdd(&acc[slot].escaped, 1);
}

// one thread per 16 bytes of output: contiguous stores (possibly into a peer's frame over NVLink)
__global__ void frame_write_pixels_kernel(unsigned P, const unsigned* __restrict__ pix_ids, const float4* __restrict__ acc, float4* __restrict__ out,
                                          int local_layout) {
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned slot = q >> 1, half = q & 1u;
    if (slot >= P) return;
    const unsigned pixel = __ldg(pix_ids + slot);
    if (pixel == 0xFFFFFFFFu && !local_layout) return;
    out[2 * (local_layout ? (size_t)slot : (size_t)pixel) + half] = __ldg(acc + 2 * (size_t)slot + half);
}

__global__ void frame_count_kernel(const unsigned* __restrict__ src, unsigned cap, unsigned* __restrict__ dst) { *dst = min(*src, cap); }

// shard in local layout -> row-major frame
__global__ void frame_untile_kernel(TileMap tm, unsigned n_elems, int spp, int rec16, const float4* __restrict__ shard, float4* __restrict__ frame) {
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const unsigned slot = e / (unsigned)spp, s = e - slot * (unsigned)spp;
    int x, y;
    if (!slot_to_pixel(tm, slot, x, y)) return;
    const size_t idx = ((size_t)y * tm.W + x) * spp + s;
    if (rec16 == 1) frame[idx] = __ldg(shard + e);
    else { frame[2 * idx] = __ldg(shard + 2 * (size_t)e); frame[2 * idx + 1] = __ldg(shard + 2 * (size_t)e + 1); }
}

int check_params(cndl_ctx* ctx, const cndl_frame_params* p, TileMap& tm) {
    if (!p) return ctx->fail(CNDL_ERR_INVALID, "null frame parameters");
    const int T = p->tile > 0 ? p->tile : 64;
    const int shards = p->shard_count > 0 ? p->shard_count : 1;
    if (p->width <= 0 || p->height <= 0 || p->spp < 1 || p->spp > 64 || p->bounces < 1 || p->bounces > kMaxBounces || T > 1024 || p->shard_index < 0 ||
        p->shard_index >= shards || p->out_format < CNDL_FRAME_OUT_HIT32 || p->out_format > CNDL_FRAME_OUT_PIXEL32)
        return ctx->fail(CNDL_ERR_INVALID, "bad frame parameters");
    if (p->out_format != CNDL_FRAME_OUT_PIXEL32 && p->bounces != 1)
        return ctx->fail(CNDL_ERR_INVALID, "the hit formats hold the first diffuse hit of every sample: bounces must be 1 (use CNDL_FRAME_OUT_PIXEL32)");
    if ((size_t)p->width * (size_t)p->height * (size_t)p->spp > 0x7FFFFFF0ull) return ctx->fail(CNDL_ERR_INVALID, "frame too large for 32-bit ray ids");
    tm = make_tile_map(p->width, p->height, T, p->shard_index, shards);
    return CNDL_OK;
}

size_t record_bytes(int fmt) { return fmt == CNDL_FRAME_OUT_HIT16 ? sizeof(cndl_hit16) : 32; }

int ensure_frame_streams(cndl_ctx* ctx) {
    for (auto& fs : ctx->frame_stream)
        if (!fs) CK(cudaStreamCreateWithFlags(&fs, cudaStreamNonBlocking));
    if (!ctx->frame_copy_stream) CK(cudaStreamCreateWithFlags(&ctx->frame_copy_stream, cudaStreamNonBlocking));
    for (auto& f : ctx->frame) {
        if (!f.traced) CK(cudaEventCreateWithFlags(&f.traced, cudaEventDisableTiming));
        if (!f.copied) CK(cudaEventCreateWithFlags(&f.copied, cudaEventDisableTiming));
        if (!f.h_counts) CK(cudaMallocHost(reinterpret_cast<void**>(&f.h_counts), 16 * sizeof(unsigned)));
    }
    return CNDL_OK;
}

}  // namespace

extern "C" {

size_t cndl_frame_record_bytes(int out_format) { return record_bytes(out_format); }

size_t cndl_frame_records(const cndl_frame_params* p) {
    if (!p || p->width <= 0 || p->height <= 0) return 0;
    return (size_t)p->width * (size_t)p->height * (size_t)(p->out_format == CNDL_FRAME_OUT_PIXEL32 ? 1 : (p->spp > 0 ? p->spp : 1));
}

size_t cndl_frame_shard_records(const cndl_frame_params* p) {
    if (!p || p->width <= 0 || p->height <= 0) return 0;
    const int T = p->tile > 0 ? p->tile : 64, shards = p->shard_count > 0 ? p->shard_count : 1;
    if (p->shard_index < 0 || p->shard_index >= shards) return 0;
    const TileMap tm = make_tile_map(p->width, p->height, T, p->shard_index, shards);
    return (size_t)tm.local_tiles * (size_t)T * (size_t)T * (size_t)(p->out_format == CNDL_FRAME_OUT_PIXEL32 ? 1 : (p->spp > 0 ? p->spp : 1));
}

int cndl_trace_frame_device(cndl_ctx* ctx, const cndl_frame_params* p, void* d_out, int slot, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    TileMap tm;
    int rc = check_params(ctx, p, tm);
    if (rc != CNDL_OK) return rc;
    if (!d_out || slot < 0 || slot >= kFrameSlots) return ctx->fail(CNDL_ERR_INVALID, "null output or bad slot");
    rc = check_ready(ctx);
    if (rc != CNDL_OK) return rc;
    if (ctx->mode != 2) return ctx->fail(CNDL_ERR_INVALID, "frame-level calls need traversal mode 2");
    CK(cudaSetDevice(ctx->device));
    rc = ensure_frame_streams(ctx);
    if (rc != CNDL_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FrameSlot& f = ctx->frame[slot];
    const size_t P = (size_t)tm.local_tiles * tm.T * tm.T;
    f.n_counts = 0;
    f.rays_traced = 0;
    if (P == 0) return CNDL_OK;
    const int spp = p->spp;
    const size_t cap = P * (size_t)spp;
    const bool pixels = p->out_format == CNDL_FRAME_OUT_PIXEL32;
    const int local = (p->flags & CNDL_FRAME_LOCAL_LAYOUT) ? 1 : 0;
    CK(f.prim_rays.ensure_scratch(P * sizeof(cndl_ray)));
    CK(f.prim_hits.ensure_scratch(P * sizeof(cndl_hit)));
    CK(f.pix_ids.ensure_scratch(P * sizeof(unsigned)));
    CK(f.rays[0].ensure_scratch(cap * sizeof(cndl_ray)));
    CK(f.hits.ensure_scratch(cap * sizeof(cndl_hit)));
    CK(f.rids[0].ensure_scratch(cap * sizeof(unsigned)));
    CK(f.gen_scratch.ensure_scratch(generate_rays_scratch_ints(P, spp) * sizeof(int)));  // bounce 0: P inputs x spp; later: cap inputs x 1
    const bool tiled = !(p->flags & CNDL_FRAME_COMPACT_RAYS);
    const size_t n_seg = (cap + kRaySegment - 1) / kRaySegment;
    // counts: [0..kMaxBounces) rays per bounce; [16 + 8 b ..) the octant cursors of bounce b; then two arrays of per-segment live counts
    constexpr size_t kCountHeader = 16 + 8 * kMaxBounces;
    CK(f.counts.ensure_scratch((kCountHeader + 2 * n_seg) * sizeof(unsigned)));
    const bool oct_lists = tiled && (p->flags & CNDL_FRAME_OCTANT_ORDER);
    if (oct_lists) CK(f.oct_list.ensure_scratch(8 * cap * sizeof(unsigned)));
    if (p->bounces > 1) {
        CK(f.rays[1].ensure_scratch(cap * sizeof(cndl_ray)));
        CK(f.rids[1].ensure_scratch(cap * sizeof(unsigned)));
    }
    if (pixels) CK(f.acc.ensure_scratch(P * sizeof(cndl_pixel)));

    Mat2 m;
    std::memcpy(m.iv, p->inv_view, 64);
    std::memcpy(m.ip, p->inv_proj, 64);
    cndl_ray* prim = static_cast<cndl_ray*>(f.prim_rays.p);
    cndl_hit* prim_hits = static_cast<cndl_hit*>(f.prim_hits.p);
    unsigned* pix_ids = static_cast<unsigned*>(f.pix_ids.p);
    unsigned* counts = static_cast<unsigned*>(f.counts.p);
    unsigned* seg[2] = {counts + kCountHeader, counts + kCountHeader + n_seg};
    const unsigned g256 = (unsigned)((P + 255) / 256);
    if (tiled) CK(cudaMemsetAsync(counts, 0, kCountHeader * sizeof(unsigned), st));
    frame_primary_kernel<<<g256, 256, 0, st>>>(m, tm, (unsigned)P, prim, pix_ids);
    ctx->launches.n++;
    rc = enqueue_trace(ctx, Q_CLOSEST, prim, P, nullptr, prim_hits, nullptr, next_counter(ctx), nullptr, nullptr, st);
    if (rc != CNDL_OK) return rc;

    const SceneView sv = scene_view(ctx);
    cndl_raygen_params g;
    std::memset(&g, 0, sizeof(g));
    g.kind = CNDL_GEN_DIFFUSE;
    g.flags = (p->flags & CNDL_FRAME_OCTANT_ORDER) ? CNDL_GEN_BUCKET_OCTANTS : 0;
    g.tmax = 1000000.0f;
    const cndl_ray* src_rays = prim;
    const cndl_hit* src_hits = prim_hits;
    const unsigned* src_count = nullptr;
    size_t src_cap = P;
    cndl_hit* hits = static_cast<cndl_hit*>(f.hits.p);
    for (int b = 0; b < p->bounces; ++b) {
        cndl_ray* out_rays = static_cast<cndl_ray*>(f.rays[b & 1].p);
        unsigned* out_ids = static_cast<unsigned*>(f.rids[b & 1].p);
        g.spp = b == 0 ? spp : 1;
        g.seed = p->seed + (uint32_t)b;
        g.offset = b == 0 ? 0.05f : 0.02f;                                   // DiffuseTrace.glsl:445, :516
        g.d_ids_in = b == 0 ? pix_ids : static_cast<const unsigned*>(f.rids[(b - 1) & 1].p);
        g.d_ids_out = out_ids;
        const int kind = b == 0 ? Q_CLOSEST_IGNORE_TRANSPARENT : Q_CLOSEST;         // :484 IntersectRayIgnoreTransparent, :518 IntersectRay
        if (tiled) {
            // one pass: rays ordered inside 1024-slot segments, dead slots at each segment's end (counts[b] accumulates the live ones)
            unsigned* cursor = counts + 16 + 8 * b;
            unsigned* list = oct_lists ? static_cast<unsigned*>(f.oct_list.p) : nullptr;
            CK(generate_rays_tiled(sv, g, src_rays, src_hits, src_cap, b == 0 ? nullptr : seg[(b - 1) & 1], out_rays, static_cast<int*>(f.gen_scratch.p),
                                   seg[b & 1], counts + b, cursor, list, cap, st, ctx->launches));
            // octant order: the launch walks the eight lists one after the other (all resident warps in one octant at a time); otherwise the
            // segments in place
            const RayOrder ro = oct_lists ? RayOrder{list, cursor, (unsigned)cap, counts + b, nullptr, nullptr} : RayOrder{nullptr, nullptr, 0, nullptr, nullptr, seg[b & 1]};
            rc = enqueue_trace(ctx, kind, out_rays, cap, nullptr, hits, nullptr, next_counter(ctx), nullptr, nullptr, st, &ro);
        } else {
            const unsigned* d_count = nullptr;
            CK(generate_rays(sv, g, src_rays, src_hits, src_cap, src_count, out_rays, nullptr, static_cast<int*>(f.gen_scratch.p), &d_count, nullptr, st,
                             ctx->launches));
            frame_count_kernel<<<1, 1, 0, st>>>(d_count, (unsigned)cap, counts + b);  // the scratch total is overwritten by the next bounce
            ctx->launches.n++;
            rc = enqueue_trace(ctx, kind, out_rays, cap, counts + b, hits, nullptr, next_counter(ctx), nullptr, nullptr, st);
        }
        if (rc != CNDL_OK) return rc;
        if (b == 0) {
            const unsigned char* keys;
            const unsigned* dest;
            generate_rays_maps(static_cast<int*>(f.gen_scratch.p), P, spp, &keys, &dest);
            if (pixels) {
                frame_first_bounce_kernel<<<g256, 256, 0, st>>>((unsigned)P, spp, keys, dest, prim_hits, hits, static_cast<cndl_pixel*>(f.acc.p));
            } else {
                if (p->out_format == CNDL_FRAME_OUT_HIT16)
                    frame_resolve_hits_kernel<true><<<(unsigned)((cap + 255) / 256), 256, 0, st>>>((unsigned)cap, spp, pix_ids, keys, dest, hits,
                                                                                                  static_cast<float4*>(d_out), local);
                else
                    frame_resolve_hits_kernel<false><<<(unsigned)((2 * cap + 255) / 256), 256, 0, st>>>((unsigned)cap, spp, pix_ids, keys, dest, hits,
                                                                                                       static_cast<float4*>(d_out), local);
            }
            ctx->launches.n++;
        } else {
            frame_accumulate_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(tm, spp, (unsigned)cap, counts + b, tiled ? seg[b & 1] : nullptr, out_ids, hits,
                                                                                 static_cast<cndl_pixel*>(f.acc.p));
            ctx->launches.n++;
        }
        src_rays = out_rays;
        src_hits = hits;
        src_count = counts + b;
        src_cap = cap;
    }
    if (pixels) {
        frame_write_pixels_kernel<<<(unsigned)((2 * P + 255) / 256), 256, 0, st>>>((unsigned)P, pix_ids, static_cast<const float4*>(f.acc.p),
                                                                                  static_cast<float4*>(d_out), local);
        ctx->launches.n++;
    }
    f.n_counts = p->bounces;
    CK(cudaMemcpyAsync(f.h_counts, counts, sizeof(unsigned) * (size_t)p->bounces, cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    return CNDL_OK;
} CNDL_CATCH

uint64_t cndl_frame_rays_traced(const cndl_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= kFrameSlots) return 0;
    const FrameSlot& f = ctx->frame[slot];
    uint64_t n = 0;
    for (int b = 0; b < f.n_counts; ++b) n += f.h_counts[b];
    return n;
}

int cndl_frame_submit(cndl_ctx* ctx, const cndl_frame_params* p, void* host_out, int slot) try {
    if (!ctx) return CNDL_ERR_INVALID;
    TileMap tm;
    int rc = check_params(ctx, p, tm);
    if (rc != CNDL_OK) return rc;
    if (!host_out || slot < 0 || slot >= kFrameSlots) return ctx->fail(CNDL_ERR_INVALID, "null output or bad slot");
    CK(cudaSetDevice(ctx->device));
    rc = ensure_frame_streams(ctx);
    if (rc != CNDL_OK) return rc;
    FrameSlot& f = ctx->frame[slot];
    if (f.pending) {
        rc = cndl_frame_wait(ctx, slot);
        if (rc != CNDL_OK) return rc;
    }
    const size_t records = (p->flags & CNDL_FRAME_LOCAL_LAYOUT) ? cndl_frame_shard_records(p) : cndl_frame_records(p);
    const size_t bytes = records * record_bytes(p->out_format);
    CK(f.out.ensure_scratch(bytes ? bytes : 16));
    const int shards = p->shard_count > 0 ? p->shard_count : 1;
    if (shards > 1 && !(p->flags & CNDL_FRAME_LOCAL_LAYOUT))  // pixels of other shards are not written: give them a defined value
        CK(cudaMemsetAsync(f.out.p, 0xFF, bytes, ctx->frame_stream[slot]));
    rc = cndl_trace_frame_device(ctx, p, f.out.p, slot, ctx->frame_stream[slot]);
    if (rc != CNDL_OK) return rc;
    CK(cudaEventRecord(f.traced, ctx->frame_stream[slot]));
    CK(cudaStreamWaitEvent(ctx->frame_copy_stream, f.traced, 0));
    CK(cudaMemcpyAsync(host_out, f.out.p, bytes, cudaMemcpyDeviceToHost, ctx->frame_copy_stream));
    CK(cudaEventRecord(f.copied, ctx->frame_copy_stream));
    f.pending = true;
    return CNDL_OK;
} CNDL_CATCH

int cndl_frame_wait(cndl_ctx* ctx, int slot) try {
    if (!ctx || slot < 0 || slot >= kFrameSlots) return CNDL_ERR_INVALID;
    FrameSlot& f = ctx->frame[slot];
    if (!f.pending) return CNDL_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(f.copied));
    f.pending = false;
    return CNDL_OK;
} CNDL_CATCH

int cndl_trace_frame(cndl_ctx* ctx, const cndl_frame_params* p, void* host_out) {
    const int rc = cndl_frame_submit(ctx, p, host_out, 0);
    return rc != CNDL_OK ? rc : cndl_frame_wait(ctx, 0);
}

int cndl_frame_untile_device(cndl_ctx* ctx, const cndl_frame_params* p, const void* d_shard, void* d_frame, void* stream) try {
    if (!ctx) return CNDL_ERR_INVALID;
    TileMap tm;
    int rc = check_params(ctx, p, tm);
    if (rc != CNDL_OK) return rc;
    if (!d_shard || !d_frame) return ctx->fail(CNDL_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(ctx->device));
    const int spp = p->out_format == CNDL_FRAME_OUT_PIXEL32 ? 1 : p->spp;
    const size_t n = (size_t)tm.local_tiles * tm.T * tm.T * (size_t)spp;
    if (n == 0) return CNDL_OK;
    frame_untile_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        tm, (unsigned)n, spp, p->out_format == CNDL_FRAME_OUT_HIT16 ? 1 : 2, static_cast<const float4*>(d_shard), static_cast<float4*>(d_frame));
    ctx->launches.n++;
    CK(cudaGetLastError());
    return CNDL_OK;
} CNDL_CATCH

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Scene replication and several devices behind one handle.

// AddObject for buffers in device memory, in three steps so that several destinations can be filled concurrently:
// reserve (validates, grows the destination buffers; may synchronise), enqueue (asynchronous copies + index rebase on the
// context's stream), finish (waits for the stream, records the object).
namespace {

int prebuilt_reserve(cndl_ctx* ctx, const void* d_nodes, size_t N, const cndl_triangle* d_tris, size_t T, const cndl_vertex* d_verts, size_t V,
                     int32_t leaf_triangle_offset) {
    if (!ctx) return CNDL_ERR_INVALID;
    if (!d_nodes || !d_tris || !d_verts || N == 0 || T == 0 || V == 0) return ctx->fail(CNDL_ERR_INVALID, "null or empty buffer");
    if ((size_t)leaf_triangle_offset != ctx->n_tris)
        return ctx->fail(CNDL_ERR_INVALID, "leaf packs embed triangle offset " + std::to_string(leaf_triangle_offset) + " but the context holds " +
                                               std::to_string(ctx->n_tris) + " triangles: replicate all objects, in insertion order, into an empty context");
    if (ctx->n_nodes + N > 0x7FFFFFF0ull || ctx->n_tris + T > (1ull << 27) || ctx->n_verts + V > 0x7FFFFFF0ull)
        return ctx->fail(CNDL_ERR_INVALID, "scene exceeds the leaf-pack limits (2^27 triangles, BVHConstructor.cpp:794)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->main_stream;
    const size_t ns = ctx->node_size;
    CK(ctx->nodes.reserve((ctx->n_nodes + N + 1) * ns, st));
    CK(ctx->tris.reserve((ctx->n_tris + T) * sizeof(cndl_triangle), st));
    CK(ctx->verts.reserve((ctx->n_verts + V) * sizeof(cndl_vertex), st));
    return CNDL_OK;
}

int prebuilt_enqueue(cndl_ctx* ctx, const void* d_nodes, size_t N, const cndl_triangle* d_tris, size_t T, const cndl_vertex* d_verts, size_t V,
                     int32_t vertex_index_base) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->main_stream;
    const size_t ns = ctx->node_size;
    char* dn = static_cast<char*>(ctx->nodes.p) + ctx->n_nodes * ns;
    char* dt = static_cast<char*>(ctx->tris.p) + ctx->n_tris * sizeof(cndl_triangle);
    char* dv = static_cast<char*>(ctx->verts.p) + ctx->n_verts * sizeof(cndl_vertex);
    CK(cudaMemcpyAsync(dn, d_nodes, N * ns, cudaMemcpyDefault, st));
    CK(cudaMemsetAsync(dn + N * ns, 0, ns, st));
    CK(cudaMemcpyAsync(dt, d_tris, T * sizeof(cndl_triangle), cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(dv, d_verts, V * sizeof(cndl_vertex), cudaMemcpyDefault, st));
    const int delta = (int)ctx->n_verts - vertex_index_base;
    if (delta != 0) launch_rebase_triangles(reinterpret_cast<int4*>(dt), T, delta, st, ctx->launches);  // Intersector.h:190-197
    CK(cudaGetLastError());
    return CNDL_OK;
}

int prebuilt_finish(cndl_ctx* ctx, uint32_t object_id, size_t N, size_t T, size_t V) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->main_stream));
    const size_t ns = ctx->node_size;
    ObjectData od;
    od.node_offset = (int)ctx->n_nodes;
    od.tri_offset = (int)ctx->n_tris;
    od.vert_offset = (int)ctx->n_verts;
    od.node_count = (int)N;
    od.tri_count = (int)T;
    od.vert_count = (int)V;
    ctx->objects[object_id] = od;
    ctx->n_nodes += N;
    ctx->n_tris += T;
    ctx->n_verts += V;
    ctx->nodes.bytes = ctx->n_nodes * ns;
    ctx->tris.bytes = ctx->n_tris * sizeof(cndl_triangle);
    ctx->verts.bytes = ctx->n_verts * sizeof(cndl_vertex);
    ctx->committed = false;
    return CNDL_OK;
}

}  // namespace

extern "C" {

int cndl_object_device_view(cndl_ctx* ctx, uint32_t object_id, const void** d_nodes, size_t* N, const cndl_triangle** d_tris, size_t* T,
                            const cndl_vertex** d_verts, size_t* V) {
    if (!ctx) return CNDL_ERR_INVALID;
    auto it = ctx->objects.find(object_id);
    if (it == ctx->objects.end()) return ctx->fail(CNDL_ERR_UNKNOWN_OBJECT, "no such object");
    const ObjectData& o = it->second;
    if (d_nodes) *d_nodes = static_cast<const char*>(ctx->nodes.p) + (size_t)o.node_offset * ctx->node_size;
    if (N) *N = (size_t)o.node_count;
    if (d_tris) *d_tris = static_cast<const cndl_triangle*>(ctx->tris.p) + o.tri_offset;
    if (T) *T = (size_t)o.tri_count;
    if (d_verts) *d_verts = static_cast<const cndl_vertex*>(ctx->verts.p) + o.vert_offset;
    if (V) *V = (size_t)o.vert_count;
    return CNDL_OK;
}

int cndl_add_prebuilt_object_device(cndl_ctx* ctx, uint32_t object_id, const void* d_nodes, size_t N, const cndl_triangle* d_tris, size_t T,
                                    const cndl_vertex* d_verts, size_t V, int32_t vertex_index_base, int32_t leaf_triangle_offset) try {
    int rc = prebuilt_reserve(ctx, d_nodes, N, d_tris, T, d_verts, V, leaf_triangle_offset);
    if (rc == CNDL_OK) rc = prebuilt_enqueue(ctx, d_nodes, N, d_tris, T, d_verts, V, vertex_index_base);
    if (rc == CNDL_OK) rc = prebuilt_finish(ctx, object_id, N, T, V);
    return rc;
} CNDL_CATCH

int cndl_clone_scene(cndl_ctx* dst, cndl_ctx* src) try {
    if (!dst || !src) return CNDL_ERR_INVALID;
    if (dst->format != src->format) return dst->fail(CNDL_ERR_INVALID, "node formats differ");
    if (dst->n_tris != 0) return dst->fail(CNDL_ERR_INVALID, "cndl_clone_scene needs an empty destination: leaf packs hold global triangle offsets");
    std::vector<std::pair<uint32_t, ObjectData>> objs(src->objects.begin(), src->objects.end());
    std::sort(objs.begin(), objs.end(), [](const auto& a, const auto& b) { return a.second.node_offset < b.second.node_offset; });
    for (const auto& kv : objs) {
        const ObjectData& o = kv.second;
        const int rc = cndl_add_prebuilt_object_device(dst, kv.first, static_cast<const char*>(src->nodes.p) + (size_t)o.node_offset * src->node_size,
                                                       (size_t)o.node_count, static_cast<const cndl_triangle*>(src->tris.p) + o.tri_offset, (size_t)o.tri_count,
                                                       static_cast<const cndl_vertex*>(src->verts.p) + o.vert_offset, (size_t)o.vert_count, o.vert_offset,
                                                       o.tri_offset);
        if (rc != CNDL_OK) return rc;
    }
    return CNDL_OK;
} CNDL_CATCH

}  // extern "C"

struct cndl_multi {
    std::vector<cndl_ctx*> ctx;
    std::vector<int> devices;
    std::string err;
    bool peer = false;                 // every device can store into the first one's memory
    bool peer_capable = false;
    std::vector<cudaEvent_t> done[kFrameSlots];  // per slot, per device: its shard's records are in the first device's frame
    std::vector<cndl::DeviceBuffer*> stage[kFrameSlots];  // without peer access: shard in local layout on its own device / on the first device
    float replicate_ms = 0.0f;
    bool pending[kFrameSlots] = {};
    unsigned long long rays_traced[kFrameSlots] = {};
    int fail(int code, const std::string& msg) { err = msg; return code; }
    int fail_from(int code, int i) { err = "device " + std::to_string(devices[(size_t)i]) + ": " + ctx[(size_t)i]->err; return code; }
};

extern "C" {

int cndl_multi_create(cndl_multi** out, int node_format, const int* devices, int n_devices) try {
    if (!out) return CNDL_ERR_INVALID;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return CNDL_ERR_INVALID;
    cndl_multi* m = new (std::nothrow) cndl_multi;
    if (!m) return CNDL_ERR_OOM;
    for (int i = 0; i < n_devices; ++i) {
        cndl_ctx* c = nullptr;
        const int rc = cndl_create(&c, node_format, devices[i]);
        if (rc != CNDL_OK) { cndl_multi_destroy(m); return rc; }
        m->ctx.push_back(c);
        m->devices.push_back(devices[i]);
    }
    // peer access both ways between the first device and every other one (stores into its frame; copies of its scene)
    m->peer = true;
    for (int i = 1; i < n_devices; ++i) {
        int a = 0, b = 0;
        if (devices[i] == devices[0]) continue;  // a second context on the same device reaches the frame directly
        cudaDeviceCanAccessPeer(&a, devices[i], devices[0]);
        cudaDeviceCanAccessPeer(&b, devices[0], devices[i]);
        if (!a || !b) { m->peer = false; continue; }
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer = false;
        cudaSetDevice(devices[0]);
        e = cudaDeviceEnablePeerAccess(devices[i], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer = false;
        cudaGetLastError();
    }
    m->peer_capable = m->peer;
    for (int s = 0; s < kFrameSlots; ++s)
        for (int i = 0; i < n_devices; ++i) {
            cudaSetDevice(devices[i]);
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cndl_multi_destroy(m); return CNDL_ERR_CUDA; }
            m->done[s].push_back(e);
            m->stage[s].push_back(new cndl::DeviceBuffer);
            m->stage[s].push_back(new cndl::DeviceBuffer);
        }
    *out = m;
    return CNDL_OK;
} CNDL_CATCH

void cndl_multi_destroy(cndl_multi* m) {
    if (!m) return;
    for (int s = 0; s < kFrameSlots; ++s) {
        for (size_t i = 0; i < m->done[s].size(); ++i) {
            cudaSetDevice(m->devices[i]);
            cudaDeviceSynchronize();
            cudaEventDestroy(m->done[s][i]);
        }
        for (size_t k = 0; k < m->stage[s].size(); ++k) {
            cudaSetDevice(m->devices[(k & 1) ? 0 : k / 2]);
            delete m->stage[s][k];
        }
    }
    for (cndl_ctx* c : m->ctx) cndl_destroy(c);
    delete m;
}

int cndl_multi_set_transport(cndl_multi* m, int transport) {
    if (!m || (transport != CNDL_TRANSPORT_PEER_STORES && transport != CNDL_TRANSPORT_STAGED_COPY)) return CNDL_ERR_INVALID;
    if (transport == CNDL_TRANSPORT_PEER_STORES && !m->peer_capable) return m->fail(CNDL_ERR_INVALID, "peer access between the devices is not available");
    m->peer = transport == CNDL_TRANSPORT_PEER_STORES;
    return CNDL_OK;
}

int cndl_multi_device_count(const cndl_multi* m) { return m ? (int)m->ctx.size() : 0; }
cndl_ctx* cndl_multi_context(cndl_multi* m, int i) { return (m && i >= 0 && (size_t)i < m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }
const char* cndl_multi_last_error(const cndl_multi* m) { return m ? m->err.c_str() : "null handle"; }
float cndl_multi_last_replicate_ms(const cndl_multi* m) { return m ? m->replicate_ms : 0.0f; }
uint64_t cndl_multi_frame_rays_traced(const cndl_multi* m, int slot) { return (m && slot >= 0 && slot < kFrameSlots) ? m->rays_traced[slot] : 0; }

int cndl_multi_add_object(cndl_multi* m, uint32_t object_id, const cndl_vertex* verts, size_t V, const uint32_t* indices, size_t I,
                          const int32_t* mesh_id_per_tri, const cndl_build_opts* opts) try {
    if (!m) return CNDL_ERR_INVALID;
    cndl_ctx* c0 = m->ctx[0];
    int rc = cndl_add_object(c0, object_id, verts, V, indices, I, mesh_id_per_tri, opts);
    if (rc != CNDL_OK) return m->fail_from(rc, 0);
    const ObjectData o = c0->objects[object_id];
    // replicate the object's reference-layout slices device to device: every destination reserves first, then all copies are in
    // flight together (one source, NVLink to every peer), then every destination is waited for.  Host wall clock around the copies.
    const void* sn = static_cast<const char*>(c0->nodes.p) + (size_t)o.node_offset * c0->node_size;
    const cndl_triangle* stp = static_cast<const cndl_triangle*>(c0->tris.p) + o.tri_offset;
    const cndl_vertex* sv = static_cast<const cndl_vertex*>(c0->verts.p) + o.vert_offset;
    const size_t N = (size_t)o.node_count, T = (size_t)o.tri_count, Vn = (size_t)o.vert_count;
    for (size_t i = 1; i < m->ctx.size(); ++i) {
        rc = prebuilt_reserve(m->ctx[i], sn, N, stp, T, sv, Vn, o.tri_offset);
        if (rc != CNDL_OK) return m->fail_from(rc, (int)i);
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 1; i < m->ctx.size(); ++i) {
        rc = prebuilt_enqueue(m->ctx[i], sn, N, stp, T, sv, Vn, o.vert_offset);
        if (rc != CNDL_OK) return m->fail_from(rc, (int)i);
    }
    for (size_t i = 1; i < m->ctx.size(); ++i) {
        rc = prebuilt_finish(m->ctx[i], object_id, N, T, Vn);
        if (rc != CNDL_OK) return m->fail_from(rc, (int)i);
    }
    m->replicate_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return CNDL_OK;
} CNDL_CATCH

int cndl_multi_commit(cndl_multi* m) {
    if (!m) return CNDL_ERR_INVALID;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        const int rc = cndl_commit(m->ctx[i], 1);
        if (rc != CNDL_OK) return m->fail_from(rc, (int)i);
    }
    return CNDL_OK;
}

int cndl_multi_push_entity(cndl_multi* m, uint32_t object_id, const float model[16], float emissive, float translucency) {
    if (!m) return CNDL_ERR_INVALID;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        const int rc = cndl_push_entity(m->ctx[i], object_id, model, emissive, translucency);
        if (rc != CNDL_OK) return m->fail_from(rc, (int)i);
    }
    return CNDL_OK;
}

int cndl_multi_buffer_entities(cndl_multi* m) {
    if (!m) return CNDL_ERR_INVALID;
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        const int rc = cndl_buffer_entities(m->ctx[i]);
        if (rc != CNDL_OK) return m->fail_from(rc, (int)i);
    }
    return CNDL_OK;
}

int cndl_multi_frame_wait(cndl_multi* m, int slot) {
    if (!m || slot < 0 || slot >= kFrameSlots) return CNDL_ERR_INVALID;
    if (!m->pending[slot]) return CNDL_OK;
    cndl_ctx* c0 = m->ctx[0];
    cudaSetDevice(m->devices[0]);
    cudaError_t e = cudaEventSynchronize(c0->frame[slot].copied);
    if (e != cudaSuccess) return m->fail(CNDL_ERR_CUDA, cudaGetErrorString(e));
    m->pending[slot] = false;
    unsigned long long n = 0;
    for (size_t i = 0; i < m->ctx.size(); ++i) {  // every device's counts arrived before its `done` event, which the copy waited on
        cudaSetDevice(m->devices[i]);
        cudaEventSynchronize(m->done[slot][i]);
        n += cndl_frame_rays_traced(m->ctx[i], slot);
    }
    m->rays_traced[slot] = n;
    return CNDL_OK;
}

int cndl_multi_frame_submit(cndl_multi* m, const cndl_frame_params* p, void* host_out, int slot) try {
    if (!m || !p || !host_out || slot < 0 || slot >= kFrameSlots) return CNDL_ERR_INVALID;
    int rc = CNDL_OK;
    if (m->pending[slot]) {
        rc = cndl_multi_frame_wait(m, slot);
        if (rc != CNDL_OK) return rc;
    }
    const int n = (int)m->ctx.size();
    cndl_ctx* c0 = m->ctx[0];
    cndl_frame_params q = *p;
    q.shard_count = n;
    q.flags &= ~(uint32_t)CNDL_FRAME_LOCAL_LAYOUT;
    const size_t bytes = cndl_frame_records(&q) * cndl_frame_record_bytes(q.out_format);
    cudaSetDevice(m->devices[0]);
    if (c0->frame[slot].out.ensure_scratch(bytes ? bytes : 16) != cudaSuccess) return m->fail(CNDL_ERR_OOM, "frame buffer");
    for (int i = 0; i < n; ++i) {
        cndl_ctx* c = m->ctx[(size_t)i];
        q.shard_index = i;
        cudaSetDevice(m->devices[(size_t)i]);
        if (!c->frame_stream[slot]) cudaStreamCreateWithFlags(&c->frame_stream[slot], cudaStreamNonBlocking);  // else: created by the first frame call on the context
        if (m->peer || i == 0) {
            // the resolve kernel of device i stores its pixels straight into the first device's frame (NVLink peer memory)
            rc = cndl_trace_frame_device(c, &q, c0->frame[slot].out.p, slot, c->frame_stream[slot]);
            if (rc != CNDL_OK) return m->fail_from(rc, i);
        } else {
            // no peer access: shard in local layout on its own device, copied to the first device, untiled there
            cndl_frame_params ql = q;
            ql.flags |= CNDL_FRAME_LOCAL_LAYOUT;
            const size_t sb = cndl_frame_shard_records(&ql) * cndl_frame_record_bytes(q.out_format);
            cndl::DeviceBuffer* own = m->stage[slot][2 * (size_t)i];
            cndl::DeviceBuffer* at0 = m->stage[slot][2 * (size_t)i + 1];
            if (own->ensure_scratch(sb ? sb : 16) != cudaSuccess) return m->fail(CNDL_ERR_OOM, "shard buffer");
            rc = cndl_trace_frame_device(c, &ql, own->p, slot, c->frame_stream[slot]);
            if (rc != CNDL_OK) return m->fail_from(rc, i);
            cudaSetDevice(m->devices[0]);
            if (at0->ensure_scratch(sb ? sb : 16) != cudaSuccess) return m->fail(CNDL_ERR_OOM, "shard buffer");
            cudaSetDevice(m->devices[(size_t)i]);
            cudaMemcpyPeerAsync(at0->p, m->devices[0], own->p, m->devices[(size_t)i], sb, c->frame_stream[slot]);
        }
        cudaEventRecord(m->done[slot][(size_t)i], c->frame_stream[slot]);
    }
    // first device: wait for every shard, (untile the staged ones,) one copy to the host
    cudaSetDevice(m->devices[0]);
    rc = CNDL_OK;
    if (!c0->frame_copy_stream) cudaStreamCreateWithFlags(&c0->frame_copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < n; ++i) cudaStreamWaitEvent(c0->frame_copy_stream, m->done[slot][(size_t)i], 0);
    if (!m->peer)
        for (int i = 1; i < n; ++i) {
            cndl_frame_params ql = q;
            ql.shard_index = i;
            rc = cndl_frame_untile_device(c0, &ql, m->stage[slot][2 * (size_t)i + 1]->p, c0->frame[slot].out.p, c0->frame_copy_stream);
            if (rc != CNDL_OK) return m->fail_from(rc, 0);
        }
    cudaMemcpyAsync(host_out, c0->frame[slot].out.p, bytes, cudaMemcpyDeviceToHost, c0->frame_copy_stream);
    cudaEventRecord(c0->frame[slot].copied, c0->frame_copy_stream);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return m->fail(CNDL_ERR_CUDA, cudaGetErrorString(e));
    m->pending[slot] = true;
    return CNDL_OK;
} CNDL_CATCH

int cndl_multi_trace_frame(cndl_multi* m, const cndl_frame_params* p, void* host_out) {
    const int rc = cndl_multi_frame_submit(m, p, host_out, 0);
    return rc != CNDL_OK ? rc : cndl_multi_frame_wait(m, 0);
}

}  // extern "C"
