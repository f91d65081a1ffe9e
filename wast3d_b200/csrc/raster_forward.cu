// Forward rasteriser for sm_100a: preprocess (K1), binning (K2-K5) and tile render (K6).
//
// What it replaces (reference paths under submodules/diff-gaussian-rasterization/):
//   preprocessCUDA            cuda_rasterizer/forward.cu:155-256
//   InclusiveSum + duplicateWithKeys + SortPairs + identifyTileRanges
//                             cuda_rasterizer/rasterizer_impl.cu:70-138,279-320
//   renderCUDA (forward)      cuda_rasterizer/forward.cu:261-390
//   Rasterizer::forward       cuda_rasterizer/rasterizer_impl.cu:198-341
//
// Design differences (results are the same, see tests/):
//  * K1 writes one 48-byte record per visible Gaussian (mean2D, depth, conic, opacity, rgb and
//    the alpha-cutoff half extents) instead of seven separate arrays; SH coefficients are
//    streamed with coalesced 16-byte loads through shared memory, only for Gaussians that
//    survive culling and only up to the active degree.
//  * The blend order (tile, depth, index) is produced by sorting the P Gaussians by depth
//    (32-bit keys) and then stably partitioning the R instances by tile id (1-2 passes over
//    8-byte pairs) instead of 6 passes over 12-byte (u64,u32) pairs.
//  * K6 stages whole records (incl. rgb and depth) per batch with cp.async, double buffered;
//    32 lanes test 32 Gaussians against the warp's 8x4 pixel block at once and the blend loop
//    only visits survivors; warps retire independently (ballot) and the block leaves when all
//    have (one __syncthreads_and per batch).
#include "project.cuh"
#include <atomic>
#include <cstdlib>

namespace w3d {

constexpr int PRE_THREADS = 128;
constexpr int PRE_WARPS = PRE_THREADS / 32;
constexpr int SH_MAX_ROW = 48;             // M = 16 coefficients x RGB
constexpr int SH_STRIDE = SH_MAX_ROW + 1;  // odd -> conflict-free per-Gaussian reads

// RAW: model-space inputs (wast3d_raster_params::raw_params) — activations applied here.
// COLOUR: true = SH -> RGB evaluated here (one kernel, as in the reference's preprocessCUDA);
// false = geometry only: rgb is left to sh_colour_kernel, which the host launches just before the
// tile render (wast3d_raster_params::colour_wait_event: the SH coefficients may still be in flight
// from the optimizer's parameter all-gather while projection, sorting and binning already run).
template <bool RAW, bool COLOUR>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_kernel(const int P, const int D, const int M, const float* __restrict__ means3D,
                  const float* __restrict__ scales, const float scale_modifier,
                  const float* __restrict__ rotations, const float* __restrict__ opacities,
                  const float* __restrict__ shs, const float* __restrict__ shs_rest,
                  const float* __restrict__ cov3D_precomp,
                  const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
                  const float* __restrict__ projmatrix, const float* __restrict__ cam_pos,
                  const int W, const int H, const float tan_fovx, const float tan_fovy,
                  const float focal_x, const float focal_y, const dim3 grid,
                  const bool prefiltered, int* __restrict__ radii, float4* __restrict__ rec,
                  uint32_t* __restrict__ depth_key, uint32_t* __restrict__ tiles_touched,
                  uint8_t* __restrict__ clamped, uint2* __restrict__ rect_out, uint32_t* __restrict__ totals,
                  const uint32_t* __restrict__ sample_bound_words, const bool cut_tiles, const bool vec3_loads) {
    __shared__ __align__(16) float s_sh[PRE_WARPS][32 * SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * PRE_THREADS + warp * 32;
    const bool live = idx < P;

    float3 p_orig = make_float3(0.f, 0.f, 0.f);

    // means3D / scales: [P,3] AoS read with coalesced 16-byte loads (load_rows3x2)
    __shared__ __align__(16) float s_v3[PRE_WARPS][192];
    const int rows_valid_w = max(0, min(32, P - warp_first));
    float3 sc_in = make_float3(0.f, 0.f, 0.f);
    if (vec3_loads) {
        load_rows3x2(means3D, cov3D_precomp == nullptr ? scales : nullptr, warp_first, rows_valid_w, lane, s_v3[warp],
                     p_orig, sc_in);
    } else if (live) {   // A/B: three strided 4-byte loads per array and lane
        p_orig = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
        if (cov3D_precomp == nullptr) sc_in = make_float3(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]);
    }
    ProjView pv;
    pv.viewmatrix = viewmatrix; pv.projmatrix = projmatrix;
    pv.W = W; pv.H = H; pv.tan_fovx = tan_fovx; pv.tan_fovy = tan_fovy; pv.focal_x = focal_x; pv.focal_y = focal_y;
    pv.grid_x = grid.x; pv.grid_y = grid.y; pv.scale_modifier = scale_modifier; pv.cut_tiles = cut_tiles;
    pv.sb = load_sample_bounds(sample_bound_words);
    Projection pr = Projection::none();
    if (live) {
        pr = project_gaussian<RAW>(p_orig, sc_in, rotations ? rotations + 4 * (size_t)idx : nullptr, make_float4(0.f, 0.f, 0.f, 0.f),
                                   opacities + idx, 0.f, cov3D_precomp ? cov3D_precomp + 6 * (size_t)idx : nullptr, pv);
        if (pr.behind && prefiltered) atomicOr(totals + 1, 1u);  // reference: printf + __trap()
    }
    const bool visible = pr.visible;
    {   // num_rendered (rasterizer_impl.cu:279-283 takes it from the end of the scan): one atomic per warp
        const uint32_t warp_tiles = __reduce_add_sync(0xffffffffu, pr.visible ? pr.n_tiles : 0u);
        if (lane == 0 && warp_tiles != 0u) atomicAdd(totals, warp_tiles);
    }

    // colour: SH -> RGB for survivors only
    float3 rgb = make_float3(0.f, 0.f, 0.f);
    unsigned clamp_bits = 0;
    if (!COLOUR) {
        // deferred: sh_colour_kernel fills rgb and the clamp bits
    } else if (colors_precomp == nullptr) {
        const unsigned need = __ballot_sync(0xffffffffu, visible);
        if (need) {
            const int rows_valid = min(32, P - warp_first);
            const int used = 3 * (D + 1) * (D + 1);
            const float3 cam = *reinterpret_cast<const float3*>(cam_pos);
            if (RAW) {
                // _features_dc [P,1,3] read per thread, _features_rest [P,M-1,3] copied linearly
                float dc[3] = {0.f, 0.f, 0.f};
                if (visible) { dc[0] = shs[3 * idx]; dc[1] = shs[3 * idx + 1]; dc[2] = shs[3 * idx + 2]; }
                const int rest_floats = 3 * (M - 1);
                if (used > 3)
                    stage_rows_linear(shs_rest + (size_t)warp_first * rest_floats, rest_floats, used - 3,
                                      rows_valid, need, s_sh[warp], lane);
                if (visible) rgb = sh_to_rgb(D, dc, s_sh[warp] + lane * rest_floats - 3, p_orig, cam, &clamp_bits);
            } else {
                const int row_floats = 3 * M;
                stage_sh_rows(shs + (size_t)warp_first * row_floats, row_floats, used, rows_valid, need,
                              s_sh[warp], SH_STRIDE, lane);
                __syncwarp();
                const float* row = s_sh[warp] + lane * SH_STRIDE;
                if (visible) rgb = sh_to_rgb(D, row, row, p_orig, cam, &clamp_bits);
            }
        }
    } else if (visible) {
        rgb = make_float3(colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]);
    }

    if (!live) return;
    store_projection(idx, pr, rgb, clamp_bits, radii, rec, depth_key, tiles_touched, clamped, rect_out);
}

// Deferred colour pass (forward.cu:20-71,241-246 for the Gaussians preprocess_kernel<.., false> kept):
// same staging and arithmetic as the fused kernel, so the render record it completes is bit-identical.
template <bool RAW>
__global__ void __launch_bounds__(PRE_THREADS)
sh_colour_kernel(const int P, const int D, const int M, const float* __restrict__ means3D,
                 const float* __restrict__ shs, const float* __restrict__ shs_rest,
                 const float* __restrict__ cam_pos, float4* __restrict__ rec, uint8_t* __restrict__ clamped) {
    __shared__ __align__(16) float s_sh[PRE_WARPS][32 * SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * PRE_THREADS + warp * 32;
    const bool visible = idx < P && (clamped[idx] & 8u) != 0;  // bit 3: render record written
    const unsigned need = __ballot_sync(0xffffffffu, visible);
    if (!need) return;
    float3 p_orig = make_float3(0.f, 0.f, 0.f);
    if (visible) p_orig = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    float3 rgb = make_float3(0.f, 0.f, 0.f);
    unsigned clamp_bits = 0;
    const int rows_valid = min(32, P - warp_first);
    const int used = 3 * (D + 1) * (D + 1);
    const float3 cam = *reinterpret_cast<const float3*>(cam_pos);
    if (RAW) {
        float dc[3] = {0.f, 0.f, 0.f};
        if (visible) { dc[0] = shs[3 * idx]; dc[1] = shs[3 * idx + 1]; dc[2] = shs[3 * idx + 2]; }
        const int rest_floats = 3 * (M - 1);
        if (used > 3)
            stage_rows_linear(shs_rest + (size_t)warp_first * rest_floats, rest_floats, used - 3,
                              rows_valid, need, s_sh[warp], lane);
        if (visible) rgb = sh_to_rgb(D, dc, s_sh[warp] + lane * rest_floats - 3, p_orig, cam, &clamp_bits);
    } else {
        const int row_floats = 3 * M;
        stage_sh_rows(shs + (size_t)warp_first * row_floats, row_floats, used, rows_valid, need,
                      s_sh[warp], SH_STRIDE, lane);
        __syncwarp();
        const float* row = s_sh[warp] + lane * SH_STRIDE;
        if (visible) rgb = sh_to_rgb(D, row, row, p_orig, cam, &clamp_bits);
    }
    if (visible) {
        float* c = reinterpret_cast<float*>(rec + 3 * (size_t)idx + 2);  // .w (cutoff half-extent y) stays
        c[0] = rgb.x;
        c[1] = rgb.y;
        c[2] = rgb.z;
        clamped[idx] = (uint8_t)(clamp_bits | 8u);
    }
}

// Bounds of the per-pixel sampling offsets (forward.cu:287 adds them to the pixel coordinate):
// words[0..3] = float_order_key of max(ox), max(-ox), max(oy), max(-oy), accumulated with atomicMax
// into zero-initialised words.  A NaN offset disables the tile cut (bounds become +inf).
__global__ void __launch_bounds__(256)
sample_bounds_kernel(size_t n_pix, const float* __restrict__ offs, uint32_t* __restrict__ words) {
    const float inf = __int_as_float(0x7f800000);
    float mx = -inf, nx = -inf, my = -inf, ny = -inf;
    // two pixels (16 bytes) per load, four independent loads in flight per thread: the kernel is a latency chain otherwise
    const size_t n2 = n_pix / 2, stride = (size_t)gridDim.x * blockDim.x;
    const bool vec = (reinterpret_cast<uintptr_t>(offs) & 15) == 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto take = [&](float4 o) {
        if (!(o.x == o.x) || !(o.y == o.y) || !(o.z == o.z) || !(o.w == o.w)) mx = nx = my = ny = inf;
        mx = fmaxf(mx, fmaxf(o.x, o.z)); nx = fmaxf(nx, fmaxf(-o.x, -o.z));
        my = fmaxf(my, fmaxf(o.y, o.w)); ny = fmaxf(ny, fmaxf(-o.y, -o.w));
    };
    if (vec) {
        const float4* o4 = reinterpret_cast<const float4*>(offs);
        for (; i + 3 * stride < n2; i += 4 * stride) {
            const float4 a = o4[i], b = o4[i + stride], c = o4[i + 2 * stride], d = o4[i + 3 * stride];
            take(a); take(b); take(c); take(d);
        }
        for (; i < n2; i += stride) take(o4[i]);
        if (blockIdx.x == 0 && threadIdx.x == 0 && (n_pix & 1)) {
            const float2 o = *reinterpret_cast<const float2*>(offs + 2 * (n_pix - 1));
            take(make_float4(o.x, o.y, o.x, o.y));
        }
    } else {
        for (; i < n_pix; i += stride) {
            const float2 o = *reinterpret_cast<const float2*>(offs + 2 * i);
            take(make_float4(o.x, o.y, o.x, o.y));
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        nx = fmaxf(nx, __shfl_xor_sync(0xffffffffu, nx, d));
        my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, d));
        ny = fmaxf(ny, __shfl_xor_sync(0xffffffffu, ny, d));
    }
    // one set of atomics per block: 4700 warps hammering four words cost more than reading the offsets
    __shared__ float s_red[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = mx; s_red[1][warp] = nx; s_red[2][warp] = my; s_red[3][warp] = ny; }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = s_red[threadIdx.x][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = fmaxf(v, s_red[threadIdx.x][w]);
        atomicMax(words + threadIdx.x, float_order_key(v));
    }
}

// rasterizer_impl.cu:54-66
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D,
                                    const float* __restrict__ viewmatrix, unsigned char* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    float3 v = xform_point_4x3(p, viewmatrix);
    present[idx] = v.z <= 0.2f ? 0 : 1;
}

// Emits (tile id, Gaussian id) for every tile of every visible Gaussian's rectangle, Gaussians in
// DEPTH order and a Gaussian's tiles row-major (like duplicateWithKeys rasterizer_impl.cu:98-109),
// so that a stable partition by tile id yields the reference's (tile, depth, index) order.
// A warp owns 32 depth-consecutive Gaussians; their instances form one contiguous output range
// [off_0, off_31 + n_31) which the 32 lanes write with coalesced stores: lane p finds its
// Gaussian with a 5-step shuffle binary search over the scanned offsets.  Persistent grid; every
// block also counts instances per tile in shared memory (flushed once), from which the per-tile
// ranges follow by a scan — identifyTileRanges (rasterizer_impl.cu:116-138) never has to read
// the sorted keys.
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_MAX_SMEM_TILES = 12288;  // 48 KB of counters; larger grids count in global memory

// FUSED: the exclusive scan of the per-Gaussian tile counts (rasterizer_impl.cu:279) happens here instead of in its own
// kernels: a block takes EMIT_GROUP depth-consecutive Gaussians by ticket (8 per thread, so the ticket / gather /
// look-back latency chain is paid once per 2048 Gaussians), scans their counts, gets the instances before it with a
// decoupled look-back over `scan_ws` (ticket, error word, one status word per group — zeroed by the caller), parks
// (offset, id, rectangle) in shared memory and then emits warp by warp as above.  !FUSED reads precomputed `offsets`
// (the preprojected path, and R >= 2^30 which the status words cannot hold).
constexpr int EMIT_PER_THREAD = 8;
constexpr int EMIT_GROUP = EMIT_THREADS * EMIT_PER_THREAD;
struct EmitGroup {
    uint32_t off[EMIT_GROUP], id[EMIT_GROUP], xy0[EMIT_GROUP], wh[EMIT_GROUP];
};

// One warp's 32 Gaussians (one per lane: first instance `off`, count `n`, id, rectangle origin, width): their instances
// form the contiguous range [off_0, max(off + n)) which the lanes write with coalesced stores, 32 slots per step.
// Gaussians without instances are squeezed out first (through `scratch`, 4 x 32 words of this warp), so the first
// instances of the remaining lanes are strictly increasing and a slot finds its Gaussian with one OR-reduction of the
// "starts in this window" bits and two popcounts instead of a shuffle binary search.
__device__ __forceinline__ void emit_warp_run(uint32_t off, uint32_t n, uint32_t id, uint32_t xy0, uint32_t w,
                                              const int lane, uint32_t* __restrict__ scratch,
                                              uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                              uint32_t* __restrict__ tile_count, uint32_t* s_count, const bool smem_hist,
                                              const uint32_t grid_x, const uint32_t capacity) {
    const unsigned has = __ballot_sync(0xffffffffu, n != 0u);
    if (has == 0u) return;
    const uint32_t end = __reduce_max_sync(0xffffffffu, n != 0u ? off + n : 0u);
    if (has != 0xffffffffu) {
        __syncwarp();
        if (n != 0u) {
            const int r = __popc(has & lanemask_lt());
            scratch[r] = off; scratch[32 + r] = id; scratch[64 + r] = xy0; scratch[96 + r] = w;
        }
        __syncwarp();
        const bool live = lane < __popc(has);
        off = live ? scratch[lane] : 0xFFFFFFFFu;
        id = scratch[32 + lane]; xy0 = scratch[64 + lane]; w = live ? scratch[96 + lane] : 1u;
    }
    // floor(t / w) = umulhi(t, ceil(2^32 / w)) while t * w < 2^32 (t < w * h: rectangles up to 1024 tiles wide and 2^22
    // tiles large qualify); inv = 0 marks the Gaussians that take the division
    const uint32_t wn = __reduce_max_sync(0xffffffffu, w);
    const uint32_t inv = w > 1u ? 0xFFFFFFFFu / w + 1u : 0u;
    const bool all_fast = wn < 1024u && end - __shfl_sync(0xffffffffu, off, 0) < (1u << 22);   // warp-uniform
    const uint32_t begin = __shfl_sync(0xffffffffu, off, 0);
    const unsigned le = 0xFFFFFFFFu >> (31 - lane);
    for (uint32_t p0 = begin; p0 < end; p0 += 32) {   // warp-uniform trip count
        const uint32_t rel = off - p0;                 // lanes that start inside this window: rel < 32
        const unsigned starts = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
        const int before = __popc(__ballot_sync(0xffffffffu, off < p0));   // lanes that started earlier (>= 1 after step 0)
        const int l = before - 1 + __popc(starts & le);                   // the last lane with off_l <= p0 + lane
        const uint32_t p = p0 + lane;
        const uint32_t o = __shfl_sync(0xffffffffu, off, l);
        const uint32_t src_xy0 = __shfl_sync(0xffffffffu, xy0, l);
        const uint32_t src_w = __shfl_sync(0xffffffffu, w, l);
        const uint32_t src_id = __shfl_sync(0xffffffffu, id, l);
        const uint32_t src_inv = __shfl_sync(0xffffffffu, inv, l);
        if (p < end && p < capacity) {
            const uint32_t t = p - o;
            const uint32_t ty = src_w == 1u ? t : (all_fast ? __umulhi(t, src_inv) : t / src_w);
            const uint32_t tx = t - ty * src_w;
            const uint32_t tile = ((src_xy0 >> 16) + ty) * grid_x + (src_xy0 & 0xFFFFu) + tx;
            keys[p] = tile;
            vals[p] = src_id;
            if (smem_hist) atomicAdd(&s_count[tile], 1u);
            else atomicAdd(&tile_count[tile], 1u);
        }
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const uint2* __restrict__ rect /* depth order */, uint32_t* __restrict__ keys,
                      uint32_t* __restrict__ vals, uint32_t* __restrict__ tile_count, const dim3 grid,
                      const int num_tiles, const bool smem_hist, const uint32_t capacity,
                      uint32_t* __restrict__ totals, uint32_t* __restrict__ scan_ws) {
    extern __shared__ __align__(16) uint32_t s_count[];   // [smem_hist ? num_tiles : 0] counters, then (FUSED) an EmitGroup
    __shared__ uint32_t s_ticket, s_base;
    __shared__ uint32_t s_scratch[EMIT_THREADS / 32][128];
    // graph-safe forward: the pair arrays hold `capacity` instances; a view that needs more raises totals[3]
    // (nothing is written out of bounds; the image of that call is incomplete and the host layer reports it)
    if (blockIdx.x == 0 && threadIdx.x == 0 && totals[0] > capacity) totals[3] = 1u;
    if (smem_hist) {
        for (int t = threadIdx.x; t < num_tiles; t += EMIT_THREADS) s_count[t] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int warps_per_block = EMIT_THREADS / 32;
    if (FUSED) {
        EmitGroup& grp = *reinterpret_cast<EmitGroup*>(s_count + (smem_hist ? (num_tiles + 3) / 4 * 4 : 0));
        const uint32_t n_groups = (uint32_t)((P + EMIT_GROUP - 1) / EMIT_GROUP);
        while (true) {
            __syncthreads();   // everybody is done with the previous group's shared arrays
            if (threadIdx.x == 0) s_ticket = atomicAdd(scan_ws, 1u);
            __syncthreads();
            const uint32_t ticket = s_ticket;
            if (ticket >= n_groups) break;
            // thread t owns Gaussians [k0, k0 + 8) of the group
            const int k0 = (int)ticket * EMIT_GROUP + threadIdx.x * EMIT_PER_THREAD;
            uint32_t ids[EMIT_PER_THREAD], cnt[EMIT_PER_THREAD], xy[EMIT_PER_THREAD], wh[EMIT_PER_THREAD];
            if (k0 + EMIT_PER_THREAD <= P) {
                const uint4 a = *reinterpret_cast<const uint4*>(order + k0), b = *reinterpret_cast<const uint4*>(order + k0 + 4);
                ids[0] = a.x; ids[1] = a.y; ids[2] = a.z; ids[3] = a.w; ids[4] = b.x; ids[5] = b.y; ids[6] = b.z; ids[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < EMIT_PER_THREAD; ++j) ids[j] = k0 + j < P ? order[k0 + j] : 0xFFFFFFFFu;
            }
            uint32_t sum = 0;
            if (k0 + EMIT_PER_THREAD <= P) {   // the rectangles K1 counted, in depth order: 64 contiguous bytes per thread
#pragma unroll
                for (int j = 0; j < EMIT_PER_THREAD; j += 2) {
                    const uint4 rr = *reinterpret_cast<const uint4*>(rect + k0 + j);
                    xy[j] = rr.x; wh[j] = rr.y; xy[j + 1] = rr.z; wh[j + 1] = rr.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < EMIT_PER_THREAD; ++j) {
                    const uint2 rc = k0 + j < P ? rect[k0 + j] : make_uint2(0u, 0u);
                    xy[j] = rc.x;
                    wh[j] = rc.y;
                }
            }
#pragma unroll
            for (int j = 0; j < EMIT_PER_THREAD; ++j) {
                cnt[j] = (wh[j] & 0xFFFFu) * (wh[j] >> 16);
                sum += cnt[j];
            }
            uint32_t group_total;
            uint32_t run = block_excl_scan<EMIT_THREADS>(sum, &group_total);
            if (threadIdx.x < 32) {
                const uint32_t e = lookback_warp(scan_ws + 2, ticket, group_total, scan_ws + 1, lane);
                if (lane == 0) s_base = e;
            }
            __syncthreads();
            run += s_base;
            uint32_t offs[EMIT_PER_THREAD];
#pragma unroll
            for (int j = 0; j < EMIT_PER_THREAD; ++j) {
                offs[j] = k0 + j < P ? run : 0xFFFFFFFFu;
                run += cnt[j];
            }
            const int slot = threadIdx.x * EMIT_PER_THREAD;
            *reinterpret_cast<uint4*>(grp.off + slot) = make_uint4(offs[0], offs[1], offs[2], offs[3]);
            *reinterpret_cast<uint4*>(grp.off + slot + 4) = make_uint4(offs[4], offs[5], offs[6], offs[7]);
            *reinterpret_cast<uint4*>(grp.id + slot) = make_uint4(ids[0], ids[1], ids[2], ids[3]);
            *reinterpret_cast<uint4*>(grp.id + slot + 4) = make_uint4(ids[4], ids[5], ids[6], ids[7]);
            *reinterpret_cast<uint4*>(grp.xy0 + slot) = make_uint4(xy[0], xy[1], xy[2], xy[3]);
            *reinterpret_cast<uint4*>(grp.xy0 + slot + 4) = make_uint4(xy[4], xy[5], xy[6], xy[7]);
            *reinterpret_cast<uint4*>(grp.wh + slot) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            *reinterpret_cast<uint4*>(grp.wh + slot + 4) = make_uint4(wh[4], wh[5], wh[6], wh[7]);
            __syncthreads();
            // warp w emits Gaussians [256 w, 256 w + 256) of the group, 32 at a time
#pragma unroll 1
            for (int sub = 0; sub < EMIT_PER_THREAD; ++sub) {
                const int sl = warp * (32 * EMIT_PER_THREAD) + sub * 32 + lane;
                const uint32_t off = grp.off[sl], rcw = grp.wh[sl];
                uint32_t w = rcw & 0xFFFFu;
                const uint32_t n = w * (rcw >> 16);
                if (n == 0) w = 1;
                emit_warp_run(off, off != 0xFFFFFFFFu ? n : 0u, grp.id[sl], grp.xy0[sl], w, lane, s_scratch[warp], keys, vals,
                              tile_count, s_count, smem_hist, grid.x, capacity);
            }
        }
    } else {
        const int n_chunks = (P + 31) / 32;
        for (int chunk = blockIdx.x * warps_per_block + warp; chunk < n_chunks; chunk += gridDim.x * warps_per_block) {
            const int k = chunk * 32 + lane;
            uint32_t id = 0, n = 0, off = 0xFFFFFFFFu, xy0 = 0, w = 1;
            if (k < P) {
                id = order[k];
                off = offsets[k];
                const uint2 rc = rect[k];   // the rectangle K1 counted (depth order)
                xy0 = rc.x;
                w = rc.y & 0xFFFFu;
                n = w * (rc.y >> 16);
                if (n == 0) w = 1;
            }
            emit_warp_run(off, n, id, xy0, w, lane, s_scratch[warp], keys, vals, tile_count, s_count, smem_hist, grid.x,
                          capacity);
        }
    }
    if (smem_hist) {
        __syncthreads();
        for (int t = threadIdx.x; t < num_tiles; t += EMIT_THREADS) {
            const uint32_t c = s_count[t];
            if (c) atomicAdd(&tile_count[t], c);
        }
    }
}

// ranges[t] = [sum of counts before t, + count) — one block, replaces identifyTileRanges.
// `capacity`: the pair arrays hold that many instances; after an overflowing graph-safe forward (flagged, the caller
// discards the step) the ranges are clamped so that the render kernels still only read initialised list entries.
__global__ void __launch_bounds__(1024)
tile_ranges_from_counts_kernel(int num_tiles, const uint32_t* __restrict__ tile_count,
                               uint2* __restrict__ ranges, const uint32_t capacity) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < num_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t c = t < num_tiles ? tile_count[t] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t v = s_warp[lane], wi = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += u;
            }
            s_warp[lane] = wi - v;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint32_t start = s_carry + s_warp[warp] + incl - c;
        if (t < num_tiles) {
            const uint32_t lo = min(start, capacity), hi = min(start + c, capacity);
            ranges[t] = hi > lo ? make_uint2(lo, hi) : make_uint2(0u, 0u);
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = start + c;
        __syncthreads();
    }
}

// Digit BASES (exclusive prefix of the global digit histogram) of the tile-id radix passes from the per-tile instance
// counts: one block per pass, no read of the keys.
__global__ void __launch_bounds__(RS_RADIX)
tile_digit_bases_kernel(int num_tiles, const uint32_t* __restrict__ tile_count, int bpp,
                        uint32_t* __restrict__ digit_base /*[passes][RS_RADIX]*/) {
    __shared__ uint32_t h[RS_RADIX];
    const int p = blockIdx.x;
    const uint32_t mask = (1u << bpp) - 1u;
    h[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < num_tiles; t += RS_RADIX) {
        const uint32_t c = tile_count[t];
        if (c) atomicAdd(&h[((uint32_t)t >> (p * bpp)) & mask], c);
    }
    __syncthreads();
    uint32_t tot;
    const uint32_t ex = block_excl_scan<RS_RADIX>(h[threadIdx.x], &tot);
    digit_base[p * RS_RADIX + threadIdx.x] = ex;
}

__global__ void tile_copy_flags_kernel(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    if (threadIdx.x == 0) *out = (a ? *a : 0u) | (b ? *b : 0u);
}
__global__ void tile_or_flag_kernel(const uint32_t* a, uint32_t* out) {
    if (threadIdx.x == 0 && *a) atomicOr(out, *a);
}
// Preprojected forward: the tile rectangles were cut with ASSUMED bounds of the sampling offsets (words: order keys
// of max ox, max -ox, max oy, max -oy); offsets outside them could see Gaussians in tiles that were not instantiated.
__global__ void check_projection_bounds_kernel(const uint32_t* __restrict__ actual, const uint32_t* __restrict__ assumed,
                                               uint32_t* __restrict__ flags) {
    if (threadIdx.x != 0) return;
    const SampleBounds a = load_sample_bounds(actual), b = load_sample_bounds(assumed);
    if (!(a.max_x <= b.max_x) || !(a.min_x >= b.min_x) || !(a.max_y <= b.max_y) || !(a.min_y >= b.min_y)) atomicOr(flags, 2u);
}
// status_dev[0..3] = {num_rendered, prefiltered violated, look-back time-out, capacity overflow}
__global__ void copy_status_kernel(const uint32_t* __restrict__ totals, uint32_t* __restrict__ status,
                                   const uint32_t* e0, const uint32_t* e1, const uint32_t* e2) {
    if (threadIdx.x >= 4) return;
    uint32_t v = totals[threadIdx.x];
    if (threadIdx.x == 2) v |= (e0 ? *e0 : 0u) | (e1 ? *e1 : 0u) | (e2 ? *e2 : 0u);   // look-back time-outs of every stage
    status[threadIdx.x] = v;
}

// ------------------------------------------------------------------ K6 -------------------
constexpr int RENDER_BATCH = 256;

__global__ void __launch_bounds__(TILE_PIX)
render_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      const int W, const int H, const float4* __restrict__ rec,
                      const float* __restrict__ bg_color, const float* __restrict__ sampling_offsets,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                      float* __restrict__ out_color, float* __restrict__ out_depth) {
    __shared__ float4 s_r0[2][RENDER_BATCH];
    __shared__ float4 s_r1[2][RENDER_BATCH];
    __shared__ float4 s_r2[2][RENDER_BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp = 8x4 pixel block inside the 16x16 tile
    const uint32_t px = blockIdx.x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = blockIdx.y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;

    float2 pixf = make_float2((float)px, (float)py);
    if (inside && sampling_offsets != nullptr) {
        const float2 o = *reinterpret_cast<const float2*>(sampling_offsets + 2 * (size_t)pix_id);
        pixf.x = (float)px + o.x;
        pixf.y = (float)py + o.y;
    }
    // bounding box of this warp's sample positions
    const float inf = __int_as_float(0x7f800000);
    float bx0 = inside ? pixf.x : inf, bx1 = inside ? pixf.x : -inf;
    float by0 = inside ? pixf.y : inf, by1 = inside ? pixf.y : -inf;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d));
        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
    }

    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const int n = (int)(range.y - range.x);
    const int rounds = (n + RENDER_BATCH - 1) / RENDER_BATCH;

    auto prefetch = [&](int b) {
        const int p = b * RENDER_BATCH + tid;
        if (p < n) {
            const uint32_t id = point_list[range.x + p];
            const float4* src = rec + 3 * (size_t)id;
            cp_async16(&s_r0[b & 1][tid], src);
            cp_async16(&s_r1[b & 1][tid], src + 1);
            cp_async16(&s_r2[b & 1][tid], src + 2);
        }
        cp_async_commit();
    };

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;

    if (rounds > 0) prefetch(0);
    for (int b = 0; b < rounds; ++b) {
        if (b + 1 < rounds) {
            prefetch(b + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        bool warp_done = __all_sync(0xffffffffu, done);
        if (!warp_done) {
            const int cnt = min(RENDER_BATCH, n - b * RENDER_BATCH);
            const float4* r0 = s_r0[b & 1];
            const float4* r1 = s_r1[b & 1];
            const float4* r2 = s_r2[b & 1];
            for (int j0 = 0; j0 < cnt && !warp_done; j0 += 32) {
                // 32 lanes test 32 Gaussians against the warp's sample box
                const int jj = j0 + lane;
                bool hit = false;
                if (jj < cnt) {
                    const float4 a = r0[jj];
                    const float hy = r2[jj].w;
                    // "not certainly outside" so that NaNs fall through to the exact test
                    hit = !((a.x + a.w < bx0) || (a.x - a.w > bx1) || (a.y + hy < by0) ||
                            (a.y - hy > by1));
                    if (hit) {  // the box overlaps: decide exactly on the ellipse
                        const float4 co = r1[jj];
                        hit = !ellipse_misses_rect(a.x, a.y, co.x, co.y, co.z, co.w, bx0, bx1, by0, by1);
                    }
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int j = j0 + __ffs(m) - 1;
                    m &= m - 1;
                    if (!done) {
                        const float4 a = r0[j];
                        const float4 con_o = r1[j];
                        // forward.cu:343-346, FMA structure pinned to the reference's SASS
                        const float dx = __fsub_rn(a.x, pixf.x);
                        const float dy = __fsub_rn(a.y, pixf.y);
                        const float sy = __fmul_rn(__fmul_rn(con_o.z, dy), dy);
                        const float sq = __fmaf_rn(dx, __fmul_rn(con_o.x, dx), sy);
                        const float cr = __fmul_rn(__fmul_rn(con_o.y, dx), dy);
                        const float power = __fmaf_rn(sq, -0.5f, -cr);
                        if (!(power > 0.0f)) {
                            const float alpha = fminf(0.99f, __fmul_rn(con_o.w, expf(power)));
                            if (!(alpha < 1.0f / 255.0f)) {
                                const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                                if (test_T < 0.0001f) {
                                    done = true;
                                } else {
                                    const float4 c = r2[j];
                                    C0 = __fmaf_rn(T, __fmul_rn(alpha, c.x), C0);
                                    C1 = __fmaf_rn(T, __fmul_rn(alpha, c.y), C1);
                                    C2 = __fmaf_rn(T, __fmul_rn(alpha, c.z), C2);
                                    Dp = __fmaf_rn(T, __fmul_rn(alpha, a.z), Dp);
                                    T = test_T;
                                    last_contributor = (uint32_t)(b * RENDER_BATCH + j + 1);
                                }
                            }
                        }
                    }
                    if (__all_sync(0xffffffffu, done)) {
                        warp_done = true;
                        break;
                    }
                }
            }
        }
        // also fences the buffer that the next iteration's prefetch will overwrite
        if (__syncthreads_and(warp_done)) break;
    }
    cp_async_wait<0>();

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const size_t HW = (size_t)H * W;
        out_color[pix_id] = __fmaf_rn(bg_color[0], T, C0);
        out_color[HW + pix_id] = __fmaf_rn(bg_color[1], T, C1);
        out_color[2 * HW + pix_id] = __fmaf_rn(bg_color[2], T, C2);
        out_depth[pix_id] = Dp;
    }
}

// ---- K6, one warp per CTA ---------------------------------------------------------------------------------
// The tile kernel above stages a tile's list once for its eight warps and pays for it with a block barrier per
// batch: a warp whose 8x4 pixels saturate early (or see few Gaussians) idles at every barrier until the slowest
// warp of the tile is through, and its registers stay resident (ncu: ~40 % of the warp-time of K6 / K7 is spent at
// that barrier, issue slots 73-81 % used).  Here every 8x4 pixel block is its own single-warp CTA: it stages the
// tile's list for itself (3 x 32 records in flight, cp.async; the extra L2 -> SM traffic is ~2 GB per launch, a few
// percent of the L2 bandwidth), synchronises with nobody, and leaves as soon as ITS pixels are done.  Same per-pixel
// instruction sequence as render_forward_kernel: images, final_T and n_contrib are bit-identical.
constexpr int WARP_STAGES = 3;

__global__ void __launch_bounds__(32)
render_forward_warp_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                           const int W, const int H, const int tiles_x, const float4* __restrict__ rec,
                           const float* __restrict__ bg_color, const float* __restrict__ sampling_offsets,
                           float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                           float* __restrict__ out_color, float* __restrict__ out_depth) {
    __shared__ float4 s_r0[WARP_STAGES][32];
    __shared__ float4 s_r1[WARP_STAGES][32];
    __shared__ float4 s_r2[WARP_STAGES][32];

    const int lane = threadIdx.x;
    // blockIdx.x: 8-pixel column block, blockIdx.y: 4-pixel row block
    const uint32_t px = blockIdx.x * 8 + (lane & 7);
    const uint32_t py = blockIdx.y * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;

    float2 pixf = make_float2((float)px, (float)py);
    if (inside && sampling_offsets != nullptr) {
        const float2 o = *reinterpret_cast<const float2*>(sampling_offsets + 2 * (size_t)pix_id);
        pixf.x = (float)px + o.x;
        pixf.y = (float)py + o.y;
    }
    const float inf = __int_as_float(0x7f800000);
    float bx0 = inside ? pixf.x : inf, bx1 = inside ? pixf.x : -inf;
    float by0 = inside ? pixf.y : inf, by1 = inside ? pixf.y : -inf;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d));
        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
    }

    const uint2 range = ranges[(blockIdx.y >> 2) * tiles_x + (blockIdx.x >> 1)];
    // a block that lies entirely outside the image (W, H not multiples of 8 / 4) has nothing to blend or to write
    const int n = __any_sync(0xffffffffu, inside) ? (int)(range.y - range.x) : 0;
    const int rounds = (n + 31) / 32;

    auto prefetch = [&](int b) {
        const int p = b * 32 + lane;
        if (b < rounds && p < n) {
            const uint32_t id = point_list[range.x + p];
            const float4* src = rec + 3 * (size_t)id;
            const int st = b % WARP_STAGES;
            cp_async16(&s_r0[st][lane], src);
            cp_async16(&s_r1[st][lane], src + 1);
            cp_async16(&s_r2[st][lane], src + 2);
        }
        cp_async_commit();   // one group per round, empty or not: the wait below counts groups
    };

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;

    prefetch(0);
    prefetch(1);
    for (int b = 0; b < rounds; ++b) {
        prefetch(b + 2);
        cp_async_wait<2>();   // round b has landed (two younger groups may still be in flight)
        __syncwarp();
        const int st = b % WARP_STAGES;
        const float4* r0 = s_r0[st];
        const float4* r1 = s_r1[st];
        const float4* r2 = s_r2[st];
        const int cnt = min(32, n - b * 32);
        bool hit = false;
        if (lane < cnt) {
            const float4 a = r0[lane];
            const float hy = r2[lane].w;
            // "not certainly outside" so that NaNs fall through to the exact test
            hit = !((a.x + a.w < bx0) || (a.x - a.w > bx1) || (a.y + hy < by0) || (a.y - hy > by1));
            if (hit) {  // the box overlaps: decide exactly on the ellipse
                const float4 co = r1[lane];
                hit = !ellipse_misses_rect(a.x, a.y, co.x, co.y, co.z, co.w, bx0, bx1, by0, by1);
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            if (!done) {
                const float4 a = r0[j];
                const float4 con_o = r1[j];
                // forward.cu:343-346, FMA structure pinned to the reference's SASS
                const float dx = __fsub_rn(a.x, pixf.x);
                const float dy = __fsub_rn(a.y, pixf.y);
                const float sy = __fmul_rn(__fmul_rn(con_o.z, dy), dy);
                const float sq = __fmaf_rn(dx, __fmul_rn(con_o.x, dx), sy);
                const float cr = __fmul_rn(__fmul_rn(con_o.y, dx), dy);
                const float power = __fmaf_rn(sq, -0.5f, -cr);
                if (!(power > 0.0f)) {
                    const float alpha = fminf(0.99f, __fmul_rn(con_o.w, expf(power)));
                    if (!(alpha < 1.0f / 255.0f)) {
                        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            const float4 c = r2[j];
                            C0 = __fmaf_rn(T, __fmul_rn(alpha, c.x), C0);
                            C1 = __fmaf_rn(T, __fmul_rn(alpha, c.y), C1);
                            C2 = __fmaf_rn(T, __fmul_rn(alpha, c.z), C2);
                            Dp = __fmaf_rn(T, __fmul_rn(alpha, a.z), Dp);
                            T = test_T;
                            last_contributor = (uint32_t)(b * 32 + j + 1);
                        }
                    }
                }
            }
        }
        // all 32 pixels saturated: checked once per round (a vote per hit cost 3.6 % of the kernel's instructions; the at
        // most 31 hits a finished warp still walks through skip their body)
        if (__all_sync(0xffffffffu, done)) break;
        __syncwarp();   // everybody is done reading stage st before round b + 3 overwrites it
    }
    cp_async_wait<0>();

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const size_t HW = (size_t)H * W;
        out_color[pix_id] = __fmaf_rn(bg_color[0], T, C0);
        out_color[HW + pix_id] = __fmaf_rn(bg_color[1], T, C1);
        out_color[2 * HW + pix_id] = __fmaf_rn(bg_color[2], T, C2);
        out_depth[pix_id] = Dp;
    }
}

// ------------------------------------------------------------------ host -----------------
// Which flavour of the sort / scan primitives each binning stage uses: 0 = multi-kernel (histogram,
// table scan, scatter), 1 = single-kernel passes with decoupled look-back.  Measured on B200
// (profiles/r02_sort_chain.md): with the windowed look-back both sorts win single-kernel (depth sort of P keys: 5
// launches instead of 20; tile partition of R pairs: no key histograms at all, the digit bases follow from the
// per-tile counts).  The offsets scan is fused into emit_instances_kernel; where it runs on its own (preprojected
// views, WAST3D_EMIT_SCAN=0) the multi-kernel flavour is the default.  WAST3D_SORT_MODE=0|1 forces one flavour
// everywhere (experiments, tests).  In the synchronous protocol a look-back time-out behind the host read of
// num_rendered (emit, tile partition: never observed, it takes a hung GPU) is only recorded in the geometry buffer's
// totals[2]; wast3d_raster_forward_async reports it through status_dev[2].
enum SortStage { STAGE_DEPTH = 0, STAGE_SCAN = 1, STAGE_TILE = 2 };
static int sort_mode(SortStage stage) {
    static const int forced = getenv("WAST3D_SORT_MODE") ? atoi(getenv("WAST3D_SORT_MODE")) : -1;
    // per-stage override (A/B measurements): WAST3D_SORT_DEPTH / WAST3D_SORT_SCAN / WAST3D_SORT_TILE = 0|1
    static const char* const names[3] = {"WAST3D_SORT_DEPTH", "WAST3D_SORT_SCAN", "WAST3D_SORT_TILE"};
    static const int per_stage[3] = {getenv(names[0]) ? atoi(getenv(names[0])) : -1,
                                     getenv(names[1]) ? atoi(getenv(names[1])) : -1,
                                     getenv(names[2]) ? atoi(getenv(names[2])) : -1};
    if (per_stage[stage] == 0 || per_stage[stage] == 1) return per_stage[stage];
    if (forced == 0 || forced == 1) return forced;
    return stage == STAGE_SCAN ? 0 : 1;
}

// Tile cut (tile_rect_cut, raster_math.cuh): 1 = instantiate a Gaussian only in the tiles that can see
// alpha >= 1/255 (default), 0 = the reference's full radius rectangles (auxiliary.h:46-56), which
// reproduces the reference's num_rendered and point list bit for bit.  WAST3D_TILE_CUT=0|1 sets the
// initial value; wast3d_set_tile_cut() changes it at run time (tests, A/B measurements).
static std::atomic<int> g_tile_cut{-1};
static int tile_cut_mode() {
    int m = g_tile_cut.load();
    if (m < 0) {
        const char* e = getenv("WAST3D_TILE_CUT");
        m = (e && atoi(e) == 0) ? 0 : 1;
        g_tile_cut.store(m);
    }
    return m;
}

static int tile_sort_passes(uint32_t num_tiles, int* bits_per_pass) {
    const int bits = bits_for(num_tiles);
    const int passes = (bits + 7) / 8;
    *bits_per_pass = (bits + passes - 1) / passes;
    return passes;
}

const uint32_t* point_list_ptr(const BinningState& b, uint32_t num_tiles) {
    int bpp;
    return (tile_sort_passes(num_tiles, &bpp) & 1) ? b.vals_b : b.vals_a;
}
int validate_params(const wast3d_raster_params* p, bool forward) {
    if (!p) return WAST3D_ERR_INVALID_ARGUMENT;
    if (p->P < 0 || p->width <= 0 || p->height <= 0) return WAST3D_ERR_INVALID_ARGUMENT;
    if (p->P == 0) return WAST3D_OK;
    if (!p->means3D || (forward && !p->opacities) || !p->viewmatrix || !p->projmatrix || !p->campos ||
        !p->background)
        return WAST3D_ERR_INVALID_ARGUMENT;
    if ((p->shs == nullptr) == (p->colors_precomp == nullptr)) return WAST3D_ERR_INVALID_ARGUMENT;
    const bool has_sr = p->scales != nullptr && p->rotations != nullptr;
    if (has_sr == (p->cov3D_precomp != nullptr)) return WAST3D_ERR_INVALID_ARGUMENT;
    if (p->shs) {
        if (p->D < 0 || p->D > 3 || p->M < (p->D + 1) * (p->D + 1) || p->M > 16)
            return WAST3D_ERR_INVALID_ARGUMENT;
    }
    if (p->raw_params) {
        if (!p->shs || !has_sr || p->colors_precomp || p->cov3D_precomp || (p->M > 1 && !p->shs_rest))
            return WAST3D_ERR_INVALID_ARGUMENT;
    }
    return WAST3D_OK;
}

}  // namespace w3d

using namespace w3d;

// capacity < 0: the reference's protocol (one blocking read of num_rendered sizes the binning buffer).
// capacity >= 1: graph-safe — nothing is read back; the binning buffer holds `capacity` instances, the R-dependent
// kernels take the instance count from device memory, status_dev receives {R, flags, time-out, overflow}.
static int raster_forward_impl(const wast3d_raster_params* prm, wast3d_alloc_fn geom_alloc,
                               void* geom_user, wast3d_alloc_fn binning_alloc,
                               void* binning_user, wast3d_alloc_fn img_alloc, void* img_user,
                               float* out_color, float* out_depth, int* radii,
                               int* num_rendered_host, long long capacity, uint32_t* status_dev, void* stream_v) {
    cudaStream_t s = (cudaStream_t)stream_v;
    int st = validate_params(prm, true);
    if (st != WAST3D_OK) return st;
    if (!geom_alloc || !binning_alloc || !img_alloc || !out_color || !out_depth || !num_rendered_host)
        return WAST3D_ERR_INVALID_ARGUMENT;
    const bool async = capacity >= 0;
    if (async && (capacity < 1 || capacity > 0x7FFFFFFFll)) return WAST3D_ERR_INVALID_ARGUMENT;
    const int P = prm->P, W = prm->width, H = prm->height;
    const bool debug = prm->debug != 0;
    const size_t N = (size_t)W * H;
    *num_rendered_host = 0;
    if (P == 0) {
        // rasterize_points.cu:69-70,83: outputs stay at their zero fill, buffers stay empty
        W3D_CUDA_TRY(cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), s));
        W3D_CUDA_TRY(cudaMemsetAsync(out_depth, 0, N * sizeof(float), s));
        if (async && status_dev) W3D_CUDA_TRY(cudaMemsetAsync(status_dev, 0, 4 * sizeof(uint32_t), s));
        return WAST3D_OK;
    }
    const dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y, 1);
    const uint32_t num_tiles = grid.x * grid.y;

    size_t geom_bytes, img_bytes;
    GeomState::carve(nullptr, P, &geom_bytes);
    ImageState::carve(nullptr, N, num_tiles, &img_bytes);
    void* geom_chunk = geom_alloc(geom_bytes, geom_user);
    void* img_chunk = img_alloc(img_bytes, img_user);
    if (!geom_chunk || !img_chunk) return WAST3D_ERR_ALLOC;
    GeomState g = GeomState::carve(geom_chunk, P, nullptr);
    ImageState im = ImageState::carve(img_chunk, N, num_tiles, nullptr);
    if (radii == nullptr) radii = g.internal_radii;

    // rasterizer_impl.cu:224-225
    const float focal_y = H / (2.0f * prm->tan_fovy);
    const float focal_x = W / (2.0f * prm->tan_fovx);

    // preprojected (wast3d_raster_backward_raw_adam_next wrote K1's outputs for this view into the buffer): words
    // 16..19 of totals carry the offset bounds that projection assumed and must survive the reset
    const bool pre = prm->preprojected != 0;
    if (pre && (!prm->raw_params || prm->colour_wait_event != nullptr)) return WAST3D_ERR_INVALID_ARGUMENT;
    W3D_CUDA_TRY(cudaMemsetAsync(g.totals, 0, (pre ? 16 : 32) * sizeof(uint32_t), s));
    const bool cut_tiles = tile_cut_mode() != 0;
    uint32_t* sample_bound_words = g.totals + 8;  // zero = no offsets
    if (pre) {
        ProfScope ps(PS_PREPROCESS, s);
        if (cut_tiles) {
            if (prm->sampling_offsets != nullptr) {
                sample_bounds_kernel<<<148 * 4, 256, 0, s>>>(N, prm->sampling_offsets, sample_bound_words);
                W3D_AFTER_LAUNCH(s, debug);
            }
            check_projection_bounds_kernel<<<1, 32, 0, s>>>(sample_bound_words, g.totals + 16, g.totals + 1);
            W3D_AFTER_LAUNCH(s, debug);
        }
    } else {
    ProfScope ps(PS_PREPROCESS, s);
    if (cut_tiles && prm->sampling_offsets != nullptr) {
        sample_bounds_kernel<<<148 * 4, 256, 0, s>>>(N, prm->sampling_offsets, sample_bound_words);
        W3D_AFTER_LAUNCH(s, debug);
    }
    // WAST3D_K1_VEC: 1 (default) = means3D / scales through coalesced 16-byte loads, 0 = strided scalar loads (A/B)
    static const bool k1_vec = !(getenv("WAST3D_K1_VEC") && atoi(getenv("WAST3D_K1_VEC")) == 0);
    // colour_wait_event: SH -> RGB moves into its own kernel behind that event (see sh_colour_kernel)
    const bool defer_colour = prm->colour_wait_event != nullptr && prm->colors_precomp == nullptr;
    auto pre = prm->raw_params ? (defer_colour ? preprocess_kernel<true, false> : preprocess_kernel<true, true>)
                               : (defer_colour ? preprocess_kernel<false, false> : preprocess_kernel<false, true>);
    pre<<<(P + PRE_THREADS - 1) / PRE_THREADS, PRE_THREADS, 0, s>>>(
        P, prm->D, prm->M, prm->means3D, prm->scales, prm->scale_modifier, prm->rotations,
        prm->opacities, prm->shs, prm->shs_rest, prm->cov3D_precomp, prm->colors_precomp, prm->viewmatrix,
        prm->projmatrix, prm->campos, W, H, prm->tan_fovx, prm->tan_fovy, focal_x, focal_y, grid,
        prm->prefiltered != 0, radii, g.rec, g.depth_key, g.tiles_touched, g.clamped, g.rect, g.totals,
        sample_bound_words, cut_tiles, k1_vec);
    W3D_AFTER_LAUNCH(s, debug);
    }

    // depth order: 4 stable 8-bit passes on the float bits (positive floats order like uints),
    // each ONE look-back kernel; the four global digit histograms come from one read of the keys
    {
    ProfScope ps(PS_DEPTH_SORT, s);
    const int shifts[4] = {0, 8, 16, 24}, bits[4] = {8, 8, 8, 8};
    uint32_t* const kin[4] = {g.depth_key, g.key_tmp, g.depth_key, g.key_tmp};
    uint32_t* const kout[4] = {g.key_tmp, g.depth_key, g.key_tmp, nullptr};
    uint32_t* const vin[4] = {nullptr, g.order_b, g.order_a, g.order_b};
    uint32_t* const vout[4] = {g.order_b, g.order_a, g.order_b, g.order_a};
    if (sort_mode(STAGE_DEPTH) == 1) {
        st = onesweep_prepare(g.sort_ws, P, 4, s);
        if (st) return st;
        st = onesweep_hist(g.depth_key, P, 4, shifts, bits, g.sort_ws, s, debug);
        if (st) return st;
    }
    for (int p = 0; p < 4; ++p) {
        // the last pass also lays the tile rectangles (and their tile counts) out in depth order
        GatherRect gather;
        if (p == 3) { gather.src = g.rect; gather.dst = g.rect_sorted; gather.cnt = g.count_sorted; }
        if (sort_mode(STAGE_DEPTH) == 1)
            st = onesweep_pass(kin[p], vin[p], kout[p], vout[p], P, shifts[p], bits[p], g.sort_ws, 4, p, s, debug,
                               nullptr, gather);
        else
            st = radix_pass_u32(kin[p], vin[p], kout[p], vout[p], P, shifts[p], bits[p], g.sort_ws,
                                g.sort_ws + rs_hist_words(P), s, debug, nullptr, gather);
        if (st) return st;
    }
    }
    // Offsets of the Gaussians' instance runs (rasterizer_impl.cu:279).  Default: fused into emit_instances_kernel
    // (K1 already summed num_rendered); WAST3D_EMIT_SCAN=0, a preprojected view (its projection ran before totals was
    // reset) and counts the look-back status words cannot hold take the stand-alone scan.
    static const bool emit_scan_env = !(getenv("WAST3D_EMIT_SCAN") && atoi(getenv("WAST3D_EMIT_SCAN")) == 0);
    bool fused_scan = emit_scan_env && !pre && (!async || capacity <= (long long)OS_VALUE);
    auto standalone_scan = [&]() -> int {
        ProfScope ps(PS_SCAN, s);
        if (sort_mode(STAGE_SCAN) == 1)
            return scan_exclusive_lookback_u32(g.count_sorted, nullptr, g.offsets, P, g.scan_ws, g.totals, s, debug);
        return scan_exclusive_u32(g.count_sorted, nullptr, g.offsets, P, g.scan_ws, g.totals, s, debug);
    };
    if (!fused_scan) {
        st = standalone_scan();
        if (st) return st;
    }
    // look-back time-outs (never expected) of the depth sort and the scan, read with num_rendered; the graph-safe
    // forward collects every stage's error word in its last kernel instead (copy_status_kernel)
    const uint32_t* err_words[4] = {nullptr, nullptr, nullptr, nullptr};
    if (sort_mode(STAGE_DEPTH) == 1) err_words[0] = onesweep_error_word(g.sort_ws, P, 4);
    if (!fused_scan && sort_mode(STAGE_SCAN) == 1) err_words[1] = g.scan_ws + 1;
    if (!async && (sort_mode(STAGE_DEPTH) == 1 || (!fused_scan && sort_mode(STAGE_SCAN) == 1))) {
        tile_copy_flags_kernel<<<1, 32, 0, s>>>(sort_mode(STAGE_DEPTH) == 1 ? onesweep_error_word(g.sort_ws, P, 4) : nullptr,
                                               !fused_scan && sort_mode(STAGE_SCAN) == 1 ? g.scan_ws + 1 : nullptr,
                                               g.totals + 2);
        W3D_AFTER_LAUNCH(s, debug);
    }

    uint32_t R;   // instances the binning buffer is carved for (== num_rendered in the synchronous protocol)
    if (!async) {
        // the one blocking read the reference also has (rasterizer_impl.cu:283)
        uint32_t host_totals[3] = {0, 0, 0};
        W3D_CUDA_TRY(cudaMemcpyAsync(host_totals, g.totals, sizeof(host_totals), cudaMemcpyDeviceToHost, s));
        W3D_CUDA_TRY(cudaStreamSynchronize(s));
        if (host_totals[1] & 1u) return WAST3D_ERR_INVALID_ARGUMENT;  // prefiltered violated
        if (host_totals[1] & 2u) return WAST3D_ERR_STALE_PROJECTION;
        if (host_totals[2]) {
            set_last_cuda_error(cudaErrorLaunchTimeout, __FILE__, __LINE__);
            return WAST3D_ERR_CUDA;
        }
        if (host_totals[0] > 0x7FFFFFFFu) return WAST3D_ERR_OVERFLOW;
        R = host_totals[0];
        if (fused_scan && R > OS_VALUE) {   // more instances than a look-back status word holds: scan on its own
            fused_scan = false;
            st = standalone_scan();
            if (st) return st;
        }
    } else {
        R = (uint32_t)capacity;
    }
    *num_rendered_host = (int)R;
    const uint32_t* n_dev = async ? g.totals : nullptr;   // device-side instance count for the R-sized kernels

    size_t bin_bytes;
    BinningState::carve(nullptr, R, &bin_bytes);
    void* bin_chunk = binning_alloc(bin_bytes, binning_user);
    if (!bin_chunk) return WAST3D_ERR_ALLOC;
    BinningState bn = BinningState::carve(bin_chunk, R, nullptr);

    if (R > 0) {
        W3D_CUDA_TRY(cudaMemsetAsync(im.tile_count, 0, num_tiles * sizeof(uint32_t), s));
        {
        ProfScope ps(PS_EMIT, s);
        const bool smem_hist = num_tiles <= (uint32_t)EMIT_MAX_SMEM_TILES;
        const size_t smem = smem_hist ? num_tiles * sizeof(uint32_t) : 0;
        static_assert(EMIT_MAX_SMEM_TILES * sizeof(uint32_t) <= 48 * 1024,
                      "emit_instances_kernel's counters must fit the default dynamic shared-memory limit");
        const int n_chunks = (P + 31) / 32;
        int blocks = 148 * 4;
        if (blocks > (n_chunks + 7) / 8) blocks = (n_chunks + 7) / 8;
        if (fused_scan) {
            W3D_CUDA_TRY(cudaMemsetAsync(g.scan_ws, 0, emit_scan_workspace_words(P) * sizeof(uint32_t), s));
            const size_t fsmem = (smem_hist ? (size_t)(num_tiles + 3) / 4 * 4 * sizeof(uint32_t) : 0) + sizeof(EmitGroup);
            // up to 80 KB of dynamic shared memory: opt in (per device, so on every launch; the call only records a number)
            W3D_CUDA_TRY(cudaFuncSetAttribute(emit_instances_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(EMIT_MAX_SMEM_TILES * sizeof(uint32_t) + sizeof(EmitGroup))));
            const int n_groups = (P + EMIT_GROUP - 1) / EMIT_GROUP;
            blocks = n_groups < 148 * 4 ? n_groups : 148 * 4;
            emit_instances_kernel<true><<<blocks, EMIT_THREADS, fsmem, s>>>(P, g.order_a, nullptr, g.rect_sorted, bn.keys_a,
                                                                         bn.vals_a, im.tile_count, grid, (int)num_tiles,
                                                                         smem_hist, R, g.totals, g.scan_ws);
            W3D_AFTER_LAUNCH(s, debug);
            // a look-back time-out (never expected) joins the flags; the synchronous protocol has read them already
            err_words[1] = g.scan_ws + 1;
            if (!async) tile_or_flag_kernel<<<1, 32, 0, s>>>(g.scan_ws + 1, g.totals + 2);
        } else {
            emit_instances_kernel<false><<<blocks, EMIT_THREADS, smem, s>>>(P, g.order_a, g.offsets, g.rect_sorted, bn.keys_a,
                                                                          bn.vals_a, im.tile_count, grid, (int)num_tiles,
                                                                          smem_hist, R, g.totals, nullptr);
        }
        W3D_AFTER_LAUNCH(s, debug);
        }
        ProfScope* pts = new ProfScope(PS_TILE_SORT, s);
        int bpp;
        const int passes = tile_sort_passes(num_tiles, &bpp);
        if (passes > TILE_SORT_MAX_PASSES) { delete pts; return WAST3D_ERR_OVERFLOW; }
        const bool tile_lookback = sort_mode(STAGE_TILE) == 1;
        if (tile_lookback) {
            st = onesweep_prepare(bn.sort_ws, R, passes, s);
            if (st) { delete pts; return st; }
            // the per-pass global digit histograms follow from the per-tile counts: no read of the keys
            tile_digit_bases_kernel<<<passes, RS_RADIX, 0, s>>>((int)num_tiles, im.tile_count, bpp,
                                                                 onesweep_digit_hist(bn.sort_ws, R, passes, 0));
            W3D_AFTER_LAUNCH(s, debug);
        }
        uint32_t *kin = bn.keys_a, *vin = bn.vals_a, *kout = bn.keys_b, *vout = bn.vals_b;
        for (int p = 0; p < passes; ++p) {
            // the last pass does not need to write the sorted tile ids: ranges come from the counts
            if (tile_lookback)
                st = onesweep_pass(kin, vin, p + 1 < passes ? kout : nullptr, vout, R, p * bpp, bpp, bn.sort_ws,
                                   passes, p, s, debug, n_dev);
            else
                st = radix_pass_u32(kin, vin, p + 1 < passes ? kout : nullptr, vout, R, p * bpp, bpp, bn.sort_ws,
                                    bn.sort_ws + rs_hist_words(R), s, debug, n_dev);
            if (st) { delete pts; return st; }
            uint32_t* t;
            t = kin; kin = kout; kout = t;
            t = vin; vin = vout; vout = t;
        }
        if (tile_lookback) {   // a look-back time-out (never expected) joins the flags of the depth sort and the scan
            err_words[2] = onesweep_error_word(bn.sort_ws, R, passes);
            if (!async) {
                tile_or_flag_kernel<<<1, 32, 0, s>>>(err_words[2], g.totals + 2);
                W3D_AFTER_LAUNCH(s, debug);
            }
        }
        delete pts;
        ProfScope ps(PS_RANGES, s);
        tile_ranges_from_counts_kernel<<<1, 1024, 0, s>>>((int)num_tiles, im.tile_count, im.ranges, R);
        W3D_AFTER_LAUNCH(s, debug);
    } else {
        W3D_CUDA_TRY(cudaMemsetAsync(im.ranges, 0, num_tiles * sizeof(uint2), s));
    }

    if (prm->colour_wait_event != nullptr) {
        // the event also orders everything after it (render, backward) behind the parameter exchange
        W3D_CUDA_TRY(cudaStreamWaitEvent(s, (cudaEvent_t)prm->colour_wait_event, 0));
        if (prm->colors_precomp == nullptr) {
            ProfScope ps(PS_PREPROCESS, s);
            auto col = prm->raw_params ? sh_colour_kernel<true> : sh_colour_kernel<false>;
            col<<<(P + PRE_THREADS - 1) / PRE_THREADS, PRE_THREADS, 0, s>>>(
                P, prm->D, prm->M, prm->means3D, prm->shs, prm->shs_rest, prm->campos, g.rec, g.clamped);
            W3D_AFTER_LAUNCH(s, debug);
        }
    }
    {
    ProfScope ps_render(PS_RENDER_FWD, s);
    // WAST3D_K6_MODE: 1 (default) = one warp (8x4 pixels) per CTA, 0 = one 16x16 tile per CTA (A/B measurements)
    static const int k6_mode = getenv("WAST3D_K6_MODE") ? atoi(getenv("WAST3D_K6_MODE")) : 1;
    if (k6_mode == 1) {
        const dim3 wgrid(2 * grid.x, 4 * grid.y, 1);   // the 8x4 blocks of whole tiles (out-of-image lanes are masked)
        render_forward_warp_kernel<<<wgrid, 32, 0, s>>>(im.ranges, point_list_ptr(bn, num_tiles), W, H, (int)grid.x, g.rec,
                                                        prm->background, prm->sampling_offsets, im.final_T,
                                                        im.n_contrib, out_color, out_depth);
    } else {
        render_forward_kernel<<<grid, TILE_PIX, 0, s>>>(im.ranges, point_list_ptr(bn, num_tiles), W, H, g.rec,
                                                        prm->background, prm->sampling_offsets, im.final_T,
                                                        im.n_contrib, out_color, out_depth);
    }
    W3D_AFTER_LAUNCH(s, debug);
    }
    if (async && status_dev) {
        copy_status_kernel<<<1, 32, 0, s>>>(g.totals, status_dev, err_words[0], err_words[1], err_words[2]);
        W3D_AFTER_LAUNCH(s, debug);
    }
    return WAST3D_OK;
}

extern "C" int wast3d_raster_forward(const wast3d_raster_params* prm, wast3d_alloc_fn geom_alloc,
                                     void* geom_user, wast3d_alloc_fn binning_alloc,
                                     void* binning_user, wast3d_alloc_fn img_alloc, void* img_user,
                                     float* out_color, float* out_depth, int* radii,
                                     int* num_rendered_host, void* stream_v) {
    return raster_forward_impl(prm, geom_alloc, geom_user, binning_alloc, binning_user, img_alloc, img_user, out_color,
                               out_depth, radii, num_rendered_host, -1, nullptr, stream_v);
}

extern "C" int wast3d_raster_forward_async(const wast3d_raster_params* prm, wast3d_alloc_fn geom_alloc,
                                           void* geom_user, wast3d_alloc_fn binning_alloc,
                                           void* binning_user, wast3d_alloc_fn img_alloc, void* img_user,
                                           float* out_color, float* out_depth, int* radii,
                                           int instance_capacity, unsigned int* status_dev, void* stream_v) {
    if (instance_capacity < 1 || !status_dev) return WAST3D_ERR_INVALID_ARGUMENT;
    int carved = 0;
    return raster_forward_impl(prm, geom_alloc, geom_user, binning_alloc, binning_user, img_alloc, img_user, out_color,
                               out_depth, radii, &carved, instance_capacity, status_dev, stream_v);
}

extern "C" int wast3d_set_tile_cut(int mode) {
    const int prev = tile_cut_mode();
    if (mode == 0 || mode == 1) g_tile_cut.store(mode);
    return prev;
}

extern "C" int wast3d_mark_visible(int P, const float* means3D, const float* viewmatrix,
                                   const float* projmatrix, unsigned char* present, void* stream_v) {
    (void)projmatrix;  // checkFrustum only uses the view-space depth (auxiliary.h:154)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return WAST3D_ERR_INVALID_ARGUMENT;
    if (P == 0) return WAST3D_OK;
    cudaStream_t s = (cudaStream_t)stream_v;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
