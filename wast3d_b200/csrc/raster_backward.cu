// Backward rasteriser for sm_100a: tile replay (K7) and the fused per-Gaussian chain rule (K8+K9).
//
// What it replaces (reference paths under submodules/diff-gaussian-rasterization/):
//   renderCUDA (backward)       cuda_rasterizer/backward.cu:413-586   (10 atomicAdd per pixel-pair)
//   computeCov2DCUDA            cuda_rasterizer/backward.cu:144-274
//   preprocessCUDA (backward)   cuda_rasterizer/backward.cu:346-410 (+ :20-139 SH, :278-341 cov3D)
//   Rasterizer::backward        cuda_rasterizer/rasterizer_impl.cu:345-446
//   the ten torch::zeros of     rasterize_points.cu:157-166
//
// Design:
//  * K7 keeps the reference's pixel-parallel back-to-front replay (same skip decisions, same
//    recurrences) but never issues a per-pixel global atomic: the 10 partial gradients of a
//    (warp, Gaussian) pair are reduced across the 32 pixels with a value-halving butterfly
//    (13 shuffles instead of 50), added to a per-batch shared-memory accumulator (one shared
//    atomic instruction per warp and Gaussian), and flushed once per (tile, Gaussian) with three
//    16-byte vector atomics into a 48-byte gradient record.  Whole (warp, Gaussian) pairs that
//    cannot contribute (alpha cutoff box misses the warp's 8x4 pixels, or the Gaussian lies
//    behind every pixel's last contributor) are skipped before any per-pixel work, and batches
//    behind the tile's last contributor are never loaded.
//  * K8 and K9 are one kernel that consumes the gradient record, recomputes Sigma3D instead of
//    reading a stored copy, and writes EVERY element of the eight output tensors (zeros for
//    culled Gaussians), so no output memset is needed.
//  Summation order across pixels differs from the reference's atomics (both are unordered);
//  gradients agree to fp32 rounding, not bitwise (tests state the tolerance per tensor).
#include "project.cuh"
#include <atomic>
#include <cmath>
#include <cstdlib>

namespace w3d {

const uint32_t* point_list_ptr(const BinningState& b, uint32_t num_tiles);
int validate_params(const wast3d_raster_params* p, bool forward);

// Deterministic backward (tests): 0 = float atomics in K7 (default), 1 = fixed summation order.
// WAST3D_DETERMINISTIC=1 sets the initial value, wast3d_set_deterministic() changes it at run time.
static std::atomic<int> g_deterministic{-1};
static int deterministic_mode() {
    int m = g_deterministic.load();
    if (m < 0) {
        const char* e = getenv("WAST3D_DETERMINISTIC");
        m = (e && atoi(e) != 0) ? 1 : 0;
        g_deterministic.store(m);
    }
    return m;
}

constexpr int BWD_BATCH = 256;
// gradient record (48 bytes per Gaussian), accumulated by K7, consumed by K8+K9:
//  0 M_x   1 M_y   2 M_xx  3 M_xy | 4 M_yy  5 M_0  6 dviewdepth  7 - | 8 dcolor.r  9 dcolor.g  10 dcolor.b  11 -
// K7 sums MOMENTS of t = G * dL/dalpha over the pixels (M_0 = sum t, M_x = sum t dx, M_y = sum t dy, M_xx = sum t dx^2,
// M_xy = sum t dx dy, M_yy = sum t dy^2; dx, dy as in backward.cu:497): with the Gaussian's conic (A, B, C) and
// opacity o, which are constant over the pixels, the reference's per-pixel gradients (backward.cu:567-583) sum to
//   dL/dmean2D.x = -o (W/2) (A M_x + B M_y)      dL/dmean2D.y = -o (H/2) (C M_y + B M_x)
//   dL/dconic    = -o/2 (M_xx, M_xy, M_yy)       dL/dopacity  = M_0
// which K8+K9 evaluates once per Gaussian (moments_to_gradients) instead of K7 once per pixel: 6 multiplies per
// (pixel, Gaussian) pair instead of 18.

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Per-pixel state of the back-to-front replay (backward.cu:463-478,529-563).
struct ReplayPixel {
    float2 pixf;
    float T, T_final;
    float accum0, accum1, accum2;
    float last_alpha, last_c0, last_c1, last_c2;
    float dpix0, dpix1, dpix2, ddepth, bg_dot_dpixel;
    uint32_t last_contributor;
};

// One (pixel, Gaussian) pair of the replay.  Returns false when the reference skips the pair (behind the pixel's
// last contributor, power > 0, alpha < 1/255); otherwise advances the pixel state and writes the ten partials
// v = {t dx, t dy, t dx^2, t dx dy, t dy^2, t, dviewdepth, dcolor.rgb}.
__device__ __forceinline__ bool replay_pair(ReplayPixel& px, uint32_t pos, const float4 a, const float4 con_o,
                                            const float4* __restrict__ c_ptr, float (&v)[10]) {
    // backward.cu:512-514: skip Gaussians behind this pixel's last contributor
    if (!(pos < px.last_contributor)) return false;
    const float dx = __fsub_rn(a.x, px.pixf.x);
    const float dy = __fsub_rn(a.y, px.pixf.y);
    const float sy = __fmul_rn(__fmul_rn(con_o.z, dy), dy);
    const float sq = __fmaf_rn(dx, __fmul_rn(con_o.x, dx), sy);
    const float cr = __fmul_rn(__fmul_rn(con_o.y, dx), dy);
    const float power = __fmaf_rn(sq, -0.5f, -cr);
    if (power > 0.0f) return false;
    const float G = expf(power);
    const float alpha = fminf(0.99f, __fmul_rn(con_o.w, G));
    if (alpha < 1.0f / 255.0f) return false;
    const float4 c = *c_ptr;
    // backward.cu:529-563.  One approximate reciprocal (MUFU.RCP, <= 1 ulp; 1 - alpha is in [0.01, 1]) serves both
    // divisions of backward.cu:530,562.
    const float inv_1ma = rcp_approx(1.f - alpha);
    px.T = px.T * inv_1ma;
    const float dchannel_dcolor = alpha * px.T;
    float dL_dalpha = 0.0f;
    px.accum0 = px.last_alpha * px.last_c0 + (1.f - px.last_alpha) * px.accum0;
    px.last_c0 = c.x;
    dL_dalpha += (c.x - px.accum0) * px.dpix0;
    v[7] = dchannel_dcolor * px.dpix0;
    px.accum1 = px.last_alpha * px.last_c1 + (1.f - px.last_alpha) * px.accum1;
    px.last_c1 = c.y;
    dL_dalpha += (c.y - px.accum1) * px.dpix1;
    v[8] = dchannel_dcolor * px.dpix1;
    px.accum2 = px.last_alpha * px.last_c2 + (1.f - px.last_alpha) * px.accum2;
    px.last_c2 = c.z;
    dL_dalpha += (c.z - px.accum2) * px.dpix2;
    v[9] = dchannel_dcolor * px.dpix2;
    v[6] = dchannel_dcolor * px.ddepth;  // backward.cu:552
    dL_dalpha *= px.T;
    px.last_alpha = alpha;
    dL_dalpha += (-px.T_final * inv_1ma) * px.bg_dot_dpixel;
    // moments of t = G dL/dalpha (see the record layout above; backward.cu:567-583 is applied per Gaussian in K8)
    const float t = G * dL_dalpha;
    const float tdx = t * dx, tdy = t * dy;
    v[0] = tdx;
    v[1] = tdy;
    v[2] = tdx * dx;
    v[3] = tdx * dy;
    v[4] = tdy * dy;
    v[5] = t;
    return true;
}

// record float index of partial i
__device__ __forceinline__ int record_index(int i) { return i < 7 ? i : i + 1; }

// Sum v[0..11] over the warp.  On return lane L holds in the result the total of slot
//   slot(L) = 6*b4 + 3*b3 + 2*b2 + b1   (b_k = bit k of L), valid when (2*b2+b1) < 3;
// lanes L and L^1 hold the same value.
__device__ __forceinline__ float warp_reduce12(const float (&v)[12], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
    float u[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float keep = h16 ? v[6 + i] : v[i];
        const float send = h16 ? v[i] : v[6 + i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float keep = h8 ? u[3 + i] : u[i];
        const float send = h8 ? u[i] : u[3 + i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float x[2];
    {
        // {w0,w1 | w2,pad}
        const float keep0 = h4 ? w[2] : w[0];
        const float send0 = h4 ? w[0] : w[2];
        x[0] = keep0 + __shfl_xor_sync(0xffffffffu, send0, 4);
        const float keep1 = h4 ? 0.0f : w[1];
        const float send1 = h4 ? w[1] : 0.0f;
        x[1] = keep1 + __shfl_xor_sync(0xffffffffu, send1, 4);
    }
    const float keep = h2 ? x[1] : x[0];
    const float send = h2 ? x[0] : x[1];
    float y = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;
}

// ---- K7 ---------------------------------------------------------------------------------------------------
// Pixel-parallel back-to-front replay of a 16x16 tile (one warp = 8x4 pixels), 256-record batches staged with
// cp.async (double buffered), one barrier per batch.  Per (warp, Gaussian) hit the ten partials of the 32 pixels
// have to be summed.  Two reduction schemes:
//  ROWS = 0  value-halving butterfly: 13 shuffles + 21 selects + 13 adds, result lanes issue one 4-byte RED each;
//  ROWS = H  (default, H = 3) transposed through shared memory: every lane stores its ten partials into ten rows of a
//            per-warp [10 H][32 + 4] array as soon as they are computed (no register pressure, nothing to select);
//            after H hits, lanes 0 .. 10 H - 1 each read one row with eight 16-byte loads, add the 32 values in a
//            fixed order and issue the RED for (hit, slot).  About 20 issue slots per hit instead of 60, the rest
//            of the cost moves to the (idle) shared-memory pipe.
// DET (tests, wast3d_set_deterministic): no float atomics; the warps park their sums in shared memory, one thread per
// instance adds them in warp order and writes inst_grad[position in the point list] (see det_gather_kernel).
constexpr int RED_STRIDE = 36;   // floats per row: 32 pixels + 4 (conflict-free 16-byte row reads by 8 lanes at a time)

template <int BATCH, int ROWS, bool DET>
struct K7Smem {
    float4 r0[2][BATCH];
    float4 r1[2][BATCH];
    float4 r2[2][BATCH];
    uint32_t id[2][BATCH];
    uint32_t warp_max[TILE_PIX / 32];
    float red[ROWS > 0 ? TILE_PIX / 32 : 1][ROWS > 0 ? 10 * ROWS : 1][RED_STRIDE];
    float part[DET ? TILE_PIX / 32 : 1][DET ? BATCH : 1][12];
};

template <int BATCH, int ROWS, bool DET, int MINB>
__global__ void __launch_bounds__(TILE_PIX, MINB)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const int W, const int H, const float* __restrict__ bg_color,
                       const float4* __restrict__ rec, const float* __restrict__ sampling_offsets,
                       const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpixels, const float* __restrict__ dL_ddepths,
                       float4* __restrict__ grad_rec, float4* __restrict__ inst_grad) {
    extern __shared__ __align__(16) uint8_t k7_smem_raw[];
    K7Smem<BATCH, ROWS, DET>& sm = *reinterpret_cast<K7Smem<BATCH, ROWS, DET>*>(k7_smem_raw);

    float* grad_f = reinterpret_cast<float*>(grad_rec);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t pxi = blockIdx.x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t pyi = blockIdx.y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;

    ReplayPixel px;
    px.pixf = make_float2((float)pxi, (float)pyi);
    if (inside && sampling_offsets != nullptr) {
        const float2 o = *reinterpret_cast<const float2*>(sampling_offsets + 2 * (size_t)pix_id);
        px.pixf.x = (float)pxi + o.x;
        px.pixf.y = (float)pyi + o.y;
    }
    const float inf = __int_as_float(0x7f800000);
    float bx0 = inside ? px.pixf.x : inf, bx1 = inside ? px.pixf.x : -inf;
    float by0 = inside ? px.pixf.y : inf, by1 = inside ? px.pixf.y : -inf;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d));
        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
    }

    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];

    // backward.cu:463-478
    px.T_final = inside ? final_Ts[pix_id] : 0.f;
    px.T = px.T_final;
    px.last_contributor = inside ? n_contrib[pix_id] : 0u;
    const size_t HW = (size_t)H * W;
    px.dpix0 = px.dpix1 = px.dpix2 = px.ddepth = 0.f;
    if (inside) {
        px.dpix0 = dL_dpixels[pix_id];
        px.dpix1 = dL_dpixels[HW + pix_id];
        px.dpix2 = dL_dpixels[2 * HW + pix_id];
        px.ddepth = dL_ddepths ? dL_ddepths[pix_id] : 0.f;
    }
    // backward.cu:560-563 (pixel constant)
    px.bg_dot_dpixel = 0.f;
    px.bg_dot_dpixel += bg_color[0] * px.dpix0;
    px.bg_dot_dpixel += bg_color[1] * px.dpix1;
    px.bg_dot_dpixel += bg_color[2] * px.dpix2;
    px.accum0 = px.accum1 = px.accum2 = 0.f;
    px.last_alpha = px.last_c0 = px.last_c1 = px.last_c2 = 0.f;

    // Nothing behind the tile's deepest last contributor can receive gradient.
    const uint32_t warp_max_last = __reduce_max_sync(0xffffffffu, px.last_contributor);
    if (lane == 0) sm.warp_max[warp] = warp_max_last;
    __syncthreads();
    uint32_t n_eff = 0;
#pragma unroll
    for (int w = 0; w < TILE_PIX / 32; ++w) n_eff = max(n_eff, sm.warp_max[w]);
    n_eff = min(n_eff, range.y - range.x);
    const int n = (int)n_eff;
    const int rounds = (n + BATCH - 1) / BATCH;

    auto prefetch = [&](int b) {
#pragma unroll
        for (int q = tid; q < BATCH; q += TILE_PIX) {
            const int p = b * BATCH + q;
            if (p < n) {
                const uint32_t id = point_list[range.x + (uint32_t)(n - 1 - p)];
                sm.id[b & 1][q] = id;
                const float4* src = rec + 3 * (size_t)id;
                cp_async16(&sm.r0[b & 1][q], src);
                cp_async16(&sm.r1[b & 1][q], src + 1);
                cp_async16(&sm.r2[b & 1][q], src + 2);
            }
        }
        cp_async_commit();
    };

    // transposed reduction state: hits parked in this warp's rows, and where their sums go
    int parked = 0;
    uint32_t dst0 = 0, dst1 = 0, dst2 = 0, dst3 = 0;   // ROWS <= 4
    float* red_w = &sm.red[ROWS > 0 ? warp : 0][0][0];
    auto flush = [&]() {
        if (ROWS == 0 || parked == 0) return;
        __syncwarp();
        if (lane < 10 * parked) {
            const float4* row = reinterpret_cast<const float4*>(red_w + lane * RED_STRIDE);
            float4 q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = row[k];
            float sum = 0.f;
            {   // fixed order: pairwise over the eight 16-byte pieces
                float p4[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) p4[k] = (q[k].x + q[k].y) + (q[k].z + q[k].w);
                sum = ((p4[0] + p4[1]) + (p4[2] + p4[3])) + ((p4[4] + p4[5]) + (p4[6] + p4[7]));
            }
            const int h = lane / 10, i = lane - 10 * h;
            const uint32_t dst = h == 0 ? dst0 : h == 1 ? dst1 : h == 2 ? dst2 : dst3;
            if (DET) sm.part[DET ? warp : 0][DET ? dst : 0][record_index(i)] = sum;   // dst = index in the batch
            else atomicAdd(grad_f + 12 * (size_t)dst + record_index(i), sum);         // dst = Gaussian id
        }
        __syncwarp();
        parked = 0;
    };

    if (rounds > 0) prefetch(0);
    for (int b = 0; b < rounds; ++b) {
        if (DET) {   // this warp's partials of the batch start at zero (a warp that skips a Gaussian adds 0)
            float* z = &sm.part[DET ? warp : 0][0][0];
            for (int q = lane; q < BATCH * 12; q += 32) z[q] = 0.f;
        }
        // one barrier per batch: batch b has landed, and every warp is done with batch b-1 whose
        // buffer the prefetch of b+1 overwrites
        cp_async_wait<0>();
        __syncthreads();
        if (b + 1 < rounds) prefetch(b + 1);

        const int cnt = min(BATCH, n - b * BATCH);
        const float4* r0 = sm.r0[b & 1];
        const float4* r1 = sm.r1[b & 1];
        const float4* r2 = sm.r2[b & 1];
        if (warp_max_last > 0) {
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int jj = j0 + lane;
                bool hit = false;
                if (jj < cnt) {
                    const uint32_t pos = (uint32_t)(n - 1 - (b * BATCH + jj));
                    const float4 a = r0[jj];
                    const float hy = r2[jj].w;
                    hit = pos < warp_max_last &&
                          !((a.x + a.w < bx0) || (a.x - a.w > bx1) || (a.y + hy < by0) ||
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
                    const uint32_t pos = (uint32_t)(n - 1 - (b * BATCH + j));
                    float v[10];
                    const bool active = replay_pair(px, pos, r0[j], r1[j], r2 + j, v);
                    if (!__any_sync(0xffffffffu, active)) continue;
                    if (ROWS > 0) {
                        float* rowp = red_w + parked * (10 * RED_STRIDE) + lane;
#pragma unroll
                        for (int i = 0; i < 10; ++i) rowp[i * RED_STRIDE] = active ? v[i] : 0.f;
                        const uint32_t dst = DET ? (uint32_t)j : sm.id[b & 1][j];
                        if (parked == 0) dst0 = dst; else if (parked == 1) dst1 = dst; else if (parked == 2) dst2 = dst; else dst3 = dst;
                        if (++parked == ROWS) flush();
                    } else {
                        float v12[12];
#pragma unroll
                        for (int i = 0; i < 10; ++i) v12[record_index(i)] = active ? v[i] : 0.f;
                        v12[7] = v12[11] = 0.f;
                        const float tot = warp_reduce12(v12, lane);
                        const int sub = ((lane >> 1) & 3);  // 2*b2 + b1
                        const int slot = 6 * ((lane >> 4) & 1) + 3 * ((lane >> 3) & 1) + sub;
                        if (!(lane & 1) && sub < 3 && slot != 7 && slot != 11) {
                            if (DET) sm.part[DET ? warp : 0][DET ? j : 0][slot] = tot;
                            else atomicAdd(grad_f + 12 * (size_t)sm.id[b & 1][j] + slot, tot);
                        }
                    }
                }
            }
        }
        if (DET) {
            flush();   // parked sums refer to positions in THIS batch
            __syncthreads();
            if (tid < cnt) {
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
#pragma unroll
                for (int w = 0; w < TILE_PIX / 32; ++w) {   // fixed order
                    const float4* q = reinterpret_cast<const float4*>(&sm.part[DET ? w : 0][DET ? tid : 0][0]);
                    const float4 x = q[0], y = q[1], z = q[2];
                    a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w;
                    a1.x += y.x; a1.y += y.y; a1.z += y.z; a1.w += y.w;
                    a2.x += z.x; a2.y += z.y; a2.z += z.z; a2.w += z.w;
                }
                const size_t pos = (size_t)range.x + (size_t)(n - 1 - (b * BATCH + tid));
                inst_grad[3 * pos + 0] = a0;
                inst_grad[3 * pos + 1] = a1;
                inst_grad[3 * pos + 2] = a2;
            }
            __syncthreads();   // before the next batch zeroes the partials
        }
    }
    flush();   // Gaussian ids were copied when the hits were parked: sums may outlive their batch
    cp_async_wait<0>();
}

// ---- K7, one warp per CTA (see render_forward_warp_kernel for the rationale) ------------------------------------
// Every 8x4 pixel block replays the tile's list for itself, back to front from ITS deepest last contributor, three
// 32-record rounds in flight, no block barrier; the ten moment sums of a hit go through the value-halving butterfly
// and one RED per slot, exactly like the tile kernel.  Per-pixel arithmetic is identical (replay_pair).
constexpr int K7W_STAGES = 3;

__global__ void __launch_bounds__(32)
render_backward_warp_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                            const int W, const int H, const int tiles_x, const float* __restrict__ bg_color,
                            const float4* __restrict__ rec, const float* __restrict__ sampling_offsets,
                            const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                            const float* __restrict__ dL_dpixels, const float* __restrict__ dL_ddepths,
                            float4* __restrict__ grad_rec) {
    __shared__ float4 s_r0[K7W_STAGES][32];
    __shared__ float4 s_r1[K7W_STAGES][32];
    __shared__ float4 s_r2[K7W_STAGES][32];
    __shared__ uint32_t s_id[K7W_STAGES][32];

    float* grad_f = reinterpret_cast<float*>(grad_rec);
    const int lane = threadIdx.x;
    const uint32_t pxi = blockIdx.x * 8 + (lane & 7);
    const uint32_t pyi = blockIdx.y * 4 + (lane >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;

    ReplayPixel px;
    px.last_contributor = inside ? n_contrib[pix_id] : 0u;
    // Nothing behind this block's deepest last contributor can receive gradient.
    const uint32_t warp_max_last = __reduce_max_sync(0xffffffffu, px.last_contributor);
    if (warp_max_last == 0) return;

    px.pixf = make_float2((float)pxi, (float)pyi);
    if (inside && sampling_offsets != nullptr) {
        const float2 o = *reinterpret_cast<const float2*>(sampling_offsets + 2 * (size_t)pix_id);
        px.pixf.x = (float)pxi + o.x;
        px.pixf.y = (float)pyi + o.y;
    }
    const float inf = __int_as_float(0x7f800000);
    float bx0 = inside ? px.pixf.x : inf, bx1 = inside ? px.pixf.x : -inf;
    float by0 = inside ? px.pixf.y : inf, by1 = inside ? px.pixf.y : -inf;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, d));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, d));
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, d));
        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, d));
    }

    const uint2 range = ranges[(blockIdx.y >> 2) * tiles_x + (blockIdx.x >> 1)];

    // backward.cu:463-478
    px.T_final = inside ? final_Ts[pix_id] : 0.f;
    px.T = px.T_final;
    const size_t HW = (size_t)H * W;
    px.dpix0 = px.dpix1 = px.dpix2 = px.ddepth = 0.f;
    if (inside) {
        px.dpix0 = dL_dpixels[pix_id];
        px.dpix1 = dL_dpixels[HW + pix_id];
        px.dpix2 = dL_dpixels[2 * HW + pix_id];
        px.ddepth = dL_ddepths ? dL_ddepths[pix_id] : 0.f;
    }
    // backward.cu:560-563 (pixel constant)
    px.bg_dot_dpixel = 0.f;
    px.bg_dot_dpixel += bg_color[0] * px.dpix0;
    px.bg_dot_dpixel += bg_color[1] * px.dpix1;
    px.bg_dot_dpixel += bg_color[2] * px.dpix2;
    px.accum0 = px.accum1 = px.accum2 = 0.f;
    px.last_alpha = px.last_c0 = px.last_c1 = px.last_c2 = 0.f;

    const int n = (int)min(warp_max_last, range.y - range.x);
    const int rounds = (n + 31) / 32;

    auto prefetch = [&](int b) {
        const int p = b * 32 + lane;
        if (b < rounds && p < n) {
            const uint32_t id = point_list[range.x + (uint32_t)(n - 1 - p)];
            const int st = b % K7W_STAGES;
            s_id[st][lane] = id;
            const float4* src = rec + 3 * (size_t)id;
            cp_async16(&s_r0[st][lane], src);
            cp_async16(&s_r1[st][lane], src + 1);
            cp_async16(&s_r2[st][lane], src + 2);
        }
        cp_async_commit();   // one group per round, empty or not
    };

    prefetch(0);
    prefetch(1);
    for (int b = 0; b < rounds; ++b) {
        prefetch(b + 2);
        cp_async_wait<2>();
        __syncwarp();
        const int st = b % K7W_STAGES;
        const float4* r0 = s_r0[st];
        const float4* r1 = s_r1[st];
        const float4* r2 = s_r2[st];
        const int cnt = min(32, n - b * 32);
        bool hit = false;
        if (lane < cnt) {
            const float4 a = r0[lane];
            const float hy = r2[lane].w;
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
            const uint32_t pos = (uint32_t)(n - 1 - (b * 32 + j));
            float v[10];
            const bool active = replay_pair(px, pos, r0[j], r1[j], r2 + j, v);
            if (!__any_sync(0xffffffffu, active)) continue;
            float v12[12];
#pragma unroll
            for (int i = 0; i < 10; ++i) v12[record_index(i)] = active ? v[i] : 0.f;
            v12[7] = v12[11] = 0.f;
            const float tot = warp_reduce12(v12, lane);
            const int sub = ((lane >> 1) & 3);  // 2*b2 + b1
            const int slot = 6 * ((lane >> 4) & 1) + 3 * ((lane >> 3) & 1) + sub;
            if (!(lane & 1) && sub < 3 && slot != 7 && slot != 11)
                atomicAdd(grad_f + 12 * (size_t)s_id[st][j] + slot, tot);
        }
        __syncwarp();   // everybody is done with stage st before round b + 3 overwrites it
    }
    cp_async_wait<0>();
}

// Deterministic mode, second half: Gaussian `id` owns tiles_touched[id] instances whose point-list positions
// sit, ascending, at inst_sorted[seg[id] ...]; their 48-byte partial gradient records are added in that order.
__global__ void __launch_bounds__(256)
det_gather_kernel(int P, const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ seg,
                  const uint32_t* __restrict__ inst_sorted, const float4* __restrict__ inst_grad,
                  float4* __restrict__ grad_rec, const uint32_t R) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= P) return;
    const uint32_t n = tiles_touched[id];
    if (n == 0) return;   // grad_rec was zero-filled
    const uint32_t first = seg[id];
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
    for (uint32_t k = 0; k < n; ++k) {
        if (first + k >= R) break;   // only after an overflowing graph-safe forward (flagged; the step is discarded)
        const size_t i = inst_sorted[first + k];
        const float4 x = inst_grad[3 * i], y = inst_grad[3 * i + 1], z = inst_grad[3 * i + 2];
        a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w;
        a1.x += y.x; a1.y += y.y; a1.z += y.z; a1.w += y.w;
        a2.x += z.x; a2.y += z.y; a2.z += z.z; a2.w += z.w;
    }
    grad_rec[3 * (size_t)id + 0] = a0;
    grad_rec[3 * (size_t)id + 1] = a1;
    grad_rec[3 * (size_t)id + 2] = a2;
}


// ------------------------------------------------------------------ K8 + K9 ---------------
constexpr int GB_THREADS = 128;
constexpr int GB_WARPS = GB_THREADS / 32;
constexpr int GB_SH_STRIDE = 49;

// Fused optimizer epilogue (ADAM = true, RAW only): the gradient of every leaf
// element is consumed where it is produced — Adam (adam.cu's arithmetic = torch's single-tensor
// update, scene/gaussian_model.py:154-163) is applied to the parameter in place, so the 236 B/Gaussian
// of gradients are neither written to nor read back from HBM and the six optimizer launches disappear.
// Valid because each parameter element is read (as a parameter) only by the warp that also updates it,
// before the update.  Culled Gaussians take the zero-gradient update (their moments still decay), like
// the dense torch.optim.Adam of the reference.
struct AdamSlot {  // one GaussianModel parameter group
    float* p;
    float* m;
    float* v;
    float step_size, inv_bc2_sqrt, one_minus_b1, b2, one_minus_b2, eps;
    const float* sched;   // optional device {step_size, inv_bc2_sqrt} read at run time (wast3d_adam_group::schedule_dev)
};
struct AdamFused {
    AdamSlot g[6];  // xyz, f_dc, f_rest, opacity, scaling, rotation
    int l2_prefetch;  // 1: each warp asks L2 for its parameter / moment tiles before the chain rule
};

struct AdamStep { float step_size, inv_bc2_sqrt; };   // the two step-dependent scalars of a group
__device__ __forceinline__ AdamStep adam_step_scalars(const AdamSlot& s) {
    AdamStep a = {s.step_size, s.inv_bc2_sqrt};
    if (s.sched != nullptr) { a.step_size = __ldg(s.sched); a.inv_bc2_sqrt = __ldg(s.sched + 1); }   // uniform
    return a;
}
__device__ __forceinline__ float adam_elem(float p, float g, float& m, float& v, const AdamSlot& s, const AdamStep& a) {
    return adam_update(p, g, m, v, s.one_minus_b1, s.b2, s.one_minus_b2, a.step_size, a.inv_bc2_sqrt, s.eps);
}
// updated parameter values are also returned in p_new (the next view's projection consumes them from registers)
template <int N>
__device__ __forceinline__ void adam_small(const AdamSlot& s, size_t base, const float (&g)[N], float (&p_new)[N]) {
    float m[N], v[N];
    const AdamStep a = adam_step_scalars(s);
#pragma unroll
    for (int k = 0; k < N; ++k) { p_new[k] = s.p[base + k]; m[k] = s.m[base + k]; v[k] = s.v[base + k]; }
#pragma unroll
    for (int k = 0; k < N; ++k) p_new[k] = adam_elem(p_new[k], g[k], m[k], v[k], s, a);
#pragma unroll
    for (int k = 0; k < N; ++k) { s.p[base + k] = p_new[k]; s.m[base + k] = m[k]; s.v[base + k] = v[k]; }
}
// Adam over a warp's contiguous block of `total` elements whose gradients sit in shared memory.
// U float4 per lane are loaded from each of p / m / v before any of them is consumed (3U independent 16-byte
// loads in flight per lane): this loop moves 76% of the fused kernel's bytes and is latency bound otherwise.
// KEEP: the updated parameters replace the gradients in shared memory (same linear layout), for a consumer in the
// same warp after a __syncwarp (the next view's SH -> RGB).
template <int U, bool KEEP>
__device__ __forceinline__ void adam_rows_linear(const AdamSlot& s, size_t base, int total,
                                                 float* s_grad, float* g_out, int lane) {
    float* P_ = s.p + base;
    float* M_ = s.m + base;
    float* V_ = s.v + base;
    const AdamStep a = adam_step_scalars(s);
    const bool vec = (total & 3) == 0 && ((((size_t)P_ | (size_t)M_ | (size_t)V_ | (size_t)s_grad) & 15) == 0) &&
                     (g_out == nullptr || (((size_t)(g_out + base)) & 15) == 0);
    if (vec) {
        const int n4 = total >> 2;
        float4* P4 = reinterpret_cast<float4*>(P_);
        float4* M4 = reinterpret_cast<float4*>(M_);
        float4* V4 = reinterpret_cast<float4*>(V_);
        for (int q0 = 0; q0 < n4; q0 += 32 * U) {
            float4 p[U], m[U], v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + 32 * u + lane;
                if (q < n4) { p[u] = __ldcs(P4 + q); m[u] = __ldcs(M4 + q); v[u] = __ldcs(V4 + q); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + 32 * u + lane;
                if (q < n4) {
                    const float4 g = *reinterpret_cast<const float4*>(s_grad + 4 * q);
                    p[u].x = adam_elem(p[u].x, g.x, m[u].x, v[u].x, s, a);
                    p[u].y = adam_elem(p[u].y, g.y, m[u].y, v[u].y, s, a);
                    p[u].z = adam_elem(p[u].z, g.z, m[u].z, v[u].z, s, a);
                    p[u].w = adam_elem(p[u].w, g.w, m[u].w, v[u].w, s, a);
                    P4[q] = p[u];
                    __stcs(M4 + q, m[u]);
                    __stcs(V4 + q, v[u]);
                    if (g_out) reinterpret_cast<float4*>(g_out + base)[q] = g;
                    if (KEEP) *reinterpret_cast<float4*>(s_grad + 4 * q) = p[u];
                }
            }
        }
    } else {
        for (int q = lane; q < total; q += 32) {
            const float g = s_grad[q];
            float m = M_[q], v = V_[q];
            const float pn = adam_elem(P_[q], g, m, v, s, a);
            P_[q] = pn;
            M_[q] = m;
            V_[q] = v;
            if (g_out) g_out[base + q] = g;
            if (KEEP) s_grad[q] = pn;
        }
    }
}

// RAW: model-space inputs (wast3d_raster_params::raw_params): scales/rotations/SH arrive
// un-activated, and the gradients written are those of the six GaussianModel leaves, i.e. this
// kernel also does the autograd backward of exp / normalize / sigmoid / cat.  In RAW mode
// dL_dsh is dL/d_features_dc [P,1,3] and dL_dsh_rest is dL/d_features_rest [P,M-1,3].
// ADAM (RAW only): apply the optimizer update in place (AdamFused); the gradient outputs of the six
// leaves become optional (non-NULL ones are still written: tests).
// NEXT (RAW + ADAM only; wast3d_next_view): after the update, every Gaussian is projected for the NEXT view from the
// parameter values this thread has just computed — K1 of the next forward pass without its 236 B/Gaussian parameter
// read and without its launch (project.cuh: the very statements preprocess_kernel executes).
struct NextProj {
    ProjView view;
    const float* campos;
    int D;
    int* radii;
    float4* rec;
    uint32_t* depth_key;
    uint32_t* tiles_touched;
    uint8_t* clamped;
    uint2* rect;
    uint32_t* bound_words;   // [4] order keys of the offset bounds the tile cut assumed (checked by the next forward)
};

template <bool RAW, bool ADAM, int ADAM_U = 4, bool NEXT = false>
__global__ void __launch_bounds__(GB_THREADS)
gaussian_backward_kernel(const int P, const int D, const int M, const float* __restrict__ means3D,
                         const int* __restrict__ radii, const float* __restrict__ shs,
                         const float* __restrict__ shs_rest, const float4* __restrict__ rec,
                         const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
                         const float* __restrict__ rotations, const float scale_modifier,
                         const float* __restrict__ cov3D_precomp, const float* __restrict__ view,
                         const float* __restrict__ proj, const float* __restrict__ campos,
                         const float h_x, const float h_y, const float tan_fovx, const float tan_fovy,
                         const float half_w, const float half_h,
                         const float4* __restrict__ grad_rec, float* __restrict__ dL_dmean2D,
                         float* __restrict__ dL_dconic_out, float* __restrict__ dL_dopacity,
                         float* __restrict__ dL_dcolor, float* __restrict__ dL_dmean3D,
                         float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh,
                         float* __restrict__ dL_dsh_rest,
                         float* __restrict__ dL_dscale, float* __restrict__ dL_drot,
                         float* __restrict__ dL_dviewdepth_out, const __grid_constant__ AdamFused af,
                         const __grid_constant__ NextProj nx) {
    __shared__ __align__(16) float s_sh[GB_WARPS][32 * GB_SH_STRIDE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * GB_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * GB_THREADS + warp * 32;
    const bool live = idx < P;
    if (RAW && ADAM && af.l2_prefetch && warp_first < P) {
        // The optimizer's operands of this warp's 32 Gaussians (parameters and both moments of the six groups,
        // 22 KB, 76% of them the SH rest rows) are needed only after the chain rule below: one lane per
        // (group, array) starts their HBM -> L2 transfer now.  Purely a hint: skipped where a tile is not
        // 16-byte granular (partial last warp with an odd row count).
        const int rows = min(32, P - warp_first);
        if (lane < 18) {
            const int grp = lane / 3, arr = lane - 3 * grp;
            const AdamSlot& sl = af.g[grp];
            const float* base = (arr == 0 ? sl.p : arr == 1 ? sl.m : sl.v);
            const int fl = grp == 2 ? 3 * (M - 1) : grp == 3 ? 1 : grp == 5 ? 4 : 3;
            l2_prefetch_rows(base, (size_t)warp_first, rows, fl);
        }
    } else if (!ADAM && af.l2_prefetch && warp_first < P && lane == 0 && shs != nullptr) {
        // without the optimizer epilogue: the SH rows are staged only after the chain rule
        const int rows = min(32, P - warp_first);
        if (RAW) l2_prefetch_rows(shs_rest, (size_t)warp_first, rows, 3 * (M - 1));
        else l2_prefetch_rows(shs, (size_t)warp_first, rows, 3 * M);
    }
    const bool vis = live && radii[idx] > 0;

    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
    float3 dmean = make_float3(0.f, 0.f, 0.f);
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float3 dscale = make_float3(0.f, 0.f, 0.f);
    float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
    float3 mean = make_float3(0.f, 0.f, 0.f);

    if (vis) {
        g0 = grad_rec[3 * (size_t)idx + 0];
        g1 = grad_rec[3 * (size_t)idx + 1];
        g2 = grad_rec[3 * (size_t)idx + 2];
        {
            // K7 accumulated moments (see the record layout at the top): backward.cu:567-583 once per Gaussian.
            // half_w, half_h = ddelx_dx, ddely_dy (backward.cu:486-487)
            const float4 co = rec[3 * (size_t)idx + 1];   // conic (A, B, C), opacity
            const float Mx = g0.x, My = g0.y, Mxx = g0.z, Mxy = g0.w, Myy = g1.x;
            g0.x = -(co.w * half_w) * (co.x * Mx + co.y * My);
            g0.y = -(co.w * half_h) * (co.z * My + co.y * Mx);
            g0.z = -0.5f * co.w * Mxx;
            g0.w = -0.5f * co.w * Mxy;
            g1.x = -0.5f * co.w * Myy;
        }
        mean = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);

        float cov6[6];
        float3 sc = make_float3(0.f, 0.f, 0.f);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        float q_denom = 1.f;
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; ++k) cov6[k] = cov3D_precomp[6 * idx + k];
        } else {
            sc = make_float3(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]);
            q = *reinterpret_cast<const float4*>(rotations + 4 * idx);
            if (RAW) {
                sc = act_exp3(sc);
                q_denom = quat_denom(q);
                q = act_normalize4(q, q_denom);
            }
            cov3d_from_scale_rot(sc, scale_modifier, q, cov6);
        }

        // ---- backward.cu:144-274 (computeCov2DCUDA)
        Cov2DCtx cx;
        const float3 cov2 = cov2d(mean, h_x, h_y, tan_fovx, tan_fovy, cov6, view, &cx);
        const float3 dL_dconic = make_float3(g0.z, g0.w, g1.x);
        const float limx = 1.3f * tan_fovx;
        const float limy = 1.3f * tan_fovy;
        const float x_grad_mul = cx.txtz < -limx || cx.txtz > limx ? 0 : 1;
        const float y_grad_mul = cx.tytz < -limy || cx.tytz > limy ? 0 : 1;
        const float3 t = cx.t;
        const M3& T = cx.T;
        const M3& Vrk = cx.Vrk;
        const float a = cov2.x, b = cov2.y, c = cov2.z;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
            dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

            dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
            dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
            dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
            dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
            dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
            dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
        }
        const float dL_dT00 = 2 * (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_da +
                              (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_db;
        const float dL_dT01 = 2 * (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_da +
                              (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_db;
        const float dL_dT02 = 2 * (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_da +
                              (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_db;
        const float dL_dT10 = 2 * (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_dc +
                              (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_db;
        const float dL_dT11 = 2 * (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_dc +
                              (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_db;
        const float dL_dT12 = 2 * (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_dc +
                              (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_db;

        // W as in cov2d(): W.c[i][j] = view[4*j + i]
        const float W00 = view[0], W01 = view[4], W02 = view[8];
        const float W10 = view[1], W11 = view[5], W12 = view[9];
        const float W20 = view[2], W21 = view[6], W22 = view[10];
        const float dL_dJ00 = W00 * dL_dT00 + W01 * dL_dT01 + W02 * dL_dT02;
        const float dL_dJ02 = W20 * dL_dT00 + W21 * dL_dT01 + W22 * dL_dT02;
        const float dL_dJ11 = W10 * dL_dT10 + W11 * dL_dT11 + W12 * dL_dT12;
        const float dL_dJ12 = W20 * dL_dT10 + W21 * dL_dT11 + W22 * dL_dT12;

        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
        dmean = xform_vec_4x3_T(make_float3(dL_dtx, dL_dty, dL_dtz), view);

        // ---- backward.cu:372-401 (projection + view-depth terms)
        const float4 m_hom = xform_point_4x4(mean, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        float3 dm2;
        dm2.x = (proj[0] * m_w - proj[3] * mul1) * g0.x + (proj[1] * m_w - proj[3] * mul2) * g0.y;
        dm2.y = (proj[4] * m_w - proj[7] * mul1) * g0.x + (proj[5] * m_w - proj[7] * mul2) * g0.y;
        dm2.z = (proj[8] * m_w - proj[11] * mul1) * g0.x + (proj[9] * m_w - proj[11] * mul2) * g0.y;
        dm2.x += view[2] * g1.z;
        dm2.y += view[6] * g1.z;
        dm2.z += view[10] * g1.z;
        dmean.x += dm2.x;
        dmean.y += dm2.y;
        dmean.z += dm2.z;

        // ---- backward.cu:278-341 (Sigma3D -> scale, quaternion)
        if (scales != nullptr) {
            const M3 R = quat_to_R(q);
            const float3 s = make_float3(scale_modifier * sc.x, scale_modifier * sc.y, scale_modifier * sc.z);
            M3 S;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) S.c[i][j] = 0.0f;
            S.c[0][0] = s.x; S.c[1][1] = s.y; S.c[2][2] = s.z;
            const M3 Mm = m3_mul(S, R);
            M3 dSig;
            dSig.c[0][0] = dcov[0];        dSig.c[0][1] = 0.5f * dcov[1]; dSig.c[0][2] = 0.5f * dcov[2];
            dSig.c[1][0] = 0.5f * dcov[1]; dSig.c[1][1] = dcov[3];        dSig.c[1][2] = 0.5f * dcov[4];
            dSig.c[2][0] = 0.5f * dcov[2]; dSig.c[2][1] = 0.5f * dcov[4]; dSig.c[2][2] = dcov[5];
            M3 M2;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) M2.c[i][j] = 2.0f * Mm.c[i][j];
            const M3 dL_dM = m3_mul(M2, dSig);
            const M3 Rt = m3_transpose(R);
            M3 dMt = m3_transpose(dL_dM);
            dscale.x = Rt.c[0][0] * dMt.c[0][0] + Rt.c[0][1] * dMt.c[0][1] + Rt.c[0][2] * dMt.c[0][2];
            dscale.y = Rt.c[1][0] * dMt.c[1][0] + Rt.c[1][1] * dMt.c[1][1] + Rt.c[1][2] * dMt.c[1][2];
            dscale.z = Rt.c[2][0] * dMt.c[2][0] + Rt.c[2][1] * dMt.c[2][1] + Rt.c[2][2] * dMt.c[2][2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dMt.c[0][k] *= s.x;
                dMt.c[1][k] *= s.y;
                dMt.c[2][k] *= s.z;
            }
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            drot.x = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) + 2 * x * (dMt.c[1][2] - dMt.c[2][1]);
            drot.y = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) + 2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
            drot.z = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) + 2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
            drot.w = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) + 2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
            if (RAW) {
                // exp backward: d/dlog_s = d/ds * s   (sc holds the activated scale)
                dscale.x *= sc.x; dscale.y *= sc.y; dscale.z *= sc.z;
                // normalize backward (x / max(|x|, eps)): (g - q_hat (q_hat . g)) / |x|; below eps the
                // denominator is the constant eps and the map is linear
                if (q_denom > 1e-12f) {
                    const float dot = q.x * drot.x + q.y * drot.y + q.z * drot.z + q.w * drot.w;
                    drot = make_float4((drot.x - q.x * dot) / q_denom, (drot.y - q.y * dot) / q_denom,
                                       (drot.z - q.z * dot) / q_denom, (drot.w - q.w * dot) / q_denom);
                } else {
                    drot = make_float4(drot.x / q_denom, drot.y / q_denom, drot.z / q_denom, drot.w / q_denom);
                }
            }
        }
    }

    float new_dc[3] = {0.f, 0.f, 0.f};   // updated _features_dc of this Gaussian (ADAM)
    // ---- backward.cu:20-139 (SH): staged in, gradients staged out through the same rows
    if (shs != nullptr) {
        const unsigned need = __ballot_sync(0xffffffffu, vis);
        const int rows_valid = min(32, P - warp_first);
        if (rows_valid > 0) {
            const int row_floats = 3 * M;
            const int rest_floats = 3 * (M - 1);
            const int used = 3 * (D + 1) * (D + 1);
            // RAW: `row` addresses the linear _features_rest block so that row[3k + c] (k >= 1) is
            // coefficient k of this Gaussian; the degree-0 term lives in registers (dc_grad).
            float* row = RAW ? s_sh[warp] + lane * rest_floats - 3 : s_sh[warp] + lane * GB_SH_STRIDE;
            float dc_grad[3] = {0.f, 0.f, 0.f};
            if (need) {
                if (RAW) {
                    if (used > 3)
                        stage_rows_linear(shs_rest + (size_t)warp_first * rest_floats, rest_floats, used - 3,
                                          rows_valid, need, s_sh[warp], lane);
                } else {
                    stage_sh_rows(shs + (size_t)warp_first * row_floats, row_floats, used, rows_valid, need,
                                  s_sh[warp], GB_SH_STRIDE, lane);
                }
            }
            __syncwarp();
            if (vis) {
                const float3 cam = make_float3(campos[0], campos[1], campos[2]);
                const float3 dir_orig = make_float3(mean.x - cam.x, mean.y - cam.y, mean.z - cam.z);
                const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
                const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
                const unsigned cb = clamped[idx];
                float dRGB[3] = {g2.x, g2.y, g2.z};
                dRGB[0] *= (cb & 1u) ? 0 : 1;
                dRGB[1] *= (cb & 2u) ? 0 : 1;
                dRGB[2] *= (cb & 4u) ? 0 : 1;

                // pass 1: d(rgb)/d(dir), needs the coefficients
                float ddir[3] = {0.f, 0.f, 0.f};
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float dx_ = 0.f, dy_ = 0.f, dz_ = 0.f;
                    if (D > 0) {
                        dx_ = -SH_C1 * row[9 + c];
                        dy_ = -SH_C1 * row[3 + c];
                        dz_ = SH_C1 * row[6 + c];
                        if (D > 1) {
                            dx_ += SH_C2[0] * y * row[12 + c] + SH_C2[2] * 2.f * -x * row[18 + c] + SH_C2[3] * z * row[21 + c] + SH_C2[4] * 2.f * x * row[24 + c];
                            dy_ += SH_C2[0] * x * row[12 + c] + SH_C2[1] * z * row[15 + c] + SH_C2[2] * 2.f * -y * row[18 + c] + SH_C2[4] * 2.f * -y * row[24 + c];
                            dz_ += SH_C2[1] * y * row[15 + c] + SH_C2[2] * 2.f * 2.f * z * row[18 + c] + SH_C2[3] * x * row[21 + c];
                            if (D > 2) {
                                dx_ += (SH_C3[0] * row[27 + c] * 3.f * 2.f * xy + SH_C3[1] * row[30 + c] * yz +
                                        SH_C3[2] * row[33 + c] * -2.f * xy + SH_C3[3] * row[36 + c] * -3.f * 2.f * xz +
                                        SH_C3[4] * row[39 + c] * (-3.f * xx + 4.f * zz - yy) +
                                        SH_C3[5] * row[42 + c] * 2.f * xz + SH_C3[6] * row[45 + c] * 3.f * (xx - yy));
                                dy_ += (SH_C3[0] * row[27 + c] * 3.f * (xx - yy) + SH_C3[1] * row[30 + c] * xz +
                                        SH_C3[2] * row[33 + c] * (-3.f * yy + 4.f * zz - xx) +
                                        SH_C3[3] * row[36 + c] * -3.f * 2.f * yz + SH_C3[4] * row[39 + c] * -2.f * xy +
                                        SH_C3[5] * row[42 + c] * -2.f * yz + SH_C3[6] * row[45 + c] * -3.f * 2.f * xy);
                                dz_ += (SH_C3[1] * row[30 + c] * xy + SH_C3[2] * row[33 + c] * 4.f * 2.f * yz +
                                        SH_C3[3] * row[36 + c] * 3.f * (2.f * zz - xx - yy) +
                                        SH_C3[4] * row[39 + c] * 4.f * 2.f * xz + SH_C3[5] * row[42 + c] * (xx - yy));
                            }
                        }
                    }
                    ddir[0] += dx_ * dRGB[c];
                    ddir[1] += dy_ * dRGB[c];
                    ddir[2] += dz_ * dRGB[c];
                }
                // dnormvdv (auxiliary.h:107-117)
                {
                    const float3 v = dir_orig;
                    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
                    const float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
                    dmean.x += ((+sum2 - v.x * v.x) * ddir[0] - v.y * v.x * ddir[1] - v.z * v.x * ddir[2]) * invsum32;
                    dmean.y += (-v.x * v.y * ddir[0] + (sum2 - v.y * v.y) * ddir[1] - v.z * v.y * ddir[2]) * invsum32;
                    dmean.z += (-v.x * v.z * ddir[0] - v.y * v.z * ddir[1] + (sum2 - v.z * v.z) * ddir[2]) * invsum32;
                }
                // pass 2: d(rgb)/d(coefficients) overwrites this Gaussian's row
                float basis[16];
                basis[0] = SH_C0;
                basis[1] = -SH_C1 * y; basis[2] = SH_C1 * z; basis[3] = -SH_C1 * x;
                basis[4] = SH_C2[0] * xy; basis[5] = SH_C2[1] * yz; basis[6] = SH_C2[2] * (2.f * zz - xx - yy);
                basis[7] = SH_C2[3] * xz; basis[8] = SH_C2[4] * (xx - yy);
                basis[9] = SH_C3[0] * y * (3.f * xx - yy); basis[10] = SH_C3[1] * xy * z;
                basis[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
                basis[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                basis[13] = SH_C3[4] * x * (4.f * zz - xx - yy); basis[14] = SH_C3[5] * z * (xx - yy);
                basis[15] = SH_C3[6] * x * (xx - 3.f * yy);
                const int ncoef = (D + 1) * (D + 1);
                if (RAW) {
                    dc_grad[0] = basis[0] * dRGB[0];
                    dc_grad[1] = basis[0] * dRGB[1];
                    dc_grad[2] = basis[0] * dRGB[2];
                }
#pragma unroll
                for (int k = RAW ? 1 : 0; k < 16; ++k) {
                    if (k < M) {
                        const float bk = k < ncoef ? basis[k] : 0.f;
                        row[3 * k + 0] = bk * dRGB[0];
                        row[3 * k + 1] = bk * dRGB[1];
                        row[3 * k + 2] = bk * dRGB[2];
                    }
                }
            } else if (lane < rows_valid) {
                for (int e = RAW ? 3 : 0; e < row_floats; ++e) row[e] = 0.f;
            }
            __syncwarp();
            if (RAW && ADAM) {
                if (live) {
                    adam_small<3>(af.g[1], 3 * (size_t)idx, dc_grad, new_dc);
                    if (dL_dsh) { dL_dsh[3 * idx] = dc_grad[0]; dL_dsh[3 * idx + 1] = dc_grad[1]; dL_dsh[3 * idx + 2] = dc_grad[2]; }
                }
                if (rest_floats > 0)
                    adam_rows_linear<ADAM_U, NEXT>(af.g[2], (size_t)warp_first * rest_floats, rows_valid * rest_floats,
                                                   s_sh[warp], dL_dsh_rest, lane);
                if (NEXT) __syncwarp();   // the updated rows are read per Gaussian below
            } else if (RAW) {
                // NULL feature outputs: the caller rebuilds the SH gradient elsewhere (colour-record exchange)
                if (live && dL_dsh) { dL_dsh[3 * idx] = dc_grad[0]; dL_dsh[3 * idx + 1] = dc_grad[1]; dL_dsh[3 * idx + 2] = dc_grad[2]; }
                if (rest_floats > 0 && dL_dsh_rest)
                    unstage_rows_linear(dL_dsh_rest + (size_t)warp_first * rest_floats, rest_floats, rows_valid,
                                        s_sh[warp], lane);
            } else {
                unstage_rows(dL_dsh + (size_t)warp_first * row_floats, row_floats, rows_valid, s_sh[warp],
                             GB_SH_STRIDE, lane);
            }
        }
    }

    if (!live) return;
    if (dL_dmean2D) {
        dL_dmean2D[3 * idx + 0] = g0.x;
        dL_dmean2D[3 * idx + 1] = g0.y;
        dL_dmean2D[3 * idx + 2] = 0.f;
    }
    if (dL_dcolor) {
        dL_dcolor[3 * idx + 0] = g2.x;
        dL_dcolor[3 * idx + 1] = g2.y;
        dL_dcolor[3 * idx + 2] = g2.z;
    }
    float dopac = g1.y;
    if (RAW) {
        // sigmoid backward: d/dlogit = d/do * o (1 - o); o is the activated opacity of the record
        float o = 0.f;
        if (vis) o = rec[3 * (size_t)idx + 1].w;
        dopac = g1.y * ((1.f - o) * o);
    }
    if (RAW && ADAM) {
        const float gx[3] = {dmean.x, dmean.y, dmean.z};
        const float go[1] = {dopac};
        const float gs[3] = {dscale.x, dscale.y, dscale.z};
        const float gr[4] = {drot.x, drot.y, drot.z, drot.w};
        float nxyz[3], nop[1], nsc[3], nrot[4];
        adam_small<3>(af.g[0], 3 * (size_t)idx, gx, nxyz);
        adam_small<1>(af.g[3], (size_t)idx, go, nop);
        adam_small<3>(af.g[4], 3 * (size_t)idx, gs, nsc);
        adam_small<4>(af.g[5], 4 * (size_t)idx, gr, nrot);
        if (NEXT) {
            // K1 of the next view on the fresh values (preprocess_kernel<RAW = true, COLOUR = true>, same statements)
            if (idx == 0) {
                nx.bound_words[0] = float_order_key(nx.view.sb.max_x);
                nx.bound_words[1] = float_order_key(-nx.view.sb.min_x);
                nx.bound_words[2] = float_order_key(nx.view.sb.max_y);
                nx.bound_words[3] = float_order_key(-nx.view.sb.min_y);
            }
            const float3 p_new = make_float3(nxyz[0], nxyz[1], nxyz[2]);
            const Projection pr = project_gaussian<true>(p_new, make_float3(nsc[0], nsc[1], nsc[2]), nullptr,
                                                         make_float4(nrot[0], nrot[1], nrot[2], nrot[3]), nullptr, nop[0],
                                                         nullptr, nx.view);
            float3 rgb = make_float3(0.f, 0.f, 0.f);
            unsigned clamp_bits = 0;
            if (pr.visible) {
                const float3 cam = make_float3(nx.campos[0], nx.campos[1], nx.campos[2]);
                rgb = sh_to_rgb(nx.D, new_dc, s_sh[warp] + lane * (3 * (M - 1)) - 3, p_new, cam, &clamp_bits);
            }
            store_projection(idx, pr, rgb, clamp_bits, nx.radii, nx.rec, nx.depth_key, nx.tiles_touched, nx.clamped, nx.rect);
        }
    }
    if (!ADAM || dL_dopacity) dL_dopacity[idx] = dopac;
    if (!ADAM || dL_dmean3D) {
        dL_dmean3D[3 * idx + 0] = dmean.x;
        dL_dmean3D[3 * idx + 1] = dmean.y;
        dL_dmean3D[3 * idx + 2] = dmean.z;
    }
    if (dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dL_dcov3D[6 * idx + k] = dcov[k];
    }
    if (!ADAM || dL_dscale) {
        dL_dscale[3 * idx + 0] = dscale.x;
        dL_dscale[3 * idx + 1] = dscale.y;
        dL_dscale[3 * idx + 2] = dscale.z;
    }
    if (!ADAM || dL_drot) *reinterpret_cast<float4*>(dL_drot + 4 * idx) = drot;
    if (dL_dconic_out) *reinterpret_cast<float4*>(dL_dconic_out + 4 * idx) = make_float4(g0.z, g0.w, 0.f, g1.x);
    if (dL_dviewdepth_out) dL_dviewdepth_out[idx] = g1.z;
}

}  // namespace w3d

using namespace w3d;

template <int BATCH, int ROWS, bool DET, int MINB, typename... Args>
static cudaError_t launch_k7(dim3 grid, cudaStream_t s, Args... args) {
    auto kern = render_backward_kernel<BATCH, ROWS, DET, MINB>;
    const size_t smem = sizeof(K7Smem<BATCH, ROWS, DET>);
    // per device and cheap: no process-wide "already set" flag (several devices per process)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, TILE_PIX, smem, s>>>(args...);
    return cudaGetLastError();
}

static int raster_backward_impl(const wast3d_raster_params* prm, int num_rendered, const int* radii,
                                void* geom_buffer, void* binning_buffer, void* img_buffer,
                                const float* dL_dpix, const float* dL_ddepth, float* dL_dmean2D,
                                float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                                float* dL_dcov3D, float* dL_dsh, float* dL_dsh_rest, float* dL_dscale,
                                float* dL_drot, float* dL_dcamViewDepth, cudaStream_t s,
                                const AdamFused* adam = nullptr, const wast3d_next_view* next = nullptr) {
    const int P = prm->P, W = prm->width, H = prm->height;
    const bool debug = prm->debug != 0;
    const size_t N = (size_t)W * H;
    const dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y, 1);
    const uint32_t num_tiles = grid.x * grid.y;
    GeomState g = GeomState::carve(geom_buffer, P, nullptr);
    ImageState im = ImageState::carve(img_buffer, N, num_tiles, nullptr);
    BinningState bn = BinningState::carve(binning_buffer, (size_t)num_rendered, nullptr);
    if (radii == nullptr) radii = g.internal_radii;

    const float focal_y = H / (2.0f * prm->tan_fovy);
    const float focal_x = W / (2.0f * prm->tan_fovx);

    {
        ProfScope ps(PS_BWD_ZERO, s);
        W3D_CUDA_TRY(cudaMemsetAsync(g.grad_rec, 0, 3 * (size_t)P * sizeof(float4), s));
    }
    if (num_rendered > 0 && deterministic_mode() != 0) {
        // test mode: fixed summation order, no float atomics (see render_backward_direct_kernel<.., true>)
        ProfScope ps(PS_RENDER_BWD, s);
        // the buffers below are sized by the ACTUAL instance count; after a graph-safe forward `num_rendered` is only
        // the capacity the binning buffer was carved for, so read the count (a test mode may synchronise)
        // While the stream is being captured into a CUDA graph nothing may be read back: the buffers then take the
        // capacity and the sort passes the device-side count.
        cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
        W3D_CUDA_TRY(cudaStreamIsCapturing(s, &capture));
        const uint32_t* n_dev = nullptr;
        size_t R = (size_t)num_rendered;
        if (capture == cudaStreamCaptureStatusNone) {
            uint32_t r_dev = 0;
            W3D_CUDA_TRY(cudaMemcpyAsync(&r_dev, g.totals, sizeof(r_dev), cudaMemcpyDeviceToHost, s));
            W3D_CUDA_TRY(cudaStreamSynchronize(s));
            if (r_dev < (uint32_t)num_rendered) R = r_dev;
        } else {
            n_dev = g.totals;
        }
        const uint32_t* plist = point_list_ptr(bn, num_tiles);
        const size_t hist_words = rs_hist_words(R), hist_scr = scan_scratch_words(hist_words);
        Carver sizer(nullptr);
        sizer.take<float4>(3 * R); sizer.take<uint32_t>(R); sizer.take<uint32_t>(R); sizer.take<uint32_t>(R);
        sizer.take<uint32_t>(R); sizer.take<uint32_t>(hist_words + hist_scr + 64); sizer.take<uint32_t>(P);
        sizer.take<uint32_t>(scan_scratch_words(P) + 64);
        void* chunk = nullptr;
        W3D_CUDA_TRY(cudaMallocAsync(&chunk, sizer.bytes(), s));
        Carver c(chunk);
        float4* inst_grad = c.take<float4>(3 * R);
        uint32_t* ka = c.take<uint32_t>(R);
        uint32_t* va = c.take<uint32_t>(R);
        uint32_t* kb = c.take<uint32_t>(R);
        uint32_t* vb = c.take<uint32_t>(R);
        uint32_t* hist = c.take<uint32_t>(hist_words + hist_scr + 64);
        uint32_t* seg = c.take<uint32_t>(P);
        uint32_t* scan_scr = c.take<uint32_t>(scan_scratch_words(P) + 64);
        int st = WAST3D_OK;
        do {
            if (cudaMemsetAsync(inst_grad, 0, 3 * R * sizeof(float4), s) != cudaSuccess) { st = WAST3D_ERR_CUDA; break; }
            if (launch_k7<64, 3, true, 1>(grid, s, im.ranges, plist, W, H, prm->background, g.rec, prm->sampling_offsets,
                                       im.final_T, im.n_contrib, dL_dpix, dL_ddepth, g.grad_rec, inst_grad) != cudaSuccess) {
                st = WAST3D_ERR_CUDA;
                break;
            }
            count_launch();
            if (cudaGetLastError() != cudaSuccess) { st = WAST3D_ERR_CUDA; break; }
            // instance positions grouped by Gaussian id, ascending inside a group (stable LSD sort on the id)
            const int bits = bits_for((uint32_t)P);
            const int passes = (bits + 7) / 8;
            const uint32_t *kin = plist, *vin = nullptr;
            uint32_t *kout = ka, *vout = va;
            for (int p = 0; p < passes && st == WAST3D_OK; ++p) {
                const int nb = (p == passes - 1) ? bits - 8 * p : 8;
                st = radix_pass_u32(kin, vin, kout, vout, R, 8 * p, nb, hist, hist + hist_words, s, debug, n_dev);
                kin = kout; vin = vout;
                kout = (kout == ka) ? kb : ka;
                vout = (vout == va) ? vb : va;
            }
            if (st != WAST3D_OK) break;
            st = scan_exclusive_u32(g.tiles_touched, nullptr, seg, P, scan_scr, nullptr, s, debug);
            if (st != WAST3D_OK) break;
            det_gather_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.tiles_touched, seg, vin, inst_grad, g.grad_rec,
                                                              (uint32_t)R);
            count_launch();
            if (cudaGetLastError() != cudaSuccess) st = WAST3D_ERR_CUDA;
        } while (0);
        cudaFreeAsync(chunk, s);
        if (st == WAST3D_ERR_CUDA) set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
        if (st != WAST3D_OK) return st;
        if (debug) W3D_CUDA_TRY(cudaStreamSynchronize(s));
    } else if (num_rendered > 0) {
        ProfScope ps(PS_RENDER_BWD, s);
        // A/B switches (measurements): WAST3D_K7_ROWS = hits parked per transposed reduction (0 = shuffle butterfly),
        // WAST3D_K7_BATCH = records staged per barrier, WAST3D_K7_MINB = resident CTAs per SM the registers are cut for
        static const int rows = getenv("WAST3D_K7_ROWS") ? atoi(getenv("WAST3D_K7_ROWS")) : 0;
        static const int batch = getenv("WAST3D_K7_BATCH") ? atoi(getenv("WAST3D_K7_BATCH")) : 512;   // measured: 512
        static const int minb = getenv("WAST3D_K7_MINB") ? atoi(getenv("WAST3D_K7_MINB")) : 4;
        const uint32_t* plist = point_list_ptr(bn, num_tiles);
#define W3D_K7(B, R, M)                                                                                         \
    launch_k7<B, R, false, M>(grid, s, im.ranges, plist, W, H, prm->background, g.rec, prm->sampling_offsets,  \
                              im.final_T, im.n_contrib, dL_dpix, dL_ddepth, g.grad_rec, nullptr)
        // WAST3D_K7_MODE: 1 (default) = one warp (8x4 pixels) per CTA, 0 = one 16x16 tile per CTA
        static const int k7_mode = getenv("WAST3D_K7_MODE") ? atoi(getenv("WAST3D_K7_MODE")) : 1;
        cudaError_t e;
        if (k7_mode == 1) {
            const dim3 wgrid(2 * grid.x, 4 * grid.y, 1);
            render_backward_warp_kernel<<<wgrid, 32, 0, s>>>(im.ranges, plist, W, H, (int)grid.x, prm->background, g.rec,
                                                             prm->sampling_offsets, im.final_T, im.n_contrib, dL_dpix,
                                                             dL_ddepth, g.grad_rec);
            e = cudaGetLastError();
        } else if (rows == 2) e = batch == 512 ? W3D_K7(512, 2, 4) : W3D_K7(256, 2, 4);
        else if (rows == 3) e = W3D_K7(256, 3, 3);
        else if (batch == 512) e = minb == 5 ? W3D_K7(512, 0, 5) : W3D_K7(512, 0, 4);
        else if (batch == 128) e = minb == 5 ? W3D_K7(128, 0, 5) : W3D_K7(128, 0, 4);
        else e = minb == 5 ? W3D_K7(256, 0, 5) : minb == 6 ? W3D_K7(256, 0, 6) : W3D_K7(256, 0, 4);
#undef W3D_K7
        W3D_CUDA_TRY(e);
        W3D_AFTER_LAUNCH(s, debug);
    }
    ProfScope ps_gb(PS_GAUSS_BWD, s);
    static const int adam_u = getenv("WAST3D_ADAM_UNROLL") ? atoi(getenv("WAST3D_ADAM_UNROLL")) : 4;
    auto gb_adam = adam_u == 1 ? gaussian_backward_kernel<true, true, 1>
                 : adam_u == 2 ? gaussian_backward_kernel<true, true, 2> : gaussian_backward_kernel<true, true, 4>;
    auto gb = prm->raw_params ? (adam ? gb_adam : gaussian_backward_kernel<true, false>)
                              : gaussian_backward_kernel<false, false>;
    AdamFused af = adam ? *adam : AdamFused{};
    af.l2_prefetch = l2_prefetch_enabled() ? 1 : 0;
    NextProj nx{};
    if (next != nullptr) {
        // project every Gaussian for the next view inside this kernel (wast3d_next_view)
        if (!adam || !prm->raw_params || !next->geom_buffer || !next->viewmatrix || !next->projmatrix || !next->campos ||
            next->width <= 0 || next->height <= 0 || next->D < 0 || next->D > 3 || prm->M < (next->D + 1) * (next->D + 1))
            return WAST3D_ERR_INVALID_ARGUMENT;
        GeomState gn = GeomState::carve(next->geom_buffer, P, nullptr);
        const dim3 ngrid((next->width + TILE_X - 1) / TILE_X, (next->height + TILE_Y - 1) / TILE_Y, 1);
        nx.view.viewmatrix = next->viewmatrix;
        nx.view.projmatrix = next->projmatrix;
        nx.view.W = next->width;
        nx.view.H = next->height;
        nx.view.tan_fovx = next->tan_fovx;
        nx.view.tan_fovy = next->tan_fovy;
        nx.view.focal_y = next->height / (2.0f * next->tan_fovy);   // rasterizer_impl.cu:224-225
        nx.view.focal_x = next->width / (2.0f * next->tan_fovx);
        nx.view.grid_x = ngrid.x;
        nx.view.grid_y = ngrid.y;
        nx.view.scale_modifier = next->scale_modifier;
        nx.view.cut_tiles = wast3d_set_tile_cut(-1) != 0;
        nx.view.sb.min_x = next->offset_min_x;
        nx.view.sb.max_x = next->offset_max_x;
        nx.view.sb.min_y = next->offset_min_y;
        nx.view.sb.max_y = next->offset_max_y;
        nx.campos = next->campos;
        nx.D = next->D;
        nx.radii = next->radii ? next->radii : gn.internal_radii;
        nx.rec = gn.rec;
        nx.depth_key = gn.depth_key;
        nx.tiles_touched = gn.tiles_touched;
        nx.clamped = gn.clamped;
        nx.rect = gn.rect;
        nx.bound_words = gn.totals + 16;
        gb = gaussian_backward_kernel<true, true, 4, true>;
    }
    gb<<<(P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, s>>>(
        P, prm->D, prm->M, prm->means3D, radii, prm->shs, prm->shs_rest, g.rec, g.clamped, prm->scales,
        prm->rotations, prm->scale_modifier, prm->cov3D_precomp, prm->viewmatrix, prm->projmatrix,
        prm->campos, focal_x, focal_y, prm->tan_fovx, prm->tan_fovy, (float)(0.5 * W), (float)(0.5 * H), g.grad_rec,
        dL_dmean2D, dL_dconic,
        dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dsh_rest, dL_dscale, dL_drot,
        dL_dcamViewDepth, af, nx);
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

extern "C" int wast3d_set_deterministic(int mode) {
    const int prev = deterministic_mode();
    if (mode == 0 || mode == 1) g_deterministic.store(mode);
    return prev;
}

extern "C" int wast3d_raster_backward(const wast3d_raster_params* prm, int num_rendered,
                                      const int* radii, void* geom_buffer, void* binning_buffer,
                                      void* img_buffer, const float* dL_dpix, const float* dL_ddepth,
                                      float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                      float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                                      float* dL_dsh, float* dL_dscale, float* dL_drot,
                                      float* dL_dcamViewDepth, void* stream_v) {
    int st = validate_params(prm, false);
    if (st != WAST3D_OK) return st;
    if (prm->raw_params) return WAST3D_ERR_INVALID_ARGUMENT;  // use wast3d_raster_backward_raw
    if (prm->P == 0) return WAST3D_OK;
    if (num_rendered < 0 || !geom_buffer || !img_buffer || !binning_buffer || !dL_dpix ||
        !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale ||
        !dL_drot || (prm->shs && prm->M > 0 && !dL_dsh))
        return WAST3D_ERR_INVALID_ARGUMENT;
    return raster_backward_impl(prm, num_rendered, radii, geom_buffer, binning_buffer, img_buffer, dL_dpix,
                                dL_ddepth, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D,
                                dL_dcov3D, dL_dsh, nullptr, dL_dscale, dL_drot, dL_dcamViewDepth,
                                (cudaStream_t)stream_v);
}

extern "C" int wast3d_raster_backward_raw(const wast3d_raster_params* prm, int num_rendered,
                                          const int* radii, void* geom_buffer, void* binning_buffer,
                                          void* img_buffer, const float* dL_dpix, const float* dL_ddepth,
                                          float* dL_dxyz, float* dL_dfeatures_dc, float* dL_dfeatures_rest,
                                          float* dL_dopacity_logit, float* dL_dlog_scale,
                                          float* dL_drotation, float* dL_dmean2D, void* stream_v) {
    int st = validate_params(prm, false);
    if (st != WAST3D_OK) return st;
    if (!prm->raw_params) return WAST3D_ERR_INVALID_ARGUMENT;
    if (prm->P == 0) return WAST3D_OK;
    // dL_dfeatures_dc / dL_dfeatures_rest may both be NULL: the SH gradients are then not written (view-parallel
    // colour-record exchange: every rank rebuilds them from 16-byte records, wast3d_b200_staged.h)
    const bool feats_all = dL_dfeatures_dc != nullptr && (prm->M <= 1 || dL_dfeatures_rest != nullptr);
    const bool feats_none = dL_dfeatures_dc == nullptr && dL_dfeatures_rest == nullptr;
    if (num_rendered < 0 || !geom_buffer || !img_buffer || !binning_buffer || !dL_dpix || !dL_dxyz ||
        !(feats_all || feats_none) || !dL_dopacity_logit || !dL_dlog_scale || !dL_drotation)
        return WAST3D_ERR_INVALID_ARGUMENT;
    return raster_backward_impl(prm, num_rendered, radii, geom_buffer, binning_buffer, img_buffer, dL_dpix,
                                dL_ddepth, dL_dmean2D, nullptr, dL_dopacity_logit, nullptr, dL_dxyz, nullptr,
                                dL_dfeatures_dc, dL_dfeatures_rest, dL_dlog_scale, dL_drotation, nullptr,
                                (cudaStream_t)stream_v);
}

extern "C" int wast3d_raster_backward_raw_adam(const wast3d_raster_params* prm, int num_rendered,
                                               const int* radii, void* geom_buffer, void* binning_buffer,
                                               void* img_buffer, const float* dL_dpix, const float* dL_ddepth,
                                               const wast3d_adam_group* groups, float* const* grads_out,
                                               float* dL_dmean2D, void* stream_v) {
    return wast3d_raster_backward_raw_adam_next(prm, num_rendered, radii, geom_buffer, binning_buffer, img_buffer, dL_dpix,
                                                dL_ddepth, groups, grads_out, dL_dmean2D, nullptr, stream_v);
}

extern "C" size_t wast3d_raster_geom_bytes(int P) {
    if (P < 0) return 0;
    size_t bytes = 0;
    GeomState::carve(nullptr, (size_t)P, &bytes);
    return bytes;
}

extern "C" int wast3d_raster_backward_raw_adam_next(const wast3d_raster_params* prm, int num_rendered,
                                                    const int* radii, void* geom_buffer, void* binning_buffer,
                                                    void* img_buffer, const float* dL_dpix, const float* dL_ddepth,
                                                    const wast3d_adam_group* groups, float* const* grads_out,
                                                    float* dL_dmean2D, const wast3d_next_view* next, void* stream_v) {
    int st = validate_params(prm, false);
    if (st != WAST3D_OK) return st;
    if (!prm->raw_params || !groups) return WAST3D_ERR_INVALID_ARGUMENT;
    if (prm->P == 0) return WAST3D_OK;
    if (num_rendered < 0 || !geom_buffer || !img_buffer || !binning_buffer || !dL_dpix)
        return WAST3D_ERR_INVALID_ARGUMENT;
    // the parameters updated are the tensors the forward read (same order as GaussianModel.training_setup)
    const float* expect[6] = {prm->means3D, prm->shs, prm->shs_rest, prm->opacities, prm->scales, prm->rotations};
    AdamFused af{};
    for (int k = 0; k < 6; ++k) {
        const wast3d_adam_group& h = groups[k];
        const bool absent = (k == 2 && prm->M <= 1);
        if (absent) continue;
        if (!h.param || h.param != expect[k] || !h.exp_avg || !h.exp_avg_sq || h.step < 1)
            return WAST3D_ERR_INVALID_ARGUMENT;
        const w3d::AdamScalars sc = w3d::adam_scalars(h.lr, h.beta1, h.beta2, h.step);
        AdamSlot& d = af.g[k];
        d.p = h.param;
        d.m = h.exp_avg;
        d.v = h.exp_avg_sq;
        d.step_size = sc.step_size;
        d.inv_bc2_sqrt = sc.inv_bc2_sqrt;
        d.one_minus_b1 = sc.one_minus_b1;
        d.b2 = sc.b2;
        d.one_minus_b2 = sc.one_minus_b2;
        d.eps = h.eps;
        d.sched = h.schedule_dev;
    }
    float* go[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (grads_out)
        for (int k = 0; k < 6; ++k) go[k] = grads_out[k];
    return raster_backward_impl(prm, num_rendered, radii, geom_buffer, binning_buffer, img_buffer, dL_dpix,
                                dL_ddepth, dL_dmean2D, nullptr, go[3], nullptr, go[0], nullptr, go[1], go[2],
                                go[4], go[5], nullptr, (cudaStream_t)stream_v, &af, next);
}
