// Per-Gaussian projection maths shared by the forward and backward rasteriser kernels.
// Expression shapes (operand order, parenthesisation) deliberately follow the reference so that
// nvcc's default FMA contraction produces the same roundings:
//   transformPoint4x3/4x4, ndc2Pix, getRect      cuda_rasterizer/auxiliary.h:41-77
//   computeCov3D                                 cuda_rasterizer/forward.cu:118-152
//   computeCov2D                                 cuda_rasterizer/forward.cu:74-113
// The reference uses GLM (column-major mat3, operator* as in glm/detail/type_mat3x3.inl:486-520);
// here a 9-float struct with the same element expressions replaces it.
#pragma once
#include "common.cuh"

namespace w3d {

__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                  -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                  0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                  -0.5900435899266435f};

// c[col][row], like glm::mat3
struct M3 {
    float c[3][3];
};

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int col = 0; col < 3; ++col)
#pragma unroll
        for (int row = 0; row < 3; ++row)
            r.c[col][row] = a.c[0][row] * b.c[col][0] + a.c[1][row] * b.c[col][1] +
                            a.c[2][row] * b.c[col][2];
    return r;
}
__device__ __forceinline__ M3 m3_transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int col = 0; col < 3; ++col)
#pragma unroll
        for (int row = 0; row < 3; ++row) r.c[col][row] = a.c[row][col];
    return r;
}

__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* __restrict__ m) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* __restrict__ m) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
                       m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}
__device__ __forceinline__ float3 xform_vec_4x3_T(const float3& p, const float* __restrict__ m) {
    return make_float3(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z,
                       m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

// auxiliary.h:41-44 — note the double-precision literals of the reference.
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// auxiliary.h:46-56
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, uint2& rect_min,
                                          uint2& rect_max, const dim3 grid) {
    rect_min = {min(grid.x, (unsigned)max((int)0, (int)((p.x - max_radius) / TILE_X))),
                min(grid.y, (unsigned)max((int)0, (int)((p.y - max_radius) / TILE_Y)))};
    rect_max = {min(grid.x, (unsigned)max((int)0, (int)((p.x + max_radius + TILE_X - 1) / TILE_X))),
                min(grid.y, (unsigned)max((int)0, (int)((p.y + max_radius + TILE_Y - 1) / TILE_Y)))};
}

// ---- tile rectangle restricted to the alpha >= 1/255 region ------------------------------------
// The reference instantiates a Gaussian in every tile of the square of side 2*radius around its
// centre (tile_rect above) although alpha = o*exp(power) can reach 1/255 (forward.cu:355) only inside
// the ellipse power >= -ln(255 o), whose axis-aligned bounding box [p - ext, p + ext] is stored in the
// render record (cutoff_extent, raster_forward.cu).  Instances in tiles whose SAMPLE positions
// (pixel + sampling offset, forward.cu:287) all lie outside that box are skipped by every pixel of
// the tile in forward and backward alike, so they are never created: the blend order of the
// remaining instances, the image and the gradients are unchanged; num_rendered shrinks.
// `sb` = bounds of the sampling offsets over the image: {min ox, max ox, min oy, max oy}.
struct SampleBounds { float min_x, max_x, min_y, max_y; };

// ordered-uint encoding of a float (monotone; 0 is below every float, used with atomicMax on a
// zero-initialised word)
__device__ __forceinline__ uint32_t float_order_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
// words[0..3] = order keys of max(ox), max(-ox), max(oy), max(-oy); all-zero words (no offsets
// tensor) decode to zero offsets.
__device__ __forceinline__ SampleBounds load_sample_bounds(const uint32_t* __restrict__ words) {
    SampleBounds b;
    const uint32_t k0 = words[0], k1 = words[1], k2 = words[2], k3 = words[3];
    b.max_x = k0 ? float_from_order_key(k0) : 0.f;
    b.min_x = k1 ? -float_from_order_key(k1) : 0.f;
    b.max_y = k2 ? float_from_order_key(k2) : 0.f;
    b.min_y = k3 ? -float_from_order_key(k3) : 0.f;
    return b;
}

__device__ __forceinline__ void tile_rect_cut(const float2 p, int max_radius, float ext_x, float ext_y,
                                              const SampleBounds sb, uint2& rect_min, uint2& rect_max,
                                              const dim3 grid) {
    tile_rect(p, max_radius, rect_min, rect_max, grid);
    if (ext_x < 0.f || ext_y < 0.f) {  // opacity < 1/255: alpha never reaches the threshold
        rect_max = rect_min;
        return;
    }
    // tile column tx holds samples with x in [16 tx + min_ox, 16 tx + 15 + max_ox]; it can be skipped
    // when that interval misses [p.x - ext_x, p.x + ext_x].  pad: fp32 rounding of the sums below.
    const float pad_x = 0.01f + 2e-6f * (fabsf(p.x) + ext_x);
    const float pad_y = 0.01f + 2e-6f * (fabsf(p.y) + ext_y);
    float lo_x = (p.x - ext_x - pad_x - (float)(TILE_X - 1) - sb.max_x) * (1.0f / TILE_X);
    float hi_x = (p.x + ext_x + pad_x - sb.min_x) * (1.0f / TILE_X);
    float lo_y = (p.y - ext_y - pad_y - (float)(TILE_Y - 1) - sb.max_y) * (1.0f / TILE_Y);
    float hi_y = (p.y + ext_y + pad_y - sb.min_y) * (1.0f / TILE_Y);
    // clamp in float first (inf extents / NaN centres fall back to the reference rectangle)
    lo_x = fminf(fmaxf(ceilf(lo_x), (float)rect_min.x), (float)rect_max.x);
    lo_y = fminf(fmaxf(ceilf(lo_y), (float)rect_min.y), (float)rect_max.y);
    hi_x = fmaxf(fminf(floorf(hi_x) + 1.0f, (float)rect_max.x), (float)rect_min.x);
    hi_y = fmaxf(fminf(floorf(hi_y) + 1.0f, (float)rect_max.y), (float)rect_min.y);
    const uint32_t x0 = (uint32_t)lo_x, x1 = (uint32_t)hi_x, y0 = (uint32_t)lo_y, y1 = (uint32_t)hi_y;
    if (x1 <= x0 || y1 <= y0) {
        rect_max = rect_min;
        return;
    }
    rect_min = {x0, y0};
    rect_max = {x1, y1};
}

// Rotation matrix from the quaternion AS GIVEN (r,x,y,z) — not normalised (forward.cu:127).
__device__ __forceinline__ M3 quat_to_R(const float4 q) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    M3 R;
    R.c[0][0] = 1.f - 2.f * (y * y + z * z);
    R.c[0][1] = 2.f * (x * y - r * z);
    R.c[0][2] = 2.f * (x * z + r * y);
    R.c[1][0] = 2.f * (x * y + r * z);
    R.c[1][1] = 1.f - 2.f * (x * x + z * z);
    R.c[1][2] = 2.f * (y * z - r * x);
    R.c[2][0] = 2.f * (x * z - r * y);
    R.c[2][1] = 2.f * (y * z + r * x);
    R.c[2][2] = 1.f - 2.f * (x * x + y * y);
    return R;
}

__device__ __forceinline__ M3 scale_rot_M(const float3 scale, float mod, const float4 q) {
    M3 S;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S.c[i][j] = 0.0f;
    S.c[0][0] = mod * scale.x;
    S.c[1][1] = mod * scale.y;
    S.c[2][2] = mod * scale.z;
    return m3_mul(S, quat_to_R(q));
}

// forward.cu:118-152: Sigma = M^T M, six upper-triangular entries
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 q,
                                                     float* cov6) {
    M3 M = scale_rot_M(scale, mod, q);
    M3 Sg = m3_mul(m3_transpose(M), M);
    cov6[0] = Sg.c[0][0];
    cov6[1] = Sg.c[0][1];
    cov6[2] = Sg.c[0][2];
    cov6[3] = Sg.c[1][1];
    cov6[4] = Sg.c[1][2];
    cov6[5] = Sg.c[2][2];
}

struct Cov2DCtx {
    float3 t;         // clamped view-space point
    float txtz, tytz; // unclamped ratios (for the backward's gradient mask)
    M3 T;             // W * J
    M3 Vrk;
};

// forward.cu:74-113 (and the recomputation at backward.cu:166-194)
__device__ __forceinline__ float3 cov2d(const float3& mean, float focal_x, float focal_y,
                                        float tan_fovx, float tan_fovy, const float* cov3D,
                                        const float* __restrict__ view, Cov2DCtx* ctx) {
    float3 t = xform_point_4x3(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;

    M3 J;
    J.c[0][0] = focal_x / t.z;
    J.c[0][1] = 0.0f;
    J.c[0][2] = -(focal_x * t.x) / (t.z * t.z);
    J.c[1][0] = 0.0f;
    J.c[1][1] = focal_y / t.z;
    J.c[1][2] = -(focal_y * t.y) / (t.z * t.z);
    J.c[2][0] = 0;
    J.c[2][1] = 0;
    J.c[2][2] = 0;

    M3 Wm;
    Wm.c[0][0] = view[0]; Wm.c[0][1] = view[4]; Wm.c[0][2] = view[8];
    Wm.c[1][0] = view[1]; Wm.c[1][1] = view[5]; Wm.c[1][2] = view[9];
    Wm.c[2][0] = view[2]; Wm.c[2][1] = view[6]; Wm.c[2][2] = view[10];

    M3 T = m3_mul(Wm, J);

    M3 Vrk;
    Vrk.c[0][0] = cov3D[0]; Vrk.c[0][1] = cov3D[1]; Vrk.c[0][2] = cov3D[2];
    Vrk.c[1][0] = cov3D[1]; Vrk.c[1][1] = cov3D[3]; Vrk.c[1][2] = cov3D[4];
    Vrk.c[2][0] = cov3D[2]; Vrk.c[2][1] = cov3D[4]; Vrk.c[2][2] = cov3D[5];

    M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
    if (ctx) {
        ctx->t = t;
        ctx->txtz = txtz;
        ctx->tytz = tytz;
        ctx->T = T;
        ctx->Vrk = Vrk;
    }
    // low-pass: every Gaussian at least one pixel wide (forward.cu:110-111)
    return make_float3(cov.c[0][0] + 0.3f, cov.c[0][1], cov.c[1][1] + 0.3f);
}

// ---- exact (warp pixel block, Gaussian) culling ------------------------------------------------
// A pixel receives a contribution only if alpha = o * exp(-q(d)) >= 1/255, q(d) = 0.5 (A dx^2 + C dy^2)
// + B dx dy (forward.cu:343-356), i.e. only if q(d) <= ln(255 o).  q is convex, so its minimum over
// the warp's sample rectangle [bx0,bx1] x [by0,by1] is 0 if the centre lies inside and otherwise
// sits on one of the four edges, where it is a clamped 1-D quadratic minimum.  Returns true when
// that minimum exceeds the (padded, see cutoff_extent) threshold: no pixel of the block can pass the
// reference's per-pixel test, so the pair is skipped exactly.  Conservative on any doubt (NaN,
// non-positive-definite conic): returns false.
__device__ __forceinline__ float edge_min_q(float e, float lo, float hi, float Pe, float Po, float B) {
    // min over t in [lo, hi] of 0.5 Pe e^2 + B e t + 0.5 Po t^2   (Po > 0)
    const float t = fminf(fmaxf(__fdividef(-B * e, Po), lo), hi);
    return 0.5f * Pe * e * e + t * (B * e + 0.5f * Po * t);
}
__device__ __forceinline__ bool ellipse_misses_rect(float cx, float cy, float A, float B, float C, float o,
                                                    float bx0, float bx1, float by0, float by1) {
    if (!(A > 0.f) || !(C > 0.f)) return false;
    const float det = A * C - B * B;
    if (!(det > 0.f)) return false;
    const float lx = bx0 - cx, hx = bx1 - cx, ly = by0 - cy, hy = by1 - cy;
    if (lx <= 0.f && hx >= 0.f && ly <= 0.f && hy >= 0.f) return false;  // centre inside
    float m = edge_min_q(lx, ly, hy, A, C, B);
    m = fminf(m, edge_min_q(hx, ly, hy, A, C, B));
    m = fminf(m, edge_min_q(ly, lx, hx, C, A, B));
    m = fminf(m, edge_min_q(hy, lx, hx, C, A, B));
    const float tau0 = __logf(255.0f * o);
    const float kappa = __fdividef(A * C, det);
    // padding: fp32 error of q (grows with the conditioning kappa), of __logf / __fdividef, and of m
    const float tau = tau0 + 2e-3f + fabsf(tau0) * (2e-3f + 8e-6f * kappa);
    return m * (1.0f - 1e-3f) - 1e-3f > tau;   // NaN -> false
}

// ---- per-warp SH staging -----------------------------------------------------------------
// A warp owns 32 consecutive Gaussians; their SH blocks are contiguous in memory
// (32 * M*3 floats).  Lanes stream that range with coalesced 16-byte (or 4-byte) loads into
// shared memory with an odd row stride, so the later one-thread-per-Gaussian reads are
// bank-conflict free.  Rows of Gaussians whose bit in `need` is clear are never touched, and
// only the first `used` floats of a row (active degree) are fetched.
__device__ __forceinline__ void stage_sh_rows(const float* __restrict__ g_rows, int row_floats,
                                              int used, int rows_valid, unsigned need,
                                              float* s_rows, int s_stride, int lane) {
    if ((row_floats & 3) == 0 && ((size_t)g_rows & 15) == 0) {
        const int row_v4 = row_floats >> 2;
        const int total = rows_valid * row_v4;
        const float4* g4 = reinterpret_cast<const float4*>(g_rows);
        for (int q = lane; q < total; q += 32) {
            const int g = q / row_v4, e = (q - g * row_v4) << 2;
            if (((need >> g) & 1u) && e < used) {
                float4 v = __ldg(g4 + q);
                float* d = s_rows + g * s_stride + e;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        }
    } else {
        const int total = rows_valid * row_floats;
        for (int q = lane; q < total; q += 32) {
            const int g = q / row_floats, e = q - g * row_floats;
            if (((need >> g) & 1u) && e < used) s_rows[g * s_stride + e] = __ldg(g_rows + q);
        }
    }
}

// Inverse: write rows from shared memory to a contiguous global range (every element of
// rows_valid rows is written; used by the backward for dL_dsh).
__device__ __forceinline__ void unstage_rows(float* __restrict__ g_rows, int row_floats,
                                             int rows_valid, const float* s_rows, int s_stride,
                                             int lane) {
    if ((row_floats & 3) == 0 && ((size_t)g_rows & 15) == 0) {
        const int row_v4 = row_floats >> 2;
        const int total = rows_valid * row_v4;
        float4* g4 = reinterpret_cast<float4*>(g_rows);
        for (int q = lane; q < total; q += 32) {
            const int g = q / row_v4, e = (q - g * row_v4) << 2;
            const float* s = s_rows + g * s_stride + e;
            g4[q] = make_float4(s[0], s[1], s[2], s[3]);
        }
    } else {
        const int total = rows_valid * row_floats;
        for (int q = lane; q < total; q += 32) {
            const int g = q / row_floats, e = q - g * row_floats;
            g_rows[q] = s_rows[g * s_stride + e];
        }
    }
}

// ---- [P,3] arrays (means3D, scales) with 16-byte loads -----------------------------------------
// The caller's tensors are AoS [P,3] (12 bytes per Gaussian): a warp's 32 rows are 384 contiguous bytes = 24
// float4.  Lanes 0-23 fetch one float4 of each array (coalesced 16-byte loads: 12 sector requests per array
// instead of the 36 of three strided 4-byte loads per lane; both arrays' loads are in flight together), the warp
// transposes through shared memory (stride-3 reads are bank-conflict free) and lane l gets row l of both.
// s_tmp: this warp's 192-float scratch, written once per kernel (no reuse, one __syncwarp).  b may be NULL.
__device__ __forceinline__ void load_rows3x2(const float* __restrict__ a, const float* __restrict__ b, int warp_first,
                                             int rows_valid, int lane, float* s_tmp, float3& va, float3& vb) {
    const float* ga = a + 3 * (size_t)warp_first;
    const float* gb = b ? b + 3 * (size_t)warp_first : nullptr;
    const bool vec = rows_valid == 32 && (((size_t)ga) & 15) == 0 && (gb == nullptr || (((size_t)gb) & 15) == 0);
    if (vec) {
        if (lane < 24) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(ga) + lane);
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gb) y = __ldg(reinterpret_cast<const float4*>(gb) + lane);
            reinterpret_cast<float4*>(s_tmp)[lane] = x;
            reinterpret_cast<float4*>(s_tmp + 96)[lane] = y;
        }
    } else {
        for (int i = lane; i < 3 * rows_valid; i += 32) {
            s_tmp[i] = __ldg(ga + i);
            s_tmp[96 + i] = gb ? __ldg(gb + i) : 0.f;
        }
    }
    __syncwarp();
    va = vb = make_float3(0.f, 0.f, 0.f);
    if (lane < rows_valid) {
        va = make_float3(s_tmp[3 * lane], s_tmp[3 * lane + 1], s_tmp[3 * lane + 2]);
        vb = make_float3(s_tmp[96 + 3 * lane], s_tmp[96 + 3 * lane + 1], s_tmp[96 + 3 * lane + 2]);
    }
}

// ---- model-space activations (raw_params mode; scene/gaussian_model.py:26-41) ----------------
// Same expressions as the torch kernels the reference runs for the GaussianModel getters:
// sigmoid = 1 / (1 + exp(-x)), exp, F.normalize = x / max(||x||_2, 1e-12).
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float3 act_exp3(float3 v) { return make_float3(expf(v.x), expf(v.y), expf(v.z)); }
__device__ __forceinline__ float quat_denom(float4 q) {
    return fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
}
__device__ __forceinline__ float4 act_normalize4(float4 q, float denom) {
    return make_float4(q.x / denom, q.y / denom, q.z / denom, q.w / denom);
}

// Linear staging for the model-space path: _features_rest rows are 3*(M-1) floats (45 at degree 3,
// odd => one-thread-per-row reads of a LINEAR copy are already bank-conflict free), so the
// warp's block of rows_valid rows is copied 1:1 into shared memory with 16-byte cp.async
// (float4 that may straddle two rows).  float4s that only cover rows outside `need`, or only
// elements >= used (inactive SH degrees), are skipped.  Ends with the data visible to the warp.
__device__ __forceinline__ void stage_rows_linear(const float* __restrict__ g_rows, int row_floats,
                                                  int used, int rows_valid, unsigned need,
                                                  float* s_rows, int lane) {
    const int total = rows_valid * row_floats;
    const int total4 = (((size_t)g_rows & 15) == 0 && row_floats >= 4) ? (total >> 2) : 0;
    // row of a float4: f / row_floats = umulhi(f, ceil(2^32 / row_floats)), exact for f < 2^16 (f < 32 rows x 48 floats)
    const uint32_t magic = 0xFFFFFFFFu / (uint32_t)max(row_floats, 2) + 1u;
    for (int q = lane; q < total4; q += 32) {
        const uint32_t f = 4u * (uint32_t)q;
        const uint32_t g = __umulhi(f, magic);
        const int e = (int)(f - g * (uint32_t)row_floats);
        const unsigned rows = need >> g;            // bit 0: row g, bit 1: row g + 1
        const bool spill = e + 3 >= row_floats;     // the float4 reaches into row g + 1
        const bool want = ((rows & 1u) && e < used) || (spill && (rows & 2u));
        if (want) cp_async16(s_rows + 4 * q, g_rows + 4 * q);
    }
    cp_async_commit();
    for (int i = (total4 << 2) + lane; i < total; i += 32) {  // unaligned base or tail
        const int gg = i / row_floats, ee = i - gg * row_floats;
        if (((need >> gg) & 1u) && ee < used) s_rows[i] = __ldg(g_rows + i);
    }
    cp_async_wait<0>();
    __syncwarp();
}

// Inverse: the linear shared-memory block (every element of rows_valid rows valid) to global.
__device__ __forceinline__ void unstage_rows_linear(float* __restrict__ g_rows, int row_floats,
                                                    int rows_valid, const float* s_rows, int lane) {
    const int total = rows_valid * row_floats;
    const int total4 = ((size_t)g_rows & 15) == 0 ? (total >> 2) : 0;
    float4* g4 = reinterpret_cast<float4*>(g_rows);
    const float4* s4 = reinterpret_cast<const float4*>(s_rows);
    for (int q = lane; q < total4; q += 32) g4[q] = s4[q];
    for (int i = (total4 << 2) + lane; i < total; i += 32) g_rows[i] = s_rows[i];
}

}  // namespace w3d
