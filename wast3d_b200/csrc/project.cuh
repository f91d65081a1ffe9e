// Per-Gaussian projection (K1) as device functions, shared by preprocess_kernel (raster_forward.cu) and by the
// optimizer-in-backward kernel, which can project every Gaussian for the NEXT view straight from the parameters it has
// just updated (raster_backward.cu, wast3d_next_view).  One source of truth: both callers execute the very same
// statements, so their outputs are bit-identical.
//   preprocessCUDA            cuda_rasterizer/forward.cu:155-256
//   computeColorFromSH        cuda_rasterizer/forward.cu:20-71
#pragma once
#include "raster_math.cuh"

namespace w3d {

// forward.cu:20-71.  `sh` points at this Gaussian's staged row (k-th coefficient at sh[3k..],
// k >= 1); the degree-0 coefficient is read from sh_dc (== sh unless the model-space path keeps
// _features_dc and _features_rest apart).
// Every multiply / add below is an explicit round-to-nearest intrinsic in the order nvcc's default contraction gives
// the reference expression (a - b*c -> fma(-b, c, a), left to right): the compiler has no freedom left, so every
// caller (K1 with colour, the deferred colour kernel, the optimizer kernel's projection of the next view) produces
// the same bits.  Inlined plain-C copies were free to contract different multiply-add pairs, which made
// "K1 + deferred colour" differ from "K1 with colour" in the last bit after an unrelated change of K1's loads
// (tests/test_peer_gpu.py compares those two schedules bit for bit).
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* sh_dc, const float* sh, float3 pos,
                                            float3 campos, unsigned* clamped_bits) {
    float3 dir = make_float3(__fsub_rn(pos.x, campos.x), __fsub_rn(pos.y, campos.y), __fsub_rn(pos.z, campos.z));
    const float len = __fsqrt_rn(__fmaf_rn(dir.z, dir.z, __fmaf_rn(dir.y, dir.y, __fmul_rn(dir.x, dir.x))));
    dir.x = __fdiv_rn(dir.x, len);
    dir.y = __fdiv_rn(dir.y, len);
    dir.z = __fdiv_rn(dir.z, len);
    const float x = dir.x, y = dir.y, z = dir.z;
    const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float r = __fmul_rn(SH_C0, sh_dc[c]);
        if (deg > 0) {
            r = __fmaf_rn(-__fmul_rn(SH_C1, y), sh[3 + c], r);
            r = __fmaf_rn(__fmul_rn(SH_C1, z), sh[6 + c], r);
            r = __fmaf_rn(-__fmul_rn(SH_C1, x), sh[9 + c], r);
            if (deg > 1) {
                r = __fmaf_rn(__fmul_rn(SH_C2[0], xy), sh[12 + c], r);
                r = __fmaf_rn(__fmul_rn(SH_C2[1], yz), sh[15 + c], r);
                r = __fmaf_rn(__fmul_rn(SH_C2[2], __fsub_rn(__fmaf_rn(2.0f, zz, -xx), yy)), sh[18 + c], r);
                r = __fmaf_rn(__fmul_rn(SH_C2[3], xz), sh[21 + c], r);
                r = __fmaf_rn(__fmul_rn(SH_C2[4], __fsub_rn(xx, yy)), sh[24 + c], r);
                if (deg > 2) {
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[0], y), __fmaf_rn(3.0f, xx, -yy)), sh[27 + c], r);
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[1], xy), z), sh[30 + c], r);
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[2], y), __fsub_rn(__fmaf_rn(4.0f, zz, -xx), yy)), sh[33 + c], r);
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[3], z),
                                            __fmaf_rn(-3.0f, yy, __fmaf_rn(2.0f, zz, -__fmul_rn(3.0f, xx)))), sh[36 + c], r);
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[4], x), __fsub_rn(__fmaf_rn(4.0f, zz, -xx), yy)), sh[39 + c], r);
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[5], z), __fsub_rn(xx, yy)), sh[42 + c], r);
                    r = __fmaf_rn(__fmul_rn(__fmul_rn(SH_C3[6], x), __fmaf_rn(-3.0f, yy, xx)), sh[45 + c], r);
                }
            }
        }
        res[c] = __fadd_rn(r, 0.5f);
    }
    unsigned bits = 0;
    if (res[0] < 0) bits |= 1u;
    if (res[1] < 0) bits |= 2u;
    if (res[2] < 0) bits |= 4u;
    *clamped_bits = bits;
    return make_float3(fmaxf(res[0], 0.0f), fmaxf(res[1], 0.0f), fmaxf(res[2], 0.0f));
}

// Half extents of the axis-aligned box around { d : alpha(d) >= 1/255 } for conic (A,B,C) and
// opacity o: alpha = o*exp(-q(d)), q = 0.5*(A dx^2 + C dy^2) + B dx dy, so the region is
// q <= tau = ln(255 o) and |dx| <= sqrt(2 tau C / det), |dy| <= sqrt(2 tau A / det).
// tau is padded for fp32 evaluation error of q (which grows with the conditioning A*C/det);
// the box is used only to skip pixel blocks where the reference's per-pixel test
// (forward.cu:355) is certain to reject.
__device__ __forceinline__ float2 cutoff_extent(float A, float B, float C, float o) {
    if (!(o >= 1.0f / 255.0f)) return make_float2(-1.0f, -1.0f);  // alpha <= o < 1/255 always
    // det = A C - B^2 without cancellation (Kahan: the rounding error of B*B is recovered with one fma), so that fp32
    // is enough: every other step loses a few ulp against pads of 1e-3 (measured: the double-precision log / sqrt of
    // the first version were 9 % of K1's instructions)
    const float w = __fmul_rn(B, B);
    const float det = __fadd_rn(__fmaf_rn(A, C, -w), __fmaf_rn(-B, B, w));
    const float inf = __int_as_float(0x7f800000);
    if (!(det > 0.f) || !(A > 0.f) || !(C > 0.f)) return make_float2(inf, inf);
    const float tau0 = logf(255.0f * o);
    const float inv_det = 1.0f / det;
    const float kappa = A * C * inv_det;
    const float tau = tau0 + 1e-3f + tau0 * (1e-3f + 4e-6f * kappa);
    const float hx = sqrtf(2.0f * tau * C * inv_det) * 1.001f + 0.05f;
    const float hy = sqrtf(2.0f * tau * A * inv_det) * 1.001f + 0.05f;
    if (!(hx == hx) || !(hy == hy)) return make_float2(inf, inf);
    return make_float2(hx, hy);
}

struct ProjView {   // per-view constants of the projection
    const float* viewmatrix;
    const float* projmatrix;
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    unsigned grid_x, grid_y;
    float scale_modifier;
    bool cut_tiles;
    SampleBounds sb;
};

struct Projection {
    bool visible, behind;
    int radius;
    uint32_t n_tiles;
    uint2 rect;           // {x0 | y0 << 16, width | height << 16} of the instantiated tiles
    float2 point_image;
    float depth;
    float3 conic;
    float opac;
    float2 ext;           // alpha >= 1/255 cutoff half extents
    __device__ static Projection none() {
        Projection p;
        p.visible = p.behind = false;
        p.radius = 0;
        p.n_tiles = 0;
        p.rect = make_uint2(0u, 0u);
        p.point_image = make_float2(0.f, 0.f);
        p.depth = 0.f;
        p.conic = make_float3(0.f, 0.f, 0.f);
        p.opac = 0.f;
        p.ext = make_float2(0.f, 0.f);
        return p;
    }
};

// forward.cu:155-256 for one Gaussian.  RAW: model-space inputs (log scales, unnormalised quaternion, opacity logit).
// The quaternion and the opacity are taken from q_ptr / op_ptr when those are non-NULL (loaded only where the
// reference's control flow needs them) and from q_val / op_val otherwise (values already in registers).
template <bool RAW>
__device__ __forceinline__ Projection project_gaussian(const float3 p_orig, const float3 sc_in,
                                                       const float* __restrict__ q_ptr, const float4 q_val,
                                                       const float* __restrict__ op_ptr, const float op_val,
                                                       const float* __restrict__ cov6_precomp, const ProjView& v) {
    Projection r = Projection::none();
    const dim3 grid(v.grid_x, v.grid_y, 1);
    // in_frustum (auxiliary.h:139-164): near plane only
    const float3 p_view = xform_point_4x3(p_orig, v.viewmatrix);
    r.depth = p_view.z;
    if (p_view.z <= 0.2f) {
        r.behind = true;
        return r;
    }
    float4 p_hom = xform_point_4x4(p_orig, v.projmatrix);
    float p_w = 1.0f / (p_hom.w + 0.0000001f);
    float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);

    float cov6[6];
    if (cov6_precomp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) cov6[k] = cov6_precomp[k];
    } else {
        float3 sc = sc_in;
        float4 q = q_ptr ? *reinterpret_cast<const float4*>(q_ptr) : q_val;
        if (RAW) {
            sc = act_exp3(sc);
            q = act_normalize4(q, quat_denom(q));
        }
        cov3d_from_scale_rot(sc, v.scale_modifier, q, cov6);
    }
    float3 cov = cov2d(p_orig, v.focal_x, v.focal_y, v.tan_fovx, v.tan_fovy, cov6, v.viewmatrix, nullptr);

    // EWA inverse (forward.cu:219-223)
    float det = (cov.x * cov.z - cov.y * cov.y);
    if (det != 0.0f) {
        float det_inv = 1.f / det;
        r.conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);
        float mid = 0.5f * (cov.x + cov.z);
        float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
        float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
        float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
        r.point_image = make_float2(ndc_to_pix(p_proj.x, v.W), ndc_to_pix(p_proj.y, v.H));
        uint2 rect_min, rect_max;
        tile_rect(r.point_image, (int)my_radius, rect_min, rect_max, grid);
        r.n_tiles = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
        if (r.n_tiles != 0) {
            r.visible = true;
            r.radius = (int)my_radius;
            const float op_in = op_ptr ? *op_ptr : op_val;
            r.opac = RAW ? act_sigmoid(op_in) : op_in;
            r.ext = cutoff_extent(r.conic.x, r.conic.y, r.conic.z, r.opac);
            if (v.cut_tiles) {
                // only the tiles that hold a sample inside the alpha >= 1/255 box are instantiated
                tile_rect_cut(r.point_image, r.radius, r.ext.x, r.ext.y, v.sb, rect_min, rect_max, grid);
                r.n_tiles = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
            }
            // the rectangle the instance emitter expands (8 bytes instead of re-deriving it from the record)
            if (r.n_tiles != 0)
                r.rect = make_uint2(rect_min.x | (rect_min.y << 16),
                                    (rect_max.x - rect_min.x) | ((rect_max.y - rect_min.y) << 16));
        } else {
            r.n_tiles = 0;
        }
    } else {
        r.conic = make_float3(0.f, 0.f, 0.f);
    }
    if (!r.visible) r.n_tiles = 0;
    return r;
}

// K1's outputs for Gaussian idx (geometry-buffer arrays of common.cuh GeomState + the radii tensor)
__device__ __forceinline__ void store_projection(int idx, const Projection& pr, float3 rgb, unsigned clamp_bits,
                                                 int* __restrict__ radii, float4* __restrict__ rec,
                                                 uint32_t* __restrict__ depth_key, uint32_t* __restrict__ tiles_touched,
                                                 uint8_t* __restrict__ clamped, uint2* __restrict__ rect_out) {
    radii[idx] = pr.radius;
    tiles_touched[idx] = pr.visible ? pr.n_tiles : 0u;
    rect_out[idx] = pr.rect;
    clamped[idx] = (uint8_t)(clamp_bits | (pr.visible ? 8u : 0u));  // bit 3: render record written
    depth_key[idx] = pr.visible ? __float_as_uint(pr.depth) : CULLED_KEY;
    if (pr.visible) {
        rec[3 * (size_t)idx + 0] = make_float4(pr.point_image.x, pr.point_image.y, pr.depth, pr.ext.x);
        rec[3 * (size_t)idx + 1] = make_float4(pr.conic.x, pr.conic.y, pr.conic.z, pr.opac);
        rec[3 * (size_t)idx + 2] = make_float4(rgb.x, rgb.y, rgb.z, pr.ext.y);
    }
}

}  // namespace w3d
