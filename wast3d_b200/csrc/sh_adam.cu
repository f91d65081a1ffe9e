// STAGED FOR ROUND 2 — not on any default path, declared in include/wast3d_b200_staged.h, no GPU run yet.
//
// View-parallel exchange of the SH features without moving SH gradients or SH parameters over NVLink
// (DESIGN.md §6, round-2 plan; algebra pinned on the CPU by oracle/sh_records.py + tests/test_sh_records_oracle.py).
//
// The SH gradient the reference's preprocess backward produces for one view
// (submodules/diff-gaussian-rasterization/cuda_rasterizer/backward.cu:20-139) is
//     dL/dsh[k][c] = basis_k(normalize(xyz - campos)) * dL/dRGB[c]     (0 for clamped channels / culled Gaussians)
// i.e. it depends on the view through the camera centre and three numbers per Gaussian.  So:
//   colour_record_kernel   packs what one view contributes: float4 {masked dL/dRGB, visible} per Gaussian
//                          (16 B instead of 4*3*M = 192 B), read from K7's gradient record;
//   sh_adam_records_kernel every rank reads the N views' records (its own and, through peer pointers, the other
//                          ranks'), rebuilds the summed gradient in rank order, and applies Adam to ALL of
//                          _features_dc / _features_rest in place (scene/gaussian_model.py:154-163) — the replicas
//                          stay bit-identical because every rank executes the same arithmetic on the same inputs.
// NVLink traffic per rank and step: (N-1) * 16 B * P received (vs ~2 * 192 B * P * (N-1)/N reduced + gathered).
#include <cmath>
#include "common.cuh"
#include "raster_math.cuh"

namespace w3d {
namespace staged {

constexpr int SA_THREADS = 128;
constexpr int SA_WARPS = SA_THREADS / 32;
constexpr int SA_MAX_VIEWS = 8;
constexpr int SA_MAX_REST = 45;  // 3 * (16 - 1)

struct Slot {  // one parameter group: arrays + torch's per-step Adam constants (common.cuh adam_scalars)
    float* p;
    float* m;
    float* v;
    float step_size, inv_bc2_sqrt, one_minus_b1, b2, one_minus_b2, eps;
};
struct ShAdamArgs {
    const float4* records[SA_MAX_VIEWS];
    float campos[SA_MAX_VIEWS][3];
    const float* xyz;
    Slot dc, rest;
    int P, D, M, views;
    float grad_scale;
    int l2_prefetch;
};

// round-to-nearest product / sum that the compiler may not contract into an FMA (device), plain on the host
__host__ __device__ inline float mul_rn(float x, float y) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(x, y);
#else
    volatile float r = x * y;
    return r;
#endif
}
__host__ __device__ inline float add_rn(float x, float y) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(x, y);
#else
    volatile float r = x + y;
    return r;
#endif
}

// Gradient of Gaussian `idx` summed over the views: dc[3] (coefficient 0) and mine[3(k-1) + c] (coefficients 1..).
// __host__ __device__ so that wast3d_staged_sh_adam_host_emulation can run the very same statements on the CPU.
__host__ __device__ inline void accumulate_views(const ShAdamArgs& a, int idx, float* mine, float dc[3]) {
    constexpr float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
    constexpr float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                             0.5462742152960396f};
    constexpr float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
    const int M = a.M, D = a.D;
    const int rest_floats = 3 * (M - 1);
    for (int q = 0; q < rest_floats; ++q) mine[q] = 0.f;
    const float px = a.xyz[3 * (size_t)idx], py = a.xyz[3 * (size_t)idx + 1], pz = a.xyz[3 * (size_t)idx + 2];
    const int ncoef = M < (D + 1) * (D + 1) ? M : (D + 1) * (D + 1);
    // all views' records first: up to SA_MAX_VIEWS independent 16-byte loads in flight per thread (seven of them
    // cross NVLink at N = 8) instead of one load -> basis -> accumulate round trip per view
    float4 recs[SA_MAX_VIEWS];
#pragma unroll
    for (int vw = 0; vw < SA_MAX_VIEWS; ++vw)
        if (vw < a.views) recs[vw] = a.records[vw][idx];
#pragma unroll 1
    for (int vw = 0; vw < a.views; ++vw) {
        float4 r = recs[0];
#pragma unroll
        for (int q = 1; q < SA_MAX_VIEWS; ++q)
            if (q == vw) r = recs[q];
        if (r.w == 0.f) continue;  // culled in this view: contributes exactly zero
        // direction and basis as gaussian_backward_kernel / backward.cu:36-128 write them
        const float dx = px - a.campos[vw][0], dy = py - a.campos[vw][1], dz = pz - a.campos[vw][2];
        const float len = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx / len, y = dy / len, z = dz / len;
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        float basis[16];
        basis[0] = C0;
        basis[1] = -C1 * y; basis[2] = C1 * z; basis[3] = -C1 * x;
        basis[4] = C2[0] * xy; basis[5] = C2[1] * yz; basis[6] = C2[2] * (2.f * zz - xx - yy);
        basis[7] = C2[3] * xz; basis[8] = C2[4] * (xx - yy);
        basis[9] = C3[0] * y * (3.f * xx - yy); basis[10] = C3[1] * xy * z;
        basis[11] = C3[2] * y * (4.f * zz - xx - yy);
        basis[12] = C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
        basis[13] = C3[4] * x * (4.f * zz - xx - yy); basis[14] = C3[5] * z * (xx - yy);
        basis[15] = C3[6] * x * (xx - 3.f * yy);
        const float rgb[3] = {r.x, r.y, r.z};
        // product rounded, then added in view order: the bits an all-reduce in rank order would produce
        for (int c = 0; c < 3; ++c) dc[c] = add_rn(dc[c], mul_rn(basis[0], rgb[c]));
        for (int k = 1; k < 16; ++k) {
            if (k < ncoef) {
                for (int c = 0; c < 3; ++c)
                    mine[3 * (k - 1) + c] = add_rn(mine[3 * (k - 1) + c], mul_rn(basis[k], rgb[c]));
            }
        }
    }
    if (a.grad_scale != 1.f) {
        for (int c = 0; c < 3; ++c) dc[c] *= a.grad_scale;
        for (int q = 0; q < rest_floats; ++q) mine[q] *= a.grad_scale;
    }
}

__device__ __forceinline__ float upd(float p, float g, float& m, float& v, const Slot& s) {
    return adam_update(p, g, m, v, s.one_minus_b1, s.b2, s.one_minus_b2, s.step_size, s.inv_bc2_sqrt, s.eps);
}

// K7's gradient record (raster_backward.cu: slots 8..10 = dL/dcolor) -> colour record of this view
__global__ void __launch_bounds__(256)
colour_record_kernel(int P, const int* __restrict__ radii, const float4* __restrict__ grad_rec,
                     const uint8_t* __restrict__ clamped, float4* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (radii[idx] > 0) {
        const float4 g = grad_rec[3 * (size_t)idx + 2];
        const unsigned cb = clamped[idx];
        r = make_float4((cb & 1u) ? 0.f : g.x, (cb & 2u) ? 0.f : g.y, (cb & 4u) ? 0.f : g.z, 1.f);
    }
    out[idx] = r;
}

__global__ void __launch_bounds__(SA_THREADS)
sh_adam_records_kernel(const __grid_constant__ ShAdamArgs a) {
    __shared__ __align__(16) float s_g[SA_WARPS][32 * SA_MAX_REST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * SA_THREADS + threadIdx.x;
    const int warp_first = blockIdx.x * SA_THREADS + warp * 32;
    const int P = a.P, M = a.M, D = a.D;
    if (warp_first >= P) return;
    const int rows_valid = min(32, P - warp_first);
    const int rest_floats = 3 * (M - 1);
    const bool live = idx < P;

    if (a.l2_prefetch && lane < 6) {  // the optimizer's operands are consumed after the record loop
        const Slot& sl = lane < 3 ? a.dc : a.rest;
        const int arr = lane % 3;
        l2_prefetch_rows(arr == 0 ? sl.p : arr == 1 ? sl.m : sl.v, (size_t)warp_first, rows_valid,
                         lane < 3 ? 3 : rest_floats);
    }

    float* mine = s_g[warp] + lane * rest_floats;  // this Gaussian's row of rest gradients: mine[3(k-1) + c]
    float dc[3] = {0.f, 0.f, 0.f};
    if (live) {
        accumulate_views(a, idx, mine, dc);
        // _features_dc [P,1,3]: three elements per Gaussian
        float p[3], m[3], v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p[c] = a.dc.p[3 * (size_t)idx + c]; m[c] = a.dc.m[3 * (size_t)idx + c]; v[c] = a.dc.v[3 * (size_t)idx + c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) p[c] = upd(p[c], dc[c], m[c], v[c], a.dc);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.dc.p[3 * (size_t)idx + c] = p[c]; a.dc.m[3 * (size_t)idx + c] = m[c]; a.dc.v[3 * (size_t)idx + c] = v[c];
        }
    }
    __syncwarp();
    if (rest_floats <= 0) return;

    // _features_rest [P,M-1,3]: the warp's 32 rows are one contiguous block; gradients sit in s_g[warp] in the same
    // linear order.  Same loop shape as raster_backward.cu adam_rows_linear (U float4 of p / m / v in flight per lane).
    constexpr int U = 4;
    const size_t base = (size_t)warp_first * rest_floats;
    const int total = rows_valid * rest_floats;
    float* P_ = a.rest.p + base;
    float* M_ = a.rest.m + base;
    float* V_ = a.rest.v + base;
    const float* s_grad = s_g[warp];
    if ((total & 3) == 0 && ((((size_t)P_ | (size_t)M_ | (size_t)V_) & 15) == 0)) {
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
                    p[u].x = upd(p[u].x, g.x, m[u].x, v[u].x, a.rest);
                    p[u].y = upd(p[u].y, g.y, m[u].y, v[u].y, a.rest);
                    p[u].z = upd(p[u].z, g.z, m[u].z, v[u].z, a.rest);
                    p[u].w = upd(p[u].w, g.w, m[u].w, v[u].w, a.rest);
                    P4[q] = p[u];
                    __stcs(M4 + q, m[u]);
                    __stcs(V4 + q, v[u]);
                }
            }
        }
    } else {
        for (int q = lane; q < total; q += 32) {
            float m = M_[q], v = V_[q];
            P_[q] = upd(P_[q], s_grad[q], m, v, a.rest);
            M_[q] = m;
            V_[q] = v;
        }
    }
}

static int fill_slot(Slot& d, const wast3d_adam_group& h) {
    if (!h.param || !h.exp_avg || !h.exp_avg_sq || h.step < 1) return WAST3D_ERR_INVALID_ARGUMENT;
    const AdamScalars sc = adam_scalars(h.lr, h.beta1, h.beta2, h.step);
    d.p = h.param;
    d.m = h.exp_avg;
    d.v = h.exp_avg_sq;
    d.step_size = sc.step_size;
    d.inv_bc2_sqrt = sc.inv_bc2_sqrt;
    d.one_minus_b1 = sc.one_minus_b1;
    d.b2 = sc.b2;
    d.one_minus_b2 = sc.one_minus_b2;
    d.eps = h.eps;
    return WAST3D_OK;
}

}  // namespace staged
}  // namespace w3d

using namespace w3d;
using namespace w3d::staged;

extern "C" int wast3d_staged_colour_records(int P, const int* radii, const void* geom_buffer, float* out_records,
                                            void* stream_v) {
    if (P < 0 || (P > 0 && (!radii || !geom_buffer || !out_records))) return WAST3D_ERR_INVALID_ARGUMENT;
    if (P == 0) return WAST3D_OK;
    cudaStream_t s = (cudaStream_t)stream_v;
    GeomState g = GeomState::carve(const_cast<void*>(geom_buffer), P, nullptr);
    colour_record_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, radii, g.grad_rec, g.clamped,
                                                         reinterpret_cast<float4*>(out_records));
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" int wast3d_staged_sh_adam_from_records(int P, int D, int M, int views, const float* const* records,
                                                  const float* campos_host, const float* xyz, float grad_scale,
                                                  const wast3d_adam_group* dc, const wast3d_adam_group* rest,
                                                  void* stream_v) {
    if (P < 0 || D < 0 || D > 3 || M < 1 || M > 16 || views < 1 || views > SA_MAX_VIEWS || !records || !campos_host ||
        !dc || (M > 1 && !rest) || (P > 0 && !xyz))
        return WAST3D_ERR_INVALID_ARGUMENT;
    if (P == 0) return WAST3D_OK;
    ShAdamArgs a{};
    for (int v = 0; v < views; ++v) {
        if (!records[v] || (((size_t)records[v]) & 15)) return WAST3D_ERR_INVALID_ARGUMENT;
        a.records[v] = reinterpret_cast<const float4*>(records[v]);
        for (int c = 0; c < 3; ++c) a.campos[v][c] = campos_host[3 * v + c];
    }
    a.xyz = xyz;
    int st = fill_slot(a.dc, *dc);
    if (st != WAST3D_OK) return st;
    if (M > 1) {
        st = fill_slot(a.rest, *rest);
        if (st != WAST3D_OK) return st;
    }
    a.P = P;
    a.D = D;
    a.M = M;
    a.views = views;
    a.grad_scale = grad_scale;
    a.l2_prefetch = l2_prefetch_enabled() ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream_v;
    ProfScope ps(PS_ADAM, s);
    sh_adam_records_kernel<<<(P + SA_THREADS - 1) / SA_THREADS, SA_THREADS, 0, s>>>(a);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

// CPU emulation of sh_adam_records_kernel for the tests (same accumulate_views statements; Adam with IEEE sqrt and
// division instead of the SFU approximations).  All pointers are HOST pointers.
extern "C" int wast3d_staged_sh_adam_host_emulation(int P, int D, int M, int views, const float* const* records,
                                                    const float* campos_host, const float* xyz, float grad_scale,
                                                    const wast3d_adam_group* dc, const wast3d_adam_group* rest) {
    if (P < 0 || D < 0 || D > 3 || M < 1 || M > 16 || views < 1 || views > SA_MAX_VIEWS || !records || !campos_host ||
        !dc || (M > 1 && !rest) || (P > 0 && !xyz))
        return WAST3D_ERR_INVALID_ARGUMENT;
    ShAdamArgs a{};
    for (int v = 0; v < views; ++v) {
        a.records[v] = reinterpret_cast<const float4*>(records[v]);
        for (int c = 0; c < 3; ++c) a.campos[v][c] = campos_host[3 * v + c];
    }
    a.xyz = xyz;
    int st = fill_slot(a.dc, *dc);
    if (st != WAST3D_OK) return st;
    if (M > 1 && (st = fill_slot(a.rest, *rest)) != WAST3D_OK) return st;
    a.P = P; a.D = D; a.M = M; a.views = views; a.grad_scale = grad_scale;
    const int rest_floats = 3 * (M - 1);
    auto host_upd = [](float p, float g, float& m, float& v, const Slot& s) {
        m = m + s.one_minus_b1 * (g - m);
        v = v * s.b2 + s.one_minus_b2 * g * g;
        return p - s.step_size * (m / (sqrtf(v) * s.inv_bc2_sqrt + s.eps));
    };
    float row[SA_MAX_REST];
    for (int idx = 0; idx < P; ++idx) {
        float g0[3] = {0.f, 0.f, 0.f};
        accumulate_views(a, idx, row, g0);
        for (int c = 0; c < 3; ++c) {
            const size_t e = 3 * (size_t)idx + c;
            a.dc.p[e] = host_upd(a.dc.p[e], g0[c], a.dc.m[e], a.dc.v[e], a.dc);
        }
        for (int q = 0; q < rest_floats; ++q) {
            const size_t e = (size_t)idx * rest_floats + q;
            a.rest.p[e] = host_upd(a.rest.p[e], row[q], a.rest.m[e], a.rest.v[e], a.rest);
        }
    }
    return WAST3D_OK;
}
