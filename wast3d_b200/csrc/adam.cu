// Fused Adam step over a flat fp32 parameter group (one pass: 4 reads + 3 writes per element).
// Replaces torch.optim.Adam as configured by GaussianModel.training_setup
// (scene/gaussian_model.py:154-163: eps = 1e-15, betas (0.9, 0.999), no weight decay, no amsgrad).
// Arithmetic follows torch's single-tensor Adam update:
//   m = m + (1-b1) (g - m);  v = b2 v + (1-b2) g g;
//   p = p - (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// (adam_update in common.cuh: m, v exact; sqrt and quotient through the SFU, update term within 4 ulp)
#include "common.cuh"

namespace w3d {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float one_minus_b1,
                                         float b2, float one_minus_b2, float step_size,
                                         float inv_bc2_sqrt, float eps) {
    p = adam_update(p, g, m, v, one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps);
}

__global__ void __launch_bounds__(256)
adam_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, float one_minus_b1, float b2, float one_minus_b2,
            float step_size, float inv_bc2_sqrt, float eps, bool vec_ok) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (vec_ok) {
        const size_t n4 = n / 4;
        float4* p4 = reinterpret_cast<float4*>(p);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* m4 = reinterpret_cast<float4*>(m);
        float4* v4 = reinterpret_cast<float4*>(v);
        for (size_t i = tid; i < n4; i += stride) {
            float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
            adam_one(pp.x, gg.x, mm.x, vv.x, one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps);
            adam_one(pp.y, gg.y, mm.y, vv.y, one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps);
            adam_one(pp.z, gg.z, mm.z, vv.z, one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps);
            adam_one(pp.w, gg.w, mm.w, vv.w, one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps);
            p4[i] = pp; m4[i] = mm; v4[i] = vv;
        }
        done = n4 * 4;
    }
    for (size_t i = done + tid; i < n; i += stride)
        adam_one(p[i], g[i], m[i], v[i], one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps);
}

}  // namespace w3d

extern "C" int wast3d_adam_step(size_t n, float* param, const float* grad, float* exp_avg,
                                float* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                                int step, void* stream_v) {
    if (n == 0) return WAST3D_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    const w3d::AdamScalars sc = w3d::adam_scalars(lr, beta1, beta2, step);
    const bool vec_ok = ((((size_t)param | (size_t)grad | (size_t)exp_avg | (size_t)exp_avg_sq) & 15) == 0);
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    w3d::ProfScope ps(w3d::PS_ADAM, s);
    w3d::adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(n, param, grad, exp_avg, exp_avg_sq, sc.one_minus_b1,
                                                      sc.b2, sc.one_minus_b2, sc.step_size, sc.inv_bc2_sqrt, eps,
                                                      vec_ok);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

namespace w3d {
// wast3d_adam_schedule_step: the optimizer's step count and bias-corrected scalars live on the device, so that a
// captured launch sequence (CUDA graph) advances them itself.  Same double-precision expressions as adam_scalars().
__global__ void adam_schedule_kernel(int ngroups, const double* __restrict__ hyper, unsigned long long* __restrict__ step,
                                     float* __restrict__ schedule) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long t = *step + 1ull;
    *step = t;
    for (int k = 0; k < ngroups; ++k) {
        const double lr = hyper[3 * k], b1 = hyper[3 * k + 1], b2 = hyper[3 * k + 2];
        const double bc1 = 1.0 - pow(b1, (double)t);
        const double bc2 = 1.0 - pow(b2, (double)t);
        schedule[2 * k] = (float)(lr / bc1);
        schedule[2 * k + 1] = (float)(1.0 / sqrt(bc2));
    }
}
}  // namespace w3d

extern "C" int wast3d_adam_schedule_step(int ngroups, const double* hyper_dev, unsigned long long* step_dev,
                                         float* schedule_dev, void* stream_v) {
    if (ngroups < 1 || !hyper_dev || !step_dev || !schedule_dev) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    w3d::adam_schedule_kernel<<<1, 32, 0, s>>>(ngroups, hyper_dev, step_dev, schedule_dev);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
