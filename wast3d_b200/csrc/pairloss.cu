// Sparse pairwise-distance kernels: the geometry regularisers of the style optimisation without the
// dense N x N matrices the reference materialises.
//
// What they replace (reference paths):
//   * aux_optimize_cluster_D_W_distance.py:253-256,278-280
//       D = torch.cdist(A, xyz); loss = torch.mean(torch.abs(D - D_target) * D_xyz_target_mask)
//     with D_xyz_target_mask the k-nearest-neighbour mask of the TARGET scene (:79-82).  Only the
//     masked entries matter, so the loss is a sum over the n*k' (row, neighbour) pairs of the mask
//     (FORMULA_CDIST: each entry is evaluated the way torch.cdist's matmul path does).
//   * notebooks/25.4.Optimize_with_SAM_masks_clean.ipynb cell 72 / 29.2.Modify_style_clusters.ipynb cell 69
//       get_descriptors: X_nns = X[idx]; torch.norm(X_nns[:,1:] - X_nns[:,0].unsqueeze(1), dim=-1)
//       loss = torch.mean(torch.square(descriptors - target))
//     (FORMULA_NORM: direct difference norm; X_nns [N,k,3] is never written).
//
// Layout: `idx` [n,k] int32 row-major neighbour lists, pairs are enumerated e = i*k + j so that idx /
// target / weight / output accesses are coalesced and the a-row is a warp broadcast; b-rows are
// gathers (a cluster or scene of a few million points = tens of MB: L2 resident on B200).
// Backward: the a-row gradient is pre-reduced over the lanes of a warp that share the row (segmented
// shuffle reduction), so a row costs ~ceil(k/32)+1 atomics instead of k; b-row gradients are one
// 3-float atomic per pair.  Loss reduction: per-block double partials summed in block order by the
// last block (deterministic value); gradients are summed by unordered float atomics (stated in tests).
#include "common.cuh"

namespace w3d {

constexpr int PL_THREADS = 256;
constexpr int PL_MAX_BLOCKS = 148 * 8;
constexpr int FORMULA_NORM = 0;   // sqrt((dx^2 + dy^2) + dz^2)                     torch.norm(a - b)
constexpr int FORMULA_CDIST = 1;  // sqrt(max(0, |a|^2 + |b|^2 - 2 a.b)) in torch.cdist's matmul order

template <int FORMULA>
__device__ __forceinline__ float pair_distance(const float ax, const float ay, const float az, const float bx,
                                               const float by, const float bz) {
    if (FORMULA == FORMULA_NORM) {
        const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
        return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    } else {
        // same operation order as nn_cost_sq (match.cu) == oracle_cdist_sq (oracle/match_oracle.c)
        const float an = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
        const float bn = __fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz));
        float acc = 0.f;
        acc = __fmaf_rn(__fmul_rn(-2.f, ax), bx, acc);
        acc = __fmaf_rn(__fmul_rn(-2.f, ay), by, acc);
        acc = __fmaf_rn(__fmul_rn(-2.f, az), bz, acc);
        acc = __fmaf_rn(an, 1.f, acc);
        acc = __fmaf_rn(1.f, bn, acc);
        return sqrtf(fmaxf(acc, 0.f));
    }
}

struct PairArgs {
    long long n;          // rows
    int k;                // neighbours per row
    const float* a;       // [Na, lda]  (first three columns used)
    int lda;
    const int32_t* center;  // [n] a-row of row i, or NULL (a-row = i)
    const float* b;       // [Nb, ldb]
    int ldb;
    const int32_t* idx;   // [n,k] b-rows
    const float* row_scale;  // [n] multiplies every distance of row i, or NULL
    const float* a2;      // optional second a operand [Na, lda2]: d = dist(a, b) + dist(a2, b)  (the rotation term
    int lda2;             // cdist(rot[:, :-1], xyz) + cdist(rot[:, 1:], xyz), aux_optimize_cluster_D_W_distance.py:254)
};

__device__ __forceinline__ float sgnf(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }

// Adds (gx,gy,gz) of all lanes with the same `row` (rows are contiguous runs of lanes; row < 0 = idle
// lane) into ga[row] with one atomic triple per run.
__device__ __forceinline__ void add_row_gradient(float gx, float gy, float gz, long long row, float* __restrict__ ga,
                                                 int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float ox = __shfl_down_sync(0xffffffffu, gx, d);
        const float oy = __shfl_down_sync(0xffffffffu, gy, d);
        const float oz = __shfl_down_sync(0xffffffffu, gz, d);
        const long long orow = __shfl_down_sync(0xffffffffu, row, d);
        if (lane + d < 32 && orow == row) {
            gx += ox; gy += oy; gz += oz;
        }
    }
    const long long prev = __shfl_up_sync(0xffffffffu, row, 1);
    if (row >= 0 && (lane == 0 || prev != row)) {
        atomicAdd(ga + 3 * row + 0, gx);
        atomicAdd(ga + 3 * row + 1, gy);
        atomicAdd(ga + 3 * row + 2, gz);
    }
}

// Gradient of one pair for an upstream gradient g on its (summed) distance: (gx,gy,gz) for the a row, (hx,hy,hz)
// for the a2 row, and the b row's share added atomically.  g (a - b) / d, 0 where d == 0 (torch's norm / cdist
// backward).
template <int FORMULA>
__device__ __forceinline__ void pair_gradients(const PairArgs& p, long long ar, long long br, float g, float& gx,
                                               float& gy, float& gz, float& hx, float& hy, float& hz,
                                               float* __restrict__ grad_b) {
    const float* pa = p.a + ar * p.lda;
    const float* pb = p.b + br * p.ldb;
    const float bx = pb[0], by = pb[1], bz = pb[2];
    {
        const float ax = pa[0], ay = pa[1], az = pa[2];
        const float d = pair_distance<FORMULA>(ax, ay, az, bx, by, bz);
        const float c = d > 0.f ? g / d : 0.f;
        gx = c * (ax - bx); gy = c * (ay - by); gz = c * (az - bz);
    }
    if (p.a2) {
        const float* pc = p.a2 + ar * p.lda2;
        const float ax = pc[0], ay = pc[1], az = pc[2];
        const float d = pair_distance<FORMULA>(ax, ay, az, bx, by, bz);
        const float c = d > 0.f ? g / d : 0.f;
        hx = c * (ax - bx); hy = c * (ay - by); hz = c * (az - bz);
    }
    if (grad_b) {
        const float tx = gx + hx, ty = gy + hy, tz = gz + hz;
        if (tx != 0.f || ty != 0.f || tz != 0.f) {
            atomicAdd(grad_b + 3 * br + 0, -tx);
            atomicAdd(grad_b + 3 * br + 1, -ty);
            atomicAdd(grad_b + 3 * br + 2, -tz);
        }
    }
}

// ---- get_descriptors: out[i,j] = row_scale[i] * |a[center[i]] - b[idx[i,j]]| -------------------------------
template <int FORMULA>
__global__ void __launch_bounds__(PL_THREADS)
pair_dist_forward_kernel(const PairArgs p, float* __restrict__ out) {
    const long long E = p.n * p.k;
    for (long long e = (long long)blockIdx.x * PL_THREADS + threadIdx.x; e < E; e += (long long)gridDim.x * PL_THREADS) {
        const long long i = e / p.k;
        const long long ar = p.center ? (long long)p.center[i] : i;
        const long long br = p.idx[e];
        const float* pa = p.a + ar * p.lda;
        const float* pb = p.b + br * p.ldb;
        float d = pair_distance<FORMULA>(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2]);
        if (p.a2) {
            const float* pc = p.a2 + ar * p.lda2;
            d += pair_distance<FORMULA>(pc[0], pc[1], pc[2], pb[0], pb[1], pb[2]);
        }
        if (p.row_scale) d *= p.row_scale[i];
        out[e] = d;
    }
}

// grad_d -> grad_a [Na,3], grad_b [Nb,3] (dense, accumulated: the caller zero-fills them).  torch's backward
// of both torch.norm and torch.cdist is g * (a - b) / d with 0 where d == 0.
template <int FORMULA>
__global__ void __launch_bounds__(PL_THREADS)
pair_dist_backward_kernel(const PairArgs p, const float* __restrict__ grad_d, float* __restrict__ grad_a,
                          float* __restrict__ grad_a2, float* __restrict__ grad_b) {
    const long long E = p.n * p.k;
    const int lane = threadIdx.x & 31;
    const long long E_pad = (E + 31) / 32 * 32;
    for (long long e = (long long)blockIdx.x * PL_THREADS + threadIdx.x; e < E_pad; e += (long long)gridDim.x * PL_THREADS) {
        float gx = 0.f, gy = 0.f, gz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
        long long ar = -1;
        if (e < E) {
            const long long i = e / p.k;
            ar = p.center ? (long long)p.center[i] : i;
            float g = grad_d[e];
            if (p.row_scale) g *= p.row_scale[i];
            pair_gradients<FORMULA>(p, ar, p.idx[e], g, gx, gy, gz, hx, hy, hz, grad_b);
        }
        if (grad_a) add_row_gradient(gx, gy, gz, ar, grad_a, lane);
        if (grad_a2 && p.a2) add_row_gradient(hx, hy, hz, ar, grad_a2, lane);
    }
}

// ---- fused loss: scale * sum_e w_e * rho(d_e - t_e), rho = |.| (mode 0) or (.)^2 (mode 1) ----------------
template <int FORMULA>
__global__ void __launch_bounds__(PL_THREADS)
pair_loss_forward_kernel(const PairArgs p, const float* __restrict__ target, const float* __restrict__ weight,
                         int mode, double scale, double* __restrict__ partials, unsigned* __restrict__ counter,
                         float* __restrict__ out_loss) {
    const long long E = p.n * p.k;
    float acc = 0.f;
    for (long long e = (long long)blockIdx.x * PL_THREADS + threadIdx.x; e < E; e += (long long)gridDim.x * PL_THREADS) {
        const float w = weight ? weight[e] : 1.f;
        if (w == 0.f) continue;
        const long long i = e / p.k;
        const long long ar = p.center ? (long long)p.center[i] : i;
        const long long br = p.idx[e];
        const float* pa = p.a + ar * p.lda;
        const float* pb = p.b + br * p.ldb;
        float d = pair_distance<FORMULA>(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2]);
        if (p.a2) {
            const float* pc = p.a2 + ar * p.lda2;
            d += pair_distance<FORMULA>(pc[0], pc[1], pc[2], pb[0], pb[1], pb[2]);
        }
        if (p.row_scale) d *= p.row_scale[i];
        const float r = d - target[e];
        acc += w * (mode == 0 ? fabsf(r) : r * r);
    }
    __shared__ float red[PL_THREADS / 32];
    __shared__ bool last;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < PL_THREADS / 32; ++w) t += (double)red[w];
        partials[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) t += ((volatile double*)partials)[b];
    *out_loss = (float)(scale * t);
    *counter = 0;
}

template <int FORMULA>
__global__ void __launch_bounds__(PL_THREADS)
pair_loss_backward_kernel(const PairArgs p, const float* __restrict__ target, const float* __restrict__ weight,
                          int mode, float scale, const float* __restrict__ grad_out, float* __restrict__ grad_a,
                          float* __restrict__ grad_a2, float* __restrict__ grad_b) {
    const long long E = p.n * p.k;
    const int lane = threadIdx.x & 31;
    const long long E_pad = (E + 31) / 32 * 32;
    const float go = (grad_out ? *grad_out : 1.f) * scale;
    for (long long e = (long long)blockIdx.x * PL_THREADS + threadIdx.x; e < E_pad; e += (long long)gridDim.x * PL_THREADS) {
        float gx = 0.f, gy = 0.f, gz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
        long long ar = -1;
        if (e < E) {
            const long long i = e / p.k;
            ar = p.center ? (long long)p.center[i] : i;
            const float w = weight ? weight[e] : 1.f;
            if (w != 0.f) {
                const long long br = p.idx[e];
                const float* pa = p.a + ar * p.lda;
                const float* pb = p.b + br * p.ldb;
                float d0 = pair_distance<FORMULA>(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2]);
                if (p.a2) {
                    const float* pc = p.a2 + ar * p.lda2;
                    d0 += pair_distance<FORMULA>(pc[0], pc[1], pc[2], pb[0], pb[1], pb[2]);
                }
                const float rs = p.row_scale ? p.row_scale[i] : 1.f;
                const float r = d0 * rs - target[e];
                const float g = go * w * rs * (mode == 0 ? sgnf(r) : 2.f * r);
                pair_gradients<FORMULA>(p, ar, br, g, gx, gy, gz, hx, hy, hz, grad_b);
            }
        }
        if (grad_a) add_row_gradient(gx, gy, gz, ar, grad_a, lane);
        if (grad_a2 && p.a2) add_row_gradient(hx, hy, hz, ar, grad_a2, lane);
    }
}

static unsigned pl_blocks(long long E) {
    long long b = (E + PL_THREADS - 1) / PL_THREADS;
    if (b > PL_MAX_BLOCKS) b = PL_MAX_BLOCKS;
    return (unsigned)(b < 1 ? 1 : b);
}

static int pl_validate(const wast3d_pair_args* q) {
    if (!q || q->n < 0 || q->k < 0 || (q->formula != 0 && q->formula != 1)) return WAST3D_ERR_INVALID_ARGUMENT;
    if (q->n == 0 || q->k == 0) return WAST3D_OK;
    if (!q->a || !q->b || !q->idx || q->lda < 3 || q->ldb < 3 || (q->a2 && q->lda2 < 3))
        return WAST3D_ERR_INVALID_ARGUMENT;
    return WAST3D_OK;
}

static PairArgs pl_args(const wast3d_pair_args* q) {
    PairArgs p;
    p.n = q->n; p.k = q->k; p.a = q->a; p.lda = q->lda; p.center = q->center; p.b = q->b; p.ldb = q->ldb;
    p.idx = q->idx; p.row_scale = q->row_scale; p.a2 = q->a2; p.lda2 = q->lda2;
    return p;
}

}  // namespace w3d

using namespace w3d;

extern "C" int wast3d_pair_dist_forward(const wast3d_pair_args* q, float* out_d, void* stream_v) {
    const int rc = pl_validate(q);
    if (rc != WAST3D_OK) return rc;
    const long long E = (long long)q->n * q->k;
    if (E == 0) return WAST3D_OK;
    if (!out_d) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    const PairArgs p = pl_args(q);
    if (q->formula == FORMULA_NORM) pair_dist_forward_kernel<FORMULA_NORM><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, out_d);
    else pair_dist_forward_kernel<FORMULA_CDIST><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, out_d);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" int wast3d_pair_dist_backward(const wast3d_pair_args* q, const float* grad_d, float* grad_a,
                                         float* grad_a2, float* grad_b, void* stream_v) {
    const int rc = pl_validate(q);
    if (rc != WAST3D_OK) return rc;
    const long long E = (long long)q->n * q->k;
    if (E == 0 || (!grad_a && !grad_a2 && !grad_b)) return WAST3D_OK;
    if (!grad_d) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    const PairArgs p = pl_args(q);
    if (q->formula == FORMULA_NORM)
        pair_dist_backward_kernel<FORMULA_NORM><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, grad_d, grad_a, grad_a2, grad_b);
    else
        pair_dist_backward_kernel<FORMULA_CDIST><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, grad_d, grad_a, grad_a2, grad_b);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" size_t wast3d_pair_loss_scratch_bytes(void) { return (size_t)PL_MAX_BLOCKS * sizeof(double) + 128; }

extern "C" int wast3d_pair_loss_forward(const wast3d_pair_args* q, const float* target, const float* weight,
                                        int mode, double scale, void* scratch, float* out_loss, void* stream_v) {
    const int rc = pl_validate(q);
    if (rc != WAST3D_OK) return rc;
    if (!out_loss || !scratch || (mode != 0 && mode != 1)) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    const long long E = (long long)q->n * q->k;
    if (E == 0) {
        W3D_CUDA_TRY(cudaMemsetAsync(out_loss, 0, sizeof(float), s));
        return WAST3D_OK;
    }
    if (!target) return WAST3D_ERR_INVALID_ARGUMENT;
    unsigned* counter = (unsigned*)scratch;
    double* partials = (double*)((char*)scratch + 128);
    const PairArgs p = pl_args(q);
    if (q->formula == FORMULA_NORM)
        pair_loss_forward_kernel<FORMULA_NORM><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, target, weight, mode, scale,
                                                                                  partials, counter, out_loss);
    else
        pair_loss_forward_kernel<FORMULA_CDIST><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, target, weight, mode, scale,
                                                                                   partials, counter, out_loss);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" int wast3d_pair_loss_backward(const wast3d_pair_args* q, const float* target, const float* weight,
                                         int mode, double scale, const float* grad_out, float* grad_a,
                                         float* grad_a2, float* grad_b, void* stream_v) {
    const int rc = pl_validate(q);
    if (rc != WAST3D_OK) return rc;
    if (mode != 0 && mode != 1) return WAST3D_ERR_INVALID_ARGUMENT;
    const long long E = (long long)q->n * q->k;
    if (E == 0 || (!grad_a && !grad_a2 && !grad_b)) return WAST3D_OK;
    if (!target) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    const PairArgs p = pl_args(q);
    if (q->formula == FORMULA_NORM)
        pair_loss_backward_kernel<FORMULA_NORM><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, target, weight, mode, (float)scale,
                                                                                   grad_out, grad_a, grad_a2, grad_b);
    else
        pair_loss_backward_kernel<FORMULA_CDIST><<<pl_blocks(E), PL_THREADS, 0, s>>>(p, target, weight, mode, (float)scale,
                                                                                    grad_out, grad_a, grad_a2, grad_b);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
