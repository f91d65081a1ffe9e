// Exact optimal-transport cost between two uniformly weighted point samples of equal size
// (sm_100a, one CTA).
//
// What it replaces: `M = ot.dist(xa, xb); loss = ot.emd2(w, w, M)` with w = 1/n
// (aux_optimize_cluster_D_W_distance.py:260-270; POT is an un-vendored, un-pinned dependency of the
// reference) — a CPU network-simplex solve plus a GPU<->CPU round trip per iteration.
// With equal uniform weights the optimal plan is a permutation (Birkhoff-von Neumann), so
// emd2 = (1/n) min_sigma sum_i M[i, sigma(i)]: a linear assignment problem, solved here exactly
// with the shortest-augmenting-path (Jonker-Volgenant / Hungarian) method, one thread per column,
// dual variables in double precision.  M is POT's `ot.dist` default (squared Euclidean):
// a2[:,None] + b2[None,:] - 2 a.b^T, clamped at 0, in fp32.
// Output: the cost (fp32) and the permutation sigma (row -> column), from which the Python layer
// builds the gradient (d cost / d M = plan = P_sigma / n, which is what POT back-propagates).
#include "common.cuh"
#include <cfloat>

namespace w3d {

constexpr int EMD_MAX_N = 1023;

__global__ void __launch_bounds__(EMD_MAX_N + 1)
emd_uniform_kernel(int n, const float* __restrict__ xa, const float* __restrict__ xb,
                   float* __restrict__ M, float* __restrict__ out_cost, int32_t* __restrict__ out_perm) {
    __shared__ double u[EMD_MAX_N + 1], v[EMD_MAX_N + 1], minv[EMD_MAX_N + 1];
    __shared__ int p[EMD_MAX_N + 1], way[EMD_MAX_N + 1];
    __shared__ unsigned char used[EMD_MAX_N + 1];
    __shared__ double s_wmin[32];
    __shared__ int s_warg[32];
    __shared__ double s_delta;
    __shared__ int s_j0, s_j1;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;

    // cost matrix in POT's operation order (ot.utils.euclidean_distances, squared=True)
    for (int idx = tid; idx < n * n; idx += T) {
        const int i = idx / n, j = idx - i * n;
        const float ax = xa[3 * i], ay = xa[3 * i + 1], az = xa[3 * i + 2];
        const float bx = xb[3 * j], by = xb[3 * j + 1], bz = xb[3 * j + 2];
        const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
        const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz));
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
        float c = __fmul_rn(-2.f, dot);
        c = __fadd_rn(c, a2);
        c = __fadd_rn(c, b2);
        M[idx] = fmaxf(c, 0.f);
    }
    for (int j = tid; j <= n; j += T) { u[j] = 0.0; v[j] = 0.0; p[j] = 0; way[j] = 0; }
    __syncthreads();

    const int j = tid;  // column owned by this thread (1..n), 0 = the virtual column
    for (int i = 1; i <= n; ++i) {
        if (tid == 0) { p[0] = i; s_j0 = 0; }
        if (j <= n) { minv[j] = DBL_MAX; used[j] = 0; }
        __syncthreads();
        while (true) {
            const int j0 = s_j0;
            if (tid == 0) used[j0] = 1;
            const int i0 = p[j0];
            __syncthreads();
            double cand = DBL_MAX;
            if (j >= 1 && j <= n && !used[j]) {
                const double cur = (double)M[(size_t)(i0 - 1) * n + (j - 1)] - u[i0] - v[j];
                if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                cand = minv[j];
            }
            // block argmin, ties to the lowest column
            int arg = j;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const double oc = __shfl_xor_sync(0xffffffffu, cand, d);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
                if (oc < cand || (oc == cand && oa < arg)) { cand = oc; arg = oa; }
            }
            if (lane == 0) { s_wmin[warp] = cand; s_warg[warp] = arg; }
            __syncthreads();
            if (warp == 0) {
                double c2 = lane < (T >> 5) ? s_wmin[lane] : DBL_MAX;
                int a2 = lane < (T >> 5) ? s_warg[lane] : 0x7fffffff;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const double oc = __shfl_xor_sync(0xffffffffu, c2, d);
                    const int oa = __shfl_xor_sync(0xffffffffu, a2, d);
                    if (oc < c2 || (oc == c2 && oa < a2)) { c2 = oc; a2 = oa; }
                }
                if (lane == 0) { s_delta = c2; s_j1 = a2; }
            }
            __syncthreads();
            const double delta = s_delta;
            const int j1 = s_j1;
            if (j <= n) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }   // used columns hold distinct rows
                else minv[j] -= delta;
            }
            __syncthreads();
            if (tid == 0) s_j0 = j1;
            const bool stop = p[j1] == 0;
            __syncthreads();
            if (stop) break;
        }
        if (tid == 0) {  // augment along the alternating path
            int j0 = s_j0;
            do {
                const int j1 = way[j0];
                p[j0] = p[j1];
                j0 = j1;
            } while (j0);
        }
        __syncthreads();
    }
    if (tid == 0) {
        double tot = 0.0;
        for (int c = 1; c <= n; ++c) tot += (double)M[(size_t)(p[c] - 1) * n + (c - 1)];
        *out_cost = (float)(tot / (double)n);
    }
    if (j >= 1 && j <= n && out_perm) out_perm[p[j] - 1] = j - 1;
}

}  // namespace w3d

using namespace w3d;

extern "C" int wast3d_emd2_uniform(int n, const float* xa, const float* xb, float* out_cost,
                                   int32_t* out_perm, void* stream_v) {
    if (n < 1 || n > EMD_MAX_N || !xa || !xb || !out_cost) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    float* M = nullptr;
    W3D_CUDA_TRY(cudaMallocAsync((void**)&M, sizeof(float) * (size_t)n * n, s));
    const int threads = ((n + 1) + 31) / 32 * 32;  // one thread per column incl. the virtual column 0
    int rc = WAST3D_OK;
    {
        ProfScope ps(PS_MATCH, s);
        emd_uniform_kernel<<<1, threads, 0, s>>>(n, xa, xb, M, out_cost, out_perm);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) {
            set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
            rc = WAST3D_ERR_CUDA;
        }
    }
    cudaFreeAsync(M, s);
    return rc;
}
