// Row-wise k smallest entries of torch.cdist(a, b) without materialising the N x M matrix (sm_100a).
//
// What it replaces (reference, all inline PyTorch on a materialised matrix — SURVEY.md §8a M1/M2):
//   D = torch.cdist(x, y); sorted, _ = torch.sort(D, 1); mask = D <= sorted[:, k-1:k]
//                                   aux_optimize_cluster_D_W_distance.py:73-82 (k = 10; 100 in v3/v4),
//                                   aux_optimize_cluster_D_W_distance2.py:269-273 (k = 20)
//   _, nn = torch.topk(torch.cdist(x, x), k=50, largest=False)   notebooks/25.4...ipynb cell 73
// Distances are torch.cdist's matmul-path values (see oracle/match_oracle.c oracle_cdist_sq),
// followed by an IEEE sqrt; the k results of a row are ordered by (distance, index), i.e. what a
// stable sort of the row returns — torch.topk leaves the order of equal distances unspecified.
//
// One warp per query row, 8 rows per CTA sharing 1024-point tiles of b staged in shared memory.
// Each warp keeps its row's k best (distance bits << 32 | index) keys sorted in shared memory;
// 32 candidates are tested per step against the current k-th key (first on the squared distance,
// so most candidates never need the sqrt) and the few survivors are inserted by the whole warp
// (ballot rank + parallel shift).
#include "common.cuh"
#include <cfloat>

namespace w3d {

constexpr int TK_WARPS = 8;
constexpr int TK_TILE = 1024;
constexpr int TK_KMAX = 128;
constexpr int TK_SLOTS = TK_KMAX / 32;

// == nn_cost_sq of match.cu / oracle_cdist_sq: desc = (x, y, z, |.|^2 as torch computes it)
__device__ __forceinline__ float tk_cost_sq(const float4 a, const float4 b) {
    float acc = 0.f;
    acc = __fmaf_rn(__fmul_rn(-2.f, a.x), b.x, acc);
    acc = __fmaf_rn(__fmul_rn(-2.f, a.y), b.y, acc);
    acc = __fmaf_rn(__fmul_rn(-2.f, a.z), b.z, acc);
    acc = __fmaf_rn(a.w, 1.f, acc);
    acc = __fmaf_rn(1.f, b.w, acc);
    return fmaxf(acc, 0.f);
}
__device__ __forceinline__ float4 tk_desc(const float* p) {
    const float x = p[0], y = p[1], z = p[2];
    return make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__global__ void __launch_bounds__(TK_WARPS * 32)
cdist_topk_kernel(int Na, int Nb, const float* __restrict__ a, const float* __restrict__ b, int k,
                  float* __restrict__ out_dist, int32_t* __restrict__ out_idx) {
    __shared__ float4 s_b[TK_TILE];
    __shared__ unsigned long long s_list[TK_WARPS][TK_KMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * TK_WARPS + warp;
    const bool row_ok = row < Na;
    unsigned long long* list = s_list[warp];
#pragma unroll
    for (int s = 0; s < TK_SLOTS; ++s) list[s * 32 + lane] = ~0ull;
    const float4 qa = row_ok ? tk_desc(a + 3 * (size_t)row) : make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned long long thresh = ~0ull;                  // key of the current k-th best
    float thresh_sq = __int_as_float(0x7f800000);       // d^2 above this cannot beat it
    __syncwarp();

    for (int base = 0; base < Nb; base += TK_TILE) {
        const int cnt = min(TK_TILE, Nb - base);
        __syncthreads();  // everyone is done with the previous tile
        for (int j = threadIdx.x; j < cnt; j += TK_WARPS * 32) s_b[j] = tk_desc(b + 3 * (size_t)(base + j));
        __syncthreads();
        if (!row_ok) continue;
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            bool cand = false;
            unsigned long long key = ~0ull;
            if (j < cnt) {
                const float sq = tk_cost_sq(qa, s_b[j]);
                if (!(sq > thresh_sq)) {
                    key = ((unsigned long long)__float_as_uint(__fsqrt_rn(sq)) << 32) | (unsigned)(base + j);
                    cand = key < thresh;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m) {  // ascending index order, one insertion at a time
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long kk = __shfl_sync(0xffffffffu, key, src);
                if (!(kk < thresh)) continue;  // the threshold tightened since the ballot
                // rank of kk in the sorted list, then shift the tail right by one
                unsigned long long mine[TK_SLOTS];
                int pos = 0;
#pragma unroll
                for (int s = 0; s < TK_SLOTS; ++s) {
                    mine[s] = list[s * 32 + lane];
                    pos += __popc(__ballot_sync(0xffffffffu, mine[s] < kk));
                }
                __syncwarp();
#pragma unroll
                for (int s = 0; s < TK_SLOTS; ++s) {
                    const int i = s * 32 + lane;
                    if (i >= pos && i + 1 < TK_KMAX) list[i + 1] = mine[s];
                }
                if (lane == 0) list[pos] = kk;
                __syncwarp();
                thresh = list[k - 1];
                const float t = __uint_as_float((unsigned)(thresh >> 32));
                thresh_sq = thresh == ~0ull ? __int_as_float(0x7f800000) : t * t * 1.000001f;
            }
        }
    }
    if (row_ok) {
        for (int i = lane; i < k; i += 32) {
            const unsigned long long e = list[i];
            out_dist[(size_t)row * k + i] = __uint_as_float((unsigned)(e >> 32));
            out_idx[(size_t)row * k + i] = (int32_t)(unsigned)(e & 0xFFFFFFFFu);
        }
    }
}

}  // namespace w3d

using namespace w3d;

extern "C" int wast3d_cdist_topk(int Na, int Nb, const float* a, const float* b, int k, float* out_dist,
                                 int32_t* out_idx, void* stream_v) {
    if (Na < 0 || Nb < 0 || k < 1 || k > TK_KMAX || k > Nb) return WAST3D_ERR_INVALID_ARGUMENT;
    if (Na == 0) return WAST3D_OK;
    if (!a || !b || !out_dist || !out_idx) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    ProfScope ps(PS_MATCH, s);
    cdist_topk_kernel<<<(Na + TK_WARPS - 1) / TK_WARPS, TK_WARPS * 32, 0, s>>>(Na, Nb, a, b, k, out_dist, out_idx);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
