// Cluster matching on sm_100a: nearest style cluster per content cluster under the squared
// 2-Wasserstein distance between Gaussians (wast3d_w2_match) and plain nearest point
// (wast3d_nn_match = argmin of torch.cdist), plus the per-cluster statistics they consume.
//
// What it replaces: there is no native reference code; the reference materialises
// torch.cdist(a, b) (a cuBLAS SGEMM over [-2a, |a|^2, 1] x [b, 1, |b|^2] plus sqrt) and takes
// argmin / min / sort of the N x M matrix (aux_optimize_cluster_D_W_distance.py:73-82,253-256;
// notebooks/10.visualize_and_fit_patch_to_multiple.ipynb cell 34; notebooks/29.2... cell 58).
// The closed-form Gaussian W2 named by the north star does not exist in the reference
// (SURVEY.md §8a M5): the cost definition and its fp32 operation order are those of
// oracle/match_oracle.c, which these kernels reproduce bit for bit.
//
// Design (the cost matrix is never written to memory):
//  1. prep: per cluster a 16-float exact descriptor (mean, cov, adj(cov), det) and a 16 x bf16
//     GEMM operand row.  With u = (mean, sqrt(tr cov)) in R^4,
//         W2^2(i,j) >= |u_i - v_j|^2 = |u_i|^2 + |v_j|^2 - 2 u_i.v_j        (Schatten-Hoelder)
//     and the right-hand side is ONE tensor-core GEMM with K = 16: u and v are split into
//     bf16 hi + lo parts (hi.hi + hi.lo + lo.hi = 12 products) and the norms ride along as
//     (|u|^2_hi, |u|^2_lo, 1, 1) x (1, 1, |v|^2_hi, |v|^2_lo), pre-scaled by (1 - 2^-11) so that
//     the GEMM output is already "lower bound minus error margin".
//  2. match: one CTA = 128 content rows (UMMA M = 128) x a range of 128-column style tiles.
//     Per tile one tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), then each of the 128
//     threads reads its row with tcgen05.ld and evaluates the exact fp32 cost only for columns
//     whose bound can still beat the row's best so far; (cost, index) is merged across column
//     ranges with a 64-bit atomicMin (cost bits high, index low => ties go to the lowest index).
//  All waits on the MMA barrier are bounded; a timeout sets an error flag instead of hanging.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cfloat>
#include <atomic>

namespace w3d {

constexpr int MT_M = 128;       // content rows per CTA == UMMA M == TMEM lanes
constexpr int MT_N = 128;       // style columns per tile == UMMA N == TMEM columns
constexpr int MT_K = 16;        // one kind::f16 UMMA K step
constexpr int W2_ITERS = 10;    // oracle/match_oracle.c
constexpr float LB_SHRINK = 1.0f - 1.0f / 2048.0f;

enum MatchMode { MODE_W2 = 0, MODE_NN = 1 };

// ---------------------------------------------------------------- exact costs (== oracle) -----
__device__ __forceinline__ float dotsym(const float* A, const float* B) {
    float d = __fmul_rn(A[0], B[0]);
    d = __fmaf_rn(A[3], B[3], d);
    d = __fmaf_rn(A[5], B[5], d);
    float o = __fmul_rn(A[1], B[1]);
    o = __fmaf_rn(A[2], B[2], o);
    o = __fmaf_rn(A[4], B[4], o);
    return __fmaf_rn(2.f, o, d);
}

// oracle_w2_descriptor
__device__ __forceinline__ void w2_descriptor(const float* mean, const float* c, float* desc) {
    desc[0] = mean[0]; desc[1] = mean[1]; desc[2] = mean[2];
#pragma unroll
    for (int k = 0; k < 6; ++k) desc[3 + k] = c[k];
    float* adj = desc + 9;
    adj[0] = __fmaf_rn(c[3], c[5], -__fmul_rn(c[4], c[4]));
    adj[1] = __fmaf_rn(c[2], c[4], -__fmul_rn(c[1], c[5]));
    adj[2] = __fmaf_rn(c[1], c[4], -__fmul_rn(c[2], c[3]));
    adj[3] = __fmaf_rn(c[0], c[5], -__fmul_rn(c[2], c[2]));
    adj[4] = __fmaf_rn(c[1], c[2], -__fmul_rn(c[0], c[4]));
    adj[5] = __fmaf_rn(c[0], c[3], -__fmul_rn(c[1], c[1]));
    float det = __fmul_rn(c[0], adj[0]);
    det = __fmaf_rn(c[1], adj[1], det);
    det = __fmaf_rn(c[2], adj[2], det);
    desc[15] = fmaxf(det, 0.f);
}

// oracle_w2_cost_desc
__device__ __forceinline__ float w2_cost(const float* d1, const float* d2) {
    const float dx = __fsub_rn(d1[0], d2[0]), dy = __fsub_rn(d1[1], d2[1]), dz = __fsub_rn(d1[2], d2[2]);
    const float dist2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    const float tr1 = __fadd_rn(__fadd_rn(d1[3], d1[6]), d1[8]);
    const float tr2 = __fadd_rn(__fadd_rn(d2[3], d2[6]), d2[8]);
    const float c2 = fmaxf(dotsym(d1 + 3, d2 + 3), 0.f);
    const float c1 = fmaxf(dotsym(d1 + 9, d2 + 9), 0.f);
    const float e3 = __fsqrt_rn(__fmul_rn(d1[15], d2[15]));
    float s = __fsqrt_rn(c2);
    const float two_e3 = __fmul_rn(2.f, e3);
#pragma unroll
    for (int it = 0; it < W2_ITERS; ++it) {
        const float e2 = __fsqrt_rn(__fmaf_rn(two_e3, s, c1));
        s = __fsqrt_rn(__fmaf_rn(2.f, e2, c2));
    }
    const float w = __fmaf_rn(-2.f, s, __fadd_rn(dist2, __fadd_rn(tr1, tr2)));
    return fmaxf(w, 0.f);
}

// oracle_cdist_sq: desc = (x, y, z, |.|^2 as torch computes it)
__device__ __forceinline__ float nn_cost_sq(const float* a, const float* b) {
    float acc = 0.f;
    acc = __fmaf_rn(__fmul_rn(-2.f, a[0]), b[0], acc);
    acc = __fmaf_rn(__fmul_rn(-2.f, a[1]), b[1], acc);
    acc = __fmaf_rn(__fmul_rn(-2.f, a[2]), b[2], acc);
    acc = __fmaf_rn(a[3], 1.f, acc);
    acc = __fmaf_rn(1.f, b[3], acc);
    return fmaxf(acc, 0.f);
}

// ---------------------------------------------------------------- prep --------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// operand row (16 bf16): content side  [uh(4) | uh(4) | ul(4) | n_h n_l 1 1]
//                        style side    [-2vh(4) | -2vl(4) | -2vh(4) | 1 1 n_h n_l]
template <int MODE>
__global__ void __launch_bounds__(128)
match_prep_kernel(int K, const float* __restrict__ mean, const float* __restrict__ cov6, bool style_side,
                  float* __restrict__ desc /*[K,16]*/, __nv_bfloat16* __restrict__ oper /*[Kpad,16]*/,
                  int Kpad, unsigned long long* __restrict__ packed /* content side: [K] row results to reset */,
                  uint32_t* __restrict__ err, unsigned long long* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && err != nullptr) {   // the content-side launch also resets the call's state words
        err[0] = 0u;
        if (stats != nullptr) stats[0] = stats[1] = stats[2] = stats[3] = 0ull;
    }
    if (i >= Kpad) return;
    if (packed != nullptr && i < K) packed[i] = ~0ull;
    __nv_bfloat16 row[16];
    if (i < K) {
        float d[16];
        float u[4];
        if (MODE == MODE_W2) {
            w2_descriptor(mean + 3 * i, cov6 + 6 * i, d);
            const float tr = __fadd_rn(__fadd_rn(d[3], d[6]), d[8]);
            u[3] = sqrtf(fmaxf(tr, 0.f));
        } else {
            const float x = mean[3 * i], y = mean[3 * i + 1], z = mean[3 * i + 2];
            d[0] = x; d[1] = y; d[2] = z;
            d[3] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));  // torch: (x^2+y^2)+z^2
#pragma unroll
            for (int k = 4; k < 16; ++k) d[k] = 0.f;
            u[3] = 0.f;
        }
        u[0] = d[0]; u[1] = d[1]; u[2] = d[2];
#pragma unroll
        for (int k = 0; k < 16; ++k) desc[16 * (size_t)i + k] = d[k];
        const float nrm = (u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]) * LB_SHRINK;
        __nv_bfloat16 nh, nl;
        split_bf16(nrm, nh, nl);
        // round the norm DOWN-ish: hi+lo may exceed nrm by < 2^-17 nrm, far inside the margin
        const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat16 h, l;
            if (!style_side) {
                split_bf16(u[k], h, l);
                row[k] = h; row[4 + k] = h; row[8 + k] = l;
            } else {
                split_bf16(-2.f * u[k], h, l);
                row[k] = h; row[4 + k] = l; row[8 + k] = h;
            }
        }
        if (!style_side) { row[12] = nh; row[13] = nl; row[14] = one; row[15] = one; }
        else { row[12] = one; row[13] = one; row[14] = nh; row[15] = nl; }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) row[k] = __float2bfloat16_rn(0.f);
    }
    uint4* dst = reinterpret_cast<uint4*>(oper + 16 * (size_t)i);
    const uint4* src = reinterpret_cast<const uint4*>(row);
    dst[0] = src[0];
    dst[1] = src[1];
}

// ---------------------------------------------------------------- tcgen05 plumbing --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes.
// Our tiles keep the two K core matrices of an 8-row group adjacent (LBO = 128 B) and
// consecutive 8-row groups 256 B apart (SBO = 256 B).
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)(128 >> 4) << 16;                // leading byte offset [16,30)
    d |= (uint64_t)(256 >> 4) << 32;                // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
    return d;                                       // layout type 0 = SWIZZLE_NONE
}
// byte offset of (row r, k element e) inside such a tile
__device__ __forceinline__ int tile_off(int r, int e) {
    return (r >> 3) * 256 + (e >> 3) * 128 + (r & 7) * 16 + (e & 7) * 2;
}
constexpr uint32_t umma_idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// ---------------------------------------------------------------- match -------------------------
// grid = (row blocks, column splits); block = 512 threads.  Thread t serves content row (t & 127) of the
// block (TMEM lane t & 127) and the 32 accumulator columns [32 (t >> 7), +32) of a tile (a warp can only
// address the TMEM lane quarter warp_id % 4, so the four warps that share a quarter split the columns).
// Two CTAs (2 x 256 TMEM columns) per SM = 32 resident warps to hide the exact evaluation's sqrt chain.
//
// Per 128 x 128 tile:
//   MMA      one tcgen05.mma into one of two TMEM buffers, issued one tile AHEAD of its epilogue, operands and
//            exact style descriptors staged with cp.async two tiles ahead (double buffered);
//   phase A  every thread compares its 32 lower bounds with the row's threshold and appends the survivors
//            (row, column) to a shared candidate list (warp-aggregated append; what does not fit the list is
//            evaluated on the spot by its owner);
//   phase B  the 512 threads evaluate the candidates' exact costs — one candidate per thread and round, so
//            the long exact evaluation (20 dependent square roots for W2) never runs with 1/32 of a warp
//            active — and merge (cost, column) into the row's packed best with a shared 64-bit atomicMin
//            (cost bits high, column low: ties go to the lowest column in any evaluation order).
// The first tile of a CTA starts best-first: every row evaluates the column with the smallest bound, which
// makes the thresholds tight before the first candidate list is built.  Column splits of the same rows share
// their progress through the global packed word (read before, published after every tile).
constexpr int MT_THREADS = 512;
constexpr int MT_COLS = MT_N / (MT_THREADS / MT_M);   // accumulator columns per thread and tile (32)
constexpr int MT_CAND = 8192;                          // candidate list entries (typical tile: < 100)
constexpr int DESC_STRIDE = 20;   // floats per descriptor row in shared memory (80 B: conflict-free float4 rows)

struct MatchSmem {
    uint8_t A[MT_M * MT_K * 2];
    uint8_t B[2][MT_N * MT_K * 2];
    float desc_s[2][MT_N * DESC_STRIDE];
    float desc_c[MT_M * DESC_STRIDE];
    unsigned long long best[MT_M];
    uint16_t cand[MT_CAND];
    uint64_t bar[2];
    uint32_t tmem;
    uint32_t n_cand;
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void load_desc(const float* s, float* d) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(s + 4 * k);
        d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
    }
}
// the packed word's cost as a threshold in lower-bound units (W2: the cost itself; NN: the squared distance,
// rounded up so that every column whose ROUNDED distance can tie is still evaluated)
template <int MODE>
__device__ __forceinline__ float thresh_of(unsigned long long packed) {
    if (packed == ~0ull) return __int_as_float(0x7f800000);
    const float c = __uint_as_float((uint32_t)(packed >> 32));
    if (MODE == MODE_W2) return c;
    return __fmul_ru(__fmul_ru(c, c), 1.0000004f);
}
template <int MODE>
__device__ __forceinline__ unsigned long long exact_packed(const float* dc, const float* ds, int j) {
    float c;
    if (MODE == MODE_W2) c = w2_cost(dc, ds);
    else c = __fsqrt_rn(nn_cost_sq(dc, ds));
    return ((unsigned long long)__float_as_uint(c) << 32) | (unsigned)j;
}

template <int MODE>
__global__ void __launch_bounds__(MT_THREADS, 2)
match_kernel(int Kc, int Ks, const float* __restrict__ desc_c, const float* __restrict__ desc_s,
             const __nv_bfloat16* __restrict__ oper_c, const __nv_bfloat16* __restrict__ oper_s,
             int tiles_per_split, unsigned long long* __restrict__ best_packed,
             unsigned long long* __restrict__ stats, float* __restrict__ lb_dump, uint32_t* __restrict__ err_flag) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    MatchSmem& sm = *reinterpret_cast<MatchSmem*>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rl = tid & (MT_M - 1);          // row within the block == TMEM lane
    const int cgrp = tid >> 7;                // which MT_COLS columns of a tile this thread reads
    const int row = blockIdx.x * MT_M + rl;
    const bool row_ok = row < Kc;
    const int n_tiles = (Ks + MT_N - 1) / MT_N;
    const int tile0 = blockIdx.y * tiles_per_split;
    const int T = min(n_tiles, tile0 + tiles_per_split) - tile0;
    if (T <= 0) return;   // uniform

    if (warp == 0) tmem_alloc(&sm.tmem, 2 * MT_N);
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.n_cand = 0;
    }
    if (tid < MT_M) {
        const uint4* src = reinterpret_cast<const uint4*>(oper_c + 16 * (size_t)row);  // oper_c is padded to the grid
        *reinterpret_cast<uint4*>(sm.A + tile_off(tid, 0)) = src[0];
        *reinterpret_cast<uint4*>(sm.A + tile_off(tid, 8)) = src[1];
        sm.best[tid] = ~0ull;
        float4* dd = reinterpret_cast<float4*>(sm.desc_c + DESC_STRIDE * tid);
        if (row_ok) {
            const float4* ds = reinterpret_cast<const float4*>(desc_c + 16 * (size_t)row);
            dd[0] = ds[0]; dd[1] = ds[1]; dd[2] = ds[2]; dd[3] = ds[3];
        } else {
            dd[0] = dd[1] = dd[2] = dd[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // stage tile (tile0 + i) into buffer i & 1: operand rows (32 B per column) by threads 0-127, the exact
    // descriptors (64 B per column) two 16-byte pieces per thread
    auto stage = [&](int i) {
        const int b = i & 1;
        const int col0 = (tile0 + i) * MT_N;
        if (tid < MT_N) {
            const __nv_bfloat16* src = oper_s + 16 * (size_t)(col0 + tid);   // padded to whole tiles
            cp_async16(sm.B[b] + tile_off(tid, 0), src);
            cp_async16(sm.B[b] + tile_off(tid, 8), src + 8);
        }
        const int c = tid >> 2, h = tid & 3;   // 512 threads x 16 bytes = 128 descriptors of 64 bytes
        if (col0 + c < Ks)
            cp_async16(sm.desc_s[b] + DESC_STRIDE * c + 4 * h, desc_s + 16 * (size_t)(col0 + c) + 4 * h);
        cp_async_commit();
    };
    stage(0);
    if (T > 1) { stage(1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = sm.tmem;
    const uint32_t idesc = umma_idesc_bf16_f32(MT_M, MT_N);
    const uint64_t adesc = umma_desc_kmajor_noswizzle(smem_u32(sm.A));
    if (tid == 0) {
        umma_bf16(tmem_base, adesc, umma_desc_kmajor_noswizzle(smem_u32(sm.B[0])), idesc, 0u);
        umma_commit(&sm.bar[0]);
    }

    unsigned long long n_exact = 0, published = ~0ull;
    uint32_t phase0 = 0u, phase1 = 0u;   // mbarrier parities of the two TMEM buffers
    bool failed = false;
    const float* my_dc = sm.desc_c + DESC_STRIDE * rl;

    for (int i = 0; i < T; ++i) {
        const int b = i & 1;
        const int col0 = (tile0 + i) * MT_N;
        // the next tile's MMA goes first: it runs while this tile's epilogue does
        if (i + 1 < T) {
            cp_async_wait<0>();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();   // also: n_cand == 0 and last tile's best[] merges are visible
        if (i + 1 < T && tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_bf16(tmem_base + (uint32_t)((b ^ 1) * MT_N), adesc,
                      umma_desc_kmajor_noswizzle(smem_u32(sm.B[b ^ 1])), idesc, 0u);
            umma_commit(&sm.bar[b ^ 1]);
        }
        // bounded wait for this tile's MMA (never hang the device on a protocol bug)
        {
            uint32_t spins = 0;
            const uint32_t parity = b ? phase1 : phase0;
            while (!mbar_try_wait(&sm.bar[b], parity)) {
                if (++spins > (1u << 22)) { failed = true; break; }
            }
            if (b) phase1 ^= 1u; else phase0 ^= 1u;
        }
        if (__syncthreads_or(failed)) { failed = true; break; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // progress of the other column splits of this row
        if (tid < MT_M && row_ok) {
            const unsigned long long g = ld_relaxed_u64(best_packed + row);
            if (g < sm.best[tid]) sm.best[tid] = g;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(b * MT_N + cgrp * MT_COLS);
        const float* dsb = sm.desc_s[b];
        static_assert(MT_COLS == 32, "one tcgen05.ld.32x32b.x32 per thread and tile");
        const int cbase = cgrp * MT_COLS;
        int kmin = -1;
        if (i == 0) {
            // best-first: exact cost of the column with the smallest bound among this thread's columns
            // (its own TMEM read: the 32 bounds are not kept in registers across the exact evaluation)
            float vmin = __int_as_float(0x7f800000);
            {
                float v0[MT_COLS];
                tmem_ld32(taddr, v0);
#pragma unroll
                for (int k = 0; k < MT_COLS; ++k)
                    if (col0 + cbase + k < Ks && v0[k] < vmin) { vmin = v0[k]; kmin = k; }
            }
            __syncthreads();   // the reads of best_packed above are merged before anybody min()s into best[]
            if (row_ok && kmin >= 0) {
                atomicMin(&sm.best[rl], exact_packed<MODE>(my_dc, dsb + DESC_STRIDE * (cbase + kmin), col0 + cbase + kmin));
                ++n_exact;
            }
        }
        __syncthreads();
        const float thresh = thresh_of<MODE>(sm.best[rl]);

        // ---- phase A: candidates whose (margin-adjusted) lower bound can still beat the row's best
        {
            float v[MT_COLS];
            tmem_ld32(taddr, v);
            if (lb_dump != nullptr && row_ok) {
#pragma unroll
                for (int k = 0; k < MT_COLS; ++k)
                    if (col0 + cbase + k < Ks) lb_dump[(size_t)row * Ks + col0 + cbase + k] = v[k];
            }
            // one compare + one bit insert per bound; validity (row, column range, the best-first column) as masks
            unsigned mask = 0;
#pragma unroll
            for (int k = 0; k < MT_COLS; ++k) mask |= (v[k] > thresh ? 0u : 1u) << k;
            const int n_valid = Ks - (col0 + cbase);
            if (n_valid < MT_COLS) mask &= n_valid > 0 ? ((1u << n_valid) - 1u) : 0u;
            if (kmin >= 0) mask &= ~(1u << kmin);
            if (!row_ok) mask = 0;
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int base = 0;
            if (lane == 31 && total) base = (int)atomicAdd(&sm.n_cand, (uint32_t)total);
            base = __shfl_sync(0xffffffffu, base, 31);
            int pos = base + incl - cnt;
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1;
                if (pos < MT_CAND) {
                    sm.cand[pos] = (uint16_t)((rl << 7) | (cbase + k));
                } else {   // list full (degenerate data: thousands of near-ties in one tile): evaluate here
                    atomicMin(&sm.best[rl], exact_packed<MODE>(my_dc, dsb + DESC_STRIDE * (cbase + k), col0 + cbase + k));
                    ++n_exact;
                }
                ++pos;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();

        // ---- phase B: exact costs, one candidate per thread and round
        const int n = min((int)sm.n_cand, MT_CAND);
        for (int k = tid; k < n; k += MT_THREADS) {
            const int cd = sm.cand[k];
            const int r = cd >> 7, cl = cd & 127;
            atomicMin(&sm.best[r], exact_packed<MODE>(sm.desc_c + DESC_STRIDE * r, dsb + DESC_STRIDE * cl, col0 + cl));
        }
        if (tid == 0) n_exact += (unsigned long long)n;
        __syncthreads();
        if (tid == 0) sm.n_cand = 0;
        if (tid < MT_M && row_ok) {
            const unsigned long long mine = sm.best[tid];
            if (mine < published) { atomicMin(best_packed + row, mine); published = mine; }
        }
        if (i + 2 < T) stage(i + 2);   // into the buffers this tile just released
    }
    cp_async_wait<0>();

    if (failed && tid == 0) atomicOr(err_flag, 1u);
    if (stats != nullptr) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_exact += __shfl_xor_sync(0xffffffffu, n_exact, d);
        if (lane == 0 && n_exact) atomicAdd(stats + 1, n_exact);
        if (tid == 0) atomicAdd(stats + 2, (unsigned long long)T);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 2 * MT_N);
}

// out_idx = -2 / cost = NaN for every row when the MMA barrier timed out (never expected): the caller sees
// the failure without a host round trip in the call itself.
__global__ void match_finalize_kernel(int Kc, const unsigned long long* __restrict__ packed,
                                      int32_t* __restrict__ out_idx, float* __restrict__ out_cost,
                                      const uint32_t* __restrict__ err, unsigned long long* __restrict__ stats,
                                      unsigned long long pairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && stats != nullptr) { stats[0] = pairs; stats[3] = *err; }
    if (i >= Kc) return;
    const unsigned long long p = packed[i];
    if (*err) {
        out_idx[i] = -2;
        if (out_cost) out_cost[i] = __int_as_float(0x7fc00000);
    } else if (p == ~0ull) {
        out_idx[i] = -1;
        if (out_cost) out_cost[i] = __int_as_float(0x7f800000);
    } else {
        out_idx[i] = (int32_t)(p & 0xFFFFFFFFu);
        if (out_cost) out_cost[i] = __uint_as_float((uint32_t)(p >> 32));
    }
}

// ---------------------------------------------------------------- cluster statistics ------------
// Two passes with double accumulators (atomicAdd(double) is native): order independent up to
// double rounding, which vanishes in the final fp32 result.
__global__ void __launch_bounds__(256)
cluster_sum_kernel(int n, int K, const float* __restrict__ pts, const int32_t* __restrict__ labels,
                   double* __restrict__ sum, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = labels[i];
    if (k < 0 || k >= K) return;
    atomicAdd(count + k, 1);
    atomicAdd(sum + 3 * k + 0, (double)pts[3 * i + 0]);
    atomicAdd(sum + 3 * k + 1, (double)pts[3 * i + 1]);
    atomicAdd(sum + 3 * k + 2, (double)pts[3 * i + 2]);
}
__global__ void cluster_mean_kernel(int K, double* __restrict__ sum, const int32_t* __restrict__ count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int c = count[k];
    for (int d = 0; d < 3; ++d) sum[3 * k + d] = c ? sum[3 * k + d] / c : 0.0;
}
__global__ void __launch_bounds__(256)
cluster_cov_kernel(int n, int K, const float* __restrict__ pts, const int32_t* __restrict__ labels,
                   const double* __restrict__ mean, double* __restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = labels[i];
    if (k < 0 || k >= K) return;
    const double dx = pts[3 * i] - mean[3 * k], dy = pts[3 * i + 1] - mean[3 * k + 1], dz = pts[3 * i + 2] - mean[3 * k + 2];
    double* a = acc + 6 * k;
    atomicAdd(a + 0, dx * dx); atomicAdd(a + 1, dx * dy); atomicAdd(a + 2, dx * dz);
    atomicAdd(a + 3, dy * dy); atomicAdd(a + 4, dy * dz); atomicAdd(a + 5, dz * dz);
}
__global__ void cluster_out_kernel(int K, const double* __restrict__ mean, const double* __restrict__ acc,
                                   const int32_t* __restrict__ count, float* __restrict__ mean_out,
                                   float* __restrict__ cov_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int c = count[k];
    for (int d = 0; d < 3; ++d) mean_out[3 * k + d] = (float)mean[3 * k + d];
    for (int d = 0; d < 6; ++d) cov_out[6 * k + d] = c ? (float)(acc[6 * k + d] / c) : 0.f;
}

// ---------------------------------------------------------------- K-Means (Lloyd) ---------------
// One Lloyd iteration = nearest-centre assignment (the MODE_NN match above: tensor-core lower bound +
// exact fp32 cdist evaluation, ties to the lowest centre) followed by the centre update below.
// Sums are double atomics (order independent after rounding to fp32).  acc: [K,3] sums, cnt: [K],
// stat[0] = labels changed, stat[1] (as double bits in stat64) = inertia.
__global__ void __launch_bounds__(256)
kmeans_accumulate_kernel(int n, const float* __restrict__ pts, const int32_t* __restrict__ new_label,
                         const float* __restrict__ dist, int32_t* __restrict__ label, double* __restrict__ acc,
                         int32_t* __restrict__ cnt, unsigned long long* __restrict__ changed,
                         double* __restrict__ inertia) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int ch = 0;
    double d2 = 0.0;
    if (i < n) {
        const int k = new_label[i];
        if (k >= 0) {   // k < 0: the match failed (error word set; the host stops the loop)
            ch = label[i] != k;
            label[i] = k;
            const float d = dist[i];
            d2 = (double)d * (double)d;
            atomicAdd(cnt + k, 1);
            atomicAdd(acc + 3 * k + 0, (double)pts[3 * i + 0]);
            atomicAdd(acc + 3 * k + 1, (double)pts[3 * i + 1]);
            atomicAdd(acc + 3 * k + 2, (double)pts[3 * i + 2]);
        }
    }
    // block-level pre-reduction of the two scalars
    __shared__ double s_d[8];
    __shared__ int s_c[8];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        ch += __shfl_xor_sync(0xffffffffu, ch, o);
    }
    if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = d2; s_c[threadIdx.x >> 5] = ch; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; int c = 0;
        for (int w = 0; w < 8; ++w) { t += s_d[w]; c += s_c[w]; }
        atomicAdd(inertia, t);
        if (c) atomicAdd(changed, (unsigned long long)c);
    }
}
// new centre = mean of members (an empty cluster keeps its centre); shift2 += |new - old|^2
__global__ void kmeans_update_kernel(int K, const double* __restrict__ acc, const int32_t* __restrict__ cnt,
                                     float* __restrict__ centers, double* __restrict__ shift2) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int c = cnt[k];
    if (c == 0) return;
    double s = 0.0;
    for (int d = 0; d < 3; ++d) {
        const float nc = (float)(acc[3 * k + d] / c);
        const double df = (double)nc - (double)centers[3 * k + d];
        s += df * df;
        centers[3 * k + d] = nc;
    }
    atomicAdd(shift2, s);
}

// ---------------------------------------------------------------- host --------------------------
struct MatchScratch {
    float* desc_c;
    float* desc_s;
    __nv_bfloat16* oper_c;
    __nv_bfloat16* oper_s;
    unsigned long long* packed;
    uint32_t* err;
    static MatchScratch carve(void* chunk, int Kc, int Ks, size_t* bytes) {
        const int Kc_pad = (Kc + MT_M - 1) / MT_M * MT_M;
        const int Ks_pad = (Ks + MT_N - 1) / MT_N * MT_N;
        Carver c(chunk);
        MatchScratch m;
        m.desc_c = c.take<float>((size_t)(Kc > 0 ? Kc : 1) * 16);
        m.desc_s = c.take<float>((size_t)(Ks > 0 ? Ks : 1) * 16);
        m.oper_c = c.take<__nv_bfloat16>((size_t)(Kc_pad > 0 ? Kc_pad : MT_M) * 16);
        m.oper_s = c.take<__nv_bfloat16>((size_t)(Ks_pad > 0 ? Ks_pad : MT_N) * 16);
        m.packed = c.take<unsigned long long>(Kc > 0 ? Kc : 1);
        m.err = c.take<uint32_t>(4);
        if (bytes) *bytes = c.bytes();
        return m;
    }
};

// Internal stream-ordered allocations (callers that pass no scratch, K-Means, cluster statistics): keep freed
// blocks in the device's default pool instead of returning them to the driver at every synchronisation
// (release threshold 0 made a call cost a cudaMalloc — milliseconds — whenever the stream had been synced).
static void keep_pool_memory() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    static std::atomic<unsigned long long> done_mask{0};
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_mask.load() & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done_mask.fetch_or(bit);
}

// No host synchronisation: the (never expected) MMA time-out is reported through out_idx = -2 / stats[3].
// err_host_check != nullptr (K-Means, which synchronises anyway) additionally receives the device address of
// the error word.
template <int MODE>
static int run_match(int Kc, int Ks, const float* mean_c, const float* cov_c, const float* mean_s,
                     const float* cov_s, int32_t* out_idx, float* out_cost, unsigned long long* stats,
                     float* lb_dump, void* scratch, size_t scratch_bytes, cudaStream_t s,
                     uint32_t** err_dev = nullptr) {
    if (Kc < 0 || Ks < 0) return WAST3D_ERR_INVALID_ARGUMENT;
    if (Kc == 0) return WAST3D_OK;
    if (!mean_c || !out_idx || (Ks > 0 && !mean_s)) return WAST3D_ERR_INVALID_ARGUMENT;
    if (MODE == MODE_W2 && (!cov_c || (Ks > 0 && !cov_s))) return WAST3D_ERR_INVALID_ARGUMENT;
    const int Kc_pad = (Kc + MT_M - 1) / MT_M * MT_M;
    const int Ks_pad = (Ks + MT_N - 1) / MT_N * MT_N;
    size_t need = 0;
    MatchScratch::carve(nullptr, Kc, Ks, &need);
    void* chunk = scratch;
    if (chunk != nullptr) {
        if (scratch_bytes < need || ((size_t)chunk & 127) != 0) return WAST3D_ERR_INVALID_ARGUMENT;
    } else {
        keep_pool_memory();
        W3D_CUDA_TRY(cudaMallocAsync(&chunk, need, s));
    }
    const MatchScratch m = MatchScratch::carve(chunk, Kc, Ks, nullptr);
    if (err_dev) *err_dev = m.err;
    int rc = WAST3D_OK;
    ProfScope ps(PS_MATCH, s);
    do {
        match_prep_kernel<MODE><<<(Kc_pad + 127) / 128, 128, 0, s>>>(Kc, mean_c, cov_c, false, m.desc_c, m.oper_c,
                                                                     Kc_pad, m.packed, m.err, stats);
        count_launch();
        if (Ks > 0) {
            match_prep_kernel<MODE><<<(Ks_pad + 127) / 128, 128, 0, s>>>(Ks, mean_s, cov_s, true, m.desc_s, m.oper_s,
                                                                         Ks_pad, nullptr, nullptr, nullptr);
            count_launch();
        }
        if (cudaGetLastError() != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
        if (Ks > 0) {
            const int row_blocks = Kc_pad / MT_M;
            const int n_tiles = Ks_pad / MT_N;
            // two CTAs (2 x 256 TMEM columns, 16 warps each) per SM: as many column splits as still fit ONE wave
            int splits = (2 * 148) / row_blocks;
            if (splits > n_tiles) splits = n_tiles;
            if (splits < 1) splits = 1;
            const int tiles_per_split = (n_tiles + splits - 1) / splits;
            splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
            // per device, cheap: no process-wide "already set" flag (several devices per process)
            if (cudaFuncSetAttribute(match_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(MatchSmem)) != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
            match_kernel<MODE><<<dim3(row_blocks, splits), MT_THREADS, sizeof(MatchSmem), s>>>(
                Kc, Ks, m.desc_c, m.desc_s, m.oper_c, m.oper_s, tiles_per_split, m.packed, stats, lb_dump, m.err);
            count_launch();
            if (cudaGetLastError() != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
        }
        match_finalize_kernel<<<(Kc + 255) / 256, 256, 0, s>>>(Kc, m.packed, out_idx, out_cost, m.err, stats,
                                                               (unsigned long long)Kc * (unsigned long long)Ks);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
    } while (0);
    if (rc == WAST3D_ERR_CUDA) set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
    if (scratch == nullptr) cudaFreeAsync(chunk, s);
    return rc;
}

}  // namespace w3d

using namespace w3d;

extern "C" size_t wast3d_match_scratch_bytes(int Kc, int Ks) {
    if (Kc < 0 || Ks < 0) return 0;
    size_t need = 0;
    MatchScratch::carve(nullptr, Kc, Ks, &need);
    return need;
}

extern "C" int wast3d_w2_match(int Kc, int Ks, const float* mean_c, const float* cov_c,
                               const float* mean_s, const float* cov_s, int32_t* out_idx,
                               float* out_cost, unsigned long long* stats, void* scratch,
                               size_t scratch_bytes, void* stream_v) {
    return run_match<MODE_W2>(Kc, Ks, mean_c, cov_c, mean_s, cov_s, out_idx, out_cost, stats, nullptr,
                              scratch, scratch_bytes, (cudaStream_t)stream_v);
}

extern "C" int wast3d_nn_match(int Na, int Nb, const float* a, const float* b, int32_t* out_idx,
                               float* out_dist, void* scratch, size_t scratch_bytes, void* stream_v) {
    return run_match<MODE_NN>(Na, Nb, a, nullptr, b, nullptr, out_idx, out_dist, nullptr, nullptr,
                              scratch, scratch_bytes, (cudaStream_t)stream_v);
}

// Test hook: also dumps the tensor-core lower-bound matrix [Kc,Ks] (see tests/test_knn_match_gpu.py).
extern "C" int wast3d_w2_match_debug(int Kc, int Ks, const float* mean_c, const float* cov_c,
                                     const float* mean_s, const float* cov_s, int32_t* out_idx,
                                     float* out_cost, unsigned long long* stats, float* lb_dump,
                                     void* stream_v) {
    return run_match<MODE_W2>(Kc, Ks, mean_c, cov_c, mean_s, cov_s, out_idx, out_cost, stats, lb_dump,
                              nullptr, 0, (cudaStream_t)stream_v);
}

extern "C" int wast3d_cluster_stats(int n, int K, const float* points, const int32_t* labels,
                                    float* mean, float* cov6, int32_t* count, void* stream_v) {
    if (n < 0 || K < 0) return WAST3D_ERR_INVALID_ARGUMENT;
    if (K == 0) return WAST3D_OK;
    if (!mean || !cov6 || !count || (n > 0 && (!points || !labels))) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    double* buf = nullptr;
    keep_pool_memory();
    W3D_CUDA_TRY(cudaMallocAsync((void**)&buf, sizeof(double) * 9 * (size_t)K, s));
    double* sum = buf;
    double* acc = buf + 3 * (size_t)K;
    int rc = WAST3D_OK;
    do {
        if (cudaMemsetAsync(buf, 0, sizeof(double) * 9 * (size_t)K, s) != cudaSuccess ||
            cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)K, s) != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
        if (n > 0) cluster_sum_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, K, points, labels, sum, count);
        cluster_mean_kernel<<<(K + 255) / 256, 256, 0, s>>>(K, sum, count);
        if (n > 0) cluster_cov_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, K, points, labels, sum, acc);
        cluster_out_kernel<<<(K + 255) / 256, 256, 0, s>>>(K, sum, acc, count, mean, cov6);
        if (cudaGetLastError() != cudaSuccess) rc = WAST3D_ERR_CUDA;
    } while (0);
    if (rc == WAST3D_ERR_CUDA) set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
    cudaFreeAsync(buf, s);
    return rc;
}

// Point-sharded statistics (BASELINE.json configs[3]): the two accumulation passes of
// wast3d_cluster_stats on caller-owned accumulators, so that N ranks can each add their share of the points and
// all-reduce 4 + 6 numbers per cluster in between (wast3d_b200/distributed.py sharded_cluster_stats).
extern "C" int wast3d_cluster_sums(int n, int K, const float* points, const int32_t* labels, double* sum3,
                                   int32_t* count, void* stream_v) {
    if (n < 0 || K < 0 || (K > 0 && (!sum3 || !count)) || (n > 0 && (!points || !labels))) return WAST3D_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return WAST3D_OK;
    cudaStream_t s = (cudaStream_t)stream_v;
    cluster_sum_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, K, points, labels, sum3, count);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
extern "C" int wast3d_cluster_scatter(int n, int K, const float* points, const int32_t* labels, const double* mean3,
                                      double* acc6, void* stream_v) {
    if (n < 0 || K < 0 || (K > 0 && (!mean3 || !acc6)) || (n > 0 && (!points || !labels))) return WAST3D_ERR_INVALID_ARGUMENT;
    if (n == 0 || K == 0) return WAST3D_OK;
    cudaStream_t s = (cudaStream_t)stream_v;
    cluster_cov_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, K, points, labels, mean3, acc6);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

// Lloyd's K-Means on 3-D points (replaces sklearn.cluster.KMeans(...).fit_predict as called by
// aux_save_clusters_clean.py:32-47 and train_st.py:54-70, for a given initialisation).
extern "C" int wast3d_kmeans_lloyd(int n, int K, const float* points, float* centers, int32_t* labels,
                                   int max_iter, double tol, double* out_inertia_host, int* out_n_iter_host,
                                   double* out_shift_host, void* stream_v) {
    if (n < 1 || K < 1 || K > n || !points || !centers || !labels || max_iter < 0 || !(tol >= 0.0))
        return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    Carver sizer(nullptr);
    sizer.take<double>((size_t)3 * K + 2); sizer.take<unsigned long long>(1); sizer.take<int32_t>(K);
    sizer.take<int32_t>(n); sizer.take<float>(n);
    const size_t match_bytes = wast3d_match_scratch_bytes(n, K);
    sizer.take<char>(match_bytes);
    void* chunk = nullptr;
    keep_pool_memory();
    W3D_CUDA_TRY(cudaMallocAsync(&chunk, sizer.bytes(), s));
    Carver c(chunk);
    double* acc = c.take<double>((size_t)3 * K + 2);
    double* inertia = acc + 3 * (size_t)K;
    double* shift2 = inertia + 1;
    unsigned long long* changed = c.take<unsigned long long>(1);
    int32_t* cnt = c.take<int32_t>(K);
    int32_t* new_label = c.take<int32_t>(n);
    float* dist = c.take<float>(n);
    void* match_scratch = c.take<char>(match_bytes);
    int rc = WAST3D_OK;
    int it = 0;
    double h_inertia = 0.0, h_shift = 0.0;
    do {
        if (cudaMemsetAsync(labels, 0xFF, sizeof(int32_t) * (size_t)n, s) != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
        // E-step, M-step, ... ; the loop always ends on an E-step against the FINAL centres so that labels and
        // centres are consistent (sklearn re-runs the E-step after a tolerance-based stop as well)
        for (;;) {
            uint32_t* err_dev = nullptr;
            rc = run_match<MODE_NN>(n, K, points, nullptr, centers, nullptr, new_label, dist, nullptr, nullptr,
                                    match_scratch, match_bytes, s, &err_dev);
            if (rc != WAST3D_OK) break;
            if (cudaMemsetAsync(acc, 0, sizeof(double) * ((size_t)3 * K + 2), s) != cudaSuccess ||
                cudaMemsetAsync(changed, 0, sizeof(unsigned long long), s) != cudaSuccess ||
                cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)K, s) != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
            kmeans_accumulate_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, points, new_label, dist, labels, acc, cnt,
                                                                      changed, inertia);
            count_launch();
            unsigned long long h_changed = 0;
            uint32_t h_err = 0;
            if (cudaMemcpyAsync(&h_err, err_dev, sizeof(h_err), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaMemcpyAsync(&h_changed, changed, sizeof(h_changed), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaMemcpyAsync(&h_inertia, inertia, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess || h_err) { rc = WAST3D_ERR_CUDA; break; }
            // stop: iteration budget used, labels stable (strict convergence), or the last update moved the
            // centres by no more than tol (then this E-step was the consistency pass)
            if (it >= max_iter || h_changed == 0 || (it > 0 && h_shift <= tol)) break;
            kmeans_update_kernel<<<(K + 127) / 128, 128, 0, s>>>(K, acc, cnt, centers, shift2);
            count_launch();
            if (cudaMemcpyAsync(&h_shift, shift2, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess) { rc = WAST3D_ERR_CUDA; break; }
            ++it;
        }
    } while (0);
    if (rc == WAST3D_ERR_CUDA) set_last_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
    cudaFreeAsync(chunk, s);
    if (out_inertia_host) *out_inertia_host = h_inertia;
    if (out_n_iter_host) *out_n_iter_host = it;
    if (out_shift_host) *out_shift_host = h_shift;
    return rc;
}
