// View-parallel optimizer step as ONE kernel over NVLink peer memory (SURVEY.md §8e "rasteriser
// step": replicated parameters, per-view gradients summed across GPUs, identical Adam update).
//
// The reference is single-GPU (scene/gaussian_model.py:149-167: torch.optim.Adam, six groups).  The
// textbook multi-GPU form is "NCCL all-reduce of the gradients, then the full Adam update on every
// rank" (wast3d_b200/distributed.py::allreduce_and_step keeps that as the baseline).  Here every rank
// owns a contiguous 1/N shard of the flat parameter arena and does, in one launch:
//
//   wait until every peer's gradients are complete            (flag exchange over NVLink)
//   for each float4 of MY shard:
//       g  = sum over ranks q (fixed order 0..N-1) of grad_q[i]     N-1 peer loads + 1 local
//       Adam update of p[i] with the shard-local moments m, v      (HBM traffic 1/N of the dense step)
//       store the new p[i] into the replica of EVERY rank           N-1 peer stores + 1 local
//   signal "my stores have landed" to every peer, wait for theirs
//
// i.e. reduce-scatter + Adam + all-gather with the same NVLink bytes as an all-reduce, but the
// optimizer's HBM traffic and state shrink by N and nothing is staged through a communication
// buffer.  Every replica receives the bit-identical value because exactly one rank computes it.
// With world == 1 this is a single-launch multi-tensor Adam.
//
// Arithmetic per element is adam.cu's (torch single-tensor Adam order).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace w3d {

constexpr int PEER_MAX_WORLD = WAST3D_PEER_MAX_WORLD;
constexpr int PEER_MAX_SEGS = WAST3D_PEER_MAX_SEGMENTS;
constexpr int PEER_THREADS = 512;

struct PeerSeg {  // device-side copy of wast3d_adam_segment with derived constants
    unsigned long long begin4, end4;
    float step_size, inv_bc2_sqrt, one_minus_b1, b2, one_minus_b2, eps;
};

struct PeerArgs {
    const float4* grads[PEER_MAX_WORLD];
    float4* params[PEER_MAX_WORLD];
    uint32_t* flags[PEER_MAX_WORLD];  // per rank: [2][PEER_MAX_WORLD] words (arrive, done), written by peers
    const float4* mc_grads;           // NVLS multicast mapping of the gradient / parameter regions of ALL
    float4* mc_params;                // ranks (NULL: use the per-rank pointers above)
    float4* m;                        // moments of MY shard, index (i - shard_begin4)
    float4* v;
    unsigned long long shard_begin4, shard_end4;
    PeerSeg segs[PEER_MAX_SEGS];
    int nsegs, world, rank;
    float grad_scale;
    uint32_t epoch;
    unsigned* block_counter;          // local, zero between launches
    int* error_word;                  // pinned host memory: set on a flag time-out
    unsigned long long timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// streaming 16-byte accesses: nothing here is re-read by this kernel
// (volatile asm: the loads of one tile are all issued, in program order, before the first use)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) { __stcs(p, v); }
// NVLS (NVSwitch multicast): one load returns the switch-side sum of the addressed float4 over every
// GPU bound to the multicast object; one store lands in every GPU's replica.
__device__ __forceinline__ float4 mc_ld_reduce_add(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Wait until every peer wrote `epoch` (or a later one) into my flag row.  Returns false on time-out.
__device__ bool wait_flags(const uint32_t* row, int world, uint32_t epoch, unsigned long long timeout_ns,
                           int* error_word) {
    const unsigned long long t0 = global_timer_ns();
    for (int q = 0; q < world; ++q) {
        unsigned spins = 0;
        while ((int32_t)(ld_acquire_sys(row + q) - epoch) < 0) {
            if ((++spins & 0x3FFu) == 0) {
                if (*(volatile int*)error_word) return false;
                if (global_timer_ns() - t0 > timeout_ns) {
                    *(volatile int*)error_word = 1 + q;
                    __threadfence_system();
                    return false;
                }
            }
        }
    }
    return true;
}

__device__ __forceinline__ void adam4(float4& p, const float4& g, float4& m, float4& v, const PeerSeg& s) {
#define W3D_ADAM1(c)                                                      \
    p.c = adam_update(p.c, g.c, m.c, v.c, s.one_minus_b1, s.b2, s.one_minus_b2, s.step_size, s.inv_bc2_sqrt, s.eps);
    W3D_ADAM1(x) W3D_ADAM1(y) W3D_ADAM1(z) W3D_ADAM1(w)
#undef W3D_ADAM1
}

// MC = true: gradients are summed by the switch (multimem.ld_reduce) and parameters broadcast by it
// (multimem.st): per GPU and direction ~1x the arena crosses NVLink instead of 2 (N-1)/N x with the
// per-peer loads and stores of MC = false.
template <int WORLD, bool MC>
__global__ void __launch_bounds__(PEER_THREADS)
peer_adam_kernel(const __grid_constant__ PeerArgs a) {
    __shared__ int ok_s;
    const int rank = a.rank;
    if (WORLD > 1) {
        // ---- phase 0: tell every peer my gradients are complete (the kernels that wrote them ran
        // earlier on this stream), then wait for theirs
        if (blockIdx.x == 0 && threadIdx.x < WORLD) {
            __threadfence_system();
            st_release_sys(a.flags[threadIdx.x] + rank, a.epoch);
        }
        if (threadIdx.x == 0) ok_s = wait_flags(a.flags[rank], WORLD, a.epoch, a.timeout_ns, a.error_word) ? 1 : 0;
        __syncthreads();
        if (!ok_s) return;
    }

    // ---- phase 1: reduce + Adam + broadcast over my shard, one segment (= parameter group) at a time.
    // NVLink latency (~2 us) x bandwidth (~0.9 TB/s per direction) needs ~2 MB of peer loads in flight per
    // GPU: every thread issues all loads of U float4 columns (U * (WORLD + 3) 16-byte loads) before it
    // touches any of them.
    // Multicast: the switch-side reduction has the longest latency and the late launch runs on a dozen CTAs only (the
    // next view's projection and sorting own the other SMs), so 8 reduced columns are in flight per thread; the local
    // parameter / moment loads follow in groups of UL columns to stay inside the register budget.
    constexpr int U = MC ? 8 : (WORLD <= 2 ? 4 : (WORLD <= 5 ? 2 : 1));
    constexpr int UL = U < 4 ? U : 4;
    constexpr int NG = MC ? 1 : WORLD;  // gradient loads per column
    const unsigned long long tile = (unsigned long long)PEER_THREADS * U;
    for (int sidx = 0; sidx < a.nsegs; ++sidx) {
        const PeerSeg seg = a.segs[sidx];
        const unsigned long long lo = seg.begin4 > a.shard_begin4 ? seg.begin4 : a.shard_begin4;
        const unsigned long long hi = seg.end4 < a.shard_end4 ? seg.end4 : a.shard_end4;
        for (unsigned long long base = lo + (unsigned long long)blockIdx.x * tile; base < hi;
             base += (unsigned long long)gridDim.x * tile) {
            float4 g[U][NG];
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned long long i = base + (unsigned long long)u * PEER_THREADS + threadIdx.x;
                const bool in = i < hi;
                if (MC) {
                    g[u][0] = in ? mc_ld_reduce_add(a.mc_grads + i) : z;
                } else {
#pragma unroll
                    for (int q = 0; q < NG; ++q) g[u][q] = in ? ld_stream(a.grads[q] + i) : z;
                }
            }
#pragma unroll
            for (int u0 = 0; u0 < U; u0 += UL) {
                float4 p[UL], m[UL], v[UL];
#pragma unroll
                for (int k = 0; k < UL; ++k) {
                    const unsigned long long i = base + (unsigned long long)(u0 + k) * PEER_THREADS + threadIdx.x;
                    const bool in = i < hi;
                    const unsigned long long j = i - a.shard_begin4;
                    p[k] = in ? ld_stream(a.params[rank] + i) : z;
                    m[k] = in ? ld_stream(a.m + j) : z;
                    v[k] = in ? ld_stream(a.v + j) : z;
                }
#pragma unroll
                for (int k = 0; k < UL; ++k) {
                    const int u = u0 + k;
                    const unsigned long long i = base + (unsigned long long)u * PEER_THREADS + threadIdx.x;
                    if (i >= hi) continue;
                    const unsigned long long j = i - a.shard_begin4;
                    float4 gs = g[u][0];
#pragma unroll
                    for (int q = 1; q < NG; ++q) {
                        gs.x += g[u][q].x; gs.y += g[u][q].y; gs.z += g[u][q].z; gs.w += g[u][q].w;
                    }
                    if (WORLD > 1) {
                        gs.x *= a.grad_scale; gs.y *= a.grad_scale; gs.z *= a.grad_scale; gs.w *= a.grad_scale;
                    }
                    adam4(p[k], gs, m[k], v[k], seg);
                    st_stream(a.m + j, m[k]);
                    st_stream(a.v + j, v[k]);
                    if (MC) {
                        mc_st(a.mc_params + i, p[k]);
                    } else {
#pragma unroll
                        for (int q = 0; q < WORLD; ++q) st_stream(a.params[q] + i, p[k]);
                    }
                }
            }
        }
    }

    if (WORLD > 1) {
        // ---- phase 2: my stores are out; the last block tells the peers and waits until every peer's
        // stores into MY replica have landed, so the next kernel on this stream sees them
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned done = atomicAdd(a.block_counter, 1u);
            __threadfence_system();
            ok_s = (done == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (!ok_s) return;
        if (threadIdx.x < WORLD) {
            __threadfence_system();
            st_release_sys(a.flags[threadIdx.x] + PEER_MAX_WORLD + rank, a.epoch);
        }
        if (threadIdx.x == 0) {
            *a.block_counter = 0;
            wait_flags(a.flags[rank] + PEER_MAX_WORLD, WORLD, a.epoch, a.timeout_ns, a.error_word);
        }
    }
}

// pinned, device-visible error word shared by all launches of this process
static int* peer_error_word() {
    static int* w = nullptr;
    if (!w) {
        if (cudaHostAlloc((void**)&w, sizeof(int), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) return nullptr;
        *w = 0;
    }
    return w;
}

}  // namespace w3d

using namespace w3d;

extern "C" size_t wast3d_peer_flag_bytes(void) { return 2 * PEER_MAX_WORLD * sizeof(uint32_t) + 64; }

extern "C" int wast3d_peer_adam_step(int world, int rank, void* const* grad_ptrs, void* const* param_ptrs,
                                     void* const* flag_ptrs, void* mc_grads, void* mc_params,
                                     float* exp_avg, float* exp_avg_sq,
                                     size_t shard_begin4, size_t shard_end4,
                                     const wast3d_adam_segment* segs, int nsegs, float grad_scale,
                                     unsigned epoch, double timeout_s, int max_ctas_arg, void* stream_v) {
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || nsegs < 0 || nsegs > PEER_MAX_SEGS ||
        !grad_ptrs || !param_ptrs || shard_end4 < shard_begin4)
        return WAST3D_ERR_INVALID_ARGUMENT;
    if (world > 1 && (!flag_ptrs || epoch == 0)) return WAST3D_ERR_INVALID_ARGUMENT;
    if (shard_end4 > shard_begin4 && (!exp_avg || !exp_avg_sq)) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    int* err = peer_error_word();
    if (!err) return WAST3D_ERR_CUDA;
    if (*(volatile int*)err) {  // an earlier launch timed out waiting for a peer: the replicas diverged
        set_last_cuda_error(cudaErrorLaunchTimeout, __FILE__, __LINE__);
        return WAST3D_ERR_CUDA;
    }
    PeerArgs a{};
    for (int q = 0; q < world; ++q) {
        if (!grad_ptrs[q] || !param_ptrs[q] || (world > 1 && !flag_ptrs[q])) return WAST3D_ERR_INVALID_ARGUMENT;
        if (((size_t)grad_ptrs[q] | (size_t)param_ptrs[q]) & 15) return WAST3D_ERR_INVALID_ARGUMENT;
        a.grads[q] = (const float4*)grad_ptrs[q];
        a.params[q] = (float4*)param_ptrs[q];
        a.flags[q] = world > 1 ? (uint32_t*)flag_ptrs[q] : nullptr;
    }
    if (((size_t)exp_avg | (size_t)exp_avg_sq | (size_t)mc_grads | (size_t)mc_params) & 15)
        return WAST3D_ERR_INVALID_ARGUMENT;
    const bool mc = world > 1 && mc_grads != nullptr && mc_params != nullptr;
    a.mc_grads = mc ? (const float4*)mc_grads : nullptr;
    a.mc_params = mc ? (float4*)mc_params : nullptr;
    a.m = (float4*)exp_avg;
    a.v = (float4*)exp_avg_sq;
    a.shard_begin4 = shard_begin4;
    a.shard_end4 = shard_end4;
    unsigned long long prev_end = 0;
    for (int k = 0; k < nsegs; ++k) {
        const wast3d_adam_segment& h = segs[k];
        if (h.end4 < h.begin4 || h.begin4 < prev_end || h.step < 1) return WAST3D_ERR_INVALID_ARGUMENT;
        prev_end = h.end4;
        const w3d::AdamScalars sc = w3d::adam_scalars(h.lr, h.beta1, h.beta2, h.step);
        PeerSeg& d = a.segs[k];
        d.begin4 = h.begin4;
        d.end4 = h.end4;
        d.step_size = sc.step_size;
        d.inv_bc2_sqrt = sc.inv_bc2_sqrt;
        d.one_minus_b1 = sc.one_minus_b1;
        d.b2 = sc.b2;
        d.one_minus_b2 = sc.one_minus_b2;
        d.eps = h.eps;
    }
    a.nsegs = nsegs;
    a.world = world;
    a.rank = rank;
    a.grad_scale = grad_scale;
    a.epoch = epoch;
    a.error_word = err;
    a.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 20.0) * 1e9);
    // the block counter lives behind my own flag rows (wast3d_peer_flag_bytes reserves it)
    a.block_counter = world > 1 ? (unsigned*)((uint32_t*)flag_ptrs[rank] + 2 * PEER_MAX_WORLD) : nullptr;

    const size_t n4 = shard_end4 - shard_begin4;
    size_t blocks = (n4 + PEER_THREADS - 1) / PEER_THREADS;
    if (blocks < 1) blocks = 1;
    ProfScope ps(PS_ADAM, s);
    switch (world) {
        // persistent grid: as many CTAs as are resident at once (register-limited, differs per WORLD)
#define W3D_PEER_LAUNCH(N, MCFLAG)                                                                        \
    {                                                                                                     \
        static int per_sm = 0;                                                                            \
        if (!per_sm) {                                                                                    \
            W3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                   \
                &per_sm, peer_adam_kernel<N, MCFLAG>, PEER_THREADS, 0));                                  \
            if (per_sm < 1) per_sm = 1;                                                                   \
        }                                                                                                 \
        {                                                                                                 \
            /* WAST3D_PEER_CTAS_PER_SM: cap the persistent grid so kernels of another stream can co-run */  \
            static const int cap = getenv("WAST3D_PEER_CTAS_PER_SM") ? atoi(getenv("WAST3D_PEER_CTAS_PER_SM")) : 0; \
            const int max_ctas = max_ctas_arg;                                                            \
            const int use = (cap > 0 && cap < per_sm) ? cap : per_sm;                                     \
            if (blocks > (size_t)148 * use) blocks = (size_t)148 * use;                                   \
            if (max_ctas > 0 && blocks > (size_t)max_ctas) blocks = (size_t)max_ctas;                     \
        }                                                                                                 \
        peer_adam_kernel<N, MCFLAG><<<(unsigned)blocks, PEER_THREADS, 0, s>>>(a);                         \
    }
#define W3D_PEER_CASE(N)                                                                                  \
    case N:                                                                                               \
        if (mc) W3D_PEER_LAUNCH(N, true) else W3D_PEER_LAUNCH(N, false)                                   \
        break;
        W3D_PEER_CASE(1) W3D_PEER_CASE(2) W3D_PEER_CASE(3) W3D_PEER_CASE(4)
        W3D_PEER_CASE(5) W3D_PEER_CASE(6) W3D_PEER_CASE(7) W3D_PEER_CASE(8)
#undef W3D_PEER_CASE
#undef W3D_PEER_LAUNCH
        default: return WAST3D_ERR_INVALID_ARGUMENT;
    }
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" int wast3d_peer_error(int reset) {
    int* err = peer_error_word();
    if (!err) return -1;
    const int v = *(volatile int*)err;
    if (reset) *(volatile int*)err = 0;
    return v;
}

// ---- peer-visible memory: cudaMalloc + CUDA IPC handles (one process per GPU) --------------------
extern "C" int wast3d_peer_alloc(size_t bytes, void** out_ptr) {
    if (!out_ptr || bytes == 0) return WAST3D_ERR_INVALID_ARGUMENT;
    W3D_CUDA_TRY(cudaMalloc(out_ptr, bytes));
    W3D_CUDA_TRY(cudaMemset(*out_ptr, 0, bytes));
    return WAST3D_OK;
}
extern "C" int wast3d_peer_export(void* ptr, unsigned char* handle64) {
    if (!ptr || !handle64) return WAST3D_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == WAST3D_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    W3D_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, sizeof(h));
    return WAST3D_OK;
}
extern "C" int wast3d_peer_import(const unsigned char* handle64, void** out_ptr) {
    if (!handle64 || !out_ptr) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    W3D_CUDA_TRY(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return WAST3D_OK;
}
extern "C" int wast3d_peer_release(void* ptr, int imported) {
    if (!ptr) return WAST3D_OK;
    if (imported) W3D_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    else W3D_CUDA_TRY(cudaFree(ptr));
    return WAST3D_OK;
}
