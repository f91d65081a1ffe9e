// Shared device/host helpers and the opaque-buffer layouts of the rasteriser.
// The layouts replace GeometryState / ImageState / BinningState of the reference
// (cuda_rasterizer/rasterizer_impl.h:29-65, rasterizer_impl.cu:155-194); only the position of
// final_T / n_contrib at the head of the image buffer is kept.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include "../../include/wast3d_b200.h"

namespace w3d {

// One instruction, no registers, no shared memory: L2 fetches `bytes` (multiple of 16, 16-byte aligned
// address) from HBM while the warp is busy with arithmetic; the later loads find the lines in L2.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Hint for a warp's contiguous tile of `rows` rows of `row_floats` floats starting at row `first`: skipped
// when the tile is not 16-byte granular (a partial last warp with an odd row count).
__device__ __forceinline__ void l2_prefetch_rows(const float* base, size_t first, int rows, int row_floats) {
    if (base == nullptr) return;
    const float* a = base + first * (size_t)row_floats;
    const uint32_t bytes = (uint32_t)(rows * row_floats * 4);
    if (bytes >= 16 && (bytes & 15) == 0 && (((size_t)a) & 15) == 0) l2_prefetch_bulk(a, bytes);
}
// WAST3D_L2_PREFETCH=0 disables the hints (A/B measurements)
inline bool l2_prefetch_enabled() {
    static const int on = getenv("WAST3D_L2_PREFETCH") ? atoi(getenv("WAST3D_L2_PREFETCH")) : 1;
    return on != 0;
}


// Hyper-parameters cross the C ABI as float, but torch.optim.Adam derives 1-beta, beta^t and lr/(1-beta1^t)
// from the Python double the user wrote (0.999, not 0.999f = 0.99900001287...): 1.0f - 0.999f is off by
// 1.3e-5 relative from torch's (float)(1 - 0.999).  as_written() returns the double with the shortest
// decimal representation that still rounds to `f` (what `repr(numpy.float32(f))` prints), i.e. the
// literal the caller passed in every practical case.
inline double as_written(float f) {
    if (!(f == f) || f == 0.0f) return (double)f;
    char buf[40];
    for (int prec = 1; prec <= 9; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, (double)f);
        const double d = strtod(buf, nullptr);
        if ((float)d == f) return d;
    }
    return (double)f;
}
struct AdamScalars {
    float step_size, inv_bc2_sqrt, one_minus_b1, b2, one_minus_b2;
};
inline AdamScalars adam_scalars(float lr, float beta1, float beta2, int step) {
    const double b1 = as_written(beta1), b2 = as_written(beta2), l = as_written(lr);
    const double bc1 = 1.0 - pow(b1, (double)step);
    const double bc2 = 1.0 - pow(b2, (double)step);
    AdamScalars a;
    a.step_size = (float)(l / bc1);
    a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    a.one_minus_b1 = (float)(1.0 - b1);
    a.b2 = (float)b2;
    a.one_minus_b2 = (float)(1.0 - b2);
    return a;
}

// One element of torch's Adam update (scene/gaussian_model.py:154-163: no weight decay, no amsgrad):
//   m = m + (1-b1)(g - m);  v = b2 v + (1-b2) g g;  p = p - step_size * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// m and v are computed exactly as torch's kernels do.  The square root and the quotient use the SFU
// approximations (sqrt.approx, MUFU.RCP; <= 2 ulp each): IEEE sqrtf and '/' call slow-path subroutines whenever
// one lane of the warp sees a zero, subnormal or tiny operand — which is every warp here (culled Gaussians have
// g = m = v = 0, eps is 1e-15) — and made the optimizer-in-backward kernel instruction bound (ncu: 963 M warp
// instructions, 12% BRA, 17 M CALLs; profiles/r01_adam_in_backward.md).  The update term differs from torch's
// by <= 4 ulp of ITSELF, i.e. far below one ulp of p.
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float one_minus_b1, float b2,
                                             float one_minus_b2, float step_size, float inv_bc2_sqrt, float eps) {
    m = m + one_minus_b1 * (g - m);
    v = v * b2 + one_minus_b2 * g * g;
    const float denom = sqrt_approx(v) * inv_bc2_sqrt + eps;
    return p - step_size * __fdividef(m, denom);
}


constexpr int TILE_X = 16;   // cuda_rasterizer/config.h:16-17
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr uint32_t CULLED_KEY = 0xFFFFFFFFu;

#define W3D_CUDA_TRY(expr)                                                       \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) {                                                 \
            w3d::set_last_cuda_error(_e, __FILE__, __LINE__);                    \
            return WAST3D_ERR_CUDA;                                              \
        }                                                                        \
    } while (0)

// After a kernel launch: always check the launch itself; with `debug` also synchronise
// (the reference's CHECK_CUDA(A, debug), auxiliary.h:166-173).
#define W3D_AFTER_LAUNCH(stream, debug)                                          \
    do {                                                                         \
        w3d::count_launch();                                                     \
        W3D_CUDA_TRY(cudaGetLastError());                                        \
        if (debug) W3D_CUDA_TRY(cudaStreamSynchronize(stream));                  \
    } while (0)

void set_last_cuda_error(cudaError_t e, const char* file, int line);
void count_launch();

// Optional per-stage timing with CUDA events on the launching stream (wast3d_profile_*):
// a ProfScope brackets the launches of one stage; it does nothing unless its slot is enabled.
enum ProfSlot {
    PS_PREPROCESS = 0, PS_DEPTH_SORT, PS_SCAN, PS_EMIT, PS_TILE_SORT, PS_RANGES, PS_RENDER_FWD,
    PS_BWD_ZERO, PS_RENDER_BWD, PS_GAUSS_BWD, PS_KNN, PS_MATCH, PS_ADAM, PS_COUNT
};
struct ProfScope {
    int slot;
    cudaStream_t stream;
    cudaEvent_t start;
    bool on;
    ProfScope(int slot, cudaStream_t s);
    ~ProfScope();
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided chunk (or a null chunk for size queries) — plays the
// role of obtain() (rasterizer_impl.h:21-27) / required<T>() (:67-73).
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* p) : base((char*)p), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 128);
        T* p = base ? (T*)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    size_t bytes() const { return align_up(off, 128) + 128; }
};

// ---- radix sort / scan scratch sizes -------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // keys per block
constexpr int RS_RADIX = 256;
inline size_t rs_num_blocks(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }
inline size_t rs_hist_words(size_t n) { return rs_num_blocks(n) * RS_RADIX; }

size_t onesweep_workspace_words(size_t n, int passes);
size_t scan_lookback_workspace_words(size_t n);
constexpr int TILE_SORT_MAX_PASSES = 3;  // tile ids of up to 24 bits

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
inline size_t scan_num_blocks(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }
// block sums for a 2-level scan (level-1 sums are themselves scanned by one block in a loop)
inline size_t scan_scratch_words(size_t n) { return scan_num_blocks(n) + 8; }
// emit_instances_kernel<FUSED>: ticket, error word, one look-back status word per group of 2048 Gaussians
inline size_t emit_scan_workspace_words(size_t P) { return (P + 2047) / 2048 + 2; }

// Per-Gaussian render record, 48 bytes, gathered with three 16-byte cp.async per instance.
//   r0 = { mean2D.x, mean2D.y, view depth, cutoff half-extent x }
//   r1 = { conic.x, conic.y, conic.z, opacity }                (forward.cu:254)
//   r2 = { r, g, b, cutoff half-extent y }
// The cutoff extents bound the region where alpha >= 1/255 can hold (forward.cu:354-356) and
// are used only to skip work that the reference would skip pixel by pixel.
struct GeomState {
    float4* rec;             // [3P]
    uint32_t* depth_key;     // [P]  float bits of view depth, CULLED_KEY if not rendered
    uint32_t* tiles_touched; // [P]
    int* internal_radii;     // [P]  used when the caller passes radii == NULL
    uint8_t* clamped;        // [P]  bit c set <=> SH colour channel c was clamped (forward.cu:67-69)
    uint32_t* key_tmp;       // [P]  radix ping-pong
    uint32_t* order_a;       // [P]  Gaussian indices, final depth order lands here
    uint32_t* order_b;       // [P]
    uint32_t* offsets;       // [P]  exclusive scan of tiles_touched in depth order
    uint32_t* sort_ws;       // [onesweep_workspace_words(P, 4)]  depth sort workspace
    uint32_t* scan_ws;       // [scan_lookback_workspace_words(P)]
    uint32_t* totals;        // [4]  {num_rendered, num_visible, ...}
    float4* grad_rec;        // [3P] backward accumulators (see raster_backward.cu)
    uint2* rect;             // [P]  instantiated tile rectangle {x0 | y0 << 16, width | height << 16}; 0 = none
    uint2* rect_sorted;      // [P]  the same in depth order (gathered by the last depth-sort pass)
    uint32_t* count_sorted;  // [P]  tiles per Gaussian in depth order

    static GeomState carve(void* chunk, size_t P, size_t* bytes) {
        Carver c(chunk);
        GeomState g;
        g.rec = c.take<float4>(3 * P);
        g.depth_key = c.take<uint32_t>(P);
        g.tiles_touched = c.take<uint32_t>(P);
        g.internal_radii = c.take<int>(P);
        g.clamped = c.take<uint8_t>(P);
        g.key_tmp = c.take<uint32_t>(P);
        g.order_a = c.take<uint32_t>(P);
        g.order_b = c.take<uint32_t>(P);
        g.offsets = c.take<uint32_t>(P);
        g.sort_ws = c.take<uint32_t>(onesweep_workspace_words(P, 4));
        g.scan_ws = c.take<uint32_t>(scan_lookback_workspace_words(P) + scan_scratch_words(P));
        g.totals = c.take<uint32_t>(32);
        g.grad_rec = c.take<float4>(3 * P);
        g.rect = c.take<uint2>(P);
        g.rect_sorted = c.take<uint2>(P);
        g.count_sorted = c.take<uint32_t>(P);
        if (bytes) *bytes = c.bytes();
        return g;
    }
};

struct ImageState {
    float* final_T;       // [N]  offset 0 (alpha = 1 - final_T)
    uint32_t* n_contrib;  // [N]
    uint2* ranges;        // [T]
    uint32_t* tile_count; // [T]  instances per tile (counted while emitting)
    static ImageState carve(void* chunk, size_t N, size_t T, size_t* bytes) {
        Carver c(chunk);
        ImageState s;
        s.final_T = c.take<float>(N);
        s.n_contrib = c.take<uint32_t>(N);
        s.ranges = c.take<uint2>(T);
        s.tile_count = c.take<uint32_t>(T);
        if (bytes) *bytes = c.bytes();
        return s;
    }
};

struct BinningState {
    uint32_t* keys_a;  // [R] tile ids in depth order, later the sorted tile ids
    uint32_t* vals_a;  // [R] Gaussian ids; after the sort: the point list
    uint32_t* keys_b;  // [R]
    uint32_t* vals_b;  // [R]
    uint32_t* sort_ws;       // [onesweep_workspace_words(R, TILE_SORT_MAX_PASSES)]
    static BinningState carve(void* chunk, size_t R, size_t* bytes) {
        Carver c(chunk);
        BinningState b;
        b.keys_a = c.take<uint32_t>(R);
        b.vals_a = c.take<uint32_t>(R);
        b.keys_b = c.take<uint32_t>(R);
        b.vals_b = c.take<uint32_t>(R);
        b.sort_ws = c.take<uint32_t>(onesweep_workspace_words(R, TILE_SORT_MAX_PASSES));
        if (bytes) *bytes = c.bytes();
        return b;
    }
};

// number of bits needed to represent values in [0, n)
inline int bits_for(uint32_t n) {
    int b = 0;
    while (b < 32 && (n - 1) >> b) ++b;
    return n <= 1 ? 1 : b;
}

// ---- sort / scan launchers (scan_sort.cu) -------------------------------------------
// out[i] = sum_{k<i} in[perm ? perm[k] : k]; total (optional) receives the full sum.
int scan_exclusive_u32(const uint32_t* in, const uint32_t* perm, uint32_t* out, size_t n,
                       uint32_t* scratch, uint32_t* total, cudaStream_t s, bool debug);
// Optional rider of a sort pass: with the sorted value v landing at position g, also dst[g] = src[v] and
// cnt[g] = width * height of that rectangle (the rasteriser's last depth pass: the tile rectangles in depth order,
// so that the offsets scan and the instance emission read them coalesced instead of gathering per Gaussian again).
struct GatherRect {
    const uint2* src = nullptr;
    uint2* dst = nullptr;
    uint32_t* cnt = nullptr;
};
// One stable LSD pass on bits [shift, shift+bits) (bits <= 8).  vals_in == NULL means iota.
// keys_out == NULL means "do not write keys" (last pass).
// n_dev (optional, device): process min(*n_dev, n) elements; n is then the capacity the launch is sized for.
int radix_pass_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, size_t n, int shift, int bits, uint32_t* hist,
                   uint32_t* scan_scratch, cudaStream_t s, bool debug, const uint32_t* n_dev = nullptr,
                   GatherRect gather = GatherRect());

// Single-pass (decoupled look-back) variants — see scan_sort.cu.  The workspace `ws` of
// onesweep_workspace_words(n, passes) words is zeroed by onesweep_prepare() once per sort;
// onesweep_hist() fills the per-pass global digit histograms from the keys (or the caller fills
// onesweep_digit_hist(...) itself), then onesweep_pass() is called once per pass.
size_t onesweep_workspace_words(size_t n, int passes);
int onesweep_prepare(uint32_t* ws, size_t n, int passes, cudaStream_t s);
int onesweep_hist(const uint32_t* keys, size_t n, int passes, const int* shifts, const int* bits,
                  uint32_t* ws, cudaStream_t s, bool debug);
int onesweep_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                  size_t n, int shift, int bits, uint32_t* ws, int passes, int pass, cudaStream_t s, bool debug,
                  const uint32_t* n_dev = nullptr, GatherRect gather = GatherRect());
uint32_t* onesweep_digit_hist(uint32_t* ws, size_t n, int passes, int pass);
int onesweep_scan_digits(uint32_t* ws, size_t n, int passes, cudaStream_t s, bool debug);
uint32_t* onesweep_error_word(uint32_t* ws, size_t n, int passes);
size_t scan_lookback_workspace_words(size_t n);
int scan_exclusive_lookback_u32(const uint32_t* in, const uint32_t* perm, uint32_t* out, size_t n,
                                uint32_t* ws, uint32_t* total, cudaStream_t s, bool debug);

// ---- small device helpers -------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
// ---- mbarrier helpers (shared::cta) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbarrier_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_addr_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// The executing thread's arrival on `bar` is triggered when all of ITS prior cp.async copies have
// landed; .noinc: the arrival was accounted for in the barrier's initial count.
__device__ __forceinline__ void cp_async_mbarrier_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- block scan and decoupled look-back (scan_sort.cu; the fused offsets scan of emit_instances_kernel) ----------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across a block of SCAN_THREADS; returns the
// exclusive prefix and writes the block total to *total (valid for all threads).
template <int THREADS = SCAN_THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[THREADS / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < THREADS / 32 ? warp_sums[lane] : 0;
        uint32_t wi = warp_incl_scan(w, lane);
        if (lane < THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == THREADS / 32 - 1) block_total = wi;
    }
    __syncthreads();
    uint32_t r = incl - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return r;
}

// Look-back status words carry a flag in bits [31:30] (1 = aggregate of this tile only, 2 = inclusive prefix up to this
// tile) and a 30-bit count; every wait is bounded and raises an error flag instead of hanging the device.
constexpr uint32_t OS_AGG = 1u << 30, OS_PREFIX = 2u << 30, OS_VALUE = (1u << 30) - 1u;
constexpr uint32_t OS_SPIN_LIMIT = 1u << 20;  // ~1 s of polling; a healthy wait is a few dozen polls

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Single-kernel exclusive scan with look-back.  ws: 2 + nblocks zero-initialised words
// (ticket, error, status[nblocks]).  The look-back is done by warp 0, 32 predecessors per step.
__device__ __forceinline__ uint32_t lookback_warp(uint32_t* status, uint32_t tile, uint32_t count,
                                                  uint32_t* err, int lane) {
    if (tile == 0) {
        if (lane == 0) st_status(status, count | OS_PREFIX);
        return 0;
    }
    if (lane == 0) st_status(status + tile, count | OS_AGG);
    uint32_t excl = 0, spins = 0;
    int look = (int)tile;  // exclusive upper end of the window
    while (true) {
        const int idx = look - 1 - lane;
        uint32_t v = idx >= 0 ? ld_status(status + idx) : OS_PREFIX;  // before tile 0: empty prefix
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (v >> 30) == 2u);
        const unsigned invalid = __ballot_sync(0xffffffffu, (v >> 30) == 0u);
        const int first = is_prefix ? __ffs(is_prefix) - 1 : 32;      // nearest predecessor with a prefix
        const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
        if (invalid & need) {
            if (++spins > OS_SPIN_LIMIT) {
                if (lane == 0) atomicOr(err, 1u);
                break;
            }
            continue;
        }
        uint32_t c = lane <= first ? (v & OS_VALUE) : 0u;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (first < 32) break;
        look -= 32;
    }
    if (lane == 0) st_status(status + tile, (excl + count) | OS_PREFIX);
    return excl;
}


}  // namespace w3d
