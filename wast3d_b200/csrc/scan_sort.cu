// Device-wide exclusive scan and one stable LSD radix-sort pass over (u32 key, u32 value) pairs.
//
// Replaces the three CUB calls of the reference binning stage:
//   cub::DeviceScan::InclusiveSum        rasterizer_impl.cu:279
//   cub::DeviceRadixSort::SortPairs      rasterizer_impl.cu:305-310   (u64 key, 6 passes)
// Here the (tile, depth) order is produced as "depth sort of P Gaussians (4 passes over P
// pairs), then a stable partition of the R instances by tile id (ceil(tile_bits/8) passes
// over R pairs)" — see raster_forward.cu — so this file only needs 32-bit keys.
//
// Two implementations of each primitive: multi-kernel (hist / 3-kernel table scan / scatter;
// no inter-block communication at all) and single-pass decoupled look-back ("onesweep", one
// kernel per pass, used by the rasteriser) — see the second half of this file.
#include "common.cuh"

namespace w3d {

// ------------------------------------------------------------------ scan ----------------
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

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm, size_t n,
                   uint32_t* __restrict__ block_sums) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += perm ? in[perm[i]] : in[i];
    }
    uint32_t total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One block scans the block sums in place (exclusive) and emits the grand total.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_spine_kernel(uint32_t* __restrict__ block_sums, size_t nblocks, uint32_t* __restrict__ total_out) {
    uint32_t carry = 0;
    for (size_t base = 0; base < nblocks; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0;
        uint32_t tot;
        uint32_t ex = block_excl_scan(v, &tot);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm,
                  uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ block_sums) {
    // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + k;
        v[k] = i < n ? (perm ? in[perm[i]] : in[i]) : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

int scan_exclusive_u32(const uint32_t* in, const uint32_t* perm, uint32_t* out, size_t n,
                       uint32_t* scratch, uint32_t* total, cudaStream_t s, bool debug) {
    if (n == 0) {
        if (total) W3D_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(uint32_t), s));
        return WAST3D_OK;
    }
    const size_t nb = scan_num_blocks(n);
    scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, perm, n, scratch);
    W3D_AFTER_LAUNCH(s, debug);
    scan_spine_kernel<<<1, SCAN_THREADS, 0, s>>>(scratch, nb, total);
    W3D_AFTER_LAUNCH(s, debug);
    scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, perm, out, n, scratch);
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

// ------------------------------------------------------------------ radix pass ----------
// Item order inside a block tile is "warp-striped": warp w owns keys
// [w*32*ITEMS, (w+1)*32*ITEMS) of the tile and its j-th load covers 32 consecutive keys, so
// loads are coalesced and the memory order is (w, j, lane) — ranking in that order is stable.

// n_dev (optional): the element count lives on the device (graph-safe forward: the instance count is never read by
// the host); `n` is then the capacity the grid was sized for and the kernel processes min(*n_dev, n) elements.
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t mask,
                  uint32_t* __restrict__ hist, unsigned nblocks, const uint32_t* __restrict__ n_dev) {
    __shared__ uint32_t h[RS_RADIX];
    if (n_dev != nullptr) n = min(n, (size_t)*n_dev);
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    // bin-major table so that one exclusive scan yields global scatter bases
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

constexpr int OS_THREADS = 512;
template <bool HAS_VALS, bool WRITE_KEYS, bool LOOKBACK>
__global__ void onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                                     int shift, uint32_t mask, const uint32_t* __restrict__ digit_hist,
                                     uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                                     uint32_t* __restrict__ err, unsigned nblocks, const uint32_t* __restrict__ n_dev);

int radix_pass_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, size_t n, int shift, int bits, uint32_t* hist,
                   uint32_t* scan_scratch, cudaStream_t s, bool debug, const uint32_t* n_dev) {
    if (n == 0) return WAST3D_OK;
    if (bits < 1 || bits > 8) return WAST3D_ERR_INVALID_ARGUMENT;
    if (n > 0xFFFFFFFFull - RS_TILE) return WAST3D_ERR_OVERFLOW;
    const unsigned nb = (unsigned)rs_num_blocks(n);
    const uint32_t mask = (1u << bits) - 1u;
    radix_hist_kernel<<<nb, RS_THREADS, 0, s>>>(keys_in, n, shift, mask, hist, nb, n_dev);
    W3D_AFTER_LAUNCH(s, debug);
    int st = scan_exclusive_u32(hist, nullptr, hist, (size_t)nb * RS_RADIX, scan_scratch, nullptr,
                                s, debug);
    if (st != WAST3D_OK) return st;
#define W3D_SCATTER(HV, WK)                                                                          \
    onesweep_pass_kernel<HV, WK, false><<<nb, OS_THREADS, 0, s>>>(keys_in, vals_in, keys_out, vals_out, n, \
                                                                  shift, mask, hist, nullptr, nullptr, nullptr, nb, n_dev)
    if (vals_in) {
        if (keys_out) W3D_SCATTER(true, true); else W3D_SCATTER(true, false);
    } else {
        if (keys_out) W3D_SCATTER(false, true); else W3D_SCATTER(false, false);
    }
#undef W3D_SCATTER
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

// ------------------------------------------------------------------ single-pass variants --
// Decoupled look-back ("onesweep"): ONE kernel per radix pass and ONE kernel per scan instead of
// hist + 3-kernel table scan + scatter.  A block takes its tile index from an atomic ticket, so
// a tile only ever waits on tiles whose blocks are already running or finished (deadlock-free
// without any assumption on block scheduling); every wait is bounded anyway and raises an error
// flag instead of hanging the device.  Status words carry the flag in bits [31:30]
// (1 = aggregate of this tile only, 2 = inclusive prefix up to this tile) and a 30-bit count.
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
// Global digit histograms of up to 4 passes from one read of the keys.
__global__ void __launch_bounds__(256)
onesweep_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int npasses, int4 shifts, int4 masks,
                     uint32_t* __restrict__ digit_hist /*[npasses][RS_RADIX]*/) {
    __shared__ uint32_t h[4][RS_RADIX];
    for (int i = threadIdx.x; i < 4 * RS_RADIX; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const int sh[4] = {shifts.x, shifts.y, shifts.z, shifts.w};
    const int mk[4] = {masks.x, masks.y, masks.z, masks.w};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const uint32_t k = keys[i];
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (p < npasses) atomicAdd(&h[p][(k >> sh[p]) & mk[p]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npasses * RS_RADIX; i += 256) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(digit_hist + i, c);
    }
}

// 512 threads x 8 keys: the per-warp ranking chain (match -> shared counter -> next key) is a
// latency chain, so the tile is spread over 16 warps with 8 dependent steps each instead of
// 8 warps with 16, and 4 blocks (64 warps) fit an SM.  The look-back wait is placed after the
// keys have been ranked and parked in shared memory, when predecessors have had time to publish.
constexpr int OS_ITEMS = 8;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
static_assert(OS_TILE == RS_TILE, "workspace sizing assumes the same tile size");

// LOOKBACK = false: the multi-kernel flavour — `digit_hist` is then the scanned bin-major table
// of per-tile scatter bases (bases[digit * nblocks + tile]) and status/ticket/err are unused.
template <bool HAS_VALS, bool WRITE_KEYS, bool LOOKBACK>
__global__ void __launch_bounds__(OS_THREADS, 3)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                     int shift, uint32_t mask, const uint32_t* __restrict__ digit_hist,
                     uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                     uint32_t* __restrict__ err, unsigned nblocks, const uint32_t* __restrict__ n_dev) {
    constexpr int WARPS = OS_THREADS / 32;
    if (!LOOKBACK && n_dev != nullptr) {   // device-side element count (see radix_hist_kernel)
        n = min(n, (size_t)*n_dev);
        if ((size_t)blockIdx.x * OS_TILE >= n) return;   // whole block past the end (uniform)
    }
    __shared__ uint16_t warp_hist[WARPS][RS_RADIX];  // <= OS_TILE, fits 16 bits
    __shared__ uint32_t bin_start[RS_RADIX];
    __shared__ uint32_t bin_base[RS_RADIX];
    __shared__ uint32_t skeys[OS_TILE];
    __shared__ uint32_t svals[OS_TILE];
    __shared__ uint32_t s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    if (LOOKBACK && threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < WARPS * RS_RADIX / 2; i += OS_THREADS)
        reinterpret_cast<uint32_t*>(&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = LOOKBACK ? s_tile : blockIdx.x;

    const size_t tile_base = (size_t)tile * OS_TILE;
    const size_t warp_base = tile_base + (size_t)warp * (32 * OS_ITEMS);
    uint32_t key[OS_ITEMS], rank[OS_ITEMS];
    unsigned peers[OS_ITEMS];
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        key[j] = i < n ? keys_in[i] : 0xFFFFFFFFu;
    }
    // Out-of-range items only exist at the very end of the last tile; give them the highest
    // digit so they rank after every real key of that digit and are simply not written.
    // All match operations are issued before the dependent shared-memory chain starts.
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        uint32_t d = i < n ? ((key[j] >> shift) & mask) : (RS_RADIX - 1);
        peers[j] = __match_any_sync(0xffffffffu, d);
    }
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        uint32_t d = i < n ? ((key[j] >> shift) & mask) : (RS_RADIX - 1);
        uint32_t before = warp_hist[warp][d];
        __syncwarp();
        if ((peers[j] & lt) == 0) warp_hist[warp][d] = (uint16_t)(before + __popc(peers[j]));
        __syncwarp();
        rank[j] = before + __popc(peers[j] & lt);
    }
    __syncthreads();
    // threads 0..255: thread t owns digit t
    const bool owner = threadIdx.x < RS_RADIX;
    uint32_t run = 0;
    if (owner) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t c = warp_hist[w][threadIdx.x];
            warp_hist[w][threadIdx.x] = (uint16_t)run;
            run += c;
        }
    }
    // Padding of the last tile was counted under digit RS_RADIX-1; it must not be published
    // (it ranks after every real key of that digit, so local positions stay right).
    const size_t remaining = n - tile_base;
    const uint32_t count_tile = remaining < (size_t)OS_TILE ? (uint32_t)remaining : (uint32_t)OS_TILE;
    uint32_t real = run;
    if (threadIdx.x == RS_RADIX - 1) real -= (uint32_t)OS_TILE - count_tile;
    uint32_t* my_status = status + (size_t)tile * RS_RADIX + threadIdx.x;
    if (LOOKBACK && owner) st_status(my_status, real | (tile == 0 ? OS_PREFIX : OS_AGG));
    uint32_t tot;
    const uint32_t start = block_excl_scan<OS_THREADS>(run, &tot);
    uint32_t dstart = 0;
    if (LOOKBACK) dstart = block_excl_scan<OS_THREADS>(owner ? digit_hist[threadIdx.x] : 0u, &tot);
    else if (owner) dstart = digit_hist[(size_t)threadIdx.x * nblocks + tile];
    if (owner) bin_start[threadIdx.x] = start;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        uint32_t d = i < n ? ((key[j] >> shift) & mask) : (RS_RADIX - 1);
        uint32_t pos = bin_start[d] + warp_hist[warp][d] + rank[j];
        skeys[pos] = key[j];
        // values are only loaded now, so they are not live across the ranking phase
        svals[pos] = i < n ? (HAS_VALS ? __ldg(vals_in + i) : (uint32_t)i) : 0u;
    }
    // look-back over earlier tiles, one digit column per thread
    if (owner) {
        uint32_t excl = 0;
        if (LOOKBACK && tile != 0) {
            uint32_t spins = 0;
            bool failed = false;
            for (uint32_t look = tile; look-- > 0 && !failed;) {
                uint32_t v;
                while (((v = ld_status(status + (size_t)look * RS_RADIX + threadIdx.x)) >> 30) == 0) {
                    if (++spins > OS_SPIN_LIMIT) { failed = true; break; }
                }
                if (failed) break;
                excl += v & OS_VALUE;
                if ((v >> 30) == 2u) break;
            }
            if (failed) atomicOr(err, 1u);
            st_status(my_status, (excl + real) | OS_PREFIX);
        }
        bin_base[threadIdx.x] = dstart + excl;
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < count_tile; p += OS_THREADS) {
        uint32_t k = skeys[p];
        uint32_t d = (k >> shift) & mask;
        size_t g = (size_t)bin_base[d] + (p - bin_start[d]);
        if (WRITE_KEYS) keys_out[g] = k;
        vals_out[g] = svals[p];
    }
}

size_t onesweep_workspace_words(size_t n, int passes) {
    return (size_t)passes * (rs_num_blocks(n) * RS_RADIX + RS_RADIX + 32) + 32;
}

// Layout of the zero-initialised workspace: [passes][RS_RADIX] digit histograms, then per pass
// {ticket (32 words, first used), status [nblocks][RS_RADIX]}, then the error word.
struct OnesweepWs {
    uint32_t *digit_hist, *err;
    uint32_t* ws;
    size_t nb;
    int passes;
    OnesweepWs(uint32_t* w, size_t n, int p) : ws(w), nb(rs_num_blocks(n)), passes(p) {
        digit_hist = ws;
        err = ws + onesweep_workspace_words(n, p) - 32;
    }
    uint32_t* ticket(int pass) const { return ws + (size_t)passes * RS_RADIX + (size_t)pass * (nb * RS_RADIX + 32); }
    uint32_t* status(int pass) const { return ticket(pass) + 32; }
};

int onesweep_prepare(uint32_t* ws, size_t n, int passes, cudaStream_t s) {
    W3D_CUDA_TRY(cudaMemsetAsync(ws, 0, onesweep_workspace_words(n, passes) * sizeof(uint32_t), s));
    return WAST3D_OK;
}

int onesweep_hist(const uint32_t* keys, size_t n, int passes, const int* shifts, const int* bits,
                  uint32_t* ws, cudaStream_t s, bool debug) {
    if (n == 0) return WAST3D_OK;
    if (passes < 1 || passes > 4) return WAST3D_ERR_INVALID_ARGUMENT;
    int sh[4] = {0, 0, 0, 0}, mk[4] = {0, 0, 0, 0};
    for (int p = 0; p < passes; ++p) {
        if (bits[p] < 1 || bits[p] > 8) return WAST3D_ERR_INVALID_ARGUMENT;
        sh[p] = shifts[p];
        mk[p] = (1 << bits[p]) - 1;
    }
    size_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    OnesweepWs w(ws, n, passes);
    onesweep_hist_kernel<<<(unsigned)blocks, 256, 0, s>>>(keys, n, passes, make_int4(sh[0], sh[1], sh[2], sh[3]),
                                                          make_int4(mk[0], mk[1], mk[2], mk[3]), w.digit_hist);
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

int onesweep_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                  size_t n, int shift, int bits, uint32_t* ws, int passes, int pass, cudaStream_t s, bool debug) {
    if (n == 0) return WAST3D_OK;
    if (bits < 1 || bits > 8 || pass < 0 || pass >= passes) return WAST3D_ERR_INVALID_ARGUMENT;
    if (n >= (size_t)OS_VALUE) return WAST3D_ERR_OVERFLOW;
    const unsigned nb = (unsigned)rs_num_blocks(n);
    const uint32_t mask = (1u << bits) - 1u;
    OnesweepWs w(ws, n, passes);
#define W3D_OS(HV, WK)                                                                                   \
    onesweep_pass_kernel<HV, WK, true><<<nb, OS_THREADS, 0, s>>>(keys_in, vals_in, keys_out, vals_out, n, shift, \
                                                                 mask, w.digit_hist + (size_t)pass * RS_RADIX,   \
                                                                 w.status(pass), w.ticket(pass), w.err, nb, nullptr)
    if (vals_in) {
        if (keys_out) W3D_OS(true, true); else W3D_OS(true, false);
    } else {
        if (keys_out) W3D_OS(false, true); else W3D_OS(false, false);
    }
#undef W3D_OS
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

uint32_t* onesweep_digit_hist(uint32_t* ws, size_t n, int passes, int pass) {
    return OnesweepWs(ws, n, passes).digit_hist + (size_t)pass * RS_RADIX;
}
uint32_t* onesweep_error_word(uint32_t* ws, size_t n, int passes) { return OnesweepWs(ws, n, passes).err; }

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

__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm,
                     uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ ws,
                     uint32_t* __restrict__ total_out, unsigned nblocks) {
    __shared__ uint32_t s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ws, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + k;
        v[k] = i < n ? (perm ? in[perm[i]] : in[i]) : 0;
        sum += v[k];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(sum, &tot);
    if (threadIdx.x < 32) {
        const uint32_t e = lookback_warp(ws + 2, tile, tot, ws + 1, threadIdx.x);
        if (threadIdx.x == 0) s_excl = e;
    }
    __syncthreads();
    ex += s_excl;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
    if (total_out && tile == nblocks - 1 && threadIdx.x == 0) *total_out = s_excl + tot;
}

size_t scan_lookback_workspace_words(size_t n) { return scan_num_blocks(n) + 2; }

int scan_exclusive_lookback_u32(const uint32_t* in, const uint32_t* perm, uint32_t* out, size_t n,
                                uint32_t* ws, uint32_t* total, cudaStream_t s, bool debug) {
    if (n == 0) {
        if (total) W3D_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(uint32_t), s));
        return WAST3D_OK;
    }
    const size_t nb = scan_num_blocks(n);
    W3D_CUDA_TRY(cudaMemsetAsync(ws, 0, scan_lookback_workspace_words(n) * sizeof(uint32_t), s));
    scan_lookback_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, perm, out, n, ws, total, (unsigned)nb);
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

}  // namespace w3d

using namespace w3d;

// Test / measurement hooks for the sort and scan primitives (no reference equivalent: the
// reference calls cub::DeviceRadixSort / cub::DeviceScan, rasterizer_impl.cu:279,305-310).
extern "C" int wast3d_test_sort_pairs(size_t n, const uint32_t* keys_in, const uint32_t* vals_in,
                                      uint32_t* keys_out, uint32_t* vals_out, int begin_bit, int end_bit,
                                      int mode, void* stream_v) {
    cudaStream_t s = (cudaStream_t)stream_v;
    if (begin_bit < 0 || end_bit > 32 || end_bit <= begin_bit || !keys_out || !vals_out || (n && !keys_in))
        return WAST3D_ERR_INVALID_ARGUMENT;
    if (n == 0) return WAST3D_OK;
    const int nbits = end_bit - begin_bit;
    const int passes = (nbits + 7) / 8;
    const int bpp = (nbits + passes - 1) / passes;
    if (passes > 4) return WAST3D_ERR_INVALID_ARGUMENT;
    // scratch: ping-pong arrays + workspace
    const size_t ws_words = onesweep_workspace_words(n, passes) + rs_hist_words(n) + scan_scratch_words(rs_hist_words(n));
    uint32_t* buf = nullptr;
    W3D_CUDA_TRY(cudaMallocAsync((void**)&buf, (2 * n + ws_words) * sizeof(uint32_t), s));
    uint32_t *ka = buf, *va = buf + n, *ws = buf + 2 * n;
    uint32_t* hist = ws + onesweep_workspace_words(n, passes);
    uint32_t* scr = hist + rs_hist_words(n);
    int st = WAST3D_OK;
    int shifts[4], bits[4];
    for (int p = 0; p < passes; ++p) {
        shifts[p] = begin_bit + p * bpp;
        bits[p] = (p == passes - 1) ? (end_bit - shifts[p]) : bpp;
    }
    if (mode >= 1) {
        st = onesweep_prepare(ws, n, passes, s);
        if (!st) st = onesweep_hist(keys_in, n, passes, shifts, bits, ws, s, false);
    }
    const uint32_t *kin = keys_in, *vin = vals_in;
    for (int p = 0; p < passes && !st; ++p) {
        // results must land in (keys_out, vals_out) after the last pass
        const bool to_out = ((passes - 1 - p) & 1) == 0;
        uint32_t* ko = to_out ? keys_out : ka;
        uint32_t* vo = to_out ? vals_out : va;
        if (mode == 0) st = radix_pass_u32(kin, vin, ko, vo, n, shifts[p], bits[p], hist, scr, s, false, nullptr);
        else st = onesweep_pass(kin, vin, ko, vo, n, shifts[p], bits[p], ws, passes, p, s, false);
        kin = ko;
        vin = vo;
    }
    if (!st && mode >= 1) {
        uint32_t h = 0;
        if (cudaMemcpyAsync(&h, onesweep_error_word(ws, n, passes), 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess || h)
            st = WAST3D_ERR_CUDA;
    }
    cudaFreeAsync(buf, s);
    return st;
}

extern "C" int wast3d_test_scan(size_t n, const uint32_t* in, const uint32_t* perm, uint32_t* out,
                                uint32_t* total, int mode, void* stream_v) {
    cudaStream_t s = (cudaStream_t)stream_v;
    if (!out && n) return WAST3D_ERR_INVALID_ARGUMENT;
    uint32_t* ws = nullptr;
    const size_t words = scan_lookback_workspace_words(n) + scan_scratch_words(n) + 8;
    W3D_CUDA_TRY(cudaMallocAsync((void**)&ws, words * sizeof(uint32_t), s));
    int st = mode == 0 ? scan_exclusive_u32(in, perm, out, n, ws, total, s, false)
                       : scan_exclusive_lookback_u32(in, perm, out, n, ws, total, s, false);
    cudaFreeAsync(ws, s);
    return st;
}

