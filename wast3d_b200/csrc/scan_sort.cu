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

// ------------------------------------------------------------------ single-pass variants --
// Decoupled look-back ("onesweep"): ONE kernel per radix pass and ONE kernel per scan instead of
// hist + 3-kernel table scan + scatter.  A block takes its tile index from an atomic ticket, so
// a tile only ever waits on tiles whose blocks are already running or finished (deadlock-free
// without any assumption on block scheduling); every wait is bounded anyway and raises an error
// flag instead of hanging the device.  Status words carry the flag in bits [31:30]
// (1 = aggregate of this tile only, 2 = inclusive prefix up to this tile) and a 30-bit count.
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
// 8 warps with 16, and 3 blocks (48 warps) fit an SM.  The values never pass through registers:
// they are staged with cp.async at kernel entry (in flight during the whole ranking phase) and
// permuted shared -> shared.  The look-back wait is placed after the keys have been ranked and
// parked in shared memory, when predecessors have had time to publish.
constexpr int OS_ITEMS = 8;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
static_assert(OS_TILE == RS_TILE, "workspace sizing assumes the same tile size");
constexpr int OS_WARPS = OS_THREADS / 32;
struct OsSmem {
    uint16_t warp_hist[OS_WARPS][RS_RADIX];  // <= OS_TILE, fits 16 bits
    uint32_t bin_start[RS_RADIX];            // first local position of a digit
    uint32_t bin_off[RS_RADIX];              // global position of a digit's first key minus bin_start
    uint32_t skeys[OS_TILE];
    uint32_t svals[OS_TILE];
    uint32_t stage[OS_TILE];                 // values in load order
    uint32_t warp_total[RS_RADIX / 32];
    uint32_t tile;
};

// LOOKBACK = false: the multi-kernel flavour — `digit_base` is then the scanned bin-major table
// of per-tile scatter bases (bases[digit * nblocks + tile]) and status/ticket/err are unused.
// LOOKBACK = true: `digit_base` holds the EXCLUSIVE prefix of the global digit histogram.
template <bool HAS_VALS, bool WRITE_KEYS, bool LOOKBACK, int LBW>
__global__ void __launch_bounds__(OS_THREADS, 3)
onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                     int shift, uint32_t mask, const uint32_t* __restrict__ digit_base,
                     uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                     uint32_t* __restrict__ err, unsigned nblocks, const uint32_t* __restrict__ n_dev,
                     const GatherRect gather) {
    extern __shared__ __align__(16) unsigned char os_smem_raw[];
    OsSmem& sm = *reinterpret_cast<OsSmem*>(os_smem_raw);
    if (n_dev != nullptr) n = min(n, (size_t)*n_dev);   // device-side element count (see radix_hist_kernel)
    if (!LOOKBACK && (size_t)blockIdx.x * OS_TILE >= n) return;   // whole block past the end (uniform)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    if (LOOKBACK && threadIdx.x == 0) sm.tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < OS_WARPS * RS_RADIX / 2; i += OS_THREADS)
        reinterpret_cast<uint32_t*>(&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = LOOKBACK ? sm.tile : blockIdx.x;
    // look-back with a device-side count: tickets past the last tile have nothing to sort and nobody waits for them
    if (LOOKBACK && (size_t)tile * OS_TILE >= n) return;

    const size_t tile_base = (size_t)tile * OS_TILE;
    const size_t remaining = n - tile_base;
    const uint32_t count_tile = remaining < (size_t)OS_TILE ? (uint32_t)remaining : (uint32_t)OS_TILE;
    const uint32_t* tkeys = keys_in + tile_base;
    const uint32_t local0 = (uint32_t)warp * (32 * OS_ITEMS) + lane;   // + 32 j: this thread's j-th item in the tile

    if (HAS_VALS) {
        const uint32_t* tvals = vals_in + tile_base;
        if ((reinterpret_cast<uintptr_t>(tvals) & 15) == 0) {
#pragma unroll
            for (int c = 0; c < OS_ITEMS / 4; ++c) {
                const uint32_t e = 4 * (c * OS_THREADS + threadIdx.x);
                if (e + 4 <= count_tile) cp_async16(sm.stage + e, tvals + e);
                else
                    for (uint32_t k = e; k < count_tile; ++k) cp_async4(sm.stage + k, tvals + k);
            }
        } else {
#pragma unroll
            for (int j = 0; j < OS_ITEMS; ++j) {
                const uint32_t e = j * OS_THREADS + threadIdx.x;
                if (e < count_tile) cp_async4(sm.stage + e, tvals + e);
            }
        }
        cp_async_commit();
    }
    uint32_t key[OS_ITEMS];
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const uint32_t li = local0 + 32 * j;
        key[j] = li < count_tile ? tkeys[li] : 0xFFFFFFFFu;
    }
    // multi-kernel flavour: this tile's scatter bases are an L2 round trip away - fetch them under the ranking phase
    // digits are < nbins = mask + 1 <= RS_RADIX: thread t < nbins owns digit t (its count, its look-back column)
    const bool owner = threadIdx.x <= mask;
    uint32_t dstart = 0;
    if (owner) dstart = LOOKBACK ? digit_base[threadIdx.x] : digit_base[(size_t)threadIdx.x * nblocks + tile];
    // Out-of-range items only exist at the very end of the last tile; give them the highest
    // digit (`mask`) so they rank after every real key of that digit and are simply not written.
    // All match operations are issued before the dependent shared-memory chain starts.
    unsigned peers[OS_ITEMS];
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const uint32_t d = local0 + 32 * j < count_tile ? ((key[j] >> shift) & mask) : mask;
        peers[j] = __match_any_sync(0xffffffffu, d);
    }
    // (one ballot per digit bit instead of MATCH.ANY was measured slower: tile partition 0.270 vs 0.241 ms at C3)
    uint32_t rank2[OS_ITEMS / 2];   // ranks inside the warp's run of the digit, two 16-bit fields per register
    uint16_t* const my_hist = sm.warp_hist[warp];
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const uint32_t d = local0 + 32 * j < count_tile ? ((key[j] >> shift) & mask) : mask;
        const uint32_t before = my_hist[d];
        __syncwarp();
        if ((peers[j] & lt) == 0) my_hist[d] = (uint16_t)(before + __popc(peers[j]));
        __syncwarp();
        const uint32_t r = before + __popc(peers[j] & lt);
        if (j & 1) rank2[j >> 1] |= r << 16; else rank2[j >> 1] = r;
    }
    __syncthreads();
    uint32_t run = 0;
    if (owner) {
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) {
            const uint32_t c = sm.warp_hist[w][threadIdx.x];
            sm.warp_hist[w][threadIdx.x] = (uint16_t)run;
            run += c;
        }
    }
    // Padding of the last tile was counted under the highest digit; it must not be published
    // (it ranks after every real key of that digit, so local positions stay right).
    uint32_t real = run;
    if (threadIdx.x == mask) real -= (uint32_t)OS_TILE - count_tile;
    uint32_t* my_status = status + (size_t)tile * RS_RADIX + threadIdx.x;
    if (LOOKBACK && owner) st_status(my_status, real | (tile == 0 ? OS_PREFIX : OS_AGG));
    // exclusive scan of the digit totals: the owners sit in warps 0..7 (run = 0 elsewhere)
    const uint32_t incl = warp_incl_scan(run, lane);
    if (warp < RS_RADIX / 32 && lane == 31) sm.warp_total[warp] = incl;
    if (HAS_VALS) cp_async_wait_all();
    __syncthreads();
    uint32_t start = incl - run;
    if (owner) {
#pragma unroll
        for (int w = 0; w < RS_RADIX / 32 - 1; ++w) start += w < warp ? sm.warp_total[w] : 0u;
        sm.bin_start[threadIdx.x] = start;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OS_ITEMS; ++j) {
        const uint32_t li = local0 + 32 * j;
        const uint32_t d = li < count_tile ? ((key[j] >> shift) & mask) : mask;
        const uint32_t r = (j & 1) ? (rank2[j >> 1] >> 16) : (rank2[j >> 1] & 0xFFFFu);
        const uint32_t pos = sm.bin_start[d] + my_hist[d] + r;
        sm.skeys[pos] = key[j];
        sm.svals[pos] = HAS_VALS ? sm.stage[li] : (uint32_t)tile_base + li;
    }
    // look-back over earlier tiles, one digit column per thread
    if (owner) {
        uint32_t excl = 0;
        if (LOOKBACK && tile != 0) {
            // LBW predecessors are read per round trip (independent loads) and consumed nearest first up to the
            // first one that has not published yet: when a whole wave of tiles starts together (the depth sort is
            // barely two waves) the inclusive prefixes then spread ~sqrt(LBW) times faster than one at a time
            uint32_t spins = 0;
            bool failed = false, done = false;
            int left = (int)tile;   // predecessors not yet summed; sp = status word of the nearest of them
            const uint32_t* sp = status + (size_t)(tile - 1) * RS_RADIX + threadIdx.x;
            while (!done && !failed) {
                uint32_t v[LBW];
#pragma unroll
                for (int i = 0; i < LBW; ++i) v[i] = i < left ? ld_status(sp - (size_t)i * RS_RADIX) : OS_PREFIX;
                int consumed = 0;
#pragma unroll
                for (int i = 0; i < LBW; ++i) {
                    const uint32_t f = v[i] >> 30;
                    if (!done && consumed == i && f != 0u) {
                        excl += v[i] & OS_VALUE;
                        consumed = i + 1;
                        done = f == 2u;
                    }
                }
                left -= consumed;
                sp -= (size_t)consumed * RS_RADIX;
                if (consumed == 0 && ++spins > OS_SPIN_LIMIT) failed = true;
            }
            if (failed) atomicOr(err, 1u);
            st_status(my_status, (excl + real) | OS_PREFIX);
        }
        sm.bin_off[threadIdx.x] = dstart + excl - start;
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < count_tile; p += OS_THREADS) {
        const uint32_t k = sm.skeys[p];
        const size_t g = (size_t)(sm.bin_off[(k >> shift) & mask] + p);
        if (WRITE_KEYS) keys_out[g] = k;
        const uint32_t v = sm.svals[p];
        vals_out[g] = v;
        if (gather.src != nullptr) {   // uniform
            const uint2 rc = __ldg(gather.src + v);
            gather.dst[g] = rc;
            gather.cnt[g] = (rc.y & 0xFFFFu) * (rc.y >> 16);
        }
    }
}

// Exclusive prefix (in place) of `passes` global digit histograms — the look-back passes take their digit bases from it.
__global__ void __launch_bounds__(RS_RADIX)
digit_scan_kernel(uint32_t* __restrict__ digit_hist) {
    uint32_t* h = digit_hist + (size_t)blockIdx.x * RS_RADIX;
    uint32_t tot;
    const uint32_t v = h[threadIdx.x];
    const uint32_t ex = block_excl_scan<RS_RADIX>(v, &tot);
    h[threadIdx.x] = ex;
}

template <bool HV, bool WK, bool LB, int LBW>
static void launch_pass(unsigned nb, cudaStream_t s, const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                        uint32_t* vals_out, size_t n, int shift, uint32_t mask, const uint32_t* digit_base,
                        uint32_t* status, uint32_t* ticket, uint32_t* err, const uint32_t* n_dev, GatherRect gather) {
    auto k = onesweep_pass_kernel<HV, WK, LB, LBW>;
    // 58 KB of dynamic shared memory: opt in (per device, so set on every launch; the call only records a number)
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsSmem));
    k<<<nb, OS_THREADS, sizeof(OsSmem), s>>>(keys_in, vals_in, keys_out, vals_out, n, shift, mask, digit_base, status,
                                             ticket, err, nb, n_dev, gather);
}
template <bool LB>
static void launch_pass_any(unsigned nb, cudaStream_t s, const uint32_t* keys_in, const uint32_t* vals_in,
                            uint32_t* keys_out, uint32_t* vals_out, size_t n, int shift, uint32_t mask,
                            const uint32_t* digit_base, uint32_t* status, uint32_t* ticket, uint32_t* err,
                            const uint32_t* n_dev, int lb_window, GatherRect gather) {
#define W3D_PASS(HV, WK)                                                                                          \
    do {                                                                                                          \
        if (!LB || lb_window <= 1)                                                                                \
            launch_pass<HV, WK, LB, 1>(nb, s, keys_in, vals_in, keys_out, vals_out, n, shift, mask, digit_base,   \
                                       status, ticket, err, n_dev, gather);                                       \
        else if (lb_window <= 4)                                                                                  \
            launch_pass<HV, WK, LB, 4>(nb, s, keys_in, vals_in, keys_out, vals_out, n, shift, mask, digit_base,   \
                                       status, ticket, err, n_dev, gather);                                       \
        else if (lb_window <= 8)                                                                                  \
            launch_pass<HV, WK, LB, 8>(nb, s, keys_in, vals_in, keys_out, vals_out, n, shift, mask, digit_base,   \
                                       status, ticket, err, n_dev, gather);                                       \
        else                                                                                                      \
            launch_pass<HV, WK, LB, 16>(nb, s, keys_in, vals_in, keys_out, vals_out, n, shift, mask, digit_base,  \
                                        status, ticket, err, n_dev, gather);                                      \
    } while (0)
    if (vals_in) {
        if (keys_out) W3D_PASS(true, true); else W3D_PASS(true, false);
    } else {
        if (keys_out) W3D_PASS(false, true); else W3D_PASS(false, false);
    }
#undef W3D_PASS
}

int radix_pass_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, size_t n, int shift, int bits, uint32_t* hist,
                   uint32_t* scan_scratch, cudaStream_t s, bool debug, const uint32_t* n_dev, GatherRect gather) {
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
    launch_pass_any<false>(nb, s, keys_in, vals_in, keys_out, vals_out, n, shift, mask, hist, nullptr, nullptr, nullptr,
                           n_dev, 1, gather);
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
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

// The look-back passes read digit BASES: turn the `passes` global digit histograms at the head of the workspace into
// their exclusive prefixes (callers that fill onesweep_digit_hist() themselves call this afterwards).
int onesweep_scan_digits(uint32_t* ws, size_t n, int passes, cudaStream_t s, bool debug) {
    if (n == 0) return WAST3D_OK;
    digit_scan_kernel<<<passes, RS_RADIX, 0, s>>>(OnesweepWs(ws, n, passes).digit_hist);
    W3D_AFTER_LAUNCH(s, debug);
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
    return onesweep_scan_digits(ws, n, passes, s, debug);
}

// WAST3D_LB_WINDOW: predecessors read per look-back round trip (1, 4, 8 or 16; default 8)
static int lookback_window() {
    static const int w = [] {
        const char* e = getenv("WAST3D_LB_WINDOW");
        int v = e ? atoi(e) : 8;
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    return w;
}

int onesweep_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                  size_t n, int shift, int bits, uint32_t* ws, int passes, int pass, cudaStream_t s, bool debug,
                  const uint32_t* n_dev, GatherRect gather) {
    if (n == 0) return WAST3D_OK;
    if (bits < 1 || bits > 8 || pass < 0 || pass >= passes) return WAST3D_ERR_INVALID_ARGUMENT;
    if (n >= (size_t)OS_VALUE) return WAST3D_ERR_OVERFLOW;
    const unsigned nb = (unsigned)rs_num_blocks(n);
    const uint32_t mask = (1u << bits) - 1u;
    OnesweepWs w(ws, n, passes);
    launch_pass_any<true>(nb, s, keys_in, vals_in, keys_out, vals_out, n, shift, mask,
                          w.digit_hist + (size_t)pass * RS_RADIX, w.status(pass), w.ticket(pass), w.err, n_dev,
                          lookback_window(), gather);
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

uint32_t* onesweep_digit_hist(uint32_t* ws, size_t n, int passes, int pass) {
    return OnesweepWs(ws, n, passes).digit_hist + (size_t)pass * RS_RADIX;
}
uint32_t* onesweep_error_word(uint32_t* ws, size_t n, int passes) { return OnesweepWs(ws, n, passes).err; }

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
        else st = onesweep_pass(kin, vin, ko, vo, n, shifts[p], bits[p], ws, passes, p, s, false, nullptr);
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

