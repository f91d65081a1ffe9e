// Device-wide exclusive scan and one stable LSD radix-sort pass over (u32 key, u32 value) pairs.
//
// Replaces the three CUB calls of the reference binning stage:
//   cub::DeviceScan::InclusiveSum        rasterizer_impl.cu:279
//   cub::DeviceRadixSort::SortPairs      rasterizer_impl.cu:305-310   (u64 key, 6 passes)
// Here the (tile, depth) order is produced as "depth sort of P Gaussians (4 passes over P
// pairs), then a stable partition of the R instances by tile id (ceil(tile_bits/8) passes
// over R pairs)" — see raster_forward.cu — so this file only needs 32-bit keys.
//
// Both primitives are multi-kernel (reduce / spine / apply) rather than single-pass
// decoupled-look-back: no inter-block spinning, so they cannot hang the device.
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
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        uint32_t wi = warp_incl_scan(w, lane);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) block_total = wi;
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

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t mask,
                  uint32_t* __restrict__ hist, unsigned nblocks) {
    __shared__ uint32_t h[RS_RADIX];
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

template <bool HAS_VALS, bool WRITE_KEYS>
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                     int shift, uint32_t mask, const uint32_t* __restrict__ bases,
                     unsigned nblocks) {
    constexpr int WARPS = RS_THREADS / 32;
    __shared__ uint32_t warp_hist[WARPS][RS_RADIX];
    __shared__ uint32_t bin_start[RS_RADIX];
    __shared__ uint32_t bin_base[RS_RADIX];
    __shared__ uint32_t skeys[RS_TILE];
    __shared__ uint32_t svals[RS_TILE];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int w = 0; w < WARPS; ++w) warp_hist[w][threadIdx.x] = 0;
    bin_base[threadIdx.x] = bases[(size_t)threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();

    const size_t tile_base = (size_t)blockIdx.x * RS_TILE;
    const size_t warp_base = tile_base + (size_t)warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        bool ok = i < n;
        key[j] = ok ? keys_in[i] : 0xFFFFFFFFu;
        val[j] = ok ? (HAS_VALS ? vals_in[i] : (uint32_t)i) : 0u;
    }
    // Out-of-range items only exist at the very end of the last tile; give them the highest
    // digit so they rank after every real key of that digit and are simply not written.
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        uint32_t d = i < n ? ((key[j] >> shift) & mask) : (RS_RADIX - 1);
        unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = warp_hist[warp][d];
        __syncwarp();
        if ((peers & lt) == 0) warp_hist[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[j] = before + __popc(peers & lt);
    }
    __syncthreads();
    // thread t owns digit t: exclusive prefix over warps, then over digits
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        uint32_t c = warp_hist[w][threadIdx.x];
        warp_hist[w][threadIdx.x] = run;
        run += c;
    }
    uint32_t tot;
    uint32_t start = block_excl_scan(run, &tot);
    bin_start[threadIdx.x] = start;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t i = warp_base + (size_t)j * 32 + lane;
        uint32_t d = i < n ? ((key[j] >> shift) & mask) : (RS_RADIX - 1);
        uint32_t pos = bin_start[d] + warp_hist[warp][d] + rank[j];
        skeys[pos] = key[j];
        svals[pos] = val[j];
    }
    __syncthreads();
    const size_t remaining = n - tile_base;
    const uint32_t count = remaining < (size_t)RS_TILE ? (uint32_t)remaining : (uint32_t)RS_TILE;
    for (uint32_t p = threadIdx.x; p < count; p += RS_THREADS) {
        uint32_t k = skeys[p];
        uint32_t d = (k >> shift) & mask;
        size_t g = (size_t)bin_base[d] + (p - bin_start[d]);
        if (WRITE_KEYS) keys_out[g] = k;
        vals_out[g] = svals[p];
    }
}

int radix_pass_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                   uint32_t* vals_out, size_t n, int shift, int bits, uint32_t* hist,
                   uint32_t* scan_scratch, cudaStream_t s, bool debug) {
    if (n == 0) return WAST3D_OK;
    if (bits < 1 || bits > 8) return WAST3D_ERR_INVALID_ARGUMENT;
    if (n > 0xFFFFFFFFull - RS_TILE) return WAST3D_ERR_OVERFLOW;
    const unsigned nb = (unsigned)rs_num_blocks(n);
    const uint32_t mask = (1u << bits) - 1u;
    radix_hist_kernel<<<nb, RS_THREADS, 0, s>>>(keys_in, n, shift, mask, hist, nb);
    W3D_AFTER_LAUNCH(s, debug);
    int st = scan_exclusive_u32(hist, nullptr, hist, (size_t)nb * RS_RADIX, scan_scratch, nullptr,
                                s, debug);
    if (st != WAST3D_OK) return st;
#define W3D_SCATTER(HV, WK)                                                                    \
    radix_scatter_kernel<HV, WK><<<nb, RS_THREADS, 0, s>>>(keys_in, vals_in, keys_out,         \
                                                           vals_out, n, shift, mask, hist, nb)
    if (vals_in) {
        if (keys_out) W3D_SCATTER(true, true); else W3D_SCATTER(true, false);
    } else {
        if (keys_out) W3D_SCATTER(false, true); else W3D_SCATTER(false, false);
    }
#undef W3D_SCATTER
    W3D_AFTER_LAUNCH(s, debug);
    return WAST3D_OK;
}

}  // namespace w3d
