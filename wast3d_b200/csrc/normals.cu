// Depth -> normal map of the train_st_normals variant (BASELINE.json configs[4]), forward and backward.
//
// Replaces, on the step's path, the torch / kornia expression of train_st_normals.py:113-123:
//   normals = kornia.geometry.depth.depth_to_normals(depth[None,None], K, normalize_points=False)
//   image_normals = (normals - amin(normals)) / (amax(normals) - amin(normals) + 1e-6)
// kornia is an un-vendored, unpinned dependency of the reference (absent from environment.yml and from
// /root/reference); its published algorithm is restated here and in oracle/normals.py:
//   xyz(u,v)   = depth(u,v) * ((u - cx)/fx, (v - cy)/fy, 1)                 depth_to_3d / unproject_points
//   a, b       = Sobel_x(xyz), Sobel_y(xyz): 3x3, kernel / 8, replicate padding, cross-correlation
//                                                                            spatial_gradient(mode="sobel", normalized=True)
//   n          = normalize(a x b, eps = 1e-12)                               torch.cross + F.normalize
// In torch this is ~25 kernels forward (meshgrid, unproject, pad, conv3d, cross, norm, clamp, div, amin, amax,
// sub, sub, add, div ...) and as many backward, each a pass over 3-6 image planes; here: two kernels forward
// (unit normals + global min/max; rescale) and three backward (two global sums; per-pixel chain rule to the Sobel
// gradients; gather to depth).  All global reductions are order-independent (integer / ordered-key atomics) or
// summed in a fixed order (double partials), so the result is deterministic.
#include "raster_math.cuh"

namespace w3d {

constexpr int NRM_THREADS = 256;
constexpr int NRM_MAX_BLOCKS = 148 * 8;

struct Intrinsics { float fx, fy, cx, cy; };

// scratch words: [0] min key, [1] max key (float_order_key), [2] block counter, [3] pad; doubles from byte 128:
// red[0..3] = S0, S1, count(min), count(max); then per-block partials [NRM_MAX_BLOCKS][2]
struct NrmScratch {
    uint32_t* words;
    double* red;
    double* partials;
    unsigned long long* counts;
    explicit NrmScratch(void* p) {
        words = (uint32_t*)p;
        red = (double*)((char*)p + 128);
        counts = (unsigned long long*)((char*)p + 192);
        partials = (double*)((char*)p + 256);
    }
};

__device__ __forceinline__ float3 point_at(const float* __restrict__ depth, int W, int H, int x, int y, Intrinsics k) {
    x = min(max(x, 0), W - 1);   // replicate padding of the xyz image: the border pixel's own point
    y = min(max(y, 0), H - 1);
    const float d = depth[(size_t)y * W + x];
    const float px = ((float)x - k.cx) / k.fx;
    const float py = ((float)y - k.cy) / k.fy;
    // explicit multiplies: the differences of neighbouring points below must not be contracted into
    // fma(px, d, -other) — that would turn "the same point twice" (replicate padding) into its rounding residue
    return make_float3(__fmul_rn(px, d), __fmul_rn(py, d), d);
}

// Sobel gradients of the point image at (x, y): a = d/dx, b = d/dy (kernels / 8).  Differences of the two
// opposite neighbours first: they are nearly equal for a smooth depth, so the subtraction is (almost) exact.
__device__ __forceinline__ void sobel_ab(const float* __restrict__ depth, int W, int H, int x, int y, Intrinsics k,
                                         float3& a, float3& b) {
    const float3 p00 = point_at(depth, W, H, x - 1, y - 1, k), p01 = point_at(depth, W, H, x, y - 1, k),
                 p02 = point_at(depth, W, H, x + 1, y - 1, k), p10 = point_at(depth, W, H, x - 1, y, k),
                 p12 = point_at(depth, W, H, x + 1, y, k), p20 = point_at(depth, W, H, x - 1, y + 1, k),
                 p21 = point_at(depth, W, H, x, y + 1, k), p22 = point_at(depth, W, H, x + 1, y + 1, k);
    a.x = ((p02.x - p00.x) + 2.f * (p12.x - p10.x) + (p22.x - p20.x)) * 0.125f;
    a.y = ((p02.y - p00.y) + 2.f * (p12.y - p10.y) + (p22.y - p20.y)) * 0.125f;
    a.z = ((p02.z - p00.z) + 2.f * (p12.z - p10.z) + (p22.z - p20.z)) * 0.125f;
    b.x = ((p20.x - p00.x) + 2.f * (p21.x - p01.x) + (p22.x - p02.x)) * 0.125f;
    b.y = ((p20.y - p00.y) + 2.f * (p21.y - p01.y) + (p22.y - p02.y)) * 0.125f;
    b.z = ((p20.z - p00.z) + 2.f * (p21.z - p01.z) + (p22.z - p02.z)) * 0.125f;
}

__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

__global__ void __launch_bounds__(NRM_THREADS)
normals_unit_kernel(int H, int W, const float* __restrict__ depth, Intrinsics k, float* __restrict__ unit,
                    uint32_t* __restrict__ words) {
    const size_t HW = (size_t)H * W;
    const float inf = __int_as_float(0x7f800000);
    float lo = inf, hi = -inf;
    for (size_t i = (size_t)blockIdx.x * NRM_THREADS + threadIdx.x; i < HW; i += (size_t)gridDim.x * NRM_THREADS) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        float3 a, b;
        sobel_ab(depth, W, H, x, y, k, a, b);
        const float3 n = cross3(a, b);
        const float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
        const float den = fmaxf(len, 1e-12f);   // F.normalize
        const float ux = n.x / den, uy = n.y / den, uz = n.z / den;
        unit[i] = ux;
        unit[HW + i] = uy;
        unit[2 * HW + i] = uz;
        lo = fminf(lo, fminf(ux, fminf(uy, uz)));
        hi = fmaxf(hi, fmaxf(ux, fmaxf(uy, uz)));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(words + 0, float_order_key(lo));
        atomicMax(words + 1, float_order_key(hi));
    }
}

__global__ void __launch_bounds__(NRM_THREADS)
normals_scale_kernel(size_t n, const float* __restrict__ unit, const uint32_t* __restrict__ words,
                     float* __restrict__ out01, float* __restrict__ minmax) {
    const float m = float_from_order_key(words[0]), M = float_from_order_key(words[1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { minmax[0] = m; minmax[1] = M; }
    const float r = (M - m) + 1e-6f;
    for (size_t i = (size_t)blockIdx.x * NRM_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NRM_THREADS)
        out01[i] = (unit[i] - m) / r;
}

// S0 = sum g, S1 = sum g (u - m), and how many elements attain the minimum / the maximum (amin / amax share their
// gradient evenly among ties, like torch).  Fixed-order two-level sum in double.
__global__ void __launch_bounds__(NRM_THREADS)
normals_bwd_reduce_kernel(size_t n, const float* __restrict__ unit, const float* __restrict__ minmax,
                          const float* __restrict__ g01, double* __restrict__ partials, uint32_t* __restrict__ counter,
                          double* __restrict__ red, unsigned long long* __restrict__ counts) {
    const float m = minmax[0], M = minmax[1];
    double s0 = 0.0, s1 = 0.0;
    unsigned cm = 0, cM = 0;
    for (size_t i = (size_t)blockIdx.x * NRM_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NRM_THREADS) {
        const float u = unit[i], g = g01[i];
        s0 += (double)g;
        s1 += (double)g * (double)(u - m);
        cm += (u == m);
        cM += (u == M);
    }
    __shared__ double r0[NRM_THREADS / 32], r1[NRM_THREADS / 32];
    __shared__ bool last;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d);
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
        cm += __shfl_xor_sync(0xffffffffu, cm, d);
        cM += __shfl_xor_sync(0xffffffffu, cM, d);
    }
    if ((threadIdx.x & 31) == 0) {
        r0[threadIdx.x >> 5] = s0;
        r1[threadIdx.x >> 5] = s1;
        if (cm) atomicAdd(counts + 0, (unsigned long long)cm);
        if (cM) atomicAdd(counts + 1, (unsigned long long)cM);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < NRM_THREADS / 32; ++w) { t0 += r0[w]; t1 += r1[w]; }
        partials[2 * (size_t)blockIdx.x] = t0;
        partials[2 * (size_t)blockIdx.x + 1] = t1;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (!last) return;
    // last block: partials added in an order fixed by the grid size (thread t takes blocks t, t + 256, ...; fixed tree)
    __shared__ double sred[2][NRM_THREADS];
    double t0 = 0.0, t1 = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += NRM_THREADS) {
        t0 += ((volatile double*)partials)[2 * (size_t)b];
        t1 += ((volatile double*)partials)[2 * (size_t)b + 1];
    }
    sred[0][threadIdx.x] = t0;
    sred[1][threadIdx.x] = t1;
    __syncthreads();
    for (int sft = NRM_THREADS / 2; sft > 0; sft >>= 1) {
        if ((int)threadIdx.x < sft) {
            sred[0][threadIdx.x] += sred[0][threadIdx.x + sft];
            sred[1][threadIdx.x] += sred[1][threadIdx.x + sft];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    red[0] = sred[0][0];
    red[1] = sred[1][0];
    red[2] = (double)((volatile unsigned long long*)counts)[0];
    red[3] = (double)((volatile unsigned long long*)counts)[1];
    *counter = 0;
}

// Per pixel: gradient of the rescaled normal -> unit normal (incl. the amin / amax terms) -> a x b -> (a, b).
__global__ void __launch_bounds__(NRM_THREADS)
normals_bwd_ab_kernel(int H, int W, const float* __restrict__ depth, Intrinsics k, const float* __restrict__ unit,
                      const float* __restrict__ minmax, const float* __restrict__ g01, const double* __restrict__ red,
                      float* __restrict__ gab /*[6,H,W]*/) {
    const size_t HW = (size_t)H * W;
    const float m = minmax[0], M = minmax[1];
    const float r = (M - m) + 1e-6f;
    const double S0 = red[0], S1 = red[1];
    const double rr = (double)r * (double)r;
    const float g_m = (float)((S1 - (double)r * S0) / rr / fmax(red[2], 1.0));   // per tied minimum element
    const float g_M = (float)(-S1 / rr / fmax(red[3], 1.0));                      // per tied maximum element
    for (size_t i = (size_t)blockIdx.x * NRM_THREADS + threadIdx.x; i < HW; i += (size_t)gridDim.x * NRM_THREADS) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        float u[3], gu[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            u[c] = unit[c * HW + i];
            gu[c] = g01[c * HW + i] / r + (u[c] == m ? g_m : 0.f) + (u[c] == M ? g_M : 0.f);
        }
        float3 a, b;
        sobel_ab(depth, W, H, x, y, k, a, b);
        const float3 n = cross3(a, b);
        const float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
        float3 gv;
        if (len > 1e-12f) {   // d(v/|v|): (g - u (u.g)) / |v|
            const float dot = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2];
            gv = make_float3((gu[0] - u[0] * dot) / len, (gu[1] - u[1] * dot) / len, (gu[2] - u[2] * dot) / len);
        } else {              // clamped denominator: v / eps is linear
            gv = make_float3(gu[0] / 1e-12f, gu[1] / 1e-12f, gu[2] / 1e-12f);
        }
        const float3 ga = cross3(b, gv), gb = cross3(gv, a);
        gab[i] = ga.x; gab[HW + i] = ga.y; gab[2 * HW + i] = ga.z;
        gab[3 * HW + i] = gb.x; gab[4 * HW + i] = gb.y; gab[5 * HW + i] = gb.z;
    }
}

// Gather: pixel q's point enters the Sobel sums of every (p, offset) with clamp(p + offset) == q (replicate
// padding: border points are used several times).  Per axis at most three such (p, offset) pairs exist.
__device__ __forceinline__ int axis_pairs(int q, int n, int* p_of, int* d_of) {
    int cnt = 0;
    for (int d = -1; d <= 1; ++d) {
        for (int p = q - 1; p <= q + 1; ++p) {
            if (p < 0 || p >= n) continue;
            if (min(max(p + d, 0), n - 1) == q) {
                if (cnt < 9) { p_of[cnt] = p; d_of[cnt] = d; ++cnt; }
            }
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(NRM_THREADS)
normals_bwd_depth_kernel(int H, int W, Intrinsics k, const float* __restrict__ gab, float* __restrict__ g_depth) {
    const size_t HW = (size_t)H * W;
    const float smooth[3] = {1.f, 2.f, 1.f}, diff[3] = {-1.f, 0.f, 1.f};
    for (size_t i = (size_t)blockIdx.x * NRM_THREADS + threadIdx.x; i < HW; i += (size_t)gridDim.x * NRM_THREADS) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        int px[9], dx[9], py[9], dy[9];
        const int nx = axis_pairs(x, W, px, dx), ny = axis_pairs(y, H, py, dy);
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix) {
                const size_t p = (size_t)py[iy] * W + px[ix];
                const float wa = smooth[dy[iy] + 1] * diff[dx[ix] + 1] * 0.125f;   // Sobel x weight
                const float wb = diff[dy[iy] + 1] * smooth[dx[ix] + 1] * 0.125f;   // Sobel y weight
                gx += wa * gab[p] + wb * gab[3 * HW + p];
                gy += wa * gab[HW + p] + wb * gab[4 * HW + p];
                gz += wa * gab[2 * HW + p] + wb * gab[5 * HW + p];
            }
        const float ux = ((float)x - k.cx) / k.fx, uy = ((float)y - k.cy) / k.fy;
        g_depth[i] = gx * ux + gy * uy + gz;
    }
}

static unsigned nrm_blocks(size_t n) {
    size_t b = (n + NRM_THREADS - 1) / NRM_THREADS;
    if (b > NRM_MAX_BLOCKS) b = NRM_MAX_BLOCKS;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace w3d

using namespace w3d;

extern "C" size_t wast3d_depth_normals_scratch_bytes(void) { return 256 + (size_t)NRM_MAX_BLOCKS * 2 * sizeof(double); }

extern "C" int wast3d_depth_normals_forward(int H, int W, const float* depth, float fx, float fy, float cx, float cy,
                                            float* normals_unit, float* normals01, float* minmax, void* scratch,
                                            void* stream_v) {
    if (H < 1 || W < 1 || !depth || !normals_unit || !normals01 || !minmax || !scratch || !(fx != 0.f) || !(fy != 0.f))
        return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    NrmScratch sc(scratch);
    const Intrinsics k{fx, fy, cx, cy};
    const size_t HW = (size_t)H * W;
    W3D_CUDA_TRY(cudaMemsetAsync(sc.words, 0xFF, 4, s));       // min key
    W3D_CUDA_TRY(cudaMemsetAsync(sc.words + 1, 0, 12, s));     // max key, counter, pad
    normals_unit_kernel<<<nrm_blocks(HW), NRM_THREADS, 0, s>>>(H, W, depth, k, normals_unit, sc.words);
    W3D_AFTER_LAUNCH(s, false);
    normals_scale_kernel<<<nrm_blocks(3 * HW), NRM_THREADS, 0, s>>>(3 * HW, normals_unit, sc.words, normals01, minmax);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" int wast3d_depth_normals_backward(int H, int W, const float* depth, float fx, float fy, float cx, float cy,
                                             const float* normals_unit, const float* minmax, const float* grad01,
                                             float* grad_ab, float* grad_depth, void* scratch, void* stream_v) {
    if (H < 1 || W < 1 || !depth || !normals_unit || !minmax || !grad01 || !grad_ab || !grad_depth || !scratch)
        return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    NrmScratch sc(scratch);
    const Intrinsics k{fx, fy, cx, cy};
    const size_t HW = (size_t)H * W;
    W3D_CUDA_TRY(cudaMemsetAsync(sc.words + 2, 0, 4, s));                          // block counter
    W3D_CUDA_TRY(cudaMemsetAsync(sc.counts, 0, 2 * sizeof(unsigned long long), s));
    normals_bwd_reduce_kernel<<<nrm_blocks(3 * HW), NRM_THREADS, 0, s>>>(3 * HW, normals_unit, minmax, grad01, sc.partials,
                                                                         sc.words + 2, sc.red, sc.counts);
    W3D_AFTER_LAUNCH(s, false);
    normals_bwd_ab_kernel<<<nrm_blocks(HW), NRM_THREADS, 0, s>>>(H, W, depth, k, normals_unit, minmax, grad01, sc.red, grad_ab);
    W3D_AFTER_LAUNCH(s, false);
    normals_bwd_depth_kernel<<<nrm_blocks(HW), NRM_THREADS, 0, s>>>(H, W, k, grad_ab, grad_depth);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
