// Fused pixel losses of the style-optimisation loop and their gradients.
//
// Replaces, on the step's path, the reference's
//   l1_loss   utils/loss_utils.py:18-19    torch.abs(out - gt).mean()
//   tv_loss   utils/loss_utils.py:213-215  0.5 * (|d/dy img|.mean() + |d/dx img|.mean())
// (train_st_normals.py:127,145) plus a depth L2 term ((depth - depth_gt)^2).mean() for the depth-loss
// variant, which in torch are ~20 elementwise/reduction kernels forward and as many backward, each a
// full pass over the image.  Here: one kernel reads the image once and produces the weighted loss
// (deterministic two-level reduction, accumulated in double), one kernel writes dL/dimg and dL/ddepth.
//
//   loss = w_l1 * mean|img - gt| + w_tv * 0.5 * (mean|img[y+1]-img[y]| + mean|img[x+1]-img[x]|)
//        + w_depth * mean((depth - depth_gt)^2)
//
// img, gt: [C,H,W]; depth, depth_gt: [H,W] (optional).  sign(0) = 0 like torch.
#include "common.cuh"

namespace w3d {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }

// Row-based loops: a block walks whole image rows (row = c * H + y), threads stride over x.  All index arithmetic
// is 32-bit with one modulo per ROW (the first version did two 64-bit divisions per ELEMENT and spent its time there).
__global__ void __launch_bounds__(LOSS_THREADS)
pixel_loss_forward_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                          const float* __restrict__ depth, const float* __restrict__ depth_gt, float w_l1,
                          float w_tv, float w_depth, double* __restrict__ partials, unsigned* __restrict__ counter,
                          float* __restrict__ out_loss) {
    const size_t HW = (size_t)H * W, n = (size_t)C * HW;
    const int rows = C * H;
    float s_l1 = 0.f, s_ty = 0.f, s_tx = 0.f, s_d = 0.f;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int y = row % H;
        const bool down = y + 1 < H;
        const bool drow = depth != nullptr && row < H;   // channel 0's rows index the [H,W] depth image as well
        const float* r = img + (size_t)row * W;
        const float* g = gt ? gt + (size_t)row * W : nullptr;
        for (int x = threadIdx.x; x < W; x += LOSS_THREADS) {
            const float v = r[x];
            if (g) s_l1 += fabsf(v - g[x]);
            if (down) s_ty += fabsf(r[x + W] - v);
            if (x + 1 < W) s_tx += fabsf(r[x + 1] - v);
            if (drow) {
                const float d = depth[(size_t)row * W + x] - depth_gt[(size_t)row * W + x];
                s_d += d * d;
            }
        }
    }
    __shared__ float red[4][LOSS_THREADS / 32];
    __shared__ double sred[4][LOSS_THREADS];
    __shared__ bool last;
    float v4[4] = {s_l1, s_ty, s_tx, s_d};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v4[k] += __shfl_xor_sync(0xffffffffu, v4[k], d);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v4[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) t += (double)red[threadIdx.x][w];
        partials[4 * (size_t)blockIdx.x + threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (!last) return;
    // The last block adds the per-block partials in an order that depends only on the grid size (thread t takes blocks
    // t, t + 256, ...; then a fixed tree): deterministic, and parallel instead of one thread walking ~1200 partials.
    {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (unsigned b = threadIdx.x; b < gridDim.x; b += LOSS_THREADS)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += ((volatile double*)partials)[4 * (size_t)b + k];
#pragma unroll
        for (int k = 0; k < 4; ++k) sred[k][threadIdx.x] = acc[k];
        __syncthreads();
        for (int sft = LOSS_THREADS / 2; sft > 0; sft >>= 1) {
            if ((int)threadIdx.x < sft)
#pragma unroll
                for (int k = 0; k < 4; ++k) sred[k][threadIdx.x] += sred[k][threadIdx.x + sft];
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        const double tot[4] = {sred[0][0], sred[1][0], sred[2][0], sred[3][0]};
        double loss = 0.0;
        if (gt) loss += (double)w_l1 * tot[0] / (double)n;
        double tv = 0.0;
        if (H > 1) tv += tot[1] / ((double)C * (H - 1) * W);
        if (W > 1) tv += tot[2] / ((double)C * H * (W - 1));
        loss += (double)w_tv * 0.5 * tv;
        if (depth) loss += (double)w_depth * tot[3] / (double)HW;
        *out_loss = (float)loss;
        *counter = 0;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
pixel_loss_backward_kernel(int C, int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                           const float* __restrict__ depth, const float* __restrict__ depth_gt, float w_l1,
                           float w_tv, float w_depth, const float* __restrict__ grad_out,
                           float* __restrict__ d_img, float* __restrict__ d_depth) {
    const size_t HW = (size_t)H * W, n = (size_t)C * HW;
    const int rows = C * H;
    const float go = grad_out ? *grad_out : 1.0f;
    const float k_l1 = gt ? go * w_l1 / (float)n : 0.f;
    const float k_ty = H > 1 ? go * w_tv * 0.5f / ((float)C * (float)(H - 1) * (float)W) : 0.f;
    const float k_tx = W > 1 ? go * w_tv * 0.5f / ((float)C * (float)H * (float)(W - 1)) : 0.f;
    const float k_d = go * w_depth * 2.0f / (float)HW;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int y = row % H;
        const bool down = y + 1 < H, up = y > 0;
        const bool drow = d_depth != nullptr && row < H;
        const size_t base = (size_t)row * W;
        const float* r = img + base;
        for (int x = threadIdx.x; x < W; x += LOSS_THREADS) {
            const float v = r[x];
            float g = 0.f;
            if (gt) g += k_l1 * sgn(v - gt[base + x]);
            // d/dv of |img[y+1]-v| is -sign(.), of |v-img[y-1]| is +sign(.)
            if (down) g -= k_ty * sgn(r[x + W] - v);
            if (up) g += k_ty * sgn(v - r[x - W]);
            if (x + 1 < W) g -= k_tx * sgn(r[x + 1] - v);
            if (x > 0) g += k_tx * sgn(v - r[x - 1]);
            d_img[base + x] = g;
            if (drow) d_depth[base + x] = depth ? k_d * (depth[base + x] - depth_gt[base + x]) : 0.f;
        }
    }
}

static unsigned loss_blocks(size_t rows) {   // one block per image row, capped at a few resident waves
    size_t b = rows;
    if (b > LOSS_MAX_BLOCKS) b = LOSS_MAX_BLOCKS;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace w3d

using namespace w3d;

extern "C" size_t wast3d_pixel_loss_scratch_bytes(void) { return (size_t)LOSS_MAX_BLOCKS * 4 * sizeof(double) + 128; }

extern "C" int wast3d_pixel_loss_forward(int C, int H, int W, const float* img, const float* gt, const float* depth,
                                         const float* depth_gt, float w_l1, float w_tv, float w_depth,
                                         void* scratch, float* out_loss, void* stream_v) {
    if (C < 1 || H < 1 || W < 1 || !img || !scratch || !out_loss || ((depth == nullptr) != (depth_gt == nullptr)))
        return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    // scratch: [counter (zero between calls; the kernel resets it) | pad to 128 | partials]
    unsigned* counter = (unsigned*)scratch;
    double* partials = (double*)((char*)scratch + 128);
    pixel_loss_forward_kernel<<<loss_blocks((size_t)C * H), LOSS_THREADS, 0, s>>>(
        C, H, W, img, gt, depth, depth_gt, w_l1, w_tv, w_depth, partials, counter, out_loss);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}

extern "C" int wast3d_pixel_loss_backward(int C, int H, int W, const float* img, const float* gt, const float* depth,
                                          const float* depth_gt, float w_l1, float w_tv, float w_depth,
                                          const float* grad_out, float* d_img, float* d_depth, void* stream_v) {
    if (C < 1 || H < 1 || W < 1 || !img || !d_img || ((depth == nullptr) != (depth_gt == nullptr)))
        return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    pixel_loss_backward_kernel<<<loss_blocks((size_t)C * H), LOSS_THREADS, 0, s>>>(
        C, H, W, img, gt, depth, depth_gt, w_l1, w_tv, w_depth, grad_out, d_img, d_depth);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
