// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Thin extern "C" wrapper around the *unmodified* reference CUDA rasteriser and
// simple-knn so that the parity tests can run the real reference on the same
// GPU, same inputs, through plain pointers (no torch needed -> compiles in ~1 min).
// The reference sources are compiled where they lie under /root/reference by
// oracle/build_ref.sh; only this wrapper lives in the repo.  Outputs go to
// oracle/_ref/ (git-ignored, but shipped to the GPU box).
//
// Wraps:
//   CudaRasterizer::Rasterizer::forward   submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer_impl.cu:198-341
//   CudaRasterizer::Rasterizer::backward  .../rasterizer_impl.cu:345-446
//   CudaRasterizer::Rasterizer::markVisible .../rasterizer_impl.cu:141-153
//   SimpleKNN::knn                        submodules/simple-knn/simple_knn.cu:185-221
//
// The host-side allocation pattern mirrors rasterize_points.cu:27-33,69-80
// (three growable byte buffers) with cudaMalloc instead of torch tensors.
#include <cstdint>
#include <cfloat>
#include <cstdio>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"
#include "simple_knn.h"

namespace {
struct RefCtx {
    char* geom = nullptr;  size_t geom_bytes = 0;
    char* bin = nullptr;   size_t bin_bytes = 0;
    char* img = nullptr;   size_t img_bytes = 0;
    int P = 0, W = 0, H = 0, R = 0;
    bool async = false;  // true: no cudaDeviceSynchronize after the call (timing with CUDA events)
};

// Grow-only, like torch's `resize_` on the three byte tensors (rasterize_points.cu:27-33): a call that
// needs no more than the buffer already holds does not touch the allocator (with torch's caching
// allocator the reference pays no cudaMalloc in steady state either).
std::function<char*(size_t)> grow(char** p, size_t* n) {
    return [p, n](size_t N) -> char* {
        if (*p == nullptr || N > *n) {
            if (*p) cudaFree(*p);
            cudaMalloc((void**)p, N ? N : 1);
            *n = N;
        }
        return *p;
    };
}
}  // namespace

extern "C" {

void* ref_ctx_create() { return new RefCtx(); }

void ref_ctx_set_async(void* h, int on) { ((RefCtx*)h)->async = on != 0; }

void ref_ctx_destroy(void* h) {
    RefCtx* c = (RefCtx*)h;
    if (!c) return;
    if (c->geom) cudaFree(c->geom);
    if (c->bin) cudaFree(c->bin);
    if (c->img) cudaFree(c->img);
    delete c;
}

// Returns num_rendered (>=0) or -1 on CUDA error.  All pointers are device pointers.
int ref_raster_forward(void* h, int P, int D, int M, const float* background, int W, int H,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* opacities, const float* scales, float scale_modifier,
                       const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                       const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                       int prefiltered, float* out_color, float* out_depth, float* sampling_offsets,
                       int* radii) {
    RefCtx* c = (RefCtx*)h;
    c->P = P; c->W = W; c->H = H;
    int R = CudaRasterizer::Rasterizer::forward(
        grow(&c->geom, &c->geom_bytes), grow(&c->bin, &c->bin_bytes), grow(&c->img, &c->img_bytes),
        P, D, M, background, W, H, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
        rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
        prefiltered != 0, out_color, out_depth, sampling_offsets, radii, false);
    c->R = R;
    if (!c->async && cudaDeviceSynchronize() != cudaSuccess) return -1;
    return R;
}

int ref_raster_backward(void* h, int P, int D, int M, int R, const float* background, int W, int H,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                        const float* dL_dpix, const float* dL_ddepth, float* dL_dmean2D,
                        float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                        float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                        float* dL_dcamViewDepth, float* sampling_offsets) {
    RefCtx* c = (RefCtx*)h;
    CudaRasterizer::Rasterizer::backward(
        P, D, M, R, background, W, H, means3D, shs, colors_precomp, scales, scale_modifier,
        rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii,
        c->geom, c->bin, c->img, dL_dpix, dL_ddepth, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
        dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dcamViewDepth, false,
        sampling_offsets);
    if (c->async) return cudaGetLastError() == cudaSuccess ? 0 : -1;
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

// Copies the reference's internal per-Gaussian and per-pixel state out to caller
// device buffers (any may be null) so K1 / K6 can be checked field by field.
// Layouts follow GeometryState/ImageState::fromChunk (rasterizer_impl.cu:155-179).
int ref_raster_export_state(void* h, float* depths, float* means2D, float* cov3D,
                            float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                            unsigned char* clamped, float* final_T, uint32_t* n_contrib,
                            uint32_t* point_list, uint32_t* ranges /* [tiles,2] */) {
    RefCtx* c = (RefCtx*)h;
    char* g = c->geom;
    CudaRasterizer::GeometryState gs = CudaRasterizer::GeometryState::fromChunk(g, c->P);
    char* i = c->img;
    CudaRasterizer::ImageState is = CudaRasterizer::ImageState::fromChunk(i, (size_t)c->W * c->H);
    const size_t P = c->P, N = (size_t)c->W * c->H;
    auto cp = [](void* d, const void* s, size_t n) {
        if (d) cudaMemcpy(d, s, n, cudaMemcpyDeviceToDevice);
    };
    cp(depths, gs.depths, 4 * P);
    cp(means2D, gs.means2D, 8 * P);
    cp(cov3D, gs.cov3D, 24 * P);
    cp(conic_opacity, gs.conic_opacity, 16 * P);
    cp(rgb, gs.rgb, 12 * P);
    cp(tiles_touched, gs.tiles_touched, 4 * P);
    cp(clamped, gs.clamped, 3 * P);
    cp(final_T, is.accum_alpha, 4 * N);
    cp(n_contrib, is.n_contrib, 4 * N);
    cp(ranges, is.ranges, 8 * (size_t)((c->W + 15) / 16) * (size_t)((c->H + 15) / 16));
    if (point_list && c->R > 0) {
        char* b = c->bin;
        CudaRasterizer::BinningState bs = CudaRasterizer::BinningState::fromChunk(b, c->R);
        cp(point_list, bs.point_list, 4 * (size_t)c->R);
    }
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

int ref_knn_dist2(int P, float* points, float* mean_dists) {
    SimpleKNN::knn(P, (float3*)points, mean_dists);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
