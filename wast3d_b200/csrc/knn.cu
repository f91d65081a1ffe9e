// Exact 3-nearest-neighbour mean squared distance on a uniform grid (sm_100a).
//
// Replaces SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-221: Morton sort, 1024-point
// boxes, pruned scan of every box) and its two blocking device->host copies (:193-200).
// Semantics kept bit for bit:
//   * neighbours are the 3 smallest squared distances to OTHER points; self is excluded by
//     index, not by distance, so duplicates give 0 (simple_knn.cu:158,177);
//   * d^2 is evaluated as fma(dz,dz, fma(dx,dx, dy*dy)) with d = other - self, which is what
//     nvcc emits for simple_knn.cu:134-135 (checked against the reference's sm_100 build:
//     of the 18 fma/sum orderings only this one reproduces its output bit for bit);
//   * result = (b0 + b1 + b2) / 3.0f with FLT_MAX for missing neighbours (:182, P < 4).
// Extension: the indices of the three neighbours, ties broken towards the lowest index.
//
// Algorithm: bounding box (device side, no host sync) -> cell size so that ~3 points share a
// cell -> points sorted by cell id (radix passes from scan_sort.cu) -> one thread per point, in
// sorted order, scans growing cubes of cells until the 3rd best distance is strictly inside the
// scanned region.  Queries that do not finish within KNN_MAX_RING rings (far outliers) are
// finished by a brute-force kernel, one block per query.
#include "common.cuh"
#include <cfloat>

namespace w3d {

constexpr int KNN_MAX_RING = 6;
constexpr int KNN_CELL_BITS = 21;        // <= 2^21 cells, 3 radix passes of 7 bits
constexpr int KNN_MAX_AXIS = 1024;

struct KnnGrid {
    float3 lo;
    float h, inv_h;
    int3 dim;
    uint32_t n_unfinished;
    uint32_t pad[3];
};

struct KnnScratch {
    KnnGrid* grid;
    float* bbox;           // [6] min xyz, max xyz (as ordered ints during the reduction)
    uint32_t* cell_key[2]; // [P]
    uint32_t* order[2];    // [P]
    float4* sorted;        // [P] xyz + original index bits
    uint32_t* cell_start;  // [2^21 + 1]
    uint32_t* cell_end;    // [2^21 + 1]
    uint32_t* rs_hist;
    uint32_t* scan_scratch;
    uint32_t* unfinished;  // [P]
    static KnnScratch carve(void* chunk, size_t P, size_t* bytes) {
        Carver c(chunk);
        KnnScratch s;
        s.grid = c.take<KnnGrid>(1);
        s.bbox = c.take<float>(8);
        s.cell_key[0] = c.take<uint32_t>(P);
        s.cell_key[1] = c.take<uint32_t>(P);
        s.order[0] = c.take<uint32_t>(P);
        s.order[1] = c.take<uint32_t>(P);
        s.sorted = c.take<float4>(P);
        s.cell_start = c.take<uint32_t>((1u << KNN_CELL_BITS) + 1);
        s.cell_end = c.take<uint32_t>((1u << KNN_CELL_BITS) + 1);
        s.rs_hist = c.take<uint32_t>(rs_hist_words(P));
        s.scan_scratch = c.take<uint32_t>(scan_scratch_words(rs_hist_words(P)));
        s.unfinished = c.take<uint32_t>(P);
        if (bytes) *bytes = c.bytes();
        return s;
    }
};

// order-preserving float <-> int so atomicMin/atomicMax work on floats
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void knn_bbox_init_kernel(int* bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = f2ord(FLT_MAX);
    else if (threadIdx.x < 6) bbox[threadIdx.x] = f2ord(-FLT_MAX);
}

__global__ void __launch_bounds__(256)
knn_bbox_kernel(int P, const float* __restrict__ pts, int* __restrict__ bbox) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bbox[k], f2ord(lo[k]));
            atomicMax(&bbox[3 + k], f2ord(hi[k]));
        }
    }
}

__global__ void knn_grid_kernel(int P, const int* __restrict__ bbox, KnnGrid* __restrict__ grid) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lo[3], ext[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = ord2f(bbox[k]);
        ext[k] = fmaxf(ord2f(bbox[3 + k]) - lo[k], 0.f);
        if (!(ext[k] < FLT_MAX)) ext[k] = 0.f;  // NaN / inf guards: collapse the axis
    }
    // cell size from the volume of the non-degenerate axes, ~3 points per cell
    double vol = 1.0;
    int nd = 0;
    for (int k = 0; k < 3; ++k)
        if (ext[k] > 0.f) { vol *= ext[k]; ++nd; }
    float h = 1.f;
    if (nd > 0) h = (float)pow(vol * 3.0 / (double)(P > 0 ? P : 1), 1.0 / nd);
    if (!(h > 0.f)) h = 1.f;
    int dim[3];
    for (int iter = 0; iter < 64; ++iter) {
        unsigned long long cells = 1;
        for (int k = 0; k < 3; ++k) {
            float n = floorf(ext[k] / h) + 1.f;
            dim[k] = n > (float)KNN_MAX_AXIS ? KNN_MAX_AXIS + 1 : (int)n;
            cells *= (unsigned long long)dim[k];
        }
        if (dim[0] <= KNN_MAX_AXIS && dim[1] <= KNN_MAX_AXIS && dim[2] <= KNN_MAX_AXIS &&
            cells <= (1ull << KNN_CELL_BITS))
            break;
        h *= 1.26f;  // ~ 2x fewer cells per step
    }
    grid->lo = make_float3(lo[0], lo[1], lo[2]);
    grid->h = h;
    grid->inv_h = 1.0f / h;
    grid->dim = make_int3(dim[0], dim[1], dim[2]);
    grid->n_unfinished = 0;
}

__device__ __forceinline__ int3 cell_of(const KnnGrid& g, float x, float y, float z) {
    int cx = (int)((x - g.lo.x) * g.inv_h), cy = (int)((y - g.lo.y) * g.inv_h),
        cz = (int)((z - g.lo.z) * g.inv_h);
    cx = min(max(cx, 0), g.dim.x - 1);
    cy = min(max(cy, 0), g.dim.y - 1);
    cz = min(max(cz, 0), g.dim.z - 1);
    return make_int3(cx, cy, cz);
}

__global__ void __launch_bounds__(256)
knn_cell_key_kernel(int P, const float* __restrict__ pts, const KnnGrid* __restrict__ grid,
                    uint32_t* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const KnnGrid g = *grid;
    const int3 c = cell_of(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    keys[i] = (uint32_t)((c.z * g.dim.y + c.y) * g.dim.x + c.x);
}

__global__ void __launch_bounds__(256)
knn_gather_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order,
                  const uint32_t* __restrict__ sorted_keys, float4* __restrict__ sorted,
                  uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_end) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const uint32_t i = order[k];
    sorted[k] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __uint_as_float(i));
    const uint32_t c = sorted_keys[k];
    if (k == 0 || sorted_keys[k - 1] != c) cell_start[c] = k;
    if (k == P - 1 || sorted_keys[k + 1] != c) cell_end[c] = k + 1;
}

struct Best3 {
    float d[3];
    uint32_t id[3];
};
// insert keeping (distance, index) lexicographic order
__device__ __forceinline__ void best3_insert(Best3& b, float dist, uint32_t id) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool better = dist < b.d[j] || (dist == b.d[j] && id < b.id[j]);
        if (better) {
            const float td = b.d[j]; const uint32_t ti = b.id[j];
            b.d[j] = dist; b.id[j] = id;
            dist = td; id = ti;
        }
    }
}
__device__ __forceinline__ float knn_dist2(float px, float py, float pz, float qx, float qy, float qz) {
    const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__global__ void __launch_bounds__(128)
knn_query_kernel(int P, const float4* __restrict__ sorted, const uint32_t* __restrict__ cell_start,
                 const uint32_t* __restrict__ cell_end, KnnGrid* __restrict__ grid,
                 float* __restrict__ out, int32_t* __restrict__ out_idx, uint32_t* __restrict__ unfinished) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const KnnGrid g = *grid;
    const float4 q = sorted[k];
    const uint32_t self = __float_as_uint(q.w);
    const int3 c = cell_of(g, q.x, q.y, q.z);
    Best3 b;
#pragma unroll
    for (int j = 0; j < 3; ++j) { b.d[j] = FLT_MAX; b.id[j] = 0xFFFFFFFFu; }
    bool done = false;
    for (int r = 0; r <= KNN_MAX_RING && !done; ++r) {
        const int z0 = max(c.z - r, 0), z1 = min(c.z + r, g.dim.z - 1);
        const int y0 = max(c.y - r, 0), y1 = min(c.y + r, g.dim.y - 1);
        const int x0 = max(c.x - r, 0), x1 = min(c.x + r, g.dim.x - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                // cells with Chebyshev distance exactly r from c: on a z or y face of the shell the
                // whole x row is new, otherwise only its two ends are
                const bool face = (z == c.z - r) || (z == c.z + r) || (y == c.y - r) || (y == c.y + r);
                const int step = (face || r == 0) ? 1 : 2 * r;
                for (int x = face ? x0 : c.x - r; x <= (face ? x1 : c.x + r); x += step) {
                    if (x < 0 || x >= g.dim.x) continue;
                    const uint32_t cell = (uint32_t)((z * g.dim.y + y) * g.dim.x + x);
                    const uint32_t s = cell_start[cell], e = cell_end[cell];
                    for (uint32_t t = s; t < e; ++t) {
                        const float4 p = sorted[t];
                        const uint32_t pid = __float_as_uint(p.w);
                        const float dist = knn_dist2(p.x, p.y, p.z, q.x, q.y, q.z);
                        // cheap reject: strictly farther than the current third best can never enter
                        // (equal distances still go through the index tie-break)
                        if (dist > b.d[2] || pid == self) continue;
                        best3_insert(b, dist, pid);
                    }
                }
            }
        // Distance from q to the nearest face of the scanned cube that still has cells behind it.
        const float inf = __int_as_float(0x7f800000);
        float safe = inf;
        if (c.x - r > 0) safe = fminf(safe, q.x - (g.lo.x + (float)(c.x - r) * g.h));
        if (c.x + r < g.dim.x - 1) safe = fminf(safe, (g.lo.x + (float)(c.x + r + 1) * g.h) - q.x);
        if (c.y - r > 0) safe = fminf(safe, q.y - (g.lo.y + (float)(c.y - r) * g.h));
        if (c.y + r < g.dim.y - 1) safe = fminf(safe, (g.lo.y + (float)(c.y + r + 1) * g.h) - q.y);
        if (c.z - r > 0) safe = fminf(safe, q.z - (g.lo.z + (float)(c.z - r) * g.h));
        if (c.z + r < g.dim.z - 1) safe = fminf(safe, (g.lo.z + (float)(c.z + r + 1) * g.h) - q.z);
        if (safe == inf) {
            done = true;  // the whole grid has been scanned
        } else {
            // shrink for rounding in cell assignment / face positions / d^2 evaluation
            safe = safe * (1.0f - 1e-4f) - 1e-3f * g.h;
            if (safe > 0.f && b.d[2] < safe * safe) done = true;
        }
    }
    if (!done) {
        const uint32_t slot = atomicAdd(&grid->n_unfinished, 1u);
        unfinished[slot] = (uint32_t)k;
        return;
    }
    out[self] = (b.d[0] + b.d[1] + b.d[2]) / 3.0f;
    if (out_idx) {
        out_idx[3 * self + 0] = (int32_t)b.id[0];
        out_idx[3 * self + 1] = (int32_t)b.id[1];
        out_idx[3 * self + 2] = (int32_t)b.id[2];
    }
}

// One block per unfinished query: brute force over all points, then a block-wide merge.
__global__ void __launch_bounds__(256)
knn_bruteforce_kernel(int P, const float4* __restrict__ sorted, const KnnGrid* __restrict__ grid,
                      const uint32_t* __restrict__ unfinished, float* __restrict__ out,
                      int32_t* __restrict__ out_idx) {
    __shared__ Best3 s_best[256];
    const uint32_t n = grid->n_unfinished;
    for (uint32_t u = blockIdx.x; u < n; u += gridDim.x) {
        const float4 q = sorted[unfinished[u]];
        const uint32_t self = __float_as_uint(q.w);
        Best3 b;
#pragma unroll
        for (int j = 0; j < 3; ++j) { b.d[j] = FLT_MAX; b.id[j] = 0xFFFFFFFFu; }
        for (int t = threadIdx.x; t < P; t += blockDim.x) {
            const float4 p = sorted[t];
            const uint32_t pid = __float_as_uint(p.w);
            if (pid == self) continue;
            best3_insert(b, knn_dist2(p.x, p.y, p.z, q.x, q.y, q.z), pid);
        }
        s_best[threadIdx.x] = b;
        __syncthreads();
        for (int off = 128; off >= 1; off >>= 1) {
            if (threadIdx.x < off) {
                Best3 mine = s_best[threadIdx.x];
                const Best3 o = s_best[threadIdx.x + off];
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (o.id[j] != 0xFFFFFFFFu || o.d[j] < FLT_MAX) best3_insert(mine, o.d[j], o.id[j]);
                s_best[threadIdx.x] = mine;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const Best3 r = s_best[0];
            out[self] = (r.d[0] + r.d[1] + r.d[2]) / 3.0f;
            if (out_idx) {
                out_idx[3 * self + 0] = (int32_t)r.id[0];
                out_idx[3 * self + 1] = (int32_t)r.id[1];
                out_idx[3 * self + 2] = (int32_t)r.id[2];
            }
        }
        __syncthreads();
    }
}

}  // namespace w3d

using namespace w3d;

extern "C" size_t wast3d_knn_scratch_bytes(int P) {
    size_t bytes = 0;
    KnnScratch::carve(nullptr, (size_t)(P > 0 ? P : 0), &bytes);
    return bytes;
}

extern "C" int wast3d_knn_dist2(int P, const float* points, float* mean_dist2, int32_t* nn_index,
                                void* scratch, size_t scratch_bytes, void* stream_v) {
    if (P < 0) return WAST3D_ERR_INVALID_ARGUMENT;
    if (P == 0) return WAST3D_OK;
    if (!points || !mean_dist2 || !scratch || scratch_bytes < wast3d_knn_scratch_bytes(P))
        return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    KnnScratch k = KnnScratch::carve(scratch, P, nullptr);
    const int nb = (P + 255) / 256;
    ProfScope ps(PS_KNN, s);

    knn_bbox_init_kernel<<<1, 32, 0, s>>>((int*)k.bbox);
    W3D_AFTER_LAUNCH(s, false);
    knn_bbox_kernel<<<min(nb, 148 * 8), 256, 0, s>>>(P, points, (int*)k.bbox);
    W3D_AFTER_LAUNCH(s, false);
    knn_grid_kernel<<<1, 32, 0, s>>>(P, (const int*)k.bbox, k.grid);
    W3D_AFTER_LAUNCH(s, false);
    knn_cell_key_kernel<<<nb, 256, 0, s>>>(P, points, k.grid, k.cell_key[0]);
    W3D_AFTER_LAUNCH(s, false);
    // 21-bit cell ids: three stable 7-bit passes, values start as iota
    int st = radix_pass_u32(k.cell_key[0], nullptr, k.cell_key[1], k.order[1], P, 0, 7, k.rs_hist, k.scan_scratch, s, false);
    if (st) return st;
    st = radix_pass_u32(k.cell_key[1], k.order[1], k.cell_key[0], k.order[0], P, 7, 7, k.rs_hist, k.scan_scratch, s, false);
    if (st) return st;
    st = radix_pass_u32(k.cell_key[0], k.order[0], k.cell_key[1], k.order[1], P, 14, 7, k.rs_hist, k.scan_scratch, s, false);
    if (st) return st;
    const size_t cells = (1u << KNN_CELL_BITS) + 1;
    W3D_CUDA_TRY(cudaMemsetAsync(k.cell_start, 0, cells * sizeof(uint32_t), s));
    W3D_CUDA_TRY(cudaMemsetAsync(k.cell_end, 0, cells * sizeof(uint32_t), s));
    knn_gather_kernel<<<nb, 256, 0, s>>>(P, points, k.order[1], k.cell_key[1], k.sorted, k.cell_start, k.cell_end);
    W3D_AFTER_LAUNCH(s, false);
    knn_query_kernel<<<(P + 127) / 128, 128, 0, s>>>(P, k.sorted, k.cell_start, k.cell_end, k.grid,
                                                      mean_dist2, nn_index, k.unfinished);
    W3D_AFTER_LAUNCH(s, false);
    knn_bruteforce_kernel<<<148 * 2, 256, 0, s>>>(P, k.sorted, k.grid, k.unfinished, mean_dist2, nn_index);
    W3D_AFTER_LAUNCH(s, false);
    return WAST3D_OK;
}
