// C-ABI housekeeping: status strings, device check, last CUDA error, state export for tests.
#include "common.cuh"
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

namespace w3d {

static thread_local char g_last_cuda_error[512] = "";

void set_last_cuda_error(cudaError_t e, const char* file, int line) {
    const char* base = strrchr(file, '/');
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s (%s:%d)", cudaGetErrorString(e),
             base ? base + 1 : file, line);
    cudaGetLastError();  // clear the sticky launch error, if it was one
}

const uint32_t* point_list_ptr(const BinningState& b, uint32_t num_tiles);

// ---- launch counter + event profiling ------------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
static unsigned g_prof_mask = 0;
struct ProfRec { int slot; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::mutex g_prof_mu;
static const char* const kSlotNames[PS_COUNT] = {
    "preprocess", "depth_sort", "scan", "emit_instances", "tile_sort", "tile_ranges", "render_forward",
    "backward_zero", "render_backward", "gaussian_backward", "knn", "match", "adam"};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

ProfScope::ProfScope(int slot_, cudaStream_t s) : slot(slot_), stream(s), start(nullptr), on(false) {
    if (!((g_prof_mask >> slot) & 1u)) return;
    if (cudaEventCreate(&start) != cudaSuccess) return;
    cudaEventRecord(start, stream);
    on = true;
}
ProfScope::~ProfScope() {
    if (!on) return;
    cudaEvent_t stop;
    if (cudaEventCreate(&stop) != cudaSuccess) { cudaEventDestroy(start); return; }
    cudaEventRecord(stop, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back({slot, start, stop});
}

__global__ void export_geom_kernel(int P, const float4* __restrict__ rec,
                                   const uint32_t* __restrict__ tiles_touched_in,
                                   const uint8_t* __restrict__ clamped_in, float* depths,
                                   float* means2D, float* conic_opacity, float* rgb,
                                   uint32_t* tiles_touched, unsigned char* clamped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = (clamped_in[i] & 8u) != 0;  // bit 3: K1 wrote a render record (radius > 0)
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 r0 = vis ? rec[3 * i] : z, r1 = vis ? rec[3 * i + 1] : z, r2 = vis ? rec[3 * i + 2] : z;
    if (depths) depths[i] = r0.z;
    if (means2D) {
        means2D[2 * i] = r0.x;
        means2D[2 * i + 1] = r0.y;
    }
    if (conic_opacity) {
        conic_opacity[4 * i] = r1.x;
        conic_opacity[4 * i + 1] = r1.y;
        conic_opacity[4 * i + 2] = r1.z;
        conic_opacity[4 * i + 3] = r1.w;
    }
    if (rgb) {
        rgb[3 * i] = r2.x;
        rgb[3 * i + 1] = r2.y;
        rgb[3 * i + 2] = r2.z;
    }
    if (tiles_touched) tiles_touched[i] = tiles_touched_in[i];
    if (clamped) {
        const unsigned b = vis ? clamped_in[i] : 0;
        clamped[3 * i] = b & 1u;
        clamped[3 * i + 1] = (b >> 1) & 1u;
        clamped[3 * i + 2] = (b >> 2) & 1u;
    }
}

}  // namespace w3d

using namespace w3d;

extern "C" const char* wast3d_strerror(int status) {
    switch (status) {
        case WAST3D_OK: return "ok";
        case WAST3D_ERR_INVALID_ARGUMENT: return "invalid argument";
        case WAST3D_ERR_CUDA: return g_last_cuda_error[0] ? g_last_cuda_error : "CUDA error";
        case WAST3D_ERR_ALLOC: return "scratch allocation callback failed";
        case WAST3D_ERR_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU fallback)";
        case WAST3D_ERR_OVERFLOW: return "instance count overflows 31 bits";
        case WAST3D_ERR_NON_RGB: return "For non-RGB, provide precomputed Gaussian colors!";
        case WAST3D_ERR_STALE_PROJECTION:
            return "preprojected forward: the sampling offsets exceed the bounds the projection assumed";
        default: return "unknown status";
    }
}

extern "C" int wast3d_abi_version(void) { return WAST3D_ABI_VERSION; }

extern "C" int wast3d_device_check(int ordinal) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || ordinal < 0 || ordinal >= n) {
        cudaGetLastError();
        return WAST3D_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    W3D_CUDA_TRY(cudaGetDeviceProperties(&prop, ordinal));
    return prop.major == 10 ? WAST3D_OK : WAST3D_ERR_NO_DEVICE;
}

extern "C" int wast3d_raster_export_state(const wast3d_raster_params* prm, int num_rendered,
                                          const void* geom_buffer, const void* binning_buffer,
                                          const void* img_buffer, float* depths, float* means2D,
                                          float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                                          unsigned char* clamped, uint32_t* point_list,
                                          uint32_t* ranges, void* stream_v) {
    if (!prm || !geom_buffer || !img_buffer) return WAST3D_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream_v;
    const int P = prm->P, W = prm->width, H = prm->height;
    if (P == 0) return WAST3D_OK;
    const dim3 grid((W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y, 1);
    const uint32_t num_tiles = grid.x * grid.y;
    GeomState g = GeomState::carve(const_cast<void*>(geom_buffer), P, nullptr);
    ImageState im = ImageState::carve(const_cast<void*>(img_buffer), (size_t)W * H, num_tiles, nullptr);
    export_geom_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, g.tiles_touched, g.clamped, depths,
                                                       means2D, conic_opacity, rgb, tiles_touched, clamped);
    W3D_AFTER_LAUNCH(s, false);
    if (point_list && num_rendered > 0) {
        if (!binning_buffer) return WAST3D_ERR_INVALID_ARGUMENT;
        BinningState bn = BinningState::carve(const_cast<void*>(binning_buffer), (size_t)num_rendered, nullptr);
        W3D_CUDA_TRY(cudaMemcpyAsync(point_list, point_list_ptr(bn, num_tiles),
                                     (size_t)num_rendered * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    if (ranges)
        W3D_CUDA_TRY(cudaMemcpyAsync(ranges, im.ranges, (size_t)num_tiles * sizeof(uint2),
                                     cudaMemcpyDeviceToDevice, s));
    return WAST3D_OK;
}

extern "C" int wast3d_profile_set(unsigned slot_mask) {
    g_prof_mask = slot_mask;
    return WAST3D_OK;
}
extern "C" int wast3d_profile_slots(void) { return PS_COUNT; }
extern "C" const char* wast3d_profile_slot_name(int slot) {
    return (slot >= 0 && slot < PS_COUNT) ? kSlotNames[slot] : "";
}
// Synchronises the recorded events, ADDS their durations (ms) and counts into the caller's arrays
// of wast3d_profile_slots() entries, and clears the records.
extern "C" int wast3d_profile_read(double* ms_per_slot, unsigned long long* scopes_per_slot) {
    std::vector<ProfRec> recs;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        recs.swap(g_prof_recs);
    }
    int rc = WAST3D_OK;
    for (auto& r : recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
            rc = WAST3D_ERR_CUDA;
            cudaGetLastError();
        } else {
            if (ms_per_slot) ms_per_slot[r.slot] += (double)ms;
            if (scopes_per_slot) scopes_per_slot[r.slot] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    return rc;
}
extern "C" unsigned long long wast3d_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}
