/* wast3d_b200_staged.h — entry points STAGED for the next round: they compile and are exported, but nothing in
 * the product calls them and they have not run on a GPU yet (their GPU test is gated by WAST3D_STAGED=1).
 * They are NOT part of the drop-in ABI of include/wast3d_b200.h.
 *
 * View-parallel exchange of the SH features through 16-byte colour records (DESIGN.md §6, round-2 plan; the
 * algebra is pinned on the CPU: oracle/sh_records.py, tests/test_sh_records_oracle.py).  There is no
 * reference counterpart (the reference is single-GPU); the gradient rebuilt is the one of
 * submodules/diff-gaussian-rasterization/cuda_rasterizer/backward.cu:20-139 summed over the views, the update is
 * torch.optim.Adam as configured in scene/gaussian_model.py:154-163. */
#ifndef WAST3D_B200_STAGED_H
#define WAST3D_B200_STAGED_H
#include "wast3d_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* After wast3d_raster_backward* of one view: out_records [P,4] = {dL/dRGB with clamped channels zeroed, visible}
 * read from that call's geometry buffer (K7's gradient record). */
int wast3d_staged_colour_records(int P, const int* radii, const void* geom_buffer, float* out_records, void* stream);
/* records: HOST array of `views` DEVICE pointers ([P,4] each; peer-mapped pointers allowed), campos_host
 * [views,3] HOST floats, xyz [P,3] the positions the views were rendered with.  Applies Adam in place to
 * _features_dc [P,1,3] (dc) and _features_rest [P,M-1,3] (rest; NULL when M == 1) with the gradient
 * grad_scale * sum_views basis(normalize(xyz - campos_v)) * record_v, views summed in array order. */
int wast3d_staged_sh_adam_from_records(int P, int D, int M, int views, const float* const* records,
                                       const float* campos_host, const float* xyz, float grad_scale,
                                       const wast3d_adam_group* dc, const wast3d_adam_group* rest, void* stream);
/* CPU emulation of the kernel behind wast3d_staged_sh_adam_from_records for tests without a GPU: the same
 * per-Gaussian accumulation statements compiled for the host, Adam with IEEE sqrt / division.  HOST pointers. */
int wast3d_staged_sh_adam_host_emulation(int P, int D, int M, int views, const float* const* records,
                                         const float* campos_host, const float* xyz, float grad_scale,
                                         const wast3d_adam_group* dc, const wast3d_adam_group* rest);
#ifdef __cplusplus
}
#endif
#endif
