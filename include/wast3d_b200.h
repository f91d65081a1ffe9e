/*
 * wast3d_b200 — C ABI of the B200-native (sm_100a) hot path of WaSt3D's style-transfer
 * optimisation step.  Plain pointers and sizes only; no torch types, no C++ exceptions.
 * Every function returns an int status (WAST3D_OK == 0); wast3d_strerror() names it.
 * All data pointers are DEVICE pointers unless the name ends in `_host`.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what the
 * reference launches on: `<<<grid,block>>>` with no stream argument, forward.cu:408,450).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repo root, facebookresearch/WaSt3D @ 786e4e1e).
 */
#ifndef WAST3D_B200_H_
#define WAST3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAST3D_ABI_VERSION 7

enum wast3d_status {
    WAST3D_OK = 0,
    WAST3D_ERR_INVALID_ARGUMENT = 1, /* bad shape / null pointer / unsupported combination   */
    WAST3D_ERR_CUDA = 2,             /* a CUDA runtime call or kernel failed                  */
    WAST3D_ERR_ALLOC = 3,            /* an allocation callback returned NULL                  */
    WAST3D_ERR_NO_DEVICE = 4,        /* no sm_100 device visible (there is no CPU fallback)   */
    WAST3D_ERR_OVERFLOW = 5,         /* instance count does not fit 31 bits                   */
    WAST3D_ERR_NON_RGB = 6,          /* reserved: non-RGB channels need precomputed colours   */
    WAST3D_ERR_STALE_PROJECTION = 7  /* preprojected forward: sampling offsets outside the bounds the projection assumed */
};

const char* wast3d_strerror(int status);
int wast3d_abi_version(void);
/* Returns WAST3D_OK iff device `ordinal` exists and is compute capability 10.x. */
int wast3d_device_check(int ordinal);

/* Growable scratch buffers.  Replaces `std::function<char*(size_t)>`
 * (submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:32-34) and the
 * torch `resize_` lambdas of rasterize_points.cu:27-33.  Must return a device pointer to at
 * least `bytes` bytes, 128-byte aligned (torch allocations are), or NULL on failure. */
typedef void* (*wast3d_alloc_fn)(size_t bytes, void* user);

/* Inputs shared by forward and backward — the union of the argument lists of
 * CudaRasterizer::Rasterizer::forward / backward (rasterizer.h:30-92).  "Not provided"
 * tensors are NULL, exactly like the size-0 tensors of the reference whose data_ptr is null
 * (diff_gaussian_rasterization/__init__.py:213-223; forward.cu:205,241). */
typedef struct wast3d_raster_params {
    int P;               /* number of Gaussians                                            */
    int D;               /* active SH degree 0..3                                          */
    int M;               /* SH coefficients per Gaussian in `shs` (0 if shs == NULL)       */
    int width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int prefiltered;     /* reference traps if a culled point is seen with this set        */
    int debug;           /* synchronise + check after every stage (auxiliary.h:166-173)    */
    const float* background;      /* [3]                                                   */
    const float* means3D;         /* [P,3]                                                 */
    const float* shs;             /* [P,M,3] or NULL                                       */
    const float* colors_precomp;  /* [P,3]  or NULL                                        */
    const float* opacities;       /* [P,1]                                                 */
    const float* scales;          /* [P,3]  or NULL                                        */
    const float* rotations;       /* [P,4]  or NULL (NOT normalised here, forward.cu:127)  */
    const float* cov3D_precomp;   /* [P,6]  or NULL                                        */
    const float* viewmatrix;      /* [16] as stored by scene/cameras.py:54 (transposed)    */
    const float* projmatrix;      /* [16] full projection, same storage                    */
    const float* campos;          /* [3]                                                   */
    const float* sampling_offsets;/* [H,W,2] per-pixel sample jitter (forward.cu:287), or
                                     NULL meaning all zero                                 */
    /* ---- model-space inputs (ABI v2) ------------------------------
     * raw_params != 0 folds the activations that gaussian_renderer/__init__.py:61-90 applies
     * through the GaussianModel getters (scene/gaussian_model.py:26-41,95-120) into K1/K9:
     *   opacities = opacity LOGITS          (sigmoid applied here)
     *   scales    = LOG scales              (exp applied here)
     *   rotations = unnormalised quaternion (x / max(|x|, 1e-12) applied here, F.normalize)
     *   shs       = _features_dc [P,1,3], shs_rest = _features_rest [P,M-1,3]
     *               (torch.cat of get_features is never materialised)
     * colors_precomp / cov3D_precomp must be NULL in this mode. */
    int raw_params;
    const float* shs_rest;        /* [P,M-1,3] or NULL when M == 1 (raw_params only)        */
    /* ---- ABI v5: overlap of the SH parameters' arrival with projection / sorting / binning -----
     * colour_wait_event (a cudaEvent_t, or NULL) is honoured by wast3d_raster_forward only: SH -> RGB
     * (forward.cu:241-246) is evaluated by a separate kernel launched just before the tile render,
     * after cudaStreamWaitEvent(stream, event).  `shs` / `shs_rest` may therefore still be written by
     * another stream (the view-parallel optimizer's parameter all-gather, wast3d_peer_adam_step on a
     * side stream) while K1's geometry part and K2-K5 run; results are bit-identical.  The wait also
     * orders the render and everything after it on `stream` behind the event. */
    void* colour_wait_event;
    /* ---- ABI v6: the geometry buffer returned by geom_alloc already holds K1's outputs for THIS view, written by
     * wast3d_raster_backward_raw_adam_next (see there); `radii` must be the pointer given there.  wast3d_raster_forward*
     * then starts at the depth sort. */
    int preprojected;
} wast3d_raster_params;

/* Replaces RasterizeGaussiansCUDA -> CudaRasterizer::Rasterizer::forward
 * (rasterize_points.cu:35-119, rasterizer_impl.cu:198-341).
 * out_color [3,H,W], out_depth [H,W], radii [P] int32 (may be NULL -> internal).
 * The three scratch buffers are opaque; keep them alive and pass them to backward.
 * One documented layout fact is kept from the reference: the image
 * buffer starts with final_T (float[H*W]) at offset 0, n_contrib (uint32[H*W]) follows at
 * the next 128-byte boundary, so alpha = 1 - imgBuffer[:4*H*W].view(float32).
 * Like the reference this performs ONE blocking device->host read of num_rendered. */
int wast3d_raster_forward(const wast3d_raster_params* prm,
                          wast3d_alloc_fn geom_alloc, void* geom_user,
                          wast3d_alloc_fn binning_alloc, void* binning_user,
                          wast3d_alloc_fn img_alloc, void* img_user,
                          float* out_color, float* out_depth, int* radii,
                          int* num_rendered_host, void* stream);

/* Graph-safe forward (ABI v6): wast3d_raster_forward without its blocking read — no cudaStreamSynchronize, no
 * device->host copy, so a whole optimisation step can be enqueued ahead of the GPU (or captured in a CUDA graph).
 * The reference sizes its binning buffer from num_rendered read back on the host (rasterizer_impl.cu:283-289); here the
 * caller states an upper bound: binning_alloc is called once for `instance_capacity` instances, the instance-sized
 * kernels take the actual count from device memory, and status_dev (DEVICE uint32[4]) receives
 *   {num_rendered, prefiltered-violated flag, look-back time-out flag, overflow flag}.
 * overflow != 0: the view needed more than instance_capacity instances; nothing was written out of bounds but the
 * image of THIS call is incomplete — the caller must discard the step and retry with a larger capacity (the host
 * layer polls the status one call later and raises).  Pass instance_capacity as `num_rendered` to the backward
 * entry points (the binning buffer is carved for it).  Results are bit-identical to wast3d_raster_forward. */
int wast3d_raster_forward_async(const wast3d_raster_params* prm,
                                wast3d_alloc_fn geom_alloc, void* geom_user,
                                wast3d_alloc_fn binning_alloc, void* binning_user,
                                wast3d_alloc_fn img_alloc, void* img_user,
                                float* out_color, float* out_depth, int* radii,
                                int instance_capacity, unsigned int* status_dev, void* stream);

/* Replaces RasterizeGaussiansBackwardCUDA -> Rasterizer::backward
 * (rasterize_points.cu:121-206, rasterizer_impl.cu:345-446).
 * Gradient outputs need NOT be zero-initialised: every element is written.
 * dL_dmean2D [P,3], dL_dcolor [P,3], dL_dopacity [P,1], dL_dmean3D [P,3], dL_dcov3D [P,6],
 * dL_dsh [P,M,3] (NULL allowed when M == 0), dL_dscale [P,3], dL_drot [P,4].
 * dL_dconic [P,2,2] and dL_dcamViewDepth [P,1] are the reference's private scratch tensors
 * (rasterize_points.cu:159,161); pass NULL unless a test wants to inspect them. */
int wast3d_raster_backward(const wast3d_raster_params* prm, int num_rendered, const int* radii,
                           void* geom_buffer, void* binning_buffer, void* img_buffer,
                           const float* dL_dpix, const float* dL_ddepth,
                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                           float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                           float* dL_dscale, float* dL_drot, float* dL_dcamViewDepth,
                           void* stream);

/* Backward of a raw_params forward: gradients with respect to the six leaf parameters of
 * GaussianModel (scene/gaussian_model.py:149-167), i.e. Rasterizer::backward followed by the
 * autograd backward of sigmoid / exp / normalize / cat that the reference runs as separate
 * torch kernels.  dL_dxyz [P,3], dL_dfeatures_dc [P,1,3], dL_dfeatures_rest [P,M-1,3],
 * dL_dopacity_logit [P,1], dL_dlog_scale [P,3], dL_drotation [P,4]; every element is written.
 * dL_dmean2D [P,3] (the viewspace_points gradient) is optional (NULL = not needed).
 * dL_dfeatures_dc and dL_dfeatures_rest may BOTH be NULL: the SH gradients are then not written (the view-parallel
 * colour-record exchange rebuilds them on every rank, include/wast3d_b200_staged.h). */
int wast3d_raster_backward_raw(const wast3d_raster_params* prm, int num_rendered, const int* radii,
                               void* geom_buffer, void* binning_buffer, void* img_buffer,
                               const float* dL_dpix, const float* dL_ddepth, float* dL_dxyz,
                               float* dL_dfeatures_dc, float* dL_dfeatures_rest,
                               float* dL_dopacity_logit, float* dL_dlog_scale, float* dL_drotation,
                               float* dL_dmean2D, void* stream);

/* wast3d_raster_backward_raw with the optimizer folded in ("the Adam update into
 * K9's epilogue"): instead of writing the six leaf gradients and running torch.optim.Adam over them
 * (scene/gaussian_model.py:149-167, GaussianModel.optimizer.step() at train_st.py:342), the per-Gaussian
 * kernel applies the Adam update to every parameter element in place as soon as its gradient is known —
 * same arithmetic as wast3d_adam_step, dense semantics (culled Gaussians take the zero-gradient update).
 * groups[6] = xyz, features_dc, features_rest, opacity, scaling, rotation, in that order; groups[k].param
 * must be the pointer the forward read (prm->means3D, shs, shs_rest, opacities, scales, rotations) and
 * groups[k].step the 1-based count of THIS update.  Only valid when this backward carries the whole
 * gradient of the step (single GPU, one rasteriser call per optimizer step).
 * grads_out: NULL, or 6 pointers (each may be NULL) that additionally receive the leaf gradients (tests). */
typedef struct wast3d_adam_group {
    float* param;
    float* exp_avg;
    float* exp_avg_sq;
    float lr, beta1, beta2, eps;
    int step;
    int reserved;
    /* ABI v7.  NULL: lr and step above give the bias-corrected step size (host values, baked into the launch).
     * Non-NULL: DEVICE [2] floats {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)} that the kernel reads when it RUNS —
     * written by wast3d_adam_schedule_step on the same stream, so that a CUDA graph of the step can be replayed
     * (lr and step are then ignored; step >= 1 is still required). */
    const float* schedule_dev;
} wast3d_adam_group;
/* One optimizer tick on the device (ABI v7): *step_dev += 1, then for every group k
 *   schedule_dev[2k] = lr_k / (1 - beta1_k^t),  schedule_dev[2k + 1] = 1 / sqrt(1 - beta2_k^t)      (t = *step_dev)
 * in double precision like torch.optim.Adam's host arithmetic (scene/gaussian_model.py:154-163 configures it);
 * hyper_dev [ngroups][3] doubles = {lr, beta1, beta2} per group, rewritten by the caller whenever a learning rate
 * changes (scene/gaussian_model.py:169-176 update_learning_rate).  One tiny kernel; capturable. */
int wast3d_adam_schedule_step(int ngroups, const double* hyper_dev, unsigned long long* step_dev,
                              float* schedule_dev, void* stream);
int wast3d_raster_backward_raw_adam(const wast3d_raster_params* prm, int num_rendered, const int* radii,
                                    void* geom_buffer, void* binning_buffer, void* img_buffer,
                                    const float* dL_dpix, const float* dL_ddepth,
                                    const wast3d_adam_group* groups, float* const* grads_out,
                                    float* dL_dmean2D, void* stream);

/* wast3d_raster_backward_raw_adam that ALSO projects every Gaussian for the next view (ABI v6).  In an optimisation
 * loop the next forward's K1 (preprocessCUDA, forward.cu:155-256) re-reads the parameters the optimizer has just
 * written; here the per-Gaussian kernel evaluates it on the updated values while they are still in registers and writes
 * K1's outputs into the NEXT call's geometry buffer: the next forward (wast3d_raster_params::preprojected = 1, same
 * buffer returned by its geom_alloc callback, same radii pointer) skips K1 and its 236 B/Gaussian parameter read.
 * Results are bit-identical to running K1 separately (one source, csrc/project.cuh).
 * The tile cut needs bounds of the next call's sampling offsets before they exist: offset_min/max (the reference
 * draws them in (-1, 0], gaussian_renderer/__init__.py:31); the next forward verifies its offsets against these
 * bounds (status WAST3D_ERR_STALE_PROJECTION).  geom_buffer: wast3d_raster_geom_bytes(P) bytes, 128-byte aligned.
 * The caller promises that nobody modifies the six parameter tensors between this call and that forward. */
typedef struct wast3d_next_view {
    int width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int D;                       /* active SH degree of the next view */
    const float* viewmatrix;     /* DEVICE [16] */
    const float* projmatrix;     /* DEVICE [16] */
    const float* campos;         /* DEVICE [3]  */
    void* geom_buffer;           /* DEVICE, the next forward's geometry buffer */
    int* radii;                  /* DEVICE [P], the next forward's radii output (NULL -> internal) */
    float offset_min_x, offset_max_x, offset_min_y, offset_max_y;
} wast3d_next_view;
size_t wast3d_raster_geom_bytes(int P);
int wast3d_raster_backward_raw_adam_next(const wast3d_raster_params* prm, int num_rendered, const int* radii,
                                         void* geom_buffer, void* binning_buffer, void* img_buffer,
                                         const float* dL_dpix, const float* dL_ddepth,
                                         const wast3d_adam_group* groups, float* const* grads_out,
                                         float* dL_dmean2D, const wast3d_next_view* next, void* stream);

/* Test/inspection hook (no reference equivalent is Python-visible; mirrors the state
 * structs of rasterizer_impl.h:29-65).  Any output may be NULL.
 * depths[P], means2D[P,2], conic_opacity[P,4], rgb[P,3], tiles_touched[P] (uint32),
 * clamped[P,3] (bytes 0/1), point_list[num_rendered] (uint32), ranges[T,2] (uint32). */
int wast3d_raster_export_state(const wast3d_raster_params* prm, int num_rendered,
                               const void* geom_buffer, const void* binning_buffer,
                               const void* img_buffer, float* depths, float* means2D,
                               float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                               unsigned char* clamped, uint32_t* point_list, uint32_t* ranges,
                               void* stream);

/* Tile instancing policy of wast3d_raster_forward (process-wide; returns the previous mode, any other
 * `mode` value only queries).  1 (default; env WAST3D_TILE_CUT) = a Gaussian is instantiated only in
 * the tiles whose sample positions can reach alpha >= 1/255 (forward.cu:355) — same image, same
 * gradients, fewer instances than duplicateWithKeys (rasterizer_impl.cu:70-113); 0 = the reference's
 * radius rectangles (auxiliary.h:46-56), which reproduces its num_rendered and point list exactly. */
int wast3d_set_tile_cut(int mode);

/* Summation order of the tile backward (process-wide; returns the previous mode, any other `mode` value only
 * queries).  0 (default; env WAST3D_DETERMINISTIC) = one float atomic per (warp, Gaussian) hit and gradient
 * slot, unordered like the reference's atomicAdds (backward.cu:576-583); 1 = no float atomics: per-instance
 * partial records are added per Gaussian in point-list order, so gradients (and everything downstream: Adam,
 * the next image) are bit-reproducible from run to run.  Slower and memory hungry (48 B per instance); meant
 * for tests that compare two schedules of the same computation bit for bit. */
int wast3d_set_deterministic(int mode);

/* Replaces markVisible -> checkFrustum (rasterize_points.cu:208-227,
 * rasterizer_impl.cu:54-66,141-153).  present is bool[P] (1 byte each). */
int wast3d_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        const float* projmatrix, unsigned char* present, void* stream);

/* Replaces distCUDA2 -> SimpleKNN::knn (submodules/simple-knn/spatial.cu:15-26,
 * simple_knn.cu:185-221): mean of the 3 smallest squared distances to OTHER points
 * (self excluded by index, simple_knn.cu:158,177).  points [P,3], mean_dist2 [P].
 * Optional extension: nn_index [P,3] int32 (ties -> lowest index), or NULL.
 * Scratch: call wast3d_knn_scratch_bytes(P) and pass a buffer of that size. */
size_t wast3d_knn_scratch_bytes(int P);
int wast3d_knn_dist2(int P, const float* points, float* mean_dist2, int32_t* nn_index,
                     void* scratch, size_t scratch_bytes, void* stream);

/* Cluster matching.  No native reference interface exists (the reference does this inline
 * with torch.cdist + argmin/min: notebooks/10.visualize_and_fit_patch_to_multiple.ipynb
 * cell 34, notebooks/29.2.Modify_style_clusters.ipynb cell 58); see SURVEY.md §8a M3/M5.
 *
 * wast3d_cluster_stats: per-cluster mean [K,3] and covariance (6 upper-tri: xx,xy,xz,yy,yz,zz)
 * of member points; labels int32 in [0,K), points [n,3].  count [K] int32.
 * Accumulated in double from sorted segments so the result is order-independent. */
int wast3d_cluster_stats(int n, int K, const float* points, const int32_t* labels,
                         float* mean, float* cov6, int32_t* count, void* stream);

/* The two accumulation passes of wast3d_cluster_stats on caller-owned accumulators (ABI v6), for statistics over
 * points that are sharded across GPUs ("segmented mean/covariance over 6 M points shard by point with a
 * reduce of 10 floats x Kc"): every rank adds its points, the host all-reduces the accumulators in between.
 *   wast3d_cluster_sums:    sum3 [K,3] double += member xyz, count [K] int32 += members        (zero them first)
 *   wast3d_cluster_scatter: acc6 [K,6] double += (x - mean)(x - mean)^T upper triangle, mean3 [K,3] double = sum / count
 * Then mean = (float) mean3, cov6 = (float)(acc6 / count): bit-identical to wast3d_cluster_stats on the union
 * of the points up to the (double) summation order. */
int wast3d_cluster_sums(int n, int K, const float* points, const int32_t* labels, double* sum3, int32_t* count,
                        void* stream);
int wast3d_cluster_scatter(int n, int K, const float* points, const int32_t* labels, const double* mean3,
                           double* acc6, void* stream);

/* wast3d_nn_match: for each query row a[i] ([Na,3]) the index of the nearest b[j] ([Nb,3])
 * = argmin_j cdist(a,b)[i,j] with ties to the lowest j; out_dist = that Euclidean distance.
 * Decided on fp32 values computed in torch.cdist's operation order (oracle/match_oracle.c).
 * scratch / scratch_bytes: see wast3d_match_scratch_bytes (ABI v6). */
int wast3d_nn_match(int Na, int Nb, const float* a, const float* b, int32_t* out_idx,
                    float* out_dist, void* scratch, size_t scratch_bytes, void* stream);

/* wast3d_cdist_topk: the k smallest entries of every row of torch.cdist(a, b) ([Na,3] x [Nb,3])
 * without materialising the matrix: out_dist [Na,k] ascending, out_idx [Na,k] int32, rows ordered by
 * (distance, index) like a stable sort.  Replaces `torch.sort(torch.cdist(x, y), 1)[:, :k]` /
 * `torch.topk(torch.cdist(x, x), k, largest=False)` and, through out_dist[:, k-1], the kNN mask
 * `D <= sorted[:, k-1:k]` (aux_optimize_cluster_D_W_distance.py:73-82, ...distance2.py:269-273,
 * notebooks/25.4.Optimize_with_SAM_masks_clean.ipynb cell 73).  1 <= k <= min(128, Nb). */
int wast3d_cdist_topk(int Na, int Nb, const float* a, const float* b, int k, float* out_dist,
                      int32_t* out_idx, void* stream);

/* wast3d_emd2_uniform: exact optimal-transport cost between two samples xa, xb ([n,3] each) with
 * uniform weights 1/n and POT's default squared-Euclidean ground cost, i.e. what
 * `ot.emd2(w, w, ot.dist(xa, xb))` returns (aux_optimize_cluster_D_W_distance.py:260-270; POT is an
 * un-vendored dependency).  With equal uniform weights the optimal plan is a permutation; out_perm[i]
 * (optional) is the column matched to row i, the plan is P/n.  1 <= n <= 1023. */
int wast3d_emd2_uniform(int n, const float* xa, const float* xb, float* out_cost, int32_t* out_perm,
                        void* stream);

/* wast3d_kmeans_lloyd: Lloyd's K-Means on 3-D points from a given initialisation — replaces
 * `sklearn.cluster.KMeans(n_clusters=K, n_init=.., max_iter=..).fit_predict(xyz)` of
 * aux_save_clusters_clean.py:32-47 / train_st.py:54-70 (scikit-learn is an un-vendored, unpinned dependency
 * whose default k-means++ initialisation is unseeded in the reference; the caller supplies `centers` [K,3]
 * initialised with the rows it chose and gets the final centres back).  Every iteration is an E-step
 * (labels[i] = argmin_j cdist(points, centers)[i,j] exactly as wast3d_nn_match decides it, ties to the lowest j)
 * and an M-step (centre = mean of its members, accumulated in double; an empty cluster keeps its centre).
 * Stops after max_iter M-steps, when no label changed, or when the summed squared centre shift of the last
 * M-step is <= tol (absolute; sklearn's `tol` is relative: pass tol_sklearn * mean(var(points, axis=0))); the
 * returned labels always come from an E-step against the returned centres.
 * Host outputs (optional): inertia = sum_i |x_i - c_label(i)|^2, number of M-steps done, last shift. */
int wast3d_kmeans_lloyd(int n, int K, const float* points, float* centers, int32_t* labels, int max_iter,
                        double tol, double* out_inertia_host, int* out_n_iter_host, double* out_shift_host,
                        void* stream);

/* Sparse pairwise-distance terms.  A "pair" is (row i, neighbour j):
 *   d[i,j] = row_scale[i] * dist(a[center ? center[i] : i, 0:3], b[idx[i,j], 0:3]),  i < n, j < k
 * formula 0: dist = sqrt((dx^2+dy^2)+dz^2) — `torch.norm(X_nns[:,1:] - X_nns[:,0].unsqueeze(1), dim=-1)` of
 *            get_descriptors (notebooks/25.4.Optimize_with_SAM_masks_clean.ipynb cell 72) with
 *            a = b = X, center = idx_full[:,0], idx = idx_full[:,1:];
 * formula 1: dist = sqrt(max(0, |a|^2 + |b|^2 - 2 a.b)) in torch.cdist's matmul-path order (the entries of
 *            `torch.cdist(A, xyz)` at aux_optimize_cluster_D_W_distance.py:253-256), center = NULL.
 * a and b are row-strided (lda, ldb floats per row, >= 3) so that `_rotation[:, :-1]` / `[:, 1:]` views
 * are passed without a copy.  center, row_scale may be NULL. */
typedef struct wast3d_pair_args {
    long long n;
    int k;
    int formula;
    const float* a;
    int lda;
    const float* b;
    int ldb;
    const int32_t* center;   /* [n] or NULL   */
    const int32_t* idx;      /* [n,k]         */
    const float* row_scale;  /* [n] or NULL   */
    const float* a2;         /* optional second a operand (same rows as a): d = dist(a,b) + dist(a2,b) — the  */
    int lda2;                /* rotation term cdist(rot[:, :-1], xyz) + cdist(rot[:, 1:], xyz) of :254-255     */
} wast3d_pair_args;
/* out_d [n,k].  backward: grad_d [n,k] -> grad_a [rows of a, 3], grad_b [rows of b, 3], dense, ACCUMULATED
 * (zero-fill them first; any may be NULL; grad_a2 [rows of a2, 3] only with a2; grad_a and grad_b may be the same
 * buffer when a == b): g * (a-b)/d, 0 at d == 0
 * like torch's norm / cdist backward.  Sums use unordered float atomics. */
int wast3d_pair_dist_forward(const wast3d_pair_args* args, float* out_d, void* stream);
int wast3d_pair_dist_backward(const wast3d_pair_args* args, const float* grad_d, float* grad_a, float* grad_a2,
                              float* grad_b, void* stream);
/* Fused loss over the pairs: loss = scale * sum_{i,j} weight[i,j] * rho(d[i,j] - target[i,j]), rho = |.| (mode 0)
 * or (.)^2 (mode 1); weight NULL = 1.  With the kNN mask of the target scene as (idx, weight) and
 * scale = 1/(n * rows_of_b) this is `torch.mean(torch.abs(D - D_target) * D_xyz_target_mask)`
 * (aux_optimize_cluster_D_W_distance.py:278-280) without any N x N matrix; with mode 1 and scale = 1/(n k) it is
 * `torch.mean(torch.square(descriptors - target))` (notebooks/25.4 cell 72).  The value is reduced in a fixed
 * order (deterministic).  scratch: wast3d_pair_loss_scratch_bytes() bytes, zero-filled once, reusable on the
 * same stream.  backward: grad_out DEVICE scalar or NULL (= 1); grad_a / grad_b as above. */
size_t wast3d_pair_loss_scratch_bytes(void);
int wast3d_pair_loss_forward(const wast3d_pair_args* args, const float* target, const float* weight, int mode,
                             double scale, void* scratch, float* out_loss, void* stream);
int wast3d_pair_loss_backward(const wast3d_pair_args* args, const float* target, const float* weight, int mode,
                              double scale, const float* grad_out, float* grad_a, float* grad_a2, float* grad_b,
                              void* stream);

/* wast3d_w2_match: content clusters (mean_c [Kc,3], cov_c [Kc,6]) against style clusters
 * (mean_s [Ks,3], cov_s [Ks,6]): out_idx[i] = argmin_j W2^2(N(mc_i,Sc_i), N(ms_j,Ss_j)),
 *   W2^2 = |m1-m2|^2 + tr(S1) + tr(S2) - 2 tr((S1^1/2 S2 S1^1/2)^1/2)      
 * ties to the lowest j; out_cost = that W2^2 (fp32, fixed operation order, see
 * oracle/match_oracle.c).  A tcgen05 GEMM over augmented bf16-split descriptors gives a
 * lower bound per pair; only pairs that can beat the row's running best get the exact
 * Bures term.  stats (optional, DEVICE [4] uint64): pairs, exact_evals, gemm_tiles, error word.
 * ABI v6 — like wast3d_knn_dist2 the call takes caller scratch: wast3d_match_scratch_bytes(Kc, Ks)
 * bytes, 128-byte aligned, contents irrelevant, reusable across calls on the same stream (NULL = a
 * stream-ordered allocation inside the call).  The call never synchronises the stream: the (never
 * expected) tensor-core barrier time-out is reported as out_idx[i] = -2, out_cost[i] = NaN for every row
 * and a non-zero stats[3]. */
size_t wast3d_match_scratch_bytes(int Kc, int Ks);
int wast3d_w2_match(int Kc, int Ks, const float* mean_c, const float* cov_c,
                    const float* mean_s, const float* cov_s, int32_t* out_idx, float* out_cost,
                    unsigned long long* stats, void* scratch, size_t scratch_bytes, void* stream);

/* Test hook: wast3d_w2_match that additionally writes the tensor-core lower-bound matrix
 * (margin already subtracted) to lb_dump [Kc,Ks]; used by tests to validate the tcgen05 path. */
int wast3d_w2_match_debug(int Kc, int Ks, const float* mean_c, const float* cov_c,
                          const float* mean_s, const float* cov_s, int32_t* out_idx,
                          float* out_cost, unsigned long long* stats, float* lb_dump, void* stream);

/* Fused Adam step over one flat fp32 parameter group (replaces torch.optim.Adam as configured
 * in scene/gaussian_model.py:154-163: betas (0.9,0.999), eps 1e-15, no weight decay).
 * step is the 1-based step count AFTER increment (bias corrections use it). */
int wast3d_adam_step(size_t n, float* param, const float* grad, float* exp_avg,
                     float* exp_avg_sq, float lr, float beta1, float beta2, float eps, int step,
                     void* stream);

/* Fused pixel losses of the style-optimisation loop (replaces the torch expressions of
 * utils/loss_utils.py:18-19 l1_loss and :213-215 tv_loss as called at train_st_normals.py:127,145, plus
 * a depth L2 term):
 *   loss = w_l1 * mean|img - gt| + w_tv * 0.5 * (mean|img[y+1]-img[y]| + mean|img[x+1]-img[x]|)
 *        + w_depth * mean((depth - depth_gt)^2)
 * img, gt [C,H,W] (gt NULL = no L1 term); depth, depth_gt [H,W] or both NULL.  scratch:
 * wast3d_pixel_loss_scratch_bytes() bytes, zero-filled once by the caller and then reusable for calls
 * on the same stream.  out_loss: one float.  The reduction order is fixed (deterministic result).
 * backward: d_img [C,H,W] and d_depth [H,W] (d_depth NULL allowed) = grad_out[0] * dloss/d(.), grad_out
 * a DEVICE scalar (NULL = 1); sign(0) = 0 like torch. */
size_t wast3d_pixel_loss_scratch_bytes(void);
int wast3d_pixel_loss_forward(int C, int H, int W, const float* img, const float* gt, const float* depth,
                              const float* depth_gt, float w_l1, float w_tv, float w_depth, void* scratch,
                              float* out_loss, void* stream);
int wast3d_pixel_loss_backward(int C, int H, int W, const float* img, const float* gt, const float* depth,
                               const float* depth_gt, float w_l1, float w_tv, float w_depth,
                               const float* grad_out, float* d_img, float* d_depth, void* stream);

/* Depth -> normal map of the train_st_normals variant (BASELINE.json configs[4]); replaces the kornia / torch
 * expression of train_st_normals.py:113-123:
 *   normals = kornia.geometry.depth.depth_to_normals(depth[None,None], K, normalize_points=False)   (unit normals)
 *   image_normals = (normals - amin(normals)) / (amax(normals) - amin(normals) + 1e-6)
 * with K = [[fx,0,cx],[0,fy,cy],[0,0,1]] (the reference hard-codes 1111, 1111, 400, 400 for its 800x800 images).
 * kornia is an un-vendored, unpinned dependency: its published algorithm (unproject with the pixel grid, 3x3 Sobel / 8
 * with replicate padding on the point image, cross product, F.normalize with eps 1e-12) is restated in
 * oracle/normals.py.  depth [H,W] -> normals_unit [3,H,W], normals01 [3,H,W], minmax [2] (amin, amax), all DEVICE.
 * backward: grad01 [3,H,W] -> grad_depth [H,W] (every element written), including the gradient that flows through
 * amin / amax (shared evenly among ties like torch); grad_ab is a [6,H,W] scratch image.  scratch:
 * wast3d_depth_normals_scratch_bytes() bytes, reusable on the same stream.  Deterministic (no float atomics). */
size_t wast3d_depth_normals_scratch_bytes(void);
int wast3d_depth_normals_forward(int H, int W, const float* depth, float fx, float fy, float cx, float cy,
                                 float* normals_unit, float* normals01, float* minmax, void* scratch, void* stream);
int wast3d_depth_normals_backward(int H, int W, const float* depth, float fx, float fy, float cx, float cy,
                                  const float* normals_unit, const float* minmax, const float* grad01,
                                  float* grad_ab, float* grad_depth, void* scratch, void* stream);

/* ---- view-parallel optimizer step over NVLink peer memory (ABI v3) ------------------
 * The reference is single-GPU: torch.optim.Adam over six groups (scene/gaussian_model.py:149-167).
 * With N GPUs rendering N views, every rank keeps a replica of one flat "arena" (all parameter tensors
 * back to back, each padded to 16 bytes, then the gradients in the same layout).  Rank r owns the
 * float4 range [shard_begin4, shard_end4) of the arena: wast3d_peer_adam_step sums the N gradient
 * replicas of that range through peer loads (fixed rank order), applies Adam with shard-local moments
 * and stores the new parameters into all N replicas — reduce-scatter + Adam + all-gather in one kernel,
 * including the flag exchange that orders it against the peers' gradient kernels.  world == 1 is a
 * single-launch multi-tensor Adam (no flags needed).
 *
 * grad_ptrs / param_ptrs / flag_ptrs: HOST arrays of `world` DEVICE pointers (entry q = rank q's
 * replica as mapped into THIS process; 16-byte aligned).  flag_ptrs[q] points at
 * wast3d_peer_flag_bytes() zero-initialised bytes of rank q.  exp_avg / exp_avg_sq: this rank's
 * moments, 4*(shard_end4-shard_begin4) floats.  segs: parameter groups as float4 ranges of the arena,
 * ascending and disjoint, `step` = 1-based step count.  grad_scale multiplies the summed gradient
 * (1/world = average).  epoch: 1, 2, 3, ... identical on all ranks and increasing per call.
 * mc_grads / mc_params: NVLS multicast addresses of the gradient / parameter regions (an NVSwitch
 * multicast object bound to every rank's arena), or NULL.  When both are given the switch sums the
 * gradients (multimem.ld_reduce) and broadcasts the parameters (multimem.st): ~1x the arena crosses
 * each NVLink direction per step instead of 2 (N-1)/N x with per-peer loads and stores.
 * A rank that waits longer than timeout_s (<= 0: 20 s) for a peer sets a sticky error
 * (wast3d_peer_error) instead of hanging the GPU.
 * max_ctas (ABI v5; 0 = as many as are resident at once): cap of the persistent grid.  A launch that is
 * meant to run BESIDE other kernels (the side-stream exchange of the SH features, see
 * wast3d_raster_params::colour_wait_event) saturates NVLink with a few dozen CTAs and must leave the
 * other SMs to the rasteriser. */
#define WAST3D_PEER_MAX_WORLD 8
#define WAST3D_PEER_MAX_SEGMENTS 8
#define WAST3D_PEER_HANDLE_BYTES 64
typedef struct wast3d_adam_segment {
    unsigned long long begin4, end4; /* float4 units within the arena */
    float lr, beta1, beta2, eps;
    int step;
    int reserved;
} wast3d_adam_segment;
size_t wast3d_peer_flag_bytes(void);
int wast3d_peer_adam_step(int world, int rank, void* const* grad_ptrs, void* const* param_ptrs,
                          void* const* flag_ptrs, void* mc_grads, void* mc_params,
                          float* exp_avg, float* exp_avg_sq,
                          size_t shard_begin4, size_t shard_end4, const wast3d_adam_segment* segs,
                          int nsegs, float grad_scale, unsigned epoch, double timeout_s, int max_ctas,
                          void* stream);
/* 0 = no error; k > 0 = timed out waiting for rank k-1.  reset != 0 clears it. */
int wast3d_peer_error(int reset);
/* Peer-visible device memory for the arena: cudaMalloc'ed (zero-filled) and shared between the
 * processes of one node through CUDA IPC handles (WAST3D_PEER_HANDLE_BYTES opaque bytes, exchanged by
 * the host layer over torch.distributed).  release: imported != 0 closes a mapping, 0 frees. */
int wast3d_peer_alloc(size_t bytes, void** out_ptr);
int wast3d_peer_export(void* ptr, unsigned char* handle64);
int wast3d_peer_import(const unsigned char* handle64, void** out_ptr);
int wast3d_peer_release(void* ptr, int imported);

/* Test / measurement hooks for the binning primitives that replace the reference's CUB calls
 * (cub::DeviceScan::InclusiveSum rasterizer_impl.cu:279, cub::DeviceRadixSort::SortPairs :305-310).
 * wast3d_test_sort_pairs: stable LSD radix sort of n (u32 key, u32 value) pairs on key bits
 * [begin_bit, end_bit); vals_in == NULL means values 0..n-1.  mode 0 = multi-kernel passes
 * (histogram / table scan / scatter), 1 = single-kernel passes with decoupled look-back (what the
 * rasteriser uses).  wast3d_test_scan: out[i] = sum_{k<i} in[perm ? perm[k] : k], *total = full sum
 * (mode 0 = three kernels, 1 = one look-back kernel). */
int wast3d_test_sort_pairs(size_t n, const uint32_t* keys_in, const uint32_t* vals_in,
                           uint32_t* keys_out, uint32_t* vals_out, int begin_bit, int end_bit,
                           int mode, void* stream);
int wast3d_test_scan(size_t n, const uint32_t* in, const uint32_t* perm, uint32_t* out,
                     uint32_t* total, int mode, void* stream);

/* Measurement hooks (bench.py): per-stage CUDA-event timing on the launching stream and a count
 * of kernel launches issued by this library.  Slots: wast3d_profile_slots() names via
 * wast3d_profile_slot_name(); enable a subset with a bit mask (0 = off, the default).
 * wast3d_profile_read ADDS elapsed milliseconds / scope counts into the caller's arrays. */
int wast3d_profile_set(unsigned slot_mask);
int wast3d_profile_slots(void);
const char* wast3d_profile_slot_name(int slot);
int wast3d_profile_read(double* ms_per_slot, unsigned long long* scopes_per_slot);
unsigned long long wast3d_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* WAST3D_B200_H_ */
