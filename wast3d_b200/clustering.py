"""Cluster creation for the style-transfer pipeline: K-Means over Gaussian centres on
the GPU and the reference's per-cluster `.npz` files.

Reference code this replaces:
  * `cluster_points(points, k_clusters)` — aux_save_clusters_clean.py:32-47 (KMeans(n_init=20, max_iter=100)) and
    train_st.py:54-70 (n_init=1, max_iter=30): scikit-learn on the CPU over the whole scene;
  * `clustering()` — aux_save_clusters_clean.py:141-166: re-centre every cluster on its centroid and save the six
    GaussianModel attributes of its members as `cluster_<i>.npz`.
The E-step is the library's exact nearest-centre kernel (wast3d_nn_match: tcgen05 lower bound + exact fp32 cdist,
ties to the lowest centre), the M-step accumulates in double; both run in wast3d_kmeans_lloyd (csrc/match.cu).
No CPU or torch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

CLUSTER_ATTRS = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity")


def kmeans_lloyd(points: torch.Tensor, init_centers: torch.Tensor, max_iter: int = 100, tol: float = 0.0):
    """Lloyd iterations from `init_centers` [K,3].  Returns (labels int64 [N], centers float32 [K,3], inertia,
    n_iter).  `tol` is the absolute bound on the summed squared centre shift (sklearn's relative `tol` times
    mean(var(points, 0)))."""
    _lib.require_device(points)
    if points.dim() != 2 or points.size(1) != 3 or init_centers.dim() != 2 or init_centers.size(1) != 3:
        raise RuntimeError("kmeans_lloyd: points [N,3] and init_centers [K,3] expected")
    n, K = int(points.size(0)), int(init_centers.size(0))
    if K < 1 or K > n:
        raise ValueError(f"n_samples={n} should be >= n_clusters={K}.")  # sklearn's message
    pts = points.detach().to(torch.float32).contiguous()
    centers = init_centers.detach().to(device=pts.device, dtype=torch.float32).clone().contiguous()
    labels = torch.empty((n,), dtype=torch.int32, device=pts.device)
    inertia, n_iter, shift = C.c_double(0.0), C.c_int(0), C.c_double(0.0)
    with torch.cuda.device(pts.device):
        rc = _lib.load().wast3d_kmeans_lloyd(n, K, pts.data_ptr(), centers.data_ptr(), labels.data_ptr(), int(max_iter),
                                             float(tol), C.byref(inertia), C.byref(n_iter), C.byref(shift),
                                             _lib.stream_ptr())
    _lib.check(rc, "kmeans_lloyd")
    return labels.long(), centers, float(inertia.value), int(n_iter.value)


def cluster_points(points, k_clusters: int, n_init: int = 1, max_iter: int = 30, tol: float = 1e-4, seed: int = 0):
    """Same contract as the reference's `cluster_points` (aux_save_clusters_clean.py:32-47): returns
    (cluster_indices [N], cluster_centers [k,3]) as numpy arrays when given numpy, tensors when given tensors.
    Initialisation: `n_init` random draws of k distinct points (seeded — the reference's k-means++ draws are
    unseeded, so memberships were never reproducible run to run); the run with the lowest inertia wins, as in
    scikit-learn."""
    as_numpy = not isinstance(points, torch.Tensor)
    pts = torch.as_tensor(np.asarray(points, dtype=np.float32)).cuda() if as_numpy else points
    n = int(pts.size(0))
    gen = torch.Generator().manual_seed(int(seed))
    var = float(pts.detach().to(torch.float32).var(dim=0, unbiased=False).mean().item()) if n > 1 else 0.0
    best = None
    for _ in range(max(1, int(n_init))):
        pick = torch.randperm(n, generator=gen)[:k_clusters].to(pts.device)
        out = kmeans_lloyd(pts, pts.detach()[pick], max_iter=max_iter, tol=tol * var)
        if best is None or out[2] < best[2]:
            best = out
    labels, centers = best[0], best[1]
    if as_numpy:
        return labels.cpu().numpy(), centers.cpu().numpy()
    return labels, centers


def save_clusters(gaussians, cluster_indices, cluster_centers, output_dir: str):
    """aux_save_clusters_clean.py:151-164: subtract every point's cluster centre from `_xyz`, then write the six
    attributes of each cluster's members to `<output_dir>/cluster_<idx>.npz` (keys as in the reference)."""
    os.makedirs(output_dir, exist_ok=True)
    labels = torch.as_tensor(np.asarray(cluster_indices)).long() if not isinstance(cluster_indices, torch.Tensor) else cluster_indices.long()
    centers = torch.as_tensor(np.asarray(cluster_centers), dtype=torch.float32) if not isinstance(cluster_centers, torch.Tensor) else cluster_centers
    xyz = gaussians._xyz.detach()
    labels = labels.to(xyz.device)
    recentred = xyz - centers.to(xyz.device)[labels]
    paths = []
    for cluster_idx in torch.unique(labels).tolist():
        members = torch.where(labels == cluster_idx)[0]
        cluster_dict = {}
        for attr in CLUSTER_ATTRS:
            src = recentred if attr == "_xyz" else getattr(gaussians, attr).detach()
            cluster_dict[attr] = src[members].cpu().numpy()
        path = os.path.join(output_dir, f"cluster_{cluster_idx}.npz")
        np.savez(path, **cluster_dict)
        paths.append(path)
    return paths


def load_cluster(path: str) -> dict:
    """One `cluster_<i>.npz` -> dict of numpy arrays with the six GaussianModel attribute keys."""
    with np.load(path) as z:
        return {k: z[k] for k in CLUSTER_ATTRS}
