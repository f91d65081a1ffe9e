"""Cluster-to-style matching — Python surface over the C ABI.

The reference has no module for this; it writes the arithmetic inline:
  * nearest centroid / nearest point: `torch.argmin(torch.cdist(a, b), -1)`
    (notebooks/10.visualize_and_fit_patch_to_multiple.ipynb cell 34) and
    `torch.min(torch.cdist(opt[b:b+10000], content[::10]), 1)` (notebooks/29.2... cell 58)
    -> `nn_match(a, b)` returns (idx int64, dist) for all rows at once, no N x M matrix;
  * cluster statistics from K-Means memberships -> `cluster_stats(points, labels, K)`;
  * closed-form Gaussian W2 nearest style cluster (named by the north star, absent from the
    reference) -> `w2_match(mean_c, cov_c, mean_s, cov_s)`.
All three run in libwast3d_b200.so (csrc/match.cu); there is no PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _lib


def _check3(t, name, cols):
    if t.dim() != 2 or t.size(1) != cols:
        raise RuntimeError(f"{name} must have dimensions (n, {cols})")


def cluster_stats(points: torch.Tensor, labels: torch.Tensor, K: int):
    """Per-cluster (mean [K,3], cov6 [K,6] = xx,xy,xz,yy,yz,zz population covariance, count [K])."""
    _check3(points, "points", 3)
    _lib.require_device(points)
    lib = _lib.load()
    dev = points.device
    n = int(points.size(0))
    mean = torch.empty((K, 3), dtype=torch.float32, device=dev)
    cov = torch.empty((K, 6), dtype=torch.float32, device=dev)
    count = torch.empty((K,), dtype=torch.int32, device=dev)
    keep: list = []
    lab = labels.to(device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        st = lib.wast3d_cluster_stats(n, int(K), _lib.fptr(points, keep), _lib.fptr(lab, keep, torch.int32),
                                      mean.data_ptr(), cov.data_ptr(), count.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "cluster_stats")
    return mean, cov, count


def nn_match(a: torch.Tensor, b: torch.Tensor):
    """(argmin_j |a_i - b_j|, that distance): what `torch.min(torch.cdist(a, b), 1)` returns
    (values, indices swapped into (idx, dist) order); ties go to the lowest index."""
    _check3(a, "a", 3)
    _check3(b, "b", 3)
    _lib.require_device(a)
    lib = _lib.load()
    Na, Nb = int(a.size(0)), int(b.size(0))
    if Nb == 0:
        raise RuntimeError("nn_match: b is empty (torch.argmin would raise as well)")
    idx = torch.empty((Na,), dtype=torch.int32, device=a.device)
    dist = torch.empty((Na,), dtype=torch.float32, device=a.device)
    keep: list = []
    with torch.cuda.device(a.device):
        st = lib.wast3d_nn_match(Na, Nb, _lib.fptr(a, keep), _lib.fptr(b, keep),
                                 idx.data_ptr() if Na else None, dist.data_ptr() if Na else None,
                                 _lib.stream_ptr())
    _lib.check(st, "nn_match")
    return idx.long(), dist


def w2_match(mean_c, cov_c, mean_s, cov_s, return_stats: bool = False, _lb_dump: bool = False):
    """Nearest style cluster per content cluster under squared Gaussian W2.

    Returns (idx int64 [Kc], cost float32 [Kc]) and, with return_stats, a dict
    {pairs, exact_evals, gemm_tiles}: how many of the Kc*Ks pairs needed the exact Bures term
    after the tensor-core lower bound.
    """
    _check3(mean_c, "mean_c", 3)
    _check3(cov_c, "cov_c", 6)
    _check3(mean_s, "mean_s", 3)
    _check3(cov_s, "cov_s", 6)
    _lib.require_device(mean_c)
    lib = _lib.load()
    dev = mean_c.device
    Kc, Ks = int(mean_c.size(0)), int(mean_s.size(0))
    if Ks == 0:
        raise RuntimeError("w2_match: no style clusters")
    idx = torch.empty((Kc,), dtype=torch.int32, device=dev)
    cost = torch.empty((Kc,), dtype=torch.float32, device=dev)
    stats = torch.zeros((4,), dtype=torch.int64, device=dev)
    lb = torch.full((Kc, Ks), float("nan"), dtype=torch.float32, device=dev) if _lb_dump else None
    keep: list = []
    args = (Kc, Ks, _lib.fptr(mean_c, keep), _lib.fptr(cov_c, keep), _lib.fptr(mean_s, keep),
            _lib.fptr(cov_s, keep), idx.data_ptr() if Kc else None, cost.data_ptr() if Kc else None,
            stats.data_ptr())
    with torch.cuda.device(dev):
        if _lb_dump:
            st = lib.wast3d_w2_match_debug(*args, lb.data_ptr(), _lib.stream_ptr())
        else:
            st = lib.wast3d_w2_match(*args, _lib.stream_ptr())
    _lib.check(st, "w2_match")
    out = (idx.long(), cost)
    if return_stats:
        s = stats.tolist()
        out = out + ({"pairs": s[0], "exact_evals": s[1], "gemm_tiles": s[2]},)
    if _lb_dump:
        out = out + (lb,)
    return out
