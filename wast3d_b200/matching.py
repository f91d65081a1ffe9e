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


# (device index) -> reusable scratch tensor of the match kernels.  Calls on one device are stream-ordered on
# torch's current stream; a caller that matches concurrently on several streams passes its own `scratch`.
_SCRATCH: dict = {}


def match_scratch(dev: torch.device, Kc: int, Ks: int) -> torch.Tensor:
    """Caller scratch of wast3d_w2_match / wast3d_nn_match (wast3d_match_scratch_bytes), cached per device."""
    need = int(_lib.load().wast3d_match_scratch_bytes(int(Kc), int(Ks)))
    key = torch.device(dev).index
    t = _SCRATCH.get(key)
    if t is None or t.numel() < need:
        t = torch.empty(_lib.bucket_bytes(need), dtype=torch.uint8, device=dev)
        _SCRATCH[key] = t
    return t


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


def cluster_sums(points: torch.Tensor, labels: torch.Tensor, K: int, sum3: torch.Tensor, count: torch.Tensor):
    """Pass 1 of the statistics on caller accumulators: sum3 [K,3] float64 += member xyz, count [K] int32 += 1."""
    _check3(points, "points", 3)
    _lib.require_device(points)
    keep: list = []
    lab = labels.to(device=points.device, dtype=torch.int32)
    with torch.cuda.device(points.device):
        st = _lib.load().wast3d_cluster_sums(int(points.size(0)), int(K), _lib.fptr(points, keep),
                                             _lib.fptr(lab, keep, torch.int32), sum3.data_ptr(), count.data_ptr(),
                                             _lib.stream_ptr())
    _lib.check(st, "cluster_sums")


def cluster_scatter(points: torch.Tensor, labels: torch.Tensor, K: int, mean3: torch.Tensor, acc6: torch.Tensor):
    """Pass 2: acc6 [K,6] float64 += (x - mean)(x - mean)^T (upper triangle) with mean3 [K,3] float64."""
    _check3(points, "points", 3)
    _lib.require_device(points)
    keep: list = []
    lab = labels.to(device=points.device, dtype=torch.int32)
    with torch.cuda.device(points.device):
        st = _lib.load().wast3d_cluster_scatter(int(points.size(0)), int(K), _lib.fptr(points, keep),
                                                _lib.fptr(lab, keep, torch.int32), mean3.data_ptr(), acc6.data_ptr(),
                                                _lib.stream_ptr())
    _lib.check(st, "cluster_scatter")


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
        ws = match_scratch(a.device, Na, Nb)
        st = lib.wast3d_nn_match(Na, Nb, _lib.fptr(a, keep), _lib.fptr(b, keep),
                                 idx.data_ptr() if Na else None, dist.data_ptr() if Na else None,
                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "nn_match")
    return idx.long(), dist


def cdist_topk(a: torch.Tensor, b: torch.Tensor, k: int):
    """(values [Na,k], indices [Na,k] int64) = the k smallest entries of each row of torch.cdist(a, b),
    ascending, equal distances ordered by index (what `torch.sort(D, 1, stable=True)` gives) — the
    N x M matrix is never written.  Call sites in the reference: the kNN masks
    `D <= torch.sort(D, 1)[0][:, k-1:k]` (aux_optimize_cluster_D_W_distance.py:79-82) and the
    descriptor neighbourhoods `torch.topk(cdist, k, largest=False)` (notebooks/25.4 cell 73)."""
    _check3(a, "a", 3)
    _check3(b, "b", 3)
    _lib.require_device(a)
    lib = _lib.load()
    Na, Nb, k = int(a.size(0)), int(b.size(0)), int(k)
    if k < 1 or k > Nb:
        raise RuntimeError("cdist_topk: selected index k out of range")  # torch.topk's message
    if k > 128:
        raise RuntimeError("cdist_topk: k > 128 is not supported")
    vals = torch.empty((Na, k), dtype=torch.float32, device=a.device)
    idx = torch.empty((Na, k), dtype=torch.int32, device=a.device)
    keep: list = []
    with torch.cuda.device(a.device):
        st = lib.wast3d_cdist_topk(Na, Nb, _lib.fptr(a, keep), _lib.fptr(b, keep), k,
                                   vals.data_ptr() if Na else None, idx.data_ptr() if Na else None,
                                   _lib.stream_ptr())
    _lib.check(st, "cdist_topk")
    return vals, idx.long()


def knn_mask_threshold(a: torch.Tensor, b: torch.Tensor, k: int):
    """Per-row threshold t_i with `torch.cdist(a, b)[i] <= t_i` == the reference's kNN mask row
    (aux_optimize_cluster_D_W_distance.py:79-82), plus the k nearest indices.  The dense mask is
    `dist_row <= t_i`; ties at the k-th distance are all inside it, as in the reference."""
    vals, idx = cdist_topk(a, b, k)
    return vals[:, k - 1].contiguous(), idx


class _Emd2Uniform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xa, xb):
        lib = _lib.load()
        n = int(xa.size(0))
        cost = torch.empty((), dtype=torch.float32, device=xa.device)
        perm = torch.empty((n,), dtype=torch.int32, device=xa.device)
        keep: list = []
        with torch.cuda.device(xa.device):
            st = lib.wast3d_emd2_uniform(n, _lib.fptr(xa.detach(), keep), _lib.fptr(xb.detach(), keep),
                                         cost.data_ptr(), perm.data_ptr(), _lib.stream_ptr())
        _lib.check(st, "emd2_uniform")
        ctx.save_for_backward(xa, xb, perm)
        ctx.mark_non_differentiable(perm)
        return cost, perm

    @staticmethod
    def backward(ctx, g_cost, _g_perm):
        # d cost / d M = plan = P_sigma / n (what POT's emd2 back-propagates); M_ij = |a_i - b_j|^2
        xa, xb, perm = ctx.saved_tensors
        n = xa.size(0)
        diff = (xa - xb[perm.long()]) * (2.0 / n) * g_cost
        gb = torch.zeros_like(xb)
        gb.index_add_(0, perm.long(), -diff)
        return diff, gb


def emd2_uniform(xa: torch.Tensor, xb: torch.Tensor, return_plan: bool = False):
    """`ot.emd2(w, w, ot.dist(xa, xb))` for uniform w = 1/n (aux_optimize_cluster_D_W_distance.py:260-270):
    exact optimal-transport cost between two equally sized samples, differentiable with respect to
    both point sets through the optimal plan.  `return_plan` also returns the permutation
    sigma (plan = P_sigma / n)."""
    _check3(xa, "xa", 3)
    _check3(xb, "xb", 3)
    if xa.size(0) != xb.size(0):
        raise RuntimeError("emd2_uniform: both samples must have the same number of points")
    _lib.require_device(xa)
    cost, perm = _Emd2Uniform.apply(xa, xb)
    return (cost, perm.long()) if return_plan else cost


def w2_match(mean_c, cov_c, mean_s, cov_s, return_stats: bool = False, _lb_dump: bool = False,
             stats_out: torch.Tensor | None = None, int32_out: tuple | None = None):
    """Nearest style cluster per content cluster under squared Gaussian W2.

    Returns (idx int64 [Kc], cost float32 [Kc]) and, with return_stats, a dict
    {pairs, exact_evals, gemm_tiles}: how many of the Kc*Ks pairs needed the exact Bures term
    after the tensor-core lower bound (reading it synchronises; pass `stats_out`, a device int64[4]
    tensor, to collect the counters without a host round trip).  The call itself never synchronises:
    a (never expected) tensor-core barrier time-out shows as idx == -2 / cost NaN.
    `int32_out` = (idx int32 [Kc], cost [Kc]) preallocated outputs: nothing is allocated or converted
    (steady-state loops, bench.py).
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
    if int32_out is not None:
        idx, cost = int32_out
        if idx.dtype != torch.int32 or cost.dtype != torch.float32 or idx.numel() != Kc or cost.numel() != Kc:
            raise RuntimeError("w2_match: int32_out must be (int32 [Kc], float32 [Kc])")
    else:
        idx = torch.empty((Kc,), dtype=torch.int32, device=dev)
        cost = torch.empty((Kc,), dtype=torch.float32, device=dev)
    want_stats = return_stats or stats_out is not None
    stats = stats_out if stats_out is not None else (torch.empty((4,), dtype=torch.int64, device=dev) if want_stats else None)
    lb = torch.full((Kc, Ks), float("nan"), dtype=torch.float32, device=dev) if _lb_dump else None
    keep: list = []
    args = (Kc, Ks, _lib.fptr(mean_c, keep), _lib.fptr(cov_c, keep), _lib.fptr(mean_s, keep),
            _lib.fptr(cov_s, keep), idx.data_ptr() if Kc else None, cost.data_ptr() if Kc else None,
            stats.data_ptr() if stats is not None else None)
    with torch.cuda.device(dev):
        if _lb_dump:
            st = lib.wast3d_w2_match_debug(*args, lb.data_ptr(), _lib.stream_ptr())
        else:
            ws = match_scratch(dev, Kc, Ks)
            st = lib.wast3d_w2_match(*args, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "w2_match")
    if int32_out is not None:
        return idx, cost
    out = (idx.long(), cost)
    if return_stats:
        s = stats.tolist()
        if s[3]:
            raise RuntimeError("w2_match: tensor-core barrier timed out on the device")
        out = out + ({"pairs": s[0], "exact_evals": s[1], "gemm_tiles": s[2]},)
    if _lb_dump:
        out = out + (lb,)
    return out
