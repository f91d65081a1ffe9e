"""Geometry regularisers of the style optimisation on sparse neighbour pairs —
Python surface over csrc/pairloss.cu with the reference's function names.

Reference code this replaces (it materialises N x N matrices / [N,k,3] gathers with torch ops):
  * notebooks/25.4.Optimize_with_SAM_masks_clean.ipynb cell 72: `get_descriptors`,
    `get_style_patch_descriptors_loss`; cell 73: neighbourhoods `topk(cdist, num_nns)[:, ::kth_nn]`;
    notebooks/29.2.Modify_style_clusters.ipynb cell 69: `descriptors_loss(normalize=...)`, cell 70: the per-step
    scale re-derivation from descriptors;
  * aux_optimize_cluster_D_W_distance.py:70-82 (pairwise-distance targets + kNN mask), :253-256, :278-280
    (`torch.mean(torch.abs(D - D_target) * D_xyz_target_mask)`).
There is no CPU or torch fallback for the distance kernels: CUDA tensors and the built library are required.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib
from .matching import cdist_topk

FORMULA_NORM, FORMULA_CDIST = 0, 1
_SCRATCH: dict = {}


def _scratch(dev: torch.device) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _SCRATCH.get(key)
    if t is None:
        t = _SCRATCH[key] = torch.zeros(int(_lib.load().wast3d_pair_loss_scratch_bytes()), dtype=torch.uint8, device=dev)
    return t


def _rows3(t: torch.Tensor, name: str) -> torch.Tensor:
    """[N,>=3] float32 CUDA tensor whose rows are `stride(0)` floats apart and unit-strided inside (so that views
    like `_rotation[:, :-1]` pass without a copy); anything else is made contiguous."""
    _lib.require_device(t)
    if t.dim() != 2 or t.size(1) != 3:
        raise RuntimeError(f"{name} must have dimensions (n, 3)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    if t.size(0) > 1 and (t.stride(1) != 1 or t.stride(0) < 3):
        t = t.contiguous()
    return t


def _idx32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.int32).contiguous()


def _args(keep, a, b, center, idx, row_scale, formula, a2=None) -> _lib.PairArgs:
    n, k = (int(v) for v in idx.shape)
    keep += [a, b, center, idx, row_scale, a2]
    p = lambda t: None if t is None or t.numel() == 0 else t.data_ptr()
    ld = lambda t: int(t.stride(0)) if t is not None and t.size(0) > 1 else 3
    return _lib.PairArgs(n, k, int(formula), p(a), ld(a), p(b), ld(b), p(center), p(idx), p(row_scale), p(a2), ld(a2))


def _check_idx(a, b, center, idx, row_scale):
    n = int(idx.size(0))
    if idx.dim() != 2 or idx.dtype != torch.int32 or not idx.is_contiguous():
        raise RuntimeError("idx must be a contiguous int32 [n,k] tensor (use descriptors.neighbour_lists)")
    if center is None and n != a.size(0):
        raise RuntimeError("idx must have one row per row of a when no centre indices are given")
    if center is not None and (center.dtype != torch.int32 or center.numel() != n):
        raise RuntimeError("center must be int32 [n]")
    if row_scale is not None and (row_scale.dtype != torch.float32 or row_scale.numel() != n):
        raise RuntimeError("row_scale must be float32 [n]")


class _PairDist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, center, idx, row_scale, formula, a2=None):
        a_, b_ = _rows3(a, "a"), _rows3(b, "b")
        a2_ = None if a2 is None else _rows3(a2, "a2")
        if a2_ is not None and a2_.size(0) != a_.size(0):
            raise RuntimeError("a2 must have as many rows as a")
        _check_idx(a_, b_, center, idx, row_scale)
        n, k = (int(v) for v in idx.shape)
        out = torch.empty((n, k), dtype=torch.float32, device=a.device)
        keep: list = []
        args = _args(keep, a_.detach(), b_.detach(), center, idx, row_scale, formula,
                     None if a2_ is None else a2_.detach())
        with torch.cuda.device(a.device):
            rc = _lib.load().wast3d_pair_dist_forward(C.byref(args), out.data_ptr() if out.numel() else None,
                                                      _lib.stream_ptr())
        _lib.check(rc, "pair_dist_forward")
        ctx.formula = formula
        ctx.same = a is b
        ctx.save_for_backward(a_, b_, center if center is not None else torch.empty(0), idx,
                              row_scale if row_scale is not None else torch.empty(0),
                              a2_ if a2_ is not None else torch.empty(0))
        return out

    @staticmethod
    def backward(ctx, grad_d):
        a, b, center, idx, row_scale, a2 = ctx.saved_tensors
        center = center if center.numel() else None
        row_scale = row_scale if row_scale.numel() else None
        a2 = a2 if a2.numel() else None
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_a2 = a2 is not None and ctx.needs_input_grad[6]
        ga = torch.zeros((a.size(0), 3), dtype=torch.float32, device=a.device) if need_a else None
        ga2 = torch.zeros((a2.size(0), 3), dtype=torch.float32, device=a.device) if need_a2 else None
        gb = ga if (ctx.same and need_a) else (torch.zeros((b.size(0), 3), dtype=torch.float32, device=a.device) if need_b else None)
        keep: list = []
        args = _args(keep, a.detach(), b.detach(), center, idx, row_scale, ctx.formula, None if a2 is None else a2.detach())
        g = grad_d.to(torch.float32).contiguous()
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(a.device):
            rc = _lib.load().wast3d_pair_dist_backward(C.byref(args), g.data_ptr() if g.numel() else None, ptr(ga),
                                                       ptr(ga2), ptr(gb), _lib.stream_ptr())
        _lib.check(rc, "pair_dist_backward")
        if ctx.same and need_a:
            return ga, None, None, None, None, None, ga2  # a is b: one buffer received both contributions
        return ga, gb, None, None, None, None, ga2


class _PairLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, center, idx, target, weight, row_scale, formula, mode, scale, a2=None):
        a_, b_ = _rows3(a, "a"), _rows3(b, "b")
        a2_ = None if a2 is None else _rows3(a2, "a2")
        if a2_ is not None and a2_.size(0) != a_.size(0):
            raise RuntimeError("a2 must have as many rows as a")
        _check_idx(a_, b_, center, idx, row_scale)
        for name, t in (("target", target), ("weight", weight)):
            if t is not None and (t.dtype != torch.float32 or tuple(t.shape) != tuple(idx.shape) or not t.is_contiguous()):
                raise RuntimeError(f"{name} must be a contiguous float32 tensor shaped like idx")
        out = torch.empty((), dtype=torch.float32, device=a.device)
        keep: list = [target, weight]
        args = _args(keep, a_.detach(), b_.detach(), center, idx, row_scale, formula,
                     None if a2_ is None else a2_.detach())
        with torch.cuda.device(a.device):
            rc = _lib.load().wast3d_pair_loss_forward(
                C.byref(args), target.data_ptr() if target.numel() else None,
                weight.data_ptr() if weight is not None and weight.numel() else None, int(mode), float(scale),
                _scratch(a.device).data_ptr(), out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "pair_loss_forward")
        ctx.cfg = (formula, int(mode), float(scale), a is b)
        e = torch.empty(0)
        ctx.save_for_backward(a_, b_, center if center is not None else e, idx, target,
                              weight if weight is not None else e, row_scale if row_scale is not None else e,
                              a2_ if a2_ is not None else e)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a, b, center, idx, target, weight, row_scale, a2 = ctx.saved_tensors
        formula, mode, scale, same = ctx.cfg
        center = center if center.numel() else None
        weight = weight if weight.numel() else None
        row_scale = row_scale if row_scale.numel() else None
        a2 = a2 if a2.numel() else None
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_a2 = a2 is not None and ctx.needs_input_grad[10]
        ga = torch.zeros((a.size(0), 3), dtype=torch.float32, device=a.device) if need_a else None
        ga2 = torch.zeros((a2.size(0), 3), dtype=torch.float32, device=a.device) if need_a2 else None
        gb = ga if (same and need_a) else (torch.zeros((b.size(0), 3), dtype=torch.float32, device=a.device) if need_b else None)
        keep: list = []
        args = _args(keep, a.detach(), b.detach(), center, idx, row_scale, formula, None if a2 is None else a2.detach())
        go = grad_out.to(torch.float32).contiguous()
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(a.device):
            rc = _lib.load().wast3d_pair_loss_backward(
                C.byref(args), target.data_ptr() if target.numel() else None, ptr(weight), mode, scale, go.data_ptr(),
                ptr(ga), ptr(ga2), ptr(gb), _lib.stream_ptr())
        _lib.check(rc, "pair_loss_backward")
        if same and need_a:
            return (ga,) + (None,) * 9 + (ga2,)
        return (ga, gb) + (None,) * 8 + (ga2,)


# ----------------------------------------------------------------------------- descriptors (rank 2)
def neighbour_lists(X: torch.Tensor, num_nns: int, kth_nn: int = 1) -> torch.Tensor:
    """`torch.topk(torch.cdist(X, X), k=num_nns, largest=False)[1][:, ::kth_nn]` (notebooks/25.4 cell 73) without the
    N x N matrix: int32 [N, ceil(num_nns / kth_nn)]; column 0 is the point itself (distance 0; among exact
    duplicates the lowest index, where torch.topk's choice is unspecified)."""
    _, idx = cdist_topk(X.detach(), X.detach(), num_nns)
    return idx[:, ::kth_nn].to(torch.int32).contiguous()


def get_descriptors(X: torch.Tensor, X_nns_indices: torch.Tensor, row_scale: torch.Tensor = None) -> torch.Tensor:
    """notebooks/25.4 cell 72 `get_descriptors`: distances from every point's first listed neighbour (itself) to its
    other listed neighbours, [N, k-1]; differentiable with respect to X.  X_nns [N,k,3] is never written.
    `row_scale` [N] (extension) multiplies row i — the `X / scaling` of notebooks/29.2 cell 70 for per-cluster
    scalings, since |x/s - y/s| = |x - y| / s."""
    idx = _idx32(X_nns_indices, X.device)
    if idx.dim() != 2 or idx.size(1) < 1:
        raise RuntimeError("X_nns_indices must have dimensions (N, k)")
    center = idx[:, 0].contiguous()
    nbr = idx[:, 1:].contiguous()
    return _PairDist.apply(X, X, center, nbr, row_scale, FORMULA_NORM)


def descriptors_loss(descriptors, descriptors_target, normalize=True):
    """notebooks/29.2 cell 69, verbatim arithmetic (the inputs are [N,k-1] tensors: small)."""
    if not normalize:
        return torch.mean(torch.square(descriptors - descriptors_target))
    descriptors_scaler = torch.mean(descriptors[:, -1])
    descriptors_target_scaler = torch.mean(descriptors_target[:, -1])
    return torch.mean(torch.square(descriptors / (1e-5 + descriptors_scaler) -
                                   descriptors_target / (1e-5 + descriptors_target_scaler)))


def descriptor_mse(X: torch.Tensor, X_nns_indices: torch.Tensor, target: torch.Tensor, row_scale=None) -> torch.Tensor:
    """`torch.mean(torch.square(get_descriptors(X, idx) - target))` in ONE kernel forward and one backward (the
    [N,k-1] descriptor tensor is not written either)."""
    idx = _idx32(X_nns_indices, X.device)
    center = idx[:, 0].contiguous()
    nbr = idx[:, 1:].contiguous()
    n_el = nbr.numel()
    tgt = target.to(torch.float32).contiguous()
    return _PairLoss.apply(X, X, center, nbr, tgt, None, row_scale, FORMULA_NORM, 1, 1.0 / max(n_el, 1))


def get_style_patch_descriptors_loss(clusters_to_opt_list: Sequence[torch.Tensor], nns_indices, target_descriptor_cluster,
                                     normalize: bool = False):
    """notebooks/25.4 cell 72 / 29.2 cell 69 with the same signature: mean over clusters of the descriptor loss.
    `nns_indices` / `target_descriptor_cluster` are either one tensor shared by all clusters or lists, one per
    cluster.  normalize=False uses the fused kernel per cluster; normalize=True needs the descriptor means and
    goes through get_descriptors + descriptors_loss.  For many clusters prefer StylePatchDescriptors (one launch)."""
    shared = not isinstance(nns_indices, (list, tuple))
    loss = 0.0
    for c, pts in enumerate(clusters_to_opt_list):
        idx = nns_indices if shared else nns_indices[c]
        tgt = target_descriptor_cluster if shared else target_descriptor_cluster[c]
        if normalize:
            loss = loss + descriptors_loss(get_descriptors(pts, idx), tgt, normalize=True)
        else:
            loss = loss + descriptor_mse(pts, idx, tgt)
    return loss / len(clusters_to_opt_list)


class StylePatchDescriptors:
    """All style-patch clusters of one optimised point set in one launch.

    notebooks/25.4 cell 73 builds, per cluster range [N_start, N_end) of `gaussians_opt._xyz`, the neighbour lists
    and target descriptors; cells 72/70 then loop over the clusters every step.  Here the per-cluster lists are
    concatenated with GLOBAL row indices and a per-pair weight 1 / (n_c (k-1) n_clusters), so that
    `loss(xyz)` == get_style_patch_descriptors_loss([xyz[s:e] / scaling_c ...], lists, targets, normalize=False)
    is one forward and one backward kernel over all pairs."""

    def __init__(self, xyz: torch.Tensor, cluster_ranges: Sequence[tuple], num_nns: int = 50, kth_nn: int = 2):
        _lib.require_device(xyz)
        xyz = xyz.detach()
        dev = xyz.device
        self.ranges = [(int(s), int(e)) for s, e in cluster_ranges]
        centers, nbrs, tgts, wts, owner = [], [], [], [], []
        ncl = len(self.ranges)
        for c, (s, e) in enumerate(self.ranges):
            pts = xyz[s:e].contiguous()
            k = min(num_nns, e - s)
            idx = neighbour_lists(pts, k, kth_nn)              # local indices, [n_c, k']
            tgt = get_descriptors(pts, idx)                    # [n_c, k'-1]
            centers.append(idx[:, 0] + s)
            nbrs.append(idx[:, 1:] + s)
            tgts.append(tgt)
            wts.append(torch.full_like(tgt, 1.0 / (max(tgt.numel(), 1) * ncl)))
            owner.append(torch.full((e - s,), c, dtype=torch.int64, device=dev))
        widths = {t.size(1) for t in nbrs}
        if len(widths) != 1:
            raise RuntimeError("StylePatchDescriptors: every cluster needs at least num_nns points")
        self.center = torch.cat(centers).to(torch.int32).contiguous()
        self.nbr = torch.cat(nbrs).to(torch.int32).contiguous()
        self.target = torch.cat(tgts).contiguous()
        self.weight = torch.cat(wts).contiguous()
        self.owner = torch.cat(owner)

    def descriptors(self, xyz: torch.Tensor, cluster_scalings: torch.Tensor = None) -> torch.Tensor:
        rs = None if cluster_scalings is None else (1.0 / cluster_scalings.to(torch.float32))[self.owner].contiguous()
        return _PairDist.apply(xyz, xyz, self.center, self.nbr, rs, FORMULA_NORM)

    def loss(self, xyz: torch.Tensor, cluster_scalings: torch.Tensor = None) -> torch.Tensor:
        rs = None if cluster_scalings is None else (1.0 / cluster_scalings.to(torch.float32))[self.owner].contiguous()
        return _PairLoss.apply(xyz, xyz, self.center, self.nbr, self.target, self.weight, rs, FORMULA_NORM, 1, 1.0)

    @torch.no_grad()
    def scalings(self, xyz: torch.Tensor, lo: float = 0.02, hi: float = 3.0) -> torch.Tensor:
        """notebooks/29.2 cell 70: per point mean(current descriptors) / (1e-8 + mean(target descriptors)), clipped."""
        cur = self.descriptors(xyz.detach())
        return torch.clip(cur.mean(-1) / (1e-8 + self.target.mean(-1)), lo, hi).unsqueeze(1)


# ----------------------------------------------------------------------------- masked cdist L1 (rank 3)
class KnnMaskPairs:
    """Sparse form of `D_target = cdist(xa, xb)` and `mask = D_target <= sort(D_target, 1)[:, k-1:k]`
    (aux_optimize_cluster_D_W_distance.py:70-82): per row the k nearest columns, their target distances and a
    0/1 weight.  `tie_slack` extra candidates per row are fetched so that every column tied with the k-th distance
    is inside the mask, as in the reference; a row with more ties than that raises."""

    def __init__(self, xa_target: torch.Tensor, xb_target: torch.Tensor, k: int = 10, tie_slack: int = 8):
        nb = int(xb_target.size(0))
        kk = min(nb, k + tie_slack)
        vals, idx = cdist_topk(xa_target.detach().contiguous(), xb_target.detach().contiguous(), kk)
        kth = vals[:, k - 1:k]
        inside = vals <= kth
        if kk < nb and bool(inside[:, -1].any()):
            raise RuntimeError("KnnMaskPairs: more ties at the k-th distance than tie_slack; raise tie_slack")
        self.k = k
        self.idx = idx.to(torch.int32).contiguous()
        self.weight = inside.to(torch.float32).contiguous()
        self.n_rows, self.n_cols = int(xa_target.size(0)), nb
        self.xb_target = xb_target.detach()

    def targets(self, a_target: torch.Tensor, a2_target: torch.Tensor = None) -> torch.Tensor:
        """Entries of `torch.cdist(a_target, xb_target)` [+ `torch.cdist(a2_target, xb_target)`] at the mask positions
        ([n, k'] float32) — the D_*_target matrices of :72-76 restricted to the mask (two operands: D_rotation_target)."""
        with torch.no_grad():
            return _PairDist.apply(a_target.detach(), self.xb_target, None, self.idx, None, FORMULA_CDIST,
                                   None if a2_target is None else a2_target.detach()).contiguous()


def masked_cdist_l1(a: torch.Tensor, b: torch.Tensor, pairs: KnnMaskPairs, target: torch.Tensor,
                    a2: torch.Tensor = None) -> torch.Tensor:
    """`torch.mean(torch.abs(torch.cdist(a, b) - D_target) * mask)` (aux_optimize_cluster_D_W_distance.py:278-280):
    the mean runs over the full n x m matrix, only the masked pairs are evaluated.  With `a2` the matrix is
    `cdist(a, b) + cdist(a2, b)` (D_rotation, :254-255).  Differentiable with respect to a, a2 and b (strided [N,3]
    views such as `_rotation[:, :-1]` are read in place)."""
    if a.size(0) != pairs.n_rows or b.size(0) != pairs.n_cols:
        raise RuntimeError("masked_cdist_l1: a / b do not have the shape the mask was built for")
    scale = 1.0 / (float(pairs.n_rows) * float(pairs.n_cols))
    return _PairLoss.apply(a, b, None, pairs.idx, target, pairs.weight, None, FORMULA_CDIST, 0, scale, a2)
