"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float32) of the reference's geometry regularisers and of
its K-Means step, for SURVEY.md §8f ranks 2-4.  Nothing under wast3d_b200/ imports this.

What is restated (paths relative to the reference root):
  * masked pairwise-distance L1  aux_optimize_cluster_D_W_distance.py:70-82 (targets + kNN mask),
                                 :253-256 (D matrices), :278-280 (loss terms)
  * local kNN descriptors        notebooks/25.4.Optimize_with_SAM_masks_clean.ipynb cell 72 (get_descriptors,
                                 get_style_patch_descriptors_loss), cell 73 (topk(num_nns)[:, ::kth_nn]);
                                 notebooks/29.2.Modify_style_clusters.ipynb cells 69-70 (normalised variant,
                                 per-step scale re-derivation)
  * K-Means                      aux_save_clusters_clean.py:32-47, train_st.py:54-70 (sklearn.cluster.KMeans:
                                 un-vendored, unpinned, unseeded init -> Lloyd from a GIVEN initialisation)

Pinning: the dense expressions are the reference's own torch code; tests/test_pairs_oracle.py evaluates them
verbatim with torch on CPU (torch is the library the reference calls) and checks these sparse restatements
against them, and checks `kmeans_lloyd` against scikit-learn's KMeans(init=<array>, n_init=1, algorithm="lloyd")
(installed here).  POT / sklearn versions are not pinned by the reference.
"""
from __future__ import annotations

import numpy as np

from . import cpu

F = np.float32


def cdist_entries(a, b, rows, cols):
    """Entries [rows[e], cols[e]] of torch.cdist(a, b) evaluated in its matmul-path order (oracle_cdist_sq,
    match_oracle.c): sqrt(max(0, fma chain of -2 a.b + |a|^2 + |b|^2))."""
    a = np.ascontiguousarray(a, F)
    b = np.ascontiguousarray(b, F)
    full = cpu.cdist(a, b, sqrt=True)  # dense, C restatement; test sizes only
    return full[rows, cols].astype(F)


def norm_entries(a, b, rows, cols):
    """|a[rows] - b[cols]| as torch.norm(x - y, dim=-1) on 3 components: sqrt((dx^2 + dy^2) + dz^2), float32."""
    a = np.asarray(a, F)
    b = np.asarray(b, F)
    d = (a[rows] - b[cols]).astype(F)
    sq = (d[:, 0] * d[:, 0]).astype(F)
    sq = (sq + (d[:, 1] * d[:, 1]).astype(F)).astype(F)
    sq = (sq + (d[:, 2] * d[:, 2]).astype(F)).astype(F)
    return np.sqrt(sq).astype(F)


def knn_mask_pairs(xa_target, xb_target, k):
    """Sparse form of the reference's mask (aux_optimize_cluster_D_W_distance.py:70-82):
        D = cdist(xa, xb); mask = D <= sort(D, 1)[:, k-1:k]
    Returns (rows, cols, target) of the masked entries in row-major order (ties at the k-th distance are all
    inside the mask, as in the reference)."""
    D = cpu.cdist(np.ascontiguousarray(xa_target, F), np.ascontiguousarray(xb_target, F), sqrt=True)
    kth = np.sort(D, axis=1)[:, k - 1:k]
    rows, cols = np.nonzero(D <= kth)
    return rows, cols, D[rows, cols].astype(F)


def masked_l1(a, b, rows, cols, target, n_rows, n_cols, a2=None):
    """torch.mean(torch.abs(D - D_target) * mask) over the FULL n_rows x n_cols matrix (:278-280) with
    D = cdist(a, b) [+ cdist(a2, b) for the rotation term, :254-255], and its gradients with respect to a, b
    (and a2) — torch.cdist backward: g (a - b) / d, 0 where d == 0.  Returns (loss, ga, gb[, ga2])."""
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d1 = cdist_entries(a, b, rows, cols)
    d = d1.copy()
    if a2 is not None:
        d2 = cdist_entries(a2, b, rows, cols)
        d = (d1 + d2).astype(F)
    r = d.astype(np.float64) - np.asarray(target, np.float64)
    scale = 1.0 / (float(n_rows) * float(n_cols))
    loss = np.abs(r).sum() * scale
    g = np.sign(r) * scale

    def grads(x64, dist):
        with np.errstate(divide="ignore", invalid="ignore"):
            c = np.where(dist > 0, g / dist.astype(np.float64), 0.0)
        diff = x64[rows] - b64[cols]
        gx = np.zeros_like(x64)
        gb = np.zeros_like(b64)
        np.add.at(gx, rows, c[:, None] * diff)
        np.add.at(gb, cols, -c[:, None] * diff)
        return gx, gb

    ga, gb = grads(a64, d1)
    if a2 is None:
        return F(loss), ga.astype(F), gb.astype(F)
    ga2, gb2 = grads(np.asarray(a2, np.float64), d2)
    return F(loss), ga.astype(F), (gb + gb2).astype(F), ga2.astype(F)


def get_descriptors(X, nns_indices):
    """notebooks/25.4 cell 72: X_nns = X[idx]; norm(X_nns[:,1:] - X_nns[:,0].unsqueeze(1), dim=-1) -> [N, k-1]."""
    idx = np.asarray(nns_indices, np.int64)
    n, k = idx.shape
    rows = np.repeat(idx[:, 0], k - 1)
    cols = idx[:, 1:].reshape(-1)
    return norm_entries(X, X, rows, cols).reshape(n, k - 1)


def descriptor_mse(X, nns_indices, target):
    """torch.mean(torch.square(get_descriptors(X, idx) - target)) and its gradient with respect to X."""
    idx = np.asarray(nns_indices, np.int64)
    n, k = idx.shape
    d = get_descriptors(X, idx).astype(np.float64)
    r = d - np.asarray(target, np.float64)
    loss = (r * r).mean()
    g = 2.0 * r / r.size
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where(d > 0, g / d, 0.0)
    X64 = np.asarray(X, np.float64)
    centre = X64[idx[:, 0]][:, None, :]
    diff = centre - X64[idx[:, 1:]]            # a - b with a = centre row
    gX = np.zeros_like(X64)
    np.add.at(gX, np.repeat(idx[:, 0], k - 1), (c[..., None] * diff).reshape(-1, 3))
    np.add.at(gX, idx[:, 1:].reshape(-1), (-c[..., None] * diff).reshape(-1, 3))
    return F(loss), gX.astype(F)


def kmeans_lloyd(points, init_centers, max_iter, tol=0.0):
    """Lloyd iterations from `init_centers`: E-step = argmin of torch.cdist in the oracle's operation order
    (ties to the lowest centre), M-step = mean of the members in float64 rounded to float32 (an empty cluster
    keeps its centre).  Stops after max_iter M-steps, when no label changed, or when the summed squared centre
    shift of the last M-step is <= tol; always ends on an E-step against the returned centres.
    Returns (labels int32, centres float32, inertia float64, n_iter)."""
    x = np.ascontiguousarray(points, F)
    c = np.ascontiguousarray(init_centers, F).copy()
    K = c.shape[0]
    labels = np.full(x.shape[0], -1, np.int32)
    it, shift = 0, 0.0
    while True:
        new, dist = cpu.nn_match(x, c)
        changed = int((new != labels).sum())
        labels = new.astype(np.int32)
        inertia = float((dist.astype(np.float64) ** 2).sum())
        if it >= max_iter or changed == 0 or (it > 0 and shift <= tol):
            break
        cnt = np.bincount(labels, minlength=K)
        s = np.zeros((K, 3), np.float64)
        np.add.at(s, labels, x.astype(np.float64))
        nc = c.copy()
        nz = cnt > 0
        nc[nz] = (s[nz] / cnt[nz, None]).astype(F)
        shift = float(((nc.astype(np.float64) - c.astype(np.float64)) ** 2).sum())
        c = nc
        it += 1
    return labels, c, inertia, it
