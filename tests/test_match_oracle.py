"""CPU: pins the matching oracle against the reference's own arithmetic (torch.cdist) and the
textbook W2 formula in float64."""
import numpy as np
import pytest
import torch

from oracle import cpu


def _torch_mm_path(a, b):
    x1 = torch.cat([a * -2, a.pow(2).sum(-1, keepdim=True), torch.ones(a.shape[0], 1)], -1)
    x2 = torch.cat([b, torch.ones(b.shape[0], 1), b.pow(2).sum(-1, keepdim=True)], -1)
    return x1.matmul(x2.mT).clamp_min(0)


@pytest.mark.parametrize("n,m,seed", [(64, 64, 0), (300, 1000, 1), (2000, 513, 2)])
def test_cdist_squared_is_bit_identical_to_torch(built, n, m, seed):
    torch.manual_seed(seed)
    a, b = torch.randn(n, 3) * 2, torch.randn(m, 3) * 1.5 + 0.3
    ours = cpu.cdist(a.numpy(), b.numpy(), sqrt=False)
    assert (ours == _torch_mm_path(a, b).numpy()).all()
    # and torch.cdist itself is that product followed by its (not correctly rounded) sqrt
    assert (torch.cdist(a, b) == _torch_mm_path(a, b).sqrt()).all()


def test_nn_match_equals_torch_argmin_up_to_sqrt_ties(built):
    torch.manual_seed(3)
    a, b = torch.randn(5000, 3), torch.randn(2000, 3)
    idx, dist = cpu.nn_match(a.numpy(), b.numpy())
    D = torch.cdist(a, b)
    t = D.argmin(1).numpy()
    diff = np.nonzero(t != idx)[0]
    sq = cpu.cdist(a.numpy(), b.numpy(), sqrt=False)
    for i in diff:  # any disagreement must be a tie created/broken by sqrt rounding (<= 4 ulp apart)
        x, y = sq[i, t[i]], sq[i, idx[i]]
        assert abs(float(x) - float(y)) <= 4 * np.spacing(np.float32(max(x, y)))
    assert len(diff) <= 5
    np.testing.assert_allclose(dist, D.min(1).values.numpy(), rtol=3e-7)


def test_nn_match_ties_go_to_lowest_index(built):
    b = np.array([[1, 0, 0], [0, 1, 0], [1, 0, 0], [-1, 0, 0]], np.float32)
    a = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    idx, dist = cpu.nn_match(a, b)
    assert idx.tolist() == [0, 0] and dist.tolist() == [1.0, 0.0]


def _sqrtm_psd(S):
    w, v = np.linalg.eigh(S)
    return (v * np.sqrt(np.maximum(w, 0))) @ v.T


def _w2_f64(m1, S1, m2, S2):
    r1 = _sqrtm_psd(S1)
    w = np.linalg.eigvalsh(r1 @ S2 @ r1)
    return ((m1 - m2) ** 2).sum() + np.trace(S1) + np.trace(S2) - 2 * np.sqrt(np.maximum(w, 0)).sum()


def _rand_cov(rng, n):
    A = rng.normal(size=(n, 3, 3)) * rng.uniform(0.05, 1.0, size=(n, 1, 3))
    return A @ A.transpose(0, 2, 1)


def _c6(S):
    return np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)


def test_w2_cost_matches_float64_formula(built):
    rng = np.random.default_rng(0)
    n = 40
    S1, S2 = _rand_cov(rng, n), _rand_cov(rng, n)
    m1, m2 = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    idx, cost, mat = cpu.w2_match(m1, _c6(S1), m2, _c6(S2), want_matrix=True)
    ref = np.array([[_w2_f64(m1[i], S1[i], m2[j], S2[j]) for j in range(n)] for i in range(n)])
    assert np.abs(mat - ref).max() <= 2e-6 * np.abs(ref).max()
    assert (ref.argmin(1) == idx).all()
    assert (cost == mat[np.arange(n), idx]).all()


def test_w2_special_cases(built):
    I6 = np.array([[1, 0, 0, 1, 0, 1]], np.float32)
    z3 = np.zeros((1, 3), np.float32)
    # identical Gaussians -> 0 ; isotropic: W2^2 = |dm|^2 + 3 (s1 - s2)^2
    _, c = cpu.w2_match(z3, I6, z3, I6)
    assert abs(c[0]) <= 1e-6
    _, c = cpu.w2_match(z3, I6 * 4.0, z3 + np.array([[1, 2, 2]], np.float32), I6)
    assert abs(c[0] - (9.0 + 3.0 * (2.0 - 1.0) ** 2)) <= 1e-5
    # degenerate (rank-deficient) covariances stay finite and close to the float64 value
    rng = np.random.default_rng(1)
    S = _rand_cov(rng, 8)
    for i in range(8):
        w, v = np.linalg.eigh(S[i]); w[0] = 0.0
        if i % 2: w[1] = 0.0
        S[i] = (v * w) @ v.T
    S2 = _rand_cov(rng, 8); m = rng.normal(size=(8, 3))
    _, _, mat = cpu.w2_match(m, _c6(S), m[::-1].copy(), _c6(S2), want_matrix=True)
    ref = np.array([[_w2_f64(m[i], S[i], m[::-1][j], S2[j]) for j in range(8)] for i in range(8)])
    assert np.isfinite(mat).all() and np.abs(mat - ref).max() <= 5e-3 * np.abs(ref).max()


def test_cluster_stats(built):
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(5000, 3)).astype(np.float32)
    lab = rng.integers(0, 17, size=5000)
    mean, cov, count = cpu.cluster_stats(pts, lab, 20)  # clusters 17..19 are empty
    for k in range(17):
        sel = pts[lab == k].astype(np.float64)
        assert count[k] == len(sel)
        np.testing.assert_allclose(mean[k], sel.mean(0), atol=1e-6)
        C = np.cov(sel.T, bias=True)
        np.testing.assert_allclose(cov[k], [C[0, 0], C[0, 1], C[0, 2], C[1, 1], C[1, 2], C[2, 2]], atol=1e-6)
    assert (count[17:] == 0).all() and (cov[17:] == 0).all()


def test_knn_oracle_small_cases(built):
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0]], np.float32)  # P < 4: FLT_MAX terms
    d, idx = cpu.knn(pts)
    assert np.isinf(d).all() or (d > 1e37).all()
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [0, 0, 0]], np.float32)
    d, idx = cpu.knn(pts)
    assert d[0] == np.float32((0 + 1 + 4) / 3.0) and idx[0].tolist() == [4, 1, 2]


@pytest.mark.parametrize("n,m,k,seed", [(50, 50, 10, 0), (300, 1000, 20, 1), (64, 4893, 100, 2), (10, 128, 128, 3)])
def test_cdist_topk_equals_torch_stable_sort(built, n, m, k, seed):
    """M2 oracle pinned on torch: k smallest of torch.cdist rows in stable-sort order
    (aux_optimize_cluster_D_W_distance.py:79-82 uses sort, notebooks/25.4 cell 73 uses topk)."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, 3)).astype(np.float32)
    b = rng.normal(size=(m, 3)).astype(np.float32)
    b[m // 2] = b[m // 3]  # an exact tie in every row
    d, i = cpu.cdist_topk(a, b, k)
    D = torch.cdist(torch.from_numpy(a), torch.from_numpy(b))
    v, ix = torch.sort(D, dim=1, stable=True)
    # torch's vectorised CPU sqrt is not correctly rounded (see oracle header): compare indices on
    # rows whose k+1 smallest squared distances are well separated, values to 1 ulp everywhere
    np.testing.assert_allclose(d, v[:, :k].numpy(), rtol=2e-7, atol=0)
    same = (i == ix[:, :k].numpy()).all(axis=1)
    assert same.mean() > 0.95
    for r in np.nonzero(~same)[0]:  # any disagreement must be between distances equal to 1 ulp
        bad = i[r] != ix[r, :k].numpy()
        assert np.abs(d[r][bad] - v[r, :k].numpy()[bad]).max() <= 2e-7 * d[r][bad].max()
    # the reference's mask D <= kth value == membership in our top-k, plus ties at the k-th value
    mask = (D <= v[:, k - 1:k]).numpy()
    ours = np.zeros_like(mask)
    np.put_along_axis(ours, i.astype(np.int64), True, axis=1)
    assert (mask | ours == mask).all() and (mask.sum(1) >= k).all()


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (100, 3), (100, 4), (300, 5)])
def test_emd2_uniform_oracle_is_optimal(built, n, seed):
    """M4: the restated ot.emd2 (uniform weights => assignment problem) against an independent exact
    solver (scipy.optimize.linear_sum_assignment) on the same ground-cost matrix, and against brute
    force for tiny n.  POT itself is not available: parity unpinned, optimality is what is checked."""
    from itertools import permutations
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, 3)).astype(np.float32)
    b = (rng.normal(size=(n, 3)) * 1.5 + 0.3).astype(np.float32)
    if n >= 7:
        b[1] = b[0]  # duplicate target points: several optimal plans, one optimal value
    cost, perm, M = cpu.emd2_uniform(a, b, want_matrix=True)
    assert sorted(perm.tolist()) == list(range(n))
    # ground cost == squared Euclidean distances (ot.dist default metric)
    D2 = ((a[:, None, :].astype(np.float64) - b[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(M, D2, rtol=1e-5, atol=1e-5)
    r, c = linear_sum_assignment(M.astype(np.float64))
    ref = M.astype(np.float64)[r, c].sum() / n
    assert abs(cost - ref) <= 1e-6 * max(ref, 1e-12)
    assert abs(M.astype(np.float64)[np.arange(n), perm].sum() / n - ref) <= 1e-9 * max(ref, 1e-12)
    if n <= 7:
        best = min(sum(M[i, s[i]] for i in range(n)) for s in permutations(range(n))) / n
        assert abs(cost - best) <= 1e-6 * max(best, 1e-12)
