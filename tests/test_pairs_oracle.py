"""CPU: pins oracle/pairs.py (sparse restatements of the geometry regularisers, Lloyd K-Means) against the
reference's own dense torch expressions evaluated verbatim on CPU, and against scikit-learn's KMeans."""
import numpy as np
import pytest
import torch

from oracle import cpu, pairs


def _scene(n, seed, dup=False):
    rng = np.random.default_rng(seed)
    xyz = (rng.normal(size=(n, 3)) * [1.0, 0.6, 0.3] + rng.integers(0, 3, size=(n, 1))).astype(np.float32)
    if dup:
        xyz[5:9] = xyz[4]          # exact duplicates: ties at distance 0
    rot = rng.normal(size=(n, 4)).astype(np.float32)
    scl = rng.normal(size=(n, 3)).astype(np.float32) - 3.0
    return xyz, rot, scl


# ---- aux_optimize_cluster_D_W_distance.py:70-82, :253-256, :278-280, verbatim torch ------------------------
def _reference_dense_losses(tx, trot, tscl, x, rot, scl, k):
    D_xyz_target = torch.cdist(tx, tx)
    D_rotation_target = torch.cdist(trot[:, :-1], tx) + torch.cdist(trot[:, 1:], tx)
    D_scaling_target = torch.cdist(tscl, tx)
    sorted_values, _ = torch.sort(D_xyz_target, dim=1)
    mask = (D_xyz_target <= sorted_values[:, k - 1:k]).to(dtype=torch.float32)
    D_xyz = torch.cdist(x, x)
    D_scaling = torch.cdist(scl, x)
    loss_D_xyz = torch.mean(torch.abs(D_xyz - D_xyz_target) * mask)
    loss_D_scaling = torch.mean(torch.abs(D_scaling - D_scaling_target) * mask)
    # the rotation term is a SUM of two cdist matrices inside one abs(); its sparse form needs both entries
    D_rotation = torch.cdist(rot[:, :-1], x) + torch.cdist(rot[:, 1:], x)
    loss_D_rotation = torch.mean(torch.abs(D_rotation - D_rotation_target) * mask)
    return mask, loss_D_xyz, loss_D_scaling, loss_D_rotation


@pytest.mark.parametrize("n,k,dup", [(200, 10, False), (333, 10, True), (64, 3, False)])
def test_masked_l1_sparse_form_equals_reference_dense_expression(built, n, k, dup):
    txyz, trot, tscl = _scene(n, 1, dup)
    xyz, rot, scl = _scene(n, 2)
    T = lambda a: torch.from_numpy(a).clone().requires_grad_(True)
    x, r, s = T(xyz), T(rot), T(scl)
    mask, l_xyz, l_scl, l_rot = _reference_dense_losses(torch.from_numpy(txyz), torch.from_numpy(trot),
                                                    torch.from_numpy(tscl), x, r, s, k)
    rows, cols, tgt = pairs.knn_mask_pairs(txyz, txyz, k)
    # the sparse mask is the dense mask
    dense = np.zeros((n, n), np.float32)
    dense[rows, cols] = 1
    assert (dense == mask.numpy()).all()
    assert (np.bincount(rows, minlength=n) >= k).all()
    # xyz term: a = b = xyz
    (l_xyz + 0 * l_scl).backward(retain_graph=True)
    loss, ga, gb = pairs.masked_l1(xyz, xyz, rows, cols, tgt, n, n)
    np.testing.assert_allclose(loss, l_xyz.item(), rtol=2e-5)
    np.testing.assert_allclose(ga + gb, x.grad.numpy(), rtol=2e-3, atol=2e-8)
    # scaling term: a = scaling, b = xyz, target = cdist(scaling_target, xyz_target) at the mask
    x.grad = None
    l_scl.backward()
    tgt_s = pairs.cdist_entries(tscl, txyz, rows, cols)
    loss, ga, gb = pairs.masked_l1(scl, xyz, rows, cols, tgt_s, n, n)
    np.testing.assert_allclose(loss, l_scl.item(), rtol=2e-5)
    np.testing.assert_allclose(ga, s.grad.numpy(), rtol=2e-3, atol=2e-8)
    np.testing.assert_allclose(gb, x.grad.numpy(), rtol=2e-3, atol=2e-8)
    # rotation term: D = cdist(rot[:, :-1], xyz) + cdist(rot[:, 1:], xyz) inside ONE abs()
    x.grad = None
    l_rot.backward()
    tgt_r = (pairs.cdist_entries(trot[:, :-1], txyz, rows, cols) + pairs.cdist_entries(trot[:, 1:], txyz, rows, cols))
    loss, ga, gb, ga2 = pairs.masked_l1(rot[:, :-1], xyz, rows, cols, tgt_r, n, n, a2=rot[:, 1:])
    np.testing.assert_allclose(loss, l_rot.item(), rtol=2e-5)
    grot = np.zeros_like(rot)
    grot[:, :-1] += ga
    grot[:, 1:] += ga2
    np.testing.assert_allclose(grot, r.grad.numpy(), rtol=2e-3, atol=2e-8)
    np.testing.assert_allclose(gb, x.grad.numpy(), rtol=2e-3, atol=2e-8)


# ---- notebooks/25.4 cells 72-73, verbatim torch ---------------------------------------------------------------
def _reference_get_descriptors(X, X_nns_indices):
    X_nns = X[X_nns_indices]
    return torch.norm(X_nns[:, 1:] - X_nns[:, 0].unsqueeze(1), dim=-1)


@pytest.mark.parametrize("n,num_nns,kth", [(300, 50, 2), (120, 10, 1), (51, 50, 7)])
def test_descriptors_equal_reference_expression(built, n, num_nns, kth):
    xyz, _, _ = _scene(n, 5)
    X = torch.from_numpy(xyz)
    distances = torch.cdist(X, X)
    _, nns = torch.topk(distances, k=num_nns, largest=False, dim=-1)
    nns = nns[:, ::kth]
    tgt = _reference_get_descriptors(X, nns).clone()
    ours = pairs.get_descriptors(xyz, nns.numpy())
    np.testing.assert_allclose(ours, tgt.numpy(), rtol=3e-7, atol=0)
    # neighbour lists from the oracle's top-k equal torch.topk's wherever distances are distinct
    _, oidx = cpu.cdist_topk(xyz, xyz, num_nns)
    same = (oidx[:, ::kth] == nns.numpy())
    assert same.mean() > 0.999
    # loss + gradient against autograd of the reference expression, on a perturbed copy
    Y = (X + 0.05 * torch.randn(n, 3, generator=torch.Generator().manual_seed(0))).requires_grad_(True)
    ref = torch.mean(torch.square(_reference_get_descriptors(Y, nns) - tgt))
    ref.backward()
    loss, gX = pairs.descriptor_mse(Y.detach().numpy(), nns.numpy(), tgt.numpy())
    np.testing.assert_allclose(loss, ref.item(), rtol=2e-5)
    np.testing.assert_allclose(gX, Y.grad.numpy(), rtol=2e-3, atol=1e-9)


# ---- K-Means against scikit-learn ------------------------------------------------------------------------------
@pytest.mark.parametrize("n,K,seed", [(3000, 30, 0), (5000, 500, 1), (400, 3, 2)])
def test_kmeans_lloyd_matches_sklearn_from_the_same_init(built, n, K, seed):
    sklearn_cluster = pytest.importorskip("sklearn.cluster")
    rng = np.random.default_rng(seed)
    blobs = rng.normal(size=(max(K // 4, 3), 3)) * 4
    x = (blobs[rng.integers(0, len(blobs), n)] + rng.normal(size=(n, 3))).astype(np.float32)
    init = x[rng.choice(n, K, replace=False)].copy()
    labels, centers, inertia, n_iter = pairs.kmeans_lloyd(x, init, max_iter=30, tol=0.0)
    km = sklearn_cluster.KMeans(n_clusters=K, init=init, n_init=1, max_iter=30, tol=0.0, algorithm="lloyd").fit(x)
    # same algorithm, different rounding (sklearn expands |x|^2 - 2 x.c in BLAS chunks): trajectories can part at
    # near-ties, so compare outcomes, not bits
    agree = (labels == km.labels_).mean()
    assert agree > 0.97, agree
    assert abs(inertia - km.inertia_) <= 2e-3 * km.inertia_
    # internal consistency: labels are the nearest centres, centres are member means (Lloyd fixed point or budget)
    idx, _ = cpu.nn_match(x, centers)
    assert (idx == labels).all()
    if n_iter < 30:
        for k in np.unique(labels):
            np.testing.assert_allclose(centers[k], x[labels == k].astype(np.float64).mean(0), rtol=1e-5, atol=1e-6)


def test_kmeans_empty_cluster_keeps_its_centre(built):
    x = np.array([[0, 0, 0], [0.1, 0, 0], [5, 5, 5], [5.1, 5, 5]], np.float32)
    init = np.array([[0, 0, 0], [5, 5, 5], [100, 100, 100]], np.float32)
    labels, centers, inertia, n_iter = pairs.kmeans_lloyd(x, init, max_iter=10)
    assert (labels == [0, 0, 1, 1]).all()
    assert (centers[2] == init[2]).all()
    np.testing.assert_allclose(centers[0], [0.05, 0, 0], atol=1e-7)
