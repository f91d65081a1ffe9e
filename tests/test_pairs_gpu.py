"""GPU parity for SURVEY.md §8f ranks 2-4 through the C ABI: sparse pairwise-distance terms (descriptors, masked
cdist L1) against the oracle and against the reference's dense torch expressions, and Lloyd K-Means against the
oracle (labels bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(n, seed, dup=False):
    rng = np.random.default_rng(seed)
    xyz = (rng.normal(size=(n, 3)) * [1.0, 0.6, 0.3] + rng.integers(0, 3, size=(n, 1))).astype(np.float32)
    if dup:
        xyz[5:9] = xyz[4]
    rot = rng.normal(size=(n, 4)).astype(np.float32)
    scl = rng.normal(size=(n, 3)).astype(np.float32) - 3.0
    return xyz, rot, scl


def _cuda(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_(True) if grad else t


# --------------------------------------------------------------------------- rank 3: masked cdist L1
@pytest.mark.parametrize("n,k,dup", [(200, 10, False), (333, 10, True), (64, 3, False), (1500, 10, False)])
def test_masked_cdist_l1_matches_oracle_and_dense_reference(built, n, k, dup):
    """aux_optimize_cluster_D_W_distance.py:70-82, :253-256, :278-280 — all three loss terms, values and gradients.
    Bars: mask / targets bit-exact with the oracle; loss <= 2e-6 relative (fp32 sums, deterministic on our side);
    gradients <= 1e-4 relative L2 (unordered float atomics)."""
    from oracle import pairs as opairs
    from tests.util import rel_l2
    from wast3d_b200.descriptors import KnnMaskPairs, masked_cdist_l1
    txyz, trot, tscl = _scene(n, 1, dup)
    xyz, rot, scl = _scene(n, 2)
    rows, cols, tgt = opairs.knn_mask_pairs(txyz, txyz, k)
    mp = KnnMaskPairs(_cuda(txyz), _cuda(txyz), k=k)
    # sparse mask == oracle mask (bit-exact, including the duplicates' ties)
    w = mp.weight.cpu().numpy() > 0
    got = sorted(zip(np.repeat(np.arange(n), mp.idx.size(1))[w.reshape(-1)].tolist(), mp.idx.cpu().numpy()[w].tolist()))
    assert got == sorted(zip(rows.tolist(), cols.tolist()))
    t_xyz = mp.targets(_cuda(txyz))
    from oracle import cpu
    D = cpu.cdist(txyz, txyz)
    assert (t_xyz.cpu().numpy() == np.take_along_axis(D, mp.idx.cpu().numpy().astype(np.int64), 1)).all()

    x, r, s = _cuda(xyz, True), _cuda(rot, True), _cuda(scl, True)
    # xyz term (a is b)
    l = masked_cdist_l1(x, x, mp, t_xyz)
    l.backward()
    ol, oga, ogb = opairs.masked_l1(xyz, xyz, rows, cols, tgt, n, n)
    np.testing.assert_allclose(l.item(), ol, rtol=2e-6)
    assert rel_l2(x.grad.cpu(), torch.from_numpy(oga + ogb)) <= 1e-4
    # scaling term
    x.grad = None
    t_s = mp.targets(_cuda(tscl))
    l = masked_cdist_l1(s, x, mp, t_s)
    l.backward()
    ol, oga, ogb = opairs.masked_l1(scl, xyz, rows, cols, opairs.cdist_entries(tscl, txyz, rows, cols), n, n)
    np.testing.assert_allclose(l.item(), ol, rtol=2e-6)
    assert rel_l2(s.grad.cpu(), torch.from_numpy(oga)) <= 1e-4
    assert rel_l2(x.grad.cpu(), torch.from_numpy(ogb)) <= 1e-4
    # rotation term: strided views of the [n,4] quaternion tensor, two operands inside one abs()
    x.grad = None
    t_r = mp.targets(_cuda(trot)[:, :-1], _cuda(trot)[:, 1:])
    l = masked_cdist_l1(r[:, :-1], x, mp, t_r, a2=r[:, 1:])
    l.backward()
    tgt_r = opairs.cdist_entries(trot[:, :-1], txyz, rows, cols) + opairs.cdist_entries(trot[:, 1:], txyz, rows, cols)
    ol, oga, ogb, oga2 = opairs.masked_l1(rot[:, :-1], xyz, rows, cols, tgt_r, n, n, a2=rot[:, 1:])
    np.testing.assert_allclose(l.item(), ol, rtol=2e-6)
    grot = np.zeros_like(rot)
    grot[:, :-1] += oga
    grot[:, 1:] += oga2
    assert rel_l2(r.grad.cpu(), torch.from_numpy(grot)) <= 1e-4
    assert rel_l2(x.grad.cpu(), torch.from_numpy(ogb)) <= 1e-4

    # and the reference's dense expression itself (torch on CPU), for the three terms at once
    if n <= 400:
        cx, cr, cs = (torch.from_numpy(v).clone().requires_grad_(True) for v in (xyz, rot, scl))
        tx, tr, ts = torch.from_numpy(txyz), torch.from_numpy(trot), torch.from_numpy(tscl)
        Dt = torch.cdist(tx, tx)
        mask = (Dt <= torch.sort(Dt, dim=1)[0][:, k - 1:k]).float()
        ref = (torch.mean(torch.abs(torch.cdist(cx, cx) - Dt) * mask) +
               torch.mean(torch.abs(torch.cdist(cr[:, :-1], cx) + torch.cdist(cr[:, 1:], cx)
                                    - (torch.cdist(tr[:, :-1], tx) + torch.cdist(tr[:, 1:], tx))) * mask) +
               torch.mean(torch.abs(torch.cdist(cs, cx) - torch.cdist(ts, tx)) * mask))
        ref.backward()
        x.grad = r.grad = s.grad = None
        ours = (masked_cdist_l1(x, x, mp, t_xyz) + masked_cdist_l1(r[:, :-1], x, mp, t_r, a2=r[:, 1:]) +
                masked_cdist_l1(s, x, mp, t_s))
        ours.backward()
        np.testing.assert_allclose(ours.item(), ref.item(), rtol=2e-5)
        for g, c in ((x.grad, cx.grad), (r.grad, cr.grad), (s.grad, cs.grad)):
            assert rel_l2(g.cpu(), c) <= 2e-3


def test_knn_mask_pairs_refuses_more_ties_than_slack(built):
    from wast3d_b200.descriptors import KnnMaskPairs
    pts = np.zeros((40, 3), np.float32)            # all points coincide: every distance ties
    with pytest.raises(RuntimeError):
        KnnMaskPairs(_cuda(pts), _cuda(pts), k=10, tie_slack=8)
    KnnMaskPairs(_cuda(pts), _cuda(pts), k=10, tie_slack=30)   # enough slack: whole rows are inside the mask


# --------------------------------------------------------------------------- rank 2: descriptors
@pytest.mark.parametrize("n,num_nns,kth", [(300, 50, 2), (120, 10, 1), (51, 50, 7), (4000, 50, 2)])
def test_descriptors_match_oracle_and_reference_expression(built, n, num_nns, kth):
    """notebooks/25.4 cells 72-73.  Bars: neighbour lists and descriptor values bit-exact with the oracle; the fused
    MSE <= 2e-6 relative; gradients <= 1e-4 relative L2."""
    from oracle import cpu, pairs as opairs
    from tests.util import rel_l2
    from wast3d_b200.descriptors import descriptor_mse, descriptors_loss, get_descriptors, neighbour_lists
    xyz, _, _ = _scene(n, 5)
    X = _cuda(xyz)
    idx = neighbour_lists(X, num_nns, kth)
    _, oidx = cpu.cdist_topk(xyz, xyz, num_nns)
    assert (idx.cpu().numpy() == oidx[:, ::kth]).all()
    tgt = get_descriptors(X, idx)
    assert (tgt.cpu().numpy() == opairs.get_descriptors(xyz, oidx[:, ::kth])).all()
    if n <= 400:  # the reference's expression with torch.topk's own lists
        Xc = torch.from_numpy(xyz)
        nns = torch.topk(torch.cdist(Xc, Xc), k=num_nns, largest=False, dim=-1)[1][:, ::kth]
        X_nns = Xc[nns]
        ref = torch.norm(X_nns[:, 1:] - X_nns[:, 0].unsqueeze(1), dim=-1)
        np.testing.assert_allclose(get_descriptors(X, nns.cuda()).cpu().numpy(), ref.numpy(), rtol=3e-7)
    y = (xyz + 0.05 * np.random.default_rng(0).normal(size=xyz.shape)).astype(np.float32)
    Y = _cuda(y, True)
    loss = descriptor_mse(Y, idx, tgt)
    loss.backward()
    ol, og = opairs.descriptor_mse(y, oidx[:, ::kth], tgt.cpu().numpy())
    np.testing.assert_allclose(loss.item(), ol, rtol=2e-6)
    assert rel_l2(Y.grad.cpu(), torch.from_numpy(og)) <= 1e-4
    # unfused route (descriptors tensor + torch reduction) gives the same loss and gradient; normalised variant runs
    g1 = Y.grad.clone()
    Y.grad = None
    l2 = descriptors_loss(get_descriptors(Y, idx), tgt, normalize=False)
    l2.backward()
    np.testing.assert_allclose(l2.item(), loss.item(), rtol=1e-5)
    assert rel_l2(Y.grad, g1) <= 1e-4
    ln = descriptors_loss(get_descriptors(Y, idx), tgt, normalize=True)
    assert torch.isfinite(ln)


def test_style_patch_descriptors_one_launch_equals_per_cluster_loop(built):
    """StylePatchDescriptors.loss == get_style_patch_descriptors_loss over the cluster list (notebooks/29.2 cell 70
    call shape, with per-cluster scalings), and the per-step scale re-derivation."""
    from tests.util import rel_l2
    from wast3d_b200.descriptors import StylePatchDescriptors, get_descriptors, get_style_patch_descriptors_loss, neighbour_lists
    rng = np.random.default_rng(3)
    sizes = [300, 257, 512]
    ranges, s0 = [], 0
    for m in sizes:
        ranges.append((s0, s0 + m))
        s0 += m
    xyz0 = (rng.normal(size=(s0, 3)) * 0.4 + np.repeat(rng.normal(size=(3, 3)) * 3, sizes, axis=0)).astype(np.float32)
    sp = StylePatchDescriptors(_cuda(xyz0), ranges, num_nns=50, kth_nn=2)
    scal = torch.tensor([1.0, 0.5, 2.0], device="cuda")
    x = _cuda((xyz0 + 0.03 * rng.normal(size=xyz0.shape)).astype(np.float32), True)
    fused = sp.loss(x, scal)
    fused.backward()
    g_fused = x.grad.clone()
    x.grad = None
    lists = [neighbour_lists(_cuda(xyz0[s:e]), 50, 2) for s, e in ranges]
    tgts = [get_descriptors(_cuda(xyz0[s:e]), l) for (s, e), l in zip(ranges, lists)]
    loop = get_style_patch_descriptors_loss([x[s:e] / scal[c] for c, (s, e) in enumerate(ranges)], lists, tgts, normalize=False)
    loop.backward()
    np.testing.assert_allclose(fused.item(), loop.item(), rtol=1e-5)
    assert rel_l2(g_fused, x.grad) <= 1e-4
    sc = sp.scalings(x)
    assert sc.shape == (s0, 1) and (sc >= 0.02).all() and (sc <= 3.0).all()
    cur = torch.cat([get_descriptors(x.detach()[s:e], l) for (s, e), l in zip(ranges, lists)])
    exp = torch.clip(cur.mean(-1) / (1e-8 + torch.cat(tgts).mean(-1)), 0.02, 3.0).unsqueeze(1)
    torch.testing.assert_close(sc, exp, rtol=1e-6, atol=0)


def test_pair_terms_edge_cases(built):
    from wast3d_b200.descriptors import KnnMaskPairs, descriptor_mse, get_descriptors, masked_cdist_l1
    # k = 1 neighbour list (only the point itself): empty descriptor rows, zero loss, zero gradient
    X = _cuda(np.random.default_rng(0).normal(size=(10, 3)).astype(np.float32), True)
    idx = torch.arange(10, device="cuda").view(10, 1)
    d = get_descriptors(X, idx)
    assert d.shape == (10, 0)
    l = descriptor_mse(X, idx, torch.empty(10, 0, device="cuda"))
    assert l.item() == 0.0
    # coincident pair: distance 0 has zero gradient (torch.norm / cdist subgradient), no NaN
    P = _cuda(np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0]], np.float32), True)
    idx = torch.tensor([[0, 1, 2], [1, 0, 2], [2, 0, 1]], device="cuda")
    l = descriptor_mse(P, idx, torch.full((3, 2), 0.5, device="cuda"))
    l.backward()
    assert torch.isfinite(P.grad).all()
    # non-CUDA input fails loudly (no CPU fallback)
    with pytest.raises(RuntimeError):
        get_descriptors(torch.zeros(4, 3), torch.zeros(4, 2, dtype=torch.long))
    # single row / single column mask
    a = _cuda(np.array([[0.0, 0.0, 0.0]], np.float32))
    mp = KnnMaskPairs(a, a, k=1)
    assert masked_cdist_l1(a, a, mp, mp.targets(a)).item() == 0.0


def test_pair_terms_full_size_properties(built):
    """C3-scale point set (3 M points, k = 10): size-independent properties — translation invariance of the
    descriptor loss, zero loss at the target, gradient sums to zero (every pair contributes +g and -g)."""
    from wast3d_b200.descriptors import descriptor_mse, get_descriptors
    from wast3d_b200.scene import synthetic_gaussians
    from wast3d_b200.simple_knn._C import knn3
    pts = torch.from_numpy(synthetic_gaussians(3_000_000, seed=1, garden=True)["xyz"]).cuda()
    _, nn = knn3(pts)                                    # exact 3-NN lists from the kNN kernel
    idx = torch.cat([torch.arange(pts.size(0), device="cuda", dtype=torch.int32).view(-1, 1), nn], 1)
    tgt = get_descriptors(pts, idx)
    assert descriptor_mse(pts, idx, tgt).item() == 0.0
    moved = (pts * 1.01).requires_grad_(True)
    l = descriptor_mse(moved, idx, tgt)
    l.backward()
    shifted = descriptor_mse(pts * 1.01 + torch.tensor([0.5, -0.25, 0.125], device="cuda"), idx, tgt)
    np.testing.assert_allclose(shifted.item(), l.item(), rtol=2e-3)
    g = moved.grad.double()
    assert g.sum(0).abs().max().item() <= 1e-6 * g.abs().sum().item()
    np.testing.assert_allclose(l.item(), (0.01 ** 2) * (tgt.double() ** 2).mean().item(), rtol=1e-3)


# --------------------------------------------------------------------------- rank 4: K-Means
@pytest.mark.parametrize("n,K,seed", [(3000, 30, 0), (5000, 500, 1), (400, 3, 2), (257, 257, 3)])
def test_kmeans_lloyd_bit_exact_with_oracle(built, n, K, seed):
    from oracle import pairs as opairs
    from wast3d_b200.clustering import kmeans_lloyd
    rng = np.random.default_rng(seed)
    blobs = rng.normal(size=(max(K // 4, 3), 3)) * 4
    x = (blobs[rng.integers(0, len(blobs), n)] + rng.normal(size=(n, 3))).astype(np.float32)
    init = x[rng.choice(n, K, replace=False)].copy()
    for max_iter in (0, 1, 30):
        labels, centers, inertia, n_iter = kmeans_lloyd(_cuda(x), _cuda(init), max_iter=max_iter, tol=0.0)
        ol, oc, oi, oit = opairs.kmeans_lloyd(x, init, max_iter=max_iter, tol=0.0)
        assert n_iter == oit
        assert (labels.cpu().numpy() == ol).all()                 # memberships: bit-exact
        assert (centers.cpu().numpy() == oc).all()                # centres: double sums rounded once
        np.testing.assert_allclose(inertia, oi, rtol=1e-6)


def test_kmeans_empty_cluster_and_errors(built):
    from wast3d_b200.clustering import kmeans_lloyd
    x = np.array([[0, 0, 0], [0.1, 0, 0], [5, 5, 5], [5.1, 5, 5]], np.float32)
    init = np.array([[0, 0, 0], [5, 5, 5], [100, 100, 100]], np.float32)
    labels, centers, inertia, n_iter = kmeans_lloyd(_cuda(x), _cuda(init), max_iter=10)
    assert labels.tolist() == [0, 0, 1, 1]
    assert (centers[2].cpu().numpy() == init[2]).all()
    with pytest.raises(ValueError):
        kmeans_lloyd(_cuda(x), _cuda(np.zeros((5, 3), np.float32)))
    with pytest.raises(RuntimeError):
        kmeans_lloyd(torch.zeros(4, 3), torch.zeros(2, 3))


def test_cluster_points_and_cluster_files_round_trip(built, tmp_path):
    """cluster_points (aux_save_clusters_clean.py:32-47 contract) + save_clusters (:151-164): every Gaussian lands in
    exactly one file, re-centred on its centroid; centroid + saved xyz reproduces the scene."""
    from wast3d_b200.clustering import CLUSTER_ATTRS, cluster_points, load_cluster, save_clusters
    from wast3d_b200.scene import GaussianModel, synthetic_gaussians
    arrs = synthetic_gaussians(20000, seed=2)
    pc = GaussianModel.from_arrays(arrs, sh_degree=3, device="cuda")
    idx, ctr = cluster_points(arrs["xyz"], 30, n_init=2, max_iter=30, seed=0)
    assert isinstance(idx, np.ndarray) and idx.shape == (20000,) and ctr.shape == (30, 3)
    idx2, _ = cluster_points(arrs["xyz"], 30, n_init=2, max_iter=30, seed=0)
    assert (idx == idx2).all()                                     # seeded: reproducible memberships
    paths = save_clusters(pc, idx, ctr, str(tmp_path / "clusters"))
    assert len(paths) == len(np.unique(idx))
    total = 0
    for p in paths:
        c = int(p.split("_")[-1].split(".")[0])
        z = load_cluster(p)
        assert set(z) == set(CLUSTER_ATTRS)
        members = np.nonzero(idx == c)[0]
        total += len(members)
        np.testing.assert_allclose(z["_xyz"] + ctr[c], arrs["xyz"][members], atol=1e-6)
        assert z["_features_rest"].shape == (len(members), 15, 3)
    assert total == 20000
