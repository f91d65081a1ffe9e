"""CPU, world_size 2 over gloo: the host logic of the two multi-GPU pieces."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wast3d_b200 import distributed as wd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 16384, 16385):
        for w in (1, 2, 3, 8):
            b = [wd.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def _grad_job(rank, world):
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.zeros(1000, 3)), torch.nn.Parameter(torch.zeros(1000, 15, 3)),
          torch.nn.Parameter(torch.zeros(1000, 1)), torch.nn.Parameter(torch.zeros(7))]
    g = torch.Generator().manual_seed(100 + rank)
    for p in ps[:3]:
        p.grad = torch.randn(p.shape, generator=g)
    # ps[3]: only rank 1 has a gradient, and a NON-CONTIGUOUS one — the ranks must still issue the same collectives
    # (the work list is derived from the parameter shapes; a missing gradient counts as zeros)
    if rank == 1:
        ps[3].grad = torch.arange(14.0)[::2]
    wd.allreduce_gradients(ps, bucket_bytes=20000, small_bytes=5000)  # in-place chunks + one packed buffer
    return [p.grad.clone() if p.grad is not None else None for p in ps]


def test_view_parallel_gradient_allreduce():
    out = _run(_grad_job)
    expect = []
    for shape in [(1000, 3), (1000, 15, 3), (1000, 1)]:
        tot = 0
        for r in range(2):
            pass
        expect.append(shape)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    sums = []
    per_rank = [[torch.randn(s, generator=gens[r]) for s in [(1000, 3), (1000, 15, 3), (1000, 1)]] for r in range(2)]
    for i in range(3):
        sums.append(per_rank[0][i] + per_rank[1][i])
    for r in range(2):
        for i in range(3):
            assert torch.equal(out[r][i], sums[i])      # two-rank fp32 sum is order independent
        assert torch.equal(out[r][3], torch.arange(14.0)[::2])


class _RangeSGD:
    """CPU stand-in with FusedAdam's begin_step()/step_range() protocol (plain SGD on a flat range)."""

    def __init__(self, params, lr):
        self.param_groups = [{"params": list(params), "lr": lr}]
        self.began = 0
        self.ranges = []

    def begin_step(self):
        self.began += 1

    def step_range(self, p, s0, e0):
        self.ranges.append((id(p), s0, e0))
        with torch.no_grad():
            p.view(-1)[s0:e0] -= self.param_groups[0]["lr"] * p.grad.view(-1)[s0:e0]

    def step(self):
        raise AssertionError("world > 1 must take the overlapped path")


def _step_job(rank, world):
    ps = [torch.nn.Parameter(torch.ones(1000, 3)), torch.nn.Parameter(torch.ones(1000, 15, 3)),
          torch.nn.Parameter(torch.ones(10, 1))]
    g = torch.Generator().manual_seed(200 + rank)
    for p in ps:
        p.grad = torch.randn(p.shape, generator=g)
    opt = _RangeSGD(ps, 0.5)
    wd.allreduce_and_step(opt, average=True, chunk_bytes=40000)
    covered = {}
    for pid, s0, e0 in opt.ranges:
        covered.setdefault(pid, []).append((s0, e0))
    full = all(sorted(covered[id(p)])[0][0] == 0 and sorted(covered[id(p)])[-1][1] == p.numel() and
               all(a[1] == b[0] for a, b in zip(sorted(covered[id(p)]), sorted(covered[id(p)])[1:])) for p in ps)
    return [p.detach().clone() for p in ps], opt.began, full


def test_overlapped_allreduce_and_step():
    out = _run(_step_job)
    gens = [torch.Generator().manual_seed(200 + r) for r in range(2)]
    shapes = [(1000, 3), (1000, 15, 3), (10, 1)]
    per_rank = [[torch.randn(s, generator=gens[r]) for s in shapes] for r in range(2)]
    for r in range(2):
        params, began, full = out[r]
        assert began == 1 and full           # one step count bump, every element updated exactly once
        for i in range(3):
            expect = torch.ones(shapes[i]) - 0.5 * ((per_rank[0][i] + per_rank[1][i]) / 2)
            assert torch.allclose(params[i], expect, rtol=0, atol=1e-6)
    for i in range(3):
        assert torch.equal(out[0][0][i], out[1][0][i])   # replicas stay bit-identical


def _match_job(rank, world):
    from oracle import cpu
    rng = np.random.default_rng(5)
    a = torch.from_numpy(rng.normal(size=(1001, 3)).astype(np.float32))   # uneven shards
    b = torch.from_numpy(rng.normal(size=(77, 3)).astype(np.float32))

    def cpu_match(a_shard, b_all):  # stands in for wast3d_b200.matching.nn_match on CPU
        i, d = cpu.nn_match(a_shard.numpy(), b_all.numpy())
        return torch.from_numpy(i).long(), torch.from_numpy(d)

    idx, cost = wd.sharded_match(cpu_match, [a], [b])
    full_i, full_d = cpu.nn_match(a.numpy(), b.numpy())
    return bool((idx.numpy() == full_i).all() and (cost.numpy() == full_d).all() and len(idx) == 1001)


def test_sharded_matching_equals_unsharded():
    assert _run(_match_job) == [True, True]


def _stats_job(rank, world):
    """C4 host logic: statistics of points sharded BY POINT over two ranks == the oracle's statistics of the union
    (two all-reduces of 4 + 6 numbers per cluster; numpy stand-ins for the CUDA accumulation kernels)."""
    from oracle import cpu
    rng = np.random.default_rng(9)
    n, K = 5003, 37
    pts = (rng.normal(size=(n, 3)) * 2.0 + 5.0).astype(np.float32)
    lab = rng.integers(0, K - 3, size=n).astype(np.int32)          # clusters K-3..K-1 stay empty
    s, e = wd.shard_bounds(n, rank, world)

    def sums(p, l, K_, sum3, count):
        np.add.at(sum3.numpy(), l.numpy(), p.numpy().astype(np.float64))
        np.add.at(count.numpy(), l.numpy(), 1)

    def scatter(p, l, K_, mean3, acc6):
        d = p.numpy().astype(np.float64) - mean3.numpy()[l.numpy()]
        m = np.stack([d[:, 0] * d[:, 0], d[:, 0] * d[:, 1], d[:, 0] * d[:, 2], d[:, 1] * d[:, 1], d[:, 1] * d[:, 2],
                      d[:, 2] * d[:, 2]], 1)
        np.add.at(acc6.numpy(), l.numpy(), m)

    mean, cov, count = wd.sharded_cluster_stats(torch.from_numpy(pts[s:e]), torch.from_numpy(lab[s:e]), K,
                                                sums_fn=sums, scatter_fn=scatter)
    om, oc, on = cpu.cluster_stats(pts, lab, K)
    return bool((count.numpy() == on).all() and np.abs(mean.numpy() - om).max() <= 1e-6 and
                np.abs(cov.numpy() - oc).max() <= 1e-6 and (cov.numpy()[K - 3:] == 0).all())


def test_point_sharded_cluster_stats_equal_unsharded():
    assert _run(_stats_job) == [True, True]


def test_view_assignment():
    cams = list(range(8))
    seen = [wd.view_for_rank(cams, step, r, 4) for step in range(2) for r in range(4)]
    assert seen == [0, 1, 2, 3, 4, 5, 6, 7]
