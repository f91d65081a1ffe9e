"""Multi-GPU pieces of the hot path.  The reference is single-GPU; these are
the only two places where the path shards naturally, one process per GPU over
torch.distributed (NCCL on the B200 box, gloo in the CPU tests):

  * view-parallel rendering: every rank holds the full parameter set and renders a different
    camera; per-Gaussian gradients are summed across ranks (`allreduce_gradients`) before an
    identical Adam step on every rank — no parameter broadcast is needed afterwards.
  * sharded cluster matching: content clusters (rows of the cost matrix) are split across
    ranks, style statistics are replicated, each rank matches its rows and the (index, cost)
    pairs are all-gathered (`sharded_match`).

Nothing here touches CUDA directly, so the host logic is testable with gloo on CPU tensors.
"""
from __future__ import annotations

from typing import Callable, Iterable, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, end) of `n` rows for `rank` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _avg_supported(group) -> bool:
    try:
        return dist.get_backend(group) == "nccl"
    except Exception:  # noqa: BLE001
        return False


def _grad_chunks(params, chunk_bytes: int, small_bytes: int):
    """Work list for the gradient all-reduce.  Every entry is (flat_view, [(param, start, end)], packed):
    large gradients are reduced IN PLACE as contiguous slices of at most chunk_bytes (no pack /
    unpack copies); gradients smaller than small_bytes are packed together into one buffer so
    that tiny tensors do not each pay a collective launch."""
    # The work list must be IDENTICAL on every rank (the sequence and sizes of the collectives): it is derived from
    # the parameters' shapes only.  A missing gradient becomes zeros and a non-contiguous one a contiguous copy
    # (allreduce_gradients materialises both before calling this).
    work, small, small_n = [], [], 0
    for p in params:
        g = p.grad
        nbytes = g.numel() * g.element_size()
        if nbytes < small_bytes:
            small.append(p)
            small_n += nbytes
            continue
        flat = g.view(-1)
        per = max(1, chunk_bytes // g.element_size())
        for s0 in range(0, flat.numel(), per):
            e0 = min(flat.numel(), s0 + per)
            work.append((flat[s0:e0], [(p, s0, e0)], False))
    if small:
        work.insert(0, (torch.cat([p.grad.reshape(-1) for p in small]), [(p, 0, p.grad.numel()) for p in small], True))
    return work


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False,
                        bucket_bytes: int = 128 << 20, async_op: bool = False, small_bytes: int = 1 << 20):
    """Sum (or average) `.grad` of every parameter across ranks.

    Gradients are reduced in place in contiguous chunks of at most `bucket_bytes` (sized for launch
    latency and for overlap with the optimizer, not for link count: on NVSwitch every peer is at
    full bandwidth); gradients below `small_bytes` share one packed buffer.  With NCCL the average
    is taken inside the collective (ReduceOp.AVG).  Returns the list of pending chunks when
    async_op=True (pass it to `finish_allreduce`, or use `allreduce_and_step`)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    params = [p for p in params if p.numel()]
    if world == 1 or not params:
        return []
    for p in params:  # same collectives on every rank, whatever the local autograd graph produced
        if p.grad is None:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
        elif not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
    use_avg = average and _avg_supported(group)
    op = dist.ReduceOp.AVG if use_avg else dist.ReduceOp.SUM
    pending = []
    for flat, owners, packed in _grad_chunks(params, bucket_bytes, small_bytes):
        h = dist.all_reduce(flat, op=op, group=group, async_op=True)
        pending.append((h, flat, owners, packed, average and not use_avg, world))
    if async_op:
        return pending
    finish_allreduce(pending)
    return []


def _finish_one(entry):
    work, flat, owners, packed, divide, world = entry
    work.wait()
    if divide:
        flat.div_(world)
    if packed:
        off = 0
        for p, s0, e0 in owners:
            n = e0 - s0
            p.grad.reshape(-1)[s0:e0].copy_(flat[off:off + n])  # contiguous by construction: a view
            off += n


def finish_allreduce(handles):
    for entry in handles:
        _finish_one(entry)


def allreduce_and_step(optimizer, params: Iterable[torch.Tensor] = None, group=None, average: bool = True,
                       chunk_bytes: int = 128 << 20):
    """View-parallel optimizer step: all-reduce the gradients chunk by chunk and apply the Adam
    update of each chunk as soon as ITS collective has finished, so the update of chunk i (HBM
    bound, compute stream) overlaps the all-reduce of chunk i+1 (NVLink, NCCL stream).

    `optimizer` must offer begin_step() and step_range(param, start, end) (optim.FusedAdam); with a
    single rank this is optimizer.step().  Every rank applies the identical update, so parameters
    stay replicated without a broadcast."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        optimizer.step()
        return
    if params is None:
        params = [p for g in optimizer.param_groups for p in g["params"]]
    pending = allreduce_gradients(params, group=group, average=average, bucket_bytes=chunk_bytes, async_op=True)
    optimizer.begin_step()
    for entry in pending:
        _finish_one(entry)
        for p, s0, e0 in entry[2]:
            optimizer.step_range(p, s0, e0)


def sharded_match(match_fn: Callable[..., Sequence[torch.Tensor]], row_tensors: Sequence[torch.Tensor],
                  replicated: Sequence[torch.Tensor], group=None):
    """Run `match_fn(*row_shard, *replicated) -> (idx, cost)` on this rank's contiguous shard of
    the row tensors and all-gather the results so every rank returns the full (idx, cost).

    `match_fn` is `wast3d_b200.matching.w2_match` / `nn_match` in production; the tests pass a
    CPU function to exercise the sharding and gather logic with gloo."""
    n = int(row_tensors[0].shape[0])
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = match_fn(*row_tensors, *replicated)
        return out[0], out[1]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    s, e = shard_bounds(n, rank, world)
    out = match_fn(*[t[s:e] for t in row_tensors], *replicated)
    idx, cost = out[0], out[1]
    cap = (n + world - 1) // world  # all_gather wants equal sizes: pad to the largest shard
    dev = idx.device
    idx_pad = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    cost_pad = torch.full((cap,), float("inf"), dtype=torch.float32, device=dev)
    idx_pad[: e - s] = idx.to(torch.int64)
    cost_pad[: e - s] = cost.to(torch.float32)
    idx_all = [torch.empty_like(idx_pad) for _ in range(world)]
    cost_all = [torch.empty_like(cost_pad) for _ in range(world)]
    dist.all_gather(idx_all, idx_pad, group=group)
    dist.all_gather(cost_all, cost_pad, group=group)
    idx_full = torch.cat([idx_all[r][: shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0]] for r in range(world)])
    cost_full = torch.cat([cost_all[r][: shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0]] for r in range(world)])
    return idx_full, cost_full


def sharded_cluster_stats(points_shard: torch.Tensor, labels_shard: torch.Tensor, K: int, group=None,
                          sums_fn: Callable = None, scatter_fn: Callable = None):
    """Per-cluster (mean [K,3], cov6 [K,6], count [K]) of points that are sharded BY POINT across the ranks
    (BASELINE.json configs[3]: 6 M Gaussians, 16 384 content clusters).  Every rank adds its
    points into [K,3] + [K] accumulators, one all-reduce; the means are then identical everywhere, every rank adds its
    centred second moments, second all-reduce of [K,6] doubles.  Two collectives of 10 numbers per cluster in total;
    the points never move.  Same two passes as the single-GPU `matching.cluster_stats`.

    `sums_fn(points, labels, K, sum3, count)` / `scatter_fn(points, labels, K, mean3, acc6)` accumulate in place;
    they default to the library kernels (wast3d_cluster_sums / wast3d_cluster_scatter) — the CPU tests pass
    stand-ins to exercise the collectives with gloo."""
    if sums_fn is None or scatter_fn is None:
        from . import matching
        sums_fn, scatter_fn = matching.cluster_sums, matching.cluster_scatter
    dev = points_shard.device
    sum3 = torch.zeros((K, 3), dtype=torch.float64, device=dev)
    count = torch.zeros((K,), dtype=torch.int32, device=dev)
    sums_fn(points_shard, labels_shard, K, sum3, count)
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        dist.all_reduce(sum3, group=group)
        dist.all_reduce(count, group=group)
    c = count.to(torch.float64).unsqueeze(1)
    mean3 = torch.where(c > 0, sum3 / c.clamp_min(1.0), torch.zeros_like(sum3))
    acc6 = torch.zeros((K, 6), dtype=torch.float64, device=dev)
    scatter_fn(points_shard, labels_shard, K, mean3, acc6)
    if multi:
        dist.all_reduce(acc6, group=group)
    cov6 = torch.where(c > 0, acc6 / c.clamp_min(1.0), torch.zeros_like(acc6))
    return mean3.to(torch.float32), cov6.to(torch.float32), count


def view_for_rank(cameras: Sequence, step: int, rank: int, world: int):
    """Camera of `rank` in the `step`-th view batch (k ranks render k different cameras)."""
    return cameras[(step * world + rank) % len(cameras)]
