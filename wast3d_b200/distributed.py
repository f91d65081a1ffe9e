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


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False,
                        bucket_bytes: int = 256 << 20, async_op: bool = False):
    """Sum (or average) `.grad` of every parameter across ranks, in flat fp32 buckets.

    Buckets are filled in the given order (pass parameters in reverse-autograd order to start
    the first collective as early as possible) and sized for launch latency, not link count:
    on NVSwitch every peer is at full bandwidth.  Returns the list of work handles when
    async_op=True (call `finish_allreduce` to wait and scatter back)."""
    params = [p for p in params if p.grad is not None]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1 or not params:
        return []
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        nbytes = p.grad.numel() * p.grad.element_size()
        if cur and cur_bytes + nbytes > bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    pending = []
    for bucket in buckets:
        if len(bucket) == 1 and bucket[0].grad.is_contiguous():
            flat = bucket[0].grad.view(-1)  # in place, no pack/unpack copy
            packed = False
        else:
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            packed = True
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        pending.append((work, flat, bucket, packed))
    if async_op:
        return [(w, f, b, k, average, world) for (w, f, b, k) in pending]
    finish_allreduce([(w, f, b, k, average, world) for (w, f, b, k) in pending])
    return []


def finish_allreduce(handles):
    for work, flat, bucket, packed, average, world in handles:
        work.wait()
        if average:
            flat.div_(world)
        if packed:
            off = 0
            for p in bucket:
                n = p.grad.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n


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


def view_for_rank(cameras: Sequence, step: int, rank: int, world: int):
    """Camera of `rank` in the `step`-th view batch (k ranks render k different cameras)."""
    return cameras[(step * world + rank) % len(cameras)]
