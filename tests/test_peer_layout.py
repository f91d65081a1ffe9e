"""CPU tests of the host logic behind the peer-memory optimizer (wast3d_b200/peer.py): arena layout,
shard bounds, segment table — and, with gloo at world_size 2, that "reduce-scatter the gradients of my
shard + Adam on the shard + all-gather the parameters" over that layout equals the dense Adam on the
averaged gradients (the algorithm csrc/peer_adam.cu executes over NVLink)."""
import torch
import torch.distributed as dist

from tests.test_distributed_gloo import _run
from wast3d_b200.peer import ArenaLayout

SHAPES = [(1001, 3), (1001, 1, 3), (1001, 15, 3), (1001, 1), (1001, 3), (1001, 4), (0,), (5,)]
LRS = [1.6e-4, 2.5e-3, 1.25e-4, 0.05, 5e-3, 1e-3, 1e-3, 1e-2]


def test_layout_padding_and_shards():
    lay = ArenaLayout([torch.Size(s).numel() for s in SHAPES])
    off = 0
    for slot, s in zip(lay.slots, SHAPES):
        n = torch.Size(s).numel()
        assert slot.begin4 == off and slot.numel == n and slot.end4 - slot.begin4 == (n + 3) // 4
        off = slot.end4
    assert lay.total4 == off
    for w in (1, 2, 3, 8):
        b = [lay.shard4(r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == lay.total4
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_layout_late_class_is_contiguous_and_sharded_per_class():
    numels = [torch.Size(s).numel() for s in SHAPES]
    late = [False, True, True, False, False, False, False, True]
    lay = ArenaLayout(numels, late)
    # slots stay in tensor order; early tensors fill [0, split4), late ones [split4, total4), no gaps
    early = sorted((sl.begin4, sl.end4) for sl, l in zip(lay.slots, late) if not l)
    lates = sorted((sl.begin4, sl.end4) for sl, l in zip(lay.slots, late) if l)
    assert early[0][0] == 0 and early[-1][1] == lay.split4 == lates[0][0] and lates[-1][1] == lay.total4
    for rng in (early, lates):
        assert all(a[1] == b[0] for a, b in zip(rng, rng[1:]))
    assert [sl.numel for sl in lay.slots] == numels
    for w in (1, 2, 3, 8):
        for cls, (lo, hi) in enumerate(((0, lay.split4), (lay.split4, lay.total4))):
            b = [lay.shard4(r, w, cls) for r in range(w)]
            assert b[0][0] == lo and b[-1][1] == hi
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    hyper = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-15) for lr in LRS]
    segs = lay.segments(hyper, [1] * len(SHAPES))
    assert all(a["end4"] <= b["begin4"] for a, b in zip(segs, segs[1:]))  # ascending for the kernel
    assert all(s["end4"] <= lay.split4 or s["begin4"] >= lay.split4 for s in segs)  # none spans the classes
    assert ArenaLayout(numels).split4 == ArenaLayout(numels).total4  # no late class: everything early


def test_segments_skip_empty_and_carry_hyperparameters():
    lay = ArenaLayout([torch.Size(s).numel() for s in SHAPES])
    hyper = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-15) for lr in LRS]
    segs = lay.segments(hyper, [3] * len(SHAPES))
    assert len(segs) == len(SHAPES) - 1  # the empty tensor has no segment
    assert all(a["end4"] <= b["begin4"] for a, b in zip(segs, segs[1:]))
    assert [s["lr"] for s in segs] == [lr for lr, sh in zip(LRS, SHAPES) if torch.Size(sh).numel()]
    assert all(s["step"] == 3 for s in segs)


def _adam_flat(p, g, m, v, lr, b1, b2, eps, t):
    m.add_((1 - b1) * (g - m))
    v.mul_(b2).add_((1 - b2) * g * g)
    p.sub_((lr / (1 - b1 ** t)) * (m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps)))


def _peer_job(rank, world):
    """Emulates peer_adam_kernel with gloo collectives over the product's layout/shard/segment logic."""
    numels = [torch.Size(s).numel() for s in SHAPES]
    lay = ArenaLayout(numels)
    g0 = torch.Generator().manual_seed(7)
    params = torch.zeros(lay.total4 * 4)
    for slot in lay.slots:
        params[4 * slot.begin4:4 * slot.begin4 + slot.numel] = torch.randn(slot.numel, generator=g0)
    s4, e4 = lay.shard4(rank, world)
    m, v = torch.zeros(4 * (e4 - s4)), torch.zeros(4 * (e4 - s4))
    hyper = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-15) for lr in LRS]
    gen = torch.Generator().manual_seed(100 + rank)
    for t in range(1, 4):
        grads = torch.zeros(lay.total4 * 4)
        for slot in lay.slots:
            grads[4 * slot.begin4:4 * slot.begin4 + slot.numel] = torch.randn(slot.numel, generator=gen)
        # "peer loads": every rank's gradient replica of MY shard, summed in rank order
        allg = [torch.empty_like(grads) for _ in range(world)]
        dist.all_gather(allg, grads)
        gs = allg[0][4 * s4:4 * e4].clone()
        for q in range(1, world):
            gs += allg[q][4 * s4:4 * e4]
        gs *= 1.0 / world
        mine = params[4 * s4:4 * e4].clone()
        for seg in lay.segments(hyper, [t] * len(SHAPES)):
            lo, hi = max(seg["begin4"], s4), min(seg["end4"], e4)
            if hi <= lo:
                continue
            a, b = 4 * (lo - s4), 4 * (hi - s4)
            _adam_flat(mine[a:b], gs[a:b], m[a:b], v[a:b], seg["lr"], seg["beta1"], seg["beta2"], seg["eps"], seg["step"])
        # "peer stores": my shard's new parameters land in every replica
        parts = [torch.empty(4 * (lay.shard4(q, world)[1] - lay.shard4(q, world)[0])) for q in range(world)]
        dist.all_gather(parts, mine) if len({p.numel() for p in parts}) == 1 else _uneven_gather(parts, mine, rank, world)
        params = torch.cat(parts)
    return params


def _uneven_gather(parts, mine, rank, world):
    for q in range(world):
        buf = mine.clone() if q == rank else parts[q]
        dist.broadcast(buf, src=q)
        parts[q] = buf if q != rank else mine
    return parts


def test_sharded_peer_adam_equals_dense_adam_gloo():
    out = _run(_peer_job, world=2)
    assert torch.equal(out[0], out[1])  # replicas stay bit-identical
    numels = [torch.Size(s).numel() for s in SHAPES]
    lay = ArenaLayout(numels)
    g0 = torch.Generator().manual_seed(7)
    ps = [torch.nn.Parameter(torch.randn(n, generator=g0)) for n in numels]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps, LRS)], lr=0.0, eps=1e-15)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    for _ in range(3):
        for p in ps:
            gr = [torch.randn(p.numel(), generator=g) for g in gens]
            p.grad = (gr[0] + gr[1]) * 0.5
        opt.step()
    for p, slot in zip(ps, lay.slots):
        got = out[0][4 * slot.begin4:4 * slot.begin4 + slot.numel]
        assert (got - p.detach()).abs().max().item() <= 1e-6 if slot.numel else True
        pad = out[0][4 * slot.begin4 + slot.numel:4 * slot.end4]
        assert (pad == 0).all()  # zero gradients keep the padding at zero
