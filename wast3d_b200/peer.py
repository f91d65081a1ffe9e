"""View-parallel optimizer over NVLink peer memory (csrc/peer_adam.cu).

The reference trains on one GPU with torch.optim.Adam over six parameter groups
(scene/gaussian_model.py:149-167).  With N GPUs rendering N views per step the textbook scheme is
"all-reduce the gradients, then every rank runs the full Adam" (distributed.allreduce_and_step, kept as
the NCCL baseline).  `PeerShardedAdam` instead keeps parameters AND gradients of all groups in one flat
peer-visible arena per rank and runs ONE kernel per step: each rank sums the N gradient replicas of its
1/N shard through peer loads, applies Adam with shard-local moments, and stores the new parameters into
all N replicas.  Same NVLink bytes as an all-reduce; optimizer HBM traffic and state divided by N; no
NCCL call on the step's critical path.

Host-side pieces that do not touch CUDA (`ArenaLayout`: padding, shard bounds, segment table) are
plain Python so the world_size-2 gloo tests can check them against an unsharded Adam.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Sequence

import torch
import torch.distributed as dist

from . import _lib


# --------------------------------------------------------------------------------- layout (host logic)
@dataclass(frozen=True)
class ArenaSlot:
    begin4: int  # float4 units from the start of the parameter (or gradient) region
    end4: int    # begin4 + ceil(numel / 4)
    numel: int


class ArenaLayout:
    """Flat layout of several fp32 tensors, each padded to a multiple of 4 floats (16-byte accesses,
    segment boundaries on float4 units), plus the balanced contiguous shard of every rank."""

    def __init__(self, numels: Sequence[int], late: Sequence[bool] | None = None):
        """`late[k]` places tensor k in the second ("late") class: the arena holds all early tensors
        first, then all late ones (order kept inside a class), and every rank owns a balanced shard of
        EACH class, so the two classes can be exchanged by two independent launches.  slots[] stays in
        the caller's tensor order."""
        late = [False] * len(numels) if late is None else [bool(x) for x in late]
        if len(late) != len(numels):
            raise ValueError("late must have one entry per tensor")
        self.slots: list[ArenaSlot] = [None] * len(numels)  # type: ignore[list-item]
        self.late = late
        off = 0
        for cls in (False, True):
            if cls:
                self.split4 = off
            for k, n in enumerate(numels):
                if late[k] != cls:
                    continue
                n = int(n)
                if n < 0:
                    raise ValueError("negative size")
                n4 = (n + 3) // 4
                self.slots[k] = ArenaSlot(off, off + n4, n)
                off += n4
        self.total4 = off

    def shard4(self, rank: int, world: int, cls: int | None = None) -> tuple[int, int]:
        """[begin4, end4) owned by `rank`: contiguous and balanced over the whole arena (cls None; a
        shard may span several parameter groups; the kernel intersects it with the segment table), or
        over the early (cls 0) / late (cls 1) class only."""
        from .distributed import shard_bounds
        if cls is None:
            return shard_bounds(self.total4, rank, world)
        lo, hi = (0, self.split4) if cls == 0 else (self.split4, self.total4)
        b, e = shard_bounds(hi - lo, rank, world)
        return lo + b, lo + e

    def segments(self, hyper: Sequence[dict], steps: Sequence[int]) -> list[dict]:
        """One Adam segment per non-empty slot: float4 range + that group's hyper-parameters."""
        out = []
        for slot, h, t in zip(self.slots, hyper, steps):
            if slot.end4 == slot.begin4:
                continue
            b1, b2 = h["betas"]
            out.append(dict(begin4=slot.begin4, end4=slot.end4, lr=float(h["lr"]), beta1=float(b1),
                            beta2=float(b2), eps=float(h["eps"]), step=int(t)))
        out.sort(key=lambda d: d["begin4"])  # the kernel wants ascending, disjoint ranges
        return out


# --------------------------------------------------------------------------------- peer-visible memory
class _RawCuda:
    """__cuda_array_interface__ holder so torch can view a raw device pointer as a uint8 tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def _wrap(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


class PeerBuffer:
    """`nbytes` of zero-filled device memory on every rank of `group`, each rank's block mapped into
    every other rank's address space.  .local is this rank's block as a uint8 tensor, .ptrs[q] the
    address of rank q's block in THIS process.

    Backends: "ipc" = cudaMalloc + CUDA IPC handles through the library's own C ABI
    (wast3d_peer_alloc/export/import); "symm" = torch.distributed._symmetric_memory, which also binds
    the blocks to an NVSwitch multicast object where the node supports it (.mc_ptr).  "auto" tries symm
    then ipc with more than two ranks, ipc then symm with two; all ranks take the same decision."""

    def __init__(self, nbytes: int, device, group=None, backend: str | None = None):
        _lib.require_device()
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._imported: list[int] = []
        self._own_ptr = None
        self._symm = None
        self.mc_ptr = 0  # NVLS multicast address of the block (symm backend on NVSwitch), 0 = none
        backend = backend or os.environ.get("WAST3D_PEER_BACKEND", "auto")
        # auto: with more than two ranks the switch-side reduction (multicast, symm backend) halves the
        # NVLink bytes, so it is tried first; with two ranks plain peer loads/stores move the same bytes
        order = ("symm", "ipc") if self.world > 2 else ("ipc", "symm")
        if self.world == 1:
            self.backend = "local"
            self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
            self.ptrs = [self.local.data_ptr()]
            return
        errors = []
        for b in (order if backend == "auto" else (backend,)):
            ok, err = True, None
            try:
                getattr(self, "_open_" + b)()
            except Exception as e:  # noqa: BLE001 - any failure moves every rank to the next backend
                ok, err = False, f"{b}: {e!r}"
            flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.backend = b
                break
            errors.append(err or f"{b}: failed on another rank")
            self._close()
        else:
            raise RuntimeError("PeerBuffer: no peer-memory backend works on this node: " + "; ".join(errors))
        dist.barrier(group=group)

    # -- backends
    def _open_ipc(self):
        lib = _lib.load()
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            _lib.check(lib.wast3d_peer_alloc(self.nbytes, C.byref(p)), "peer_alloc")
            self._own_ptr = int(p.value)
            h = (C.c_ubyte * 64)()
            _lib.check(lib.wast3d_peer_export(self._own_ptr, h), "peer_export")
            mine = torch.tensor(list(h), dtype=torch.uint8, device=self.device)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=self.group)
            allh = allh.cpu().view(self.world, 64)
            self.ptrs = []
            for q in range(self.world):
                if q == self.rank:
                    self.ptrs.append(self._own_ptr)
                    continue
                hq = (C.c_ubyte * 64)(*allh[q].tolist())
                pq = C.c_void_p()
                _lib.check(lib.wast3d_peer_import(hq, C.byref(pq)), f"peer_import(rank {q})")
                self._imported.append(int(pq.value))
                self.ptrs.append(int(pq.value))
            self.local = _wrap(self._own_ptr, self.nbytes, self.device)

    def _open_symm(self):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        with torch.cuda.device(self.device):
            t = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            hdl = symm_mem.rendezvous(t, grp.group_name)
            t.zero_()
            torch.cuda.synchronize()
        self._symm = (t, hdl)
        self.local = t
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        self.mc_ptr = 0 if os.environ.get("WAST3D_PEER_MULTICAST", "1") == "0" else mc

    def _close(self):
        lib = _lib.load()
        for p in self._imported:
            lib.wast3d_peer_release(p, 1)
        self._imported = []
        if self._own_ptr is not None:
            self.local = None
            lib.wast3d_peer_release(self._own_ptr, 0)
            self._own_ptr = None
        self._symm = None
        self.mc_ptr = 0

    def close(self):
        """Unmap the peers and free the block.  Call after all ranks stopped using it."""
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        self._close()


# --------------------------------------------------------------------------------- the optimizer
class GradSink:
    """Where the rasteriser's backward writes the leaf gradients (model_render.py): views of the
    arena's gradient region, one per parameter.  `fresh` is True until the first backward after
    zero_grad(); later backwards of the same step accumulate through autograd as usual."""

    def __init__(self, views: dict):
        self.views = views  # id(param) -> gradient view with the parameter's shape
        self.fresh = True

    def view_for(self, p):
        return self.views.get(id(p))


class PeerShardedAdam(torch.optim.Optimizer):
    """torch.optim.Adam's interface (param_groups with per-group lr / betas / eps) over the peer arena.

    Construction moves every parameter's storage into the arena (`p.data` becomes a view; the
    nn.Parameter objects stay the same) and creates `grad_sink`.  step() launches the fused
    reduce + Adam + broadcast kernel once.  Gradients are averaged over ranks when average=True
    (mean of the per-view losses), summed otherwise.  Replicas stay bit-identical: each element is
    computed by exactly one rank.  Moments are sharded: state_dict-style access goes through
    `exp_avg` / `exp_avg_sq` (this rank's shard, flat)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, group=None, average=True,
                 backend: str | None = None, timeout_s: float = 20.0, late_params=None):
        """`late_params` (optional list of parameters, e.g. the SH features): these are exchanged by a
        second launch on a side stream; step() returns with only the other ("early") parameters ordered
        on the current stream, and `take_late_event()` hands the consumer the event behind which the late
        ones are valid (render() passes it to the rasteriser, whose colour kernel is the only reader of
        the features: their all-gather then overlaps projection, sorting and binning of the next view)."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.group = group
        self.average = average
        self.timeout_s = timeout_s
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise RuntimeError("PeerShardedAdam: at most 8 ranks (one NVSwitch node)")
        self._params = [p for g in self.param_groups for p in g["params"]]
        self._hyper_of = [g for g in self.param_groups for _ in g["params"]]
        if not self._params:
            raise ValueError("no parameters")
        dev = self._params[0].device
        for p in self._params:
            _lib.require_device(p)
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise RuntimeError("PeerShardedAdam: parameters must be contiguous float32 on one CUDA device")
        self.device = dev
        late_ids = {id(p) for p in (late_params or [])}
        late = [id(p) in late_ids for p in self._params]
        if late_ids - {id(p) for p in self._params}:
            raise ValueError("late_params must be a subset of the optimised parameters")
        self.layout = ArenaLayout([p.numel() for p in self._params], late)
        self.overlap_late = any(late) and not all(late)
        lib = _lib.load()
        # two flag sets: the early and the late launch of a step may be in flight at the same time
        self._flag_bytes = (int(lib.wast3d_peer_flag_bytes()) + 255) // 256 * 256
        head = 2 * self._flag_bytes
        region = self.layout.total4 * 16
        self.buffer = PeerBuffer(head + 2 * region, dev, group=group, backend=backend)
        base = self.buffer.local
        self._param_flat = base[head:head + region].view(torch.float32)
        self._grad_flat = base[head + region:head + 2 * region].view(torch.float32)
        views = {}
        with torch.no_grad():
            for p, slot in zip(self._params, self.layout.slots):
                pv = self._param_flat[4 * slot.begin4:4 * slot.begin4 + slot.numel].view(p.shape)
                pv.copy_(p.data)
                p.data = pv
                views[id(p)] = self._grad_flat[4 * slot.begin4:4 * slot.begin4 + slot.numel].view(p.shape)
        self.grad_sink = GradSink(views)
        # shards: one over the whole arena, or one per class (early, late)
        if self.overlap_late:
            self.shards = [self.layout.shard4(self.rank, self.world, 0), self.layout.shard4(self.rank, self.world, 1)]
        else:
            self.shards = [self.layout.shard4(self.rank, self.world)]
        self.shard = self.shards[0]
        n = 4 * sum(e - b for b, e in self.shards)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)      # [early shard | late shard]
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._steps = [0] * len(self._params)
        self._epoch = 0
        W = self.world
        self._grad_ptrs = (C.c_void_p * W)(*[q + head + region for q in self.buffer.ptrs])
        self._param_ptrs = (C.c_void_p * W)(*[q + head for q in self.buffer.ptrs])
        self._flag_ptrs = [(C.c_void_p * W)(*[q + k * self._flag_bytes for q in self.buffer.ptrs]) for k in (0, 1)]
        mc = self.buffer.mc_ptr if W > 1 else 0
        self._mc_params = (mc + head) if mc else None
        self._mc_grads = (mc + head + region) if mc else None
        self.multicast = bool(mc)
        # WAST3D_PEER_SIDE_PRIORITY: CUDA stream priority of the late launch (0 = default, -1.. = higher)
        prio = int(os.environ.get("WAST3D_PEER_SIDE_PRIORITY", "0"))
        self._side = torch.cuda.Stream(device=dev, priority=prio) if self.overlap_late else None
        self._early_done = torch.cuda.Event() if self.overlap_late else None
        self._late_event = torch.cuda.Event() if self.overlap_late else None
        self._late_pending = False
        # persistent-grid cap of the late launch: it runs beside the next view's projection / sorting / binning;
        # its duration barely changes between 148 and 12-20 CTAs (NVLink/NVSwitch bound), the slowdown of the kernels
        # beside it does: every request it keeps in flight queues in front of theirs (profiles/r01_overlap.md)
        # defaults = the measured configurations: multicast 12 (N = 4 and N = 8), plain peer loads / stores 37 at N = 2
        # and 20 at N = 4
        # end of round 2 (8 reduced columns in flight per thread; the binning chain beside the exchange is latency-bound
        # and suffers from every extra CTA): N = 8, multicast: 4 / 6 / 8 / 12 / 20 CTAs = 3.50 / 3.24 / 3.22 / 3.45 /
        # 3.57 ms per step; N = 4 (twice the shard per rank): 8 / 12 CTAs = 3.55 / 3.30 ms (profiles/r02_scaling.md)
        default_ctas = (12 if W <= 4 else 8) if self.multicast else (37 if W <= 2 else 20)
        self.late_ctas = int(os.environ.get("WAST3D_PEER_LATE_CTAS", str(default_ctas)))
        if W > 1:  # replicas start identical: rank 0's values win (the reference has one copy)
            dist.broadcast(self._param_flat, src=dist.get_global_rank(group, 0) if group is not None else 0,
                           group=group)
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)

    # -- late class
    def take_late_event(self):
        """The event behind which the late parameters of the last step() are valid, or None.  The taker
        must make its stream wait for it before reading them (the rasteriser does: ABI v5
        colour_wait_event); after that the current stream is ordered and nobody else needs to wait."""
        if not self._late_pending:
            return None
        self._late_pending = False
        return self._late_event

    def sync(self):
        """Order the current stream behind the late launch (for readers other than render())."""
        ev = self.take_late_event()
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _launch(self, cls: int, segs, scale: float):
        b4, e4 = self.shards[cls]
        lo, hi = (0, self.layout.total4) if not self.overlap_late else \
            ((0, self.layout.split4) if cls == 0 else (self.layout.split4, self.layout.total4))
        segs = [s for s in segs if s["begin4"] >= lo and s["end4"] <= hi]
        arr = (_lib.AdamSegment * max(1, len(segs)))()
        for k, s in enumerate(segs):
            arr[k] = _lib.AdamSegment(s["begin4"], s["end4"], s["lr"], s["beta1"], s["beta2"], s["eps"], s["step"], 0)
        moff = 16 * sum(e - b for b, e in self.shards[:cls])  # bytes into the moment arrays
        rc = _lib.load().wast3d_peer_adam_step(
            self.world, self.rank, self._grad_ptrs, self._param_ptrs, self._flag_ptrs[cls],
            self._mc_grads, self._mc_params, self.exp_avg.data_ptr() + moff, self.exp_avg_sq.data_ptr() + moff,
            b4, e4, arr, len(segs), scale, self._epoch, float(self.timeout_s),
            self.late_ctas if (cls == 1 and self.overlap_late) else 0, _lib.stream_ptr())
        _lib.check(rc, "peer_adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.sync()  # a late launch nobody consumed: everything below touches the arena
        # gradients that were produced outside the sink (plain autograd) are copied into the arena
        for p in self._params:
            gv = self.grad_sink.view_for(p)
            if p.grad is None:
                gv.zero_()  # torch.optim.Adam skips such parameters; peers may still have gradients
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
        for k in range(len(self._params)):
            self._steps[k] += 1
        segs = self.layout.segments(self._hyper_of, self._steps)
        self._epoch += 1
        scale = 1.0 / self.world if self.average else 1.0
        with torch.cuda.device(self.device):
            self._launch(0, segs, scale)
            if self.overlap_late:
                # the late launch starts when the early one is done (two persistent, flag-spinning grids
                # of one GPU never compete for the SMs) and runs beside whatever the caller enqueues next
                main = torch.cuda.current_stream(self.device)
                self._early_done.record(main)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._early_done)
                    self._launch(1, segs, scale)
                    self._late_event.record(self._side)
                self._late_pending = True
        return loss

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=True)  # the arena views are overwritten by the next backward
        self.grad_sink.fresh = True

    # -- checkpointing (the reference's GaussianModel.capture / restore go through optimizer.state_dict(),
    #    scene/gaussian_model.py:58-77).  The moments live in flat per-rank shards, not in self.state: expose them
    #    so a resume restores them instead of silently resetting Adam.
    def state_dict(self):
        d = super().state_dict()
        d["peer_sharded"] = {"world": self.world, "rank": self.rank, "shards": [tuple(s) for s in self.shards],
                             "exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                             "steps": list(self._steps), "epoch": int(self._epoch)}
        return d

    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        mine = sd.pop("peer_sharded", None)
        if mine is None:
            raise RuntimeError("PeerShardedAdam.load_state_dict: not a PeerShardedAdam checkpoint (the sharded "
                               "moments are missing); resuming would silently reset Adam's state")
        if mine["world"] != self.world or mine["rank"] != self.rank or \
                [tuple(s) for s in mine["shards"]] != [tuple(s) for s in self.shards]:
            raise RuntimeError("PeerShardedAdam.load_state_dict: checkpoint was written with a different world size, "
                               "rank or arena layout")
        super().load_state_dict(sd)
        with torch.no_grad():
            self.exp_avg.copy_(mine["exp_avg"].to(self.device))
            self.exp_avg_sq.copy_(mine["exp_avg_sq"].to(self.device))
        self._steps = list(mine["steps"])
        self._epoch = int(mine["epoch"])

    def check_peers(self):
        """Raise if a step timed out waiting for a peer (sticky flag set by the kernel)."""
        e = int(_lib.load().wast3d_peer_error(0))
        if e:
            raise RuntimeError(f"PeerShardedAdam: timed out waiting for rank {e - 1}; replicas have diverged")

    def close(self):
        self.buffer.close()
