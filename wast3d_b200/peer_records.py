"""Colour-record exchange of the SH features for the view-parallel optimizer (bench.py --sync records).

View-parallel optimizer in which the SH features never cross NVLink as gradients or parameters:

  * xyz / opacity / scaling / rotation (11 of the 59 floats per Gaussian) go through `peer.PeerShardedAdam`
    exactly as today (reduce-scatter + Adam + all-gather in one kernel, on the current stream);
  * for `_features_dc` / `_features_rest` every rank publishes a 16-byte colour record per Gaussian of ITS view
    (csrc/sh_adam.cu `colour_record_kernel`, written by the rasteriser's backward into peer-visible memory) and
    then rebuilds the gradient summed over all N views from the N records, xyz and the N camera centres, and
    applies Adam to all P Gaussians locally (`sh_adam_records_kernel`) on a side stream.  Every rank executes
    the same arithmetic on the same inputs, so the replicas stay bit-identical without a parameter all-gather.

The side-stream work is ordered like `PeerShardedAdam`'s late launch: `take_late_event()` hands render() the
event its colour kernel waits for.  The ranks are ordered around the record reads by two empty-shard launches of
`wast3d_peer_adam_step` on the second flag set (phase 0 + phase 2 of that kernel form an all-rank barrier on
the device).  The algebra is pinned on the CPU: oracle/sh_records.py, tests/test_sh_records_oracle.py.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .peer import PeerBuffer, PeerShardedAdam


class _AdamGroup(C.Structure):  # struct wast3d_adam_group (include/wast3d_b200.h)
    _fields_ = [("param", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int), ("reserved", C.c_int),
                ("schedule_dev", C.c_void_p)]


def _staged(lib):
    """ctypes signatures of include/wast3d_b200_staged.h (kept out of _lib.SIGNATURES: not part of the drop-in ABI)."""
    f = lib.wast3d_staged_colour_records
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    g = lib.wast3d_staged_sh_adam_from_records
    g.restype = C.c_int
    g.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                  C.c_void_p, C.c_void_p]
    return f, g


class PeerRecordAdam(torch.optim.Optimizer):
    """torch.optim.Adam's interface over [peer-sharded early parameters] + [record-rebuilt, replicated features].

    params: the usual list of group dicts; `xyz`, `features_dc`, `features_rest` name the three tensors the
    record path needs.  Before every step() the caller passes the camera centres of ALL ranks' views of this
    step with `set_view_centres` (a [world, 3] sequence of floats, identical on every rank)."""

    def __init__(self, params, xyz, features_dc, features_rest, sh_degree_fn, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 group=None, average=True, backend: str | None = None, timeout_s: float = 20.0):
        groups = [dict(g) if not isinstance(g, dict) else g for g in params]
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps))
        feat_ids = {id(features_dc), id(features_rest)}
        self._feat_groups = [g for g in self.param_groups if any(id(p) in feat_ids for p in g["params"])]
        early_groups = [g for g in self.param_groups if not any(id(p) in feat_ids for p in g["params"])]
        if len(self._feat_groups) != 2 or not early_groups:
            raise ValueError("PeerRecordAdam: features_dc and features_rest must each be a parameter group of their own")
        # the SAME dict objects: learning-rate changes on self.param_groups reach the inner optimizer
        self.early = PeerShardedAdam(early_groups, lr=lr, betas=betas, eps=eps, group=group, average=average,
                                     backend=backend, timeout_s=timeout_s)
        self.group, self.average, self.timeout_s = group, average, timeout_s
        self.world, self.rank, self.device = self.early.world, self.early.rank, self.early.device
        self.xyz, self.f_dc, self.f_rest = xyz, features_dc, features_rest
        self.sh_degree_fn = sh_degree_fn  # () -> active SH degree of the views just rendered
        P = int(xyz.size(0))
        self.P = P
        dev = self.device
        # this rank's colour records, readable by every peer
        self.records = PeerBuffer(max(16, 16 * P), dev, group=group, backend=backend)
        self.record_view = self.records.local[:16 * P].view(torch.float32).view(P, 4)
        self._record_ptrs = (C.c_void_p * self.world)(*self.records.ptrs)
        # K8+K9 does not write the SH gradients on this path (NULL outputs, wast3d_raster_backward_raw): the sink only
        # needs placeholder views so that the rasteriser's backward takes its arena route for the other leaves
        self._scratch = {id(features_dc): torch.empty(0, device=dev), id(features_rest): torch.empty(0, device=dev)}
        self.grad_sink = self.early.grad_sink
        self.grad_sink.views.update(self._scratch)
        self.grad_sink.record_out = self.record_view  # model_render._RasterizeModel.backward fills it
        z = torch.zeros_like
        self.m_dc, self.v_dc, self.m_rest, self.v_rest = z(features_dc), z(features_dc), z(features_rest), z(features_rest)
        self.xyz_old = torch.empty_like(xyz)
        self._feat_steps = 0
        self._barrier_epoch = 0
        self._centres = None
        self._side = torch.cuda.Stream(device=dev)
        self._bwd_done = torch.cuda.Event()
        self._late_event = torch.cuda.Event()
        self._late_pending = False
        self.overlap_late = True
        self.multicast = self.early.multicast
        self.buffer = self.early.buffer

    # -- the same consumer protocol as PeerShardedAdam
    def take_late_event(self):
        if not self._late_pending:
            return None
        self._late_pending = False
        return self._late_event

    def sync(self):
        ev = self.take_late_event()
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def set_view_centres(self, centres):
        c = torch.as_tensor(centres, dtype=torch.float32).reshape(-1, 3).cpu().contiguous()
        if c.shape[0] != self.world:
            raise ValueError(f"set_view_centres: need one camera centre per rank ({self.world}), got {c.shape[0]}")
        self._centres = c

    def _barrier(self):
        """Empty-shard launch of the peer kernel on the second flag set: all ranks arrive, all ranks leave."""
        if self.world == 1:
            return
        e = self.early
        self._barrier_epoch += 1
        arr = (_lib.AdamSegment * 1)()
        rc = _lib.load().wast3d_peer_adam_step(
            self.world, self.rank, e._grad_ptrs, e._param_ptrs, e._flag_ptrs[1], None, None, None, None,
            0, 0, arr, 0, 1.0, self._barrier_epoch, float(self.timeout_s), 1, _lib.stream_ptr())
        _lib.check(rc, "peer barrier")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._centres is None:
            raise RuntimeError("PeerRecordAdam.step: call set_view_centres() with this step's camera centres first")
        self.sync()
        lib = _lib.load()
        _, sh_adam = _staged(lib)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            self.xyz_old.copy_(self.xyz.detach())  # the views were rendered with these positions
            self._bwd_done.record(main)        # gradients, this rank's records and the xyz snapshot are complete
            self.early.step()                  # xyz / opacity / scaling / rotation: one peer launch, current stream
            self._feat_steps += 1
            hd, hr = self._feat_groups
            if hd["params"][0] is not self.f_dc:
                hd, hr = hr, hd
            def grp(p, m, v, h):
                b1, b2 = h["betas"]
                return _AdamGroup(p.data_ptr() if p.numel() else None, m.data_ptr() if p.numel() else None,
                                  v.data_ptr() if p.numel() else None, float(h["lr"]), float(b1), float(b2),
                                  float(h["eps"]), self._feat_steps, 0, None)
            gd = grp(self.f_dc, self.m_dc, self.v_dc, hd)
            gr = grp(self.f_rest, self.m_rest, self.v_rest, hr)
            M = 1 + (int(self.f_rest.size(1)) if self.f_rest.numel() else 0)
            scale = 1.0 / self.world if self.average else 1.0
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._bwd_done)
                self._barrier()                # every rank's records are complete
                rc = sh_adam(self.P, int(self.sh_degree_fn()), M, self.world, self._record_ptrs,
                             self._centres.data_ptr(), self.xyz_old.data_ptr(), scale, C.byref(gd),
                             C.byref(gr) if M > 1 else None, _lib.stream_ptr())
                _lib.check(rc, "sh_adam_from_records")
                self._barrier()                # every rank is done reading my records
                self._late_event.record(self._side)
            self._late_pending = True
        self._centres = None
        return loss

    def zero_grad(self, set_to_none: bool = True):
        self.early.zero_grad(set_to_none)
        for p in (self.f_dc, self.f_rest):
            p.grad = None

    def check_peers(self):
        self.early.check_peers()

    def close(self):
        self.sync()
        self.early.close()
        self.records.close()


def write_colour_records(sink, P, radii, geom, stream_ptr):
    """Called by model_render._RasterizeModel.backward after the backward kernels when the grad sink carries a
    `record_out` buffer: this view's 16-byte colour records."""
    out = getattr(sink, "record_out", None)
    if out is None or P == 0:
        return
    colour_records, _ = _staged(_lib.load())
    rc = colour_records(int(P), radii.data_ptr(), geom.data_ptr(), out.data_ptr(), stream_ptr)
    _lib.check(rc, "colour_records")
