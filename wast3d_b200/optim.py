"""Fused Adam (csrc/adam.cu) with torch.optim.Adam's interface for the reference's six
parameter groups (scene/gaussian_model.py:154-163).  One kernel per parameter tensor, one pass
over (param, grad, exp_avg, exp_avg_sq); same update rule as torch's single-tensor Adam.

`BackwardFusedAdam` goes one step further: the rasteriser's per-Gaussian backward
kernel applies the update itself (wast3d_raster_backward_raw_adam), so the leaf gradients never touch HBM
and optimizer.step() has nothing left to launch."""
from __future__ import annotations

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def begin_step(self):
        """Advance the step count of every parameter that has a gradient (once per optimizer step);
        step_range() then applies the update to any sub-range of a parameter."""
        self._group_of = {}
        for group in self.param_groups:
            for p in group["params"]:
                self._group_of[p] = group
                if p.grad is None:
                    continue
                _lib.require_device(p)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous float32")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1

    @torch.no_grad()
    def step_range(self, p, start: int = 0, end: int = None):
        """Adam update of the flat element range [start, end) of parameter p (after begin_step())."""
        if p.grad is None:
            return
        end = p.numel() if end is None else end
        if end <= start:
            return
        group, st = self._group_of[p], self.state[p]
        b1, b2 = group["betas"]
        g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
        off = 4 * start
        with torch.cuda.device(p.device):
            rc = _lib.load().wast3d_adam_step(
                end - start, p.data_ptr() + off, g.data_ptr() + off, st["exp_avg"].data_ptr() + off,
                st["exp_avg_sq"].data_ptr() + off, float(group["lr"]), float(b1), float(b2),
                float(group["eps"]), int(st["step"]), _lib.stream_ptr())
        _lib.check(rc, "adam_step")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.begin_step()
        for group in self.param_groups:
            for p in group["params"]:
                self.step_range(p)
        return loss


class AdamInBackwardSink:
    """Handed to model_render.rasterize_model as `grad_sink`: tells the rasteriser's backward to apply
    the optimizer update in its per-Gaussian kernel.  `fresh` is True until the first backward after
    zero_grad(); a second backward in the same optimizer step is refused (its gradients would have to be
    summed with the first one's before the update)."""

    def __init__(self, optimizer):
        self.fused_adam = optimizer
        self.fresh = True

    def view_for(self, p):  # peer.GradSink interface: no gradient storage here
        return None


class BackwardFusedAdam(FusedAdam):
    """FusedAdam whose update is applied inside the rasteriser's backward kernel.

    Requirements: exactly the six GaussianModel leaves, one per group, in training_setup()'s order
    (xyz, f_dc, f_rest, opacity, scaling, rotation); one render()+backward() per step(); one GPU (the
    backward must carry the step's whole gradient).  step() only finishes the bookkeeping; if the
    gradients arrived through plain autograd instead (p.grad set), it falls back to FusedAdam's dense
    kernels over those gradients — same arithmetic, two more passes over HBM."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        if len(self.param_groups) != 6 or any(len(g["params"]) != 1 for g in self.param_groups):
            raise ValueError("BackwardFusedAdam: expected the six GaussianModel parameter groups, one tensor each")
        self.grad_sink = AdamInBackwardSink(self)
        self._applied = False
        self.capture_grads = False  # test hook: keep the leaf gradients of the last backward in .last_grads
        self.last_grads = None
        self._schedule = None       # device_schedule(): [6, 2] step-dependent scalars read by the kernel at run time
        self._hyper = self._step_dev = self._hyper_host = None
        self._next_view = None      # request: project for this camera inside the next backward (prefetch_view)
        self.projection = None      # result: pre-filled geometry buffer + radii of the view projected last

    def prefetch_view(self, viewpoint_camera, scaling_modifier: float = 1.0, sh_degree=None,
                      offset_bounds=(-1.0, 0.0, -1.0, 0.0)):
        """Announce the camera of the NEXT render() before this step's loss.backward().  The per-Gaussian backward
        kernel then also projects every Gaussian for that view from the parameter values it has just updated
        (wast3d_raster_backward_raw_adam_next), and the next render() of that camera starts at the depth sort: K1 and
        its re-read of all parameters disappear from the step.  Contract: nothing else modifies the parameters between
        this backward and that render(); a render() of any other camera / size simply projects as usual.
        `offset_bounds` = (min x, max x, min y, max y) of the next call's sampling offsets; the default is the range of
        the reference's own draw, `rand * -1` (gaussian_renderer/__init__.py:31); the next forward verifies it.
        `sh_degree`: the model's active SH degree at the next render (default: unchanged)."""
        self._next_view = dict(cam=viewpoint_camera, scale=float(scaling_modifier), D=sh_degree,
                               bounds=tuple(float(b) for b in offset_bounds))

    @torch.no_grad()
    def device_schedule(self, on: bool = True):
        """Keep the step count and the bias-corrected step sizes on the DEVICE (wast3d_adam_schedule_step): the launch
        sequence of a step then carries no step-dependent host value and can be replayed as a CUDA graph
        (wast3d_b200.graphed.GraphedStep).  Learning rates are read from param_groups at every sync_hyper() call
        (GraphedStep does it before each replay); betas from the groups at this call."""
        if not on:
            self._schedule = self._hyper = self._step_dev = self._hyper_host = None
            return
        p0 = self.param_groups[0]["params"][0]
        steps = {int(self.state[g["params"][0]].get("step", 0)) if self.state[g["params"][0]] else 0 for g in self.param_groups}
        if len(steps) != 1:
            raise RuntimeError("BackwardFusedAdam.device_schedule: the six groups must be at the same step")
        self._schedule = torch.zeros(6, 2, dtype=torch.float32, device=p0.device)
        self._step_dev = torch.full((1,), steps.pop(), dtype=torch.int64, device=p0.device)
        self._hyper_host = torch.zeros(6, 3, dtype=torch.float64, pin_memory=True)
        self._hyper = torch.zeros(6, 3, dtype=torch.float64, device=p0.device)
        self.sync_hyper(force=True)

    @torch.no_grad()
    def sync_hyper(self, force: bool = False):
        """Copy the groups' current lr / betas to the device table when they changed (device_schedule mode)."""
        if self._schedule is None:
            return
        rows = [[float(g["lr"]), float(g["betas"][0]), float(g["betas"][1])] for g in self.param_groups]
        new = torch.tensor(rows, dtype=torch.float64)
        if force or not torch.equal(new, self._hyper_host):
            self._hyper_host.copy_(new)
            self._hyper.copy_(self._hyper_host, non_blocking=True)

    def advance_host_steps(self, n: int = 1):
        """Book-keeping after a graph replay: the device counter advanced, state['step'] follows."""
        for g in self.param_groups:
            st = self.state[g["params"][0]]
            if st:
                st["step"] += n

    @torch.no_grad()
    def adam_groups(self, leaves):
        """ctypes array of wast3d_adam_group for this update (advances the step counts)."""
        mine = [g["params"][0] for g in self.param_groups]
        if len(leaves) != 6 or any(a is not b for a, b in zip(mine, leaves)):
            raise RuntimeError("BackwardFusedAdam: the rasterised tensors are not this optimizer's parameters "
                               "(order: xyz, f_dc, f_rest, opacity, scaling, rotation)")
        arr = (_lib.AdamGroup * 6)()
        for k, (group, p) in enumerate(zip(self.param_groups, mine)):
            _lib.require_device(p)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("BackwardFusedAdam: parameters must be contiguous float32")
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            b1, b2 = group["betas"]
            ptr = lambda t: t.data_ptr() if t.numel() else None
            arr[k] = _lib.AdamGroup(ptr(p), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), float(group["lr"]), float(b1),
                                    float(b2), float(group["eps"]), int(st["step"]), 0,
                                    self._schedule[k].data_ptr() if self._schedule is not None else None)
        if self._schedule is not None:
            # the kernel reads lr / (1 - beta1^t) and 1 / sqrt(1 - beta2^t) from the device: advance them on this stream
            _lib.check(_lib.load().wast3d_adam_schedule_step(6, self._hyper.data_ptr(), self._step_dev.data_ptr(),
                                                            self._schedule.data_ptr(), _lib.stream_ptr()),
                       "adam_schedule_step")
        return arr

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("BackwardFusedAdam: closures are not supported")
        if self._applied:
            # the update already happened inside the rasteriser's backward, from the render gradient alone: a
            # gradient that autograd left in .grad (a regulariser on the leaves in the same loss.backward())
            # was not part of it
            stray = [i for i, g in enumerate(self.param_groups) if g["params"][0].grad is not None]
            if stray:
                self._applied = False
                raise RuntimeError(
                    "BackwardFusedAdam: parameter group(s) %s received a gradient outside the rasteriser's backward; "
                    "the in-backward update only sees the render gradient. Use training_setup(fused=True) when the "
                    "loss has terms that reach the leaves without passing through render()." % stray)
            self._applied = False
            return None
        return super().step()

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        self.grad_sink.fresh = True
