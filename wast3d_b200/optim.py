"""Fused Adam (csrc/adam.cu) with torch.optim.Adam's interface for the reference's six
parameter groups (scene/gaussian_model.py:154-163).  One kernel per parameter tensor, one pass
over (param, grad, exp_avg, exp_avg_sq); same update rule as torch's single-tensor Adam."""
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
