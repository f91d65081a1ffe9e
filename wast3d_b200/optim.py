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
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
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
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                with torch.cuda.device(p.device):
                    rc = lib.wast3d_adam_step(p.numel(), p.data_ptr(), g.data_ptr(),
                                              st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                              float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                              int(st["step"]), _lib.stream_ptr())
                _lib.check(rc, "adam_step")
        return loss
