"""Pixel losses of the style-optimisation loop with the reference's names (utils/loss_utils.py:18-19
`l1_loss`, :213-215 `tv_loss`; called at train_st_normals.py:127,145), computed by the fused kernels of
csrc/loss.cu: one pass over the image forward, one pass backward, instead of ~40 torch kernels.
No CPU or torch fallback: CUDA tensors and the built library are required."""
from __future__ import annotations

import torch

from . import _lib

_SCRATCH: dict = {}


def _scratch(dev: torch.device) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _SCRATCH.get(key)
    if t is None:
        t = _SCRATCH[key] = torch.zeros(int(_lib.load().wast3d_pixel_loss_scratch_bytes()), dtype=torch.uint8, device=dev)
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


class _PixelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, depth, gt, depth_gt, w_l1, w_tv, w_depth):
        _lib.require_device(img)
        if img.dim() != 3 or img.dtype != torch.float32:
            raise RuntimeError("pixel_loss: img must be float32 [C,H,W]")
        img = img.contiguous()
        C, H, W = (int(v) for v in img.shape)
        for name, t, shape in (("gt", gt, (C, H, W)), ("depth", depth, (H, W)), ("depth_gt", depth_gt, (H, W))):
            if t is not None and (tuple(t.shape) != shape or t.dtype != torch.float32 or not t.is_cuda):
                raise RuntimeError(f"pixel_loss: {name} must be a float32 CUDA tensor of shape {shape}")
        if (depth is None) != (depth_gt is None):
            raise RuntimeError("pixel_loss: depth and depth_gt go together")
        gt = None if gt is None else gt.contiguous()
        depth = None if depth is None else depth.contiguous()
        depth_gt = None if depth_gt is None else depth_gt.contiguous()
        out = torch.empty((), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            rc = _lib.load().wast3d_pixel_loss_forward(
                C, H, W, img.data_ptr(), _ptr(gt), _ptr(depth), _ptr(depth_gt), float(w_l1), float(w_tv),
                float(w_depth), _scratch(img.device).data_ptr(), out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "pixel_loss_forward")
        ctx.weights = (float(w_l1), float(w_tv), float(w_depth))
        ctx.has = (gt is not None, depth is not None)
        e = torch.empty(0, device=img.device)
        ctx.save_for_backward(img, gt if gt is not None else e, depth if depth is not None else e,
                              depth_gt if depth_gt is not None else e)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        img, gt, depth, depth_gt = ctx.saved_tensors
        has_gt, has_depth = ctx.has
        C, H, W = (int(v) for v in img.shape)
        d_img = torch.empty_like(img)
        d_depth = torch.empty_like(depth) if has_depth and ctx.needs_input_grad[1] else None
        go = grad_out.to(torch.float32).contiguous()
        w_l1, w_tv, w_depth = ctx.weights
        with torch.cuda.device(img.device):
            rc = _lib.load().wast3d_pixel_loss_backward(
                C, H, W, img.data_ptr(), gt.data_ptr() if has_gt else None, depth.data_ptr() if has_depth else None,
                depth_gt.data_ptr() if has_depth else None, w_l1, w_tv, w_depth, go.data_ptr(), d_img.data_ptr(),
                _ptr(d_depth), _lib.stream_ptr())
        _lib.check(rc, "pixel_loss_backward")
        return d_img, d_depth, None, None, None, None, None


def pixel_loss(img, gt=None, depth=None, depth_gt=None, w_l1: float = 1.0, w_tv: float = 0.0, w_depth: float = 0.0):
    """w_l1 * l1_loss(img, gt) + w_tv * tv_loss(img) + w_depth * mean((depth - depth_gt)**2) in one pass."""
    return _PixelLoss.apply(img, depth, gt, depth_gt, w_l1, w_tv, w_depth)


def l1_loss(network_output, gt):
    """utils/loss_utils.py:18-19."""
    return pixel_loss(network_output, gt, w_l1=1.0)


def tv_loss(img):
    """utils/loss_utils.py:213-215."""
    return pixel_loss(img, None, w_l1=0.0, w_tv=1.0)
