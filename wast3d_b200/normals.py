"""Depth -> normal map of the train_st_normals variant (train_st_normals.py:108-123), fused.

The reference writes it with kornia and torch:

    K = torch.tensor([[1111, 0, 400], [0, 1111, 400], [0, 0, 1]])[None]
    normals = kornia.geometry.depth.depth_to_normals(depth=image_depth[None, None], camera_matrix=K,
                                                     normalize_points=False)
    image_normals = normals.squeeze(0)
    mins, maxs = torch.amin(image_normals, (0, 1, 2), True), torch.amax(image_normals, (0, 1, 2), True)
    image_normals = (image_normals - mins) / (maxs - mins + 1e-6)

`depth_to_normals01(depth, fx, fy, cx, cy)` returns the same [3,H,W] image from two kernels forward and three
backward (csrc/normals.cu), differentiable with respect to the depth image (which is how the normal-style loss
reaches the Gaussians through the rasteriser's depth output).  No CPU or torch fallback.
"""
from __future__ import annotations

import math

import torch

from . import _lib

_SCRATCH: dict = {}


def _scratch(dev: torch.device) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _SCRATCH.get(key)
    if t is None:
        t = _SCRATCH[key] = torch.zeros(int(_lib.load().wast3d_depth_normals_scratch_bytes()), dtype=torch.uint8, device=dev)
    return t


class _DepthNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, fx, fy, cx, cy):
        _lib.require_device(depth)
        if depth.dim() != 2 or depth.dtype != torch.float32:
            raise RuntimeError("depth_to_normals01: depth must be float32 [H,W]")
        depth = depth.contiguous()
        H, W = int(depth.shape[0]), int(depth.shape[1])
        dev = depth.device
        unit = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        minmax = torch.empty((2,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().wast3d_depth_normals_forward(H, W, depth.data_ptr(), float(fx), float(fy), float(cx), float(cy),
                                                          unit.data_ptr(), out.data_ptr(), minmax.data_ptr(),
                                                          _scratch(dev).data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "depth_normals_forward")
        ctx.k = (float(fx), float(fy), float(cx), float(cy))
        ctx.save_for_backward(depth, unit, minmax)
        return out

    @staticmethod
    def backward(ctx, g):
        depth, unit, minmax = ctx.saved_tensors
        H, W = int(depth.shape[0]), int(depth.shape[1])
        dev = depth.device
        g = g.to(torch.float32).contiguous()
        gab = torch.empty((6, H, W), dtype=torch.float32, device=dev)
        gd = torch.empty((H, W), dtype=torch.float32, device=dev)
        fx, fy, cx, cy = ctx.k
        with torch.cuda.device(dev):
            rc = _lib.load().wast3d_depth_normals_backward(H, W, depth.data_ptr(), fx, fy, cx, cy, unit.data_ptr(),
                                                           minmax.data_ptr(), g.data_ptr(), gab.data_ptr(), gd.data_ptr(),
                                                           _scratch(dev).data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "depth_normals_backward")
        return gd, None, None, None, None


def depth_to_normals01(depth: torch.Tensor, fx: float, fy: float, cx: float, cy: float) -> torch.Tensor:
    """[3,H,W] normal image rescaled to [0,1] by its global min / max, from a depth image [H,W] and pinhole
    intrinsics — train_st_normals.py:113-123 in one call."""
    return _DepthNormals.apply(depth, fx, fy, cx, cy)


def intrinsics_for(camera):
    """(fx, fy, cx, cy) of a scene.Camera the way the reference's hard-coded K relates to its 800x800 / FoVx = 0.6911
    cameras (1111 = 800 / (2 tan(FoVx / 2)), 400 = 800 / 2; train_st_normals.py:114-116)."""
    W, H = int(camera.image_width), int(camera.image_height)
    return (W / (2.0 * math.tan(camera.FoVx * 0.5)), H / (2.0 * math.tan(camera.FoVy * 0.5)), W / 2.0, H / 2.0)
