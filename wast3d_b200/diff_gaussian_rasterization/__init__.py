"""Drop-in for the reference's `diff_gaussian_rasterization` Python package
(submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py).

Public surface kept verbatim so `gaussian_renderer/__init__.py:14` and the `train_st*` scripts
work unchanged:

    GaussianRasterizationSettings   12-field NamedTuple          (reference :173-185)
    GaussianRasterizer              nn.Module, forward(...)      (reference :187-238)
    rasterize_gaussians             functional entry             (reference :21-46)

Semantics reproduced from the reference:
  * returns (color[3,H,W], depth[H,W], radii[P]) and back-propagates both image gradients;
  * `cam_view_depth` is accepted but unused and gets no gradient (:66-87, :152-164);
  * exactly one of shs / colors_precomp, and scales+rotations xor cov3D_precomp (:207-211);
  * "not provided" tensors become empty tensors before crossing into `_C` (:213-223);
  * with `debug` the arguments are snapshotted to snapshot_fw.dump / snapshot_bw.dump when the
    native call raises (:90-97, :141-148).
One extension: `sampling_offsets=None` means zero jitter (the reference cannot pass None
through pybind).
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _snapshot(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def _call_native(fn, args, debug: bool, dump_name: str, what: str):
    if not debug:
        return fn(*args)
    saved = _snapshot(args)  # taken before the call so a crash cannot corrupt it
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_name)
        print(f"\nAn error occured in {what}. Please forward {dump_name} for debugging.")
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings, cam_view_depth, sampling_offsets):
        rs = raster_settings
        if sampling_offsets is None:
            sampling_offsets = torch.empty(0)
        native_args = (
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
            rs.debug, sampling_offsets)
        (num_rendered, color, depth, radii, geom_buf, binning_buf, img_buf) = _call_native(
            _C.rasterize_gaussians, native_args, rs.debug, "snapshot_fw.dump", "forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii,
                              sh, geom_buf, binning_buf, img_buf, sampling_offsets)
        ctx.mark_non_differentiable(radii)
        return color, depth, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth, _grad_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom_buf,
         binning_buf, img_buf, sampling_offsets) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, rs.image_height, rs.image_width),
                                         dtype=means3D.dtype, device=means3D.device)
        if grad_out_depth is None:
            grad_out_depth = torch.zeros((rs.image_height, rs.image_width),
                                         dtype=means3D.dtype, device=means3D.device)
        native_args = (
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color,
            grad_out_depth, sh, rs.sh_degree, rs.campos, geom_buf, ctx.num_rendered, binning_buf,
            img_buf, rs.debug, sampling_offsets)
        (g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot) = _call_native(
            _C.rasterize_gaussians_backward, native_args, rs.debug, "snapshot_bw.dump", "backward")
        # order of forward()'s inputs; settings, cam_view_depth and sampling_offsets get None
        return (g_means3D, g_means2D, g_sh, g_colors, g_opacity, g_scales, g_rot, g_cov3D,
                None, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings, cam_view_depth, sampling_offsets):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings, cam_view_depth,
                                     sampling_offsets)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points in front of the near plane (z_view > 0.2)."""
        rs = self.raster_settings
        with torch.no_grad():
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, cam_view_depth=None, sampling_offsets=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr_part = scales is not None or rotations is not None
        has_sr_full = scales is not None and rotations is not None
        if (not has_sr_full and cov3D_precomp is None) or (has_sr_part and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or '
                            'precomputed 3D covariance!')
        empty = torch.Tensor([])
        return rasterize_gaussians(
            means3D, means2D,
            empty if shs is None else shs,
            empty if colors_precomp is None else colors_precomp,
            opacities,
            empty if scales is None else scales,
            empty if rotations is None else rotations,
            empty if cov3D_precomp is None else cov3D_precomp,
            self.raster_settings, cam_view_depth, sampling_offsets)
