"""`render()` glue with the reference's signature and return dict
(gaussian_renderer/__init__.py:18-115), calling the B200 rasteriser.

`pc` is anything with the GaussianModel getters used by the reference
(get_xyz, get_opacity, get_scaling, get_rotation, get_features, get_covariance,
active_sh_degree, max_sh_degree); `viewpoint_camera` anything with the Camera attributes
(FoVx, FoVy, image_height, image_width, world_view_transform, full_proj_transform,
camera_center) — see scene/cameras.py:17-59.
"""
from __future__ import annotations

import math

import torch

from .diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from .model_render import model_supports_fusion, rasterize_model
from .sh import eval_sh


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0,
           override_color=None, sampling_offsets=None):
    """Render the scene; background tensor must be on the GPU.

    `sampling_offsets` (extension) lets a caller pin the per-pixel jitter; by default it is
    drawn like the reference does: -rand(H, W, 2) in (-1, 0] (gaussian_renderer/__init__.py:31).

    When `pc` is a GaussianModel with the reference's activations and neither Python-side SH nor
    covariance is requested, the getters' sigmoid / exp / normalize / cat are folded into the
    kernels (model_render.py); `pipe.fused_activations = False` forces the reference's op-by-op
    call chain.  Both produce the same dict.
    """
    xyz = pc.get_xyz
    dev = xyz.device
    H, W = int(viewpoint_camera.image_height), int(viewpoint_camera.image_width)
    if sampling_offsets is None:
        sampling_offsets = torch.rand(H, W, 2, device=dev).mul_(-1)
    fused = (override_color is None and not pipe.convert_SHs_python and not pipe.compute_cov3D_python
             and getattr(pipe, "fused_activations", True) and model_supports_fusion(pc))
    # zero tensors that only exist to carry gradients out (reference :26-36)
    if fused:
        # the model-space op never reads means2D: a stride-0 view of one zero row carries the gradient just as well
        # (viewspace_points.grad is still the dense [P,3] tensor), without a 36 MB fill + add per render
        screenspace_points = torch.zeros((1, xyz.shape[1]), dtype=xyz.dtype, device=dev).expand(xyz.shape[0], -1)
        screenspace_points.requires_grad_(True)
        cam_view_depth = None
    else:
        screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
        cam_view_depth = torch.zeros(xyz.shape[:-1] + (1,), dtype=xyz.dtype, device=dev,
                                     requires_grad=True) + 0
        try:
            screenspace_points.retain_grad()
            cam_view_depth.retain_grad()
        except Exception:
            pass

    settings = GaussianRasterizationSettings(
        image_height=H, image_width=W,
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=pipe.debug)
    # view-parallel optimizer with overlapped feature exchange (peer.PeerShardedAdam late_params): the SH
    # features of the last step() may still be in flight; only code behind this event may read them
    take = getattr(getattr(pc, "optimizer", None), "take_late_event", None)
    features_ready = take() if take is not None else None
    if fused:
        image, depth, radii = rasterize_model(
            xyz, screenspace_points, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling,
            pc._rotation, settings, sampling_offsets, getattr(pc, "grad_sink", None), features_ready)
        return {"render": image, "depth": depth, "viewspace_points": screenspace_points,
                "visibility_filter": radii > 0, "radii": radii}

    if features_ready is not None:  # op-by-op path: torch kernels read the features right away
        torch.cuda.current_stream(dev).wait_event(features_ready)
    rasterizer = GaussianRasterizer(raster_settings=settings)

    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation

    shs = colors_precomp = None
    if override_color is not None:
        colors_precomp = override_color
    elif pipe.convert_SHs_python:
        feats = pc.get_features
        shs_view = feats.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
        dir_pp = xyz - viewpoint_camera.camera_center.repeat(feats.shape[0], 1)
        dir_pp = dir_pp / dir_pp.norm(dim=1, keepdim=True)
        colors_precomp = torch.clamp_min(eval_sh(pc.active_sh_degree, shs_view, dir_pp) + 0.5, 0.0)
    else:
        shs = pc.get_features

    image, depth, radii = rasterizer(
        means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=pc.get_opacity, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
        cam_view_depth=cam_view_depth, sampling_offsets=sampling_offsets)

    return {"render": image, "depth": depth, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}
