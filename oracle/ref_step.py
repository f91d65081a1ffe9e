"""TEST / BASELINE INFRASTRUCTURE ONLY — one style-optimisation iteration exactly as the reference runs
it on a GPU: the UNMODIFIED reference CUDA rasteriser (oracle/_ref/libwast3d_ref.so, built by
oracle/build_ref.sh from /root/reference) behind the reference's own Python glue, restated here:

  * `_RefRasterize`    = `_RasterizeGaussians` (submodules/diff-gaussian-rasterization/
                         diff_gaussian_rasterization/__init__.py:50-171): forward saves the inputs, backward
                         returns the 11-tuple of gradients in the reference's order; the ten zero-initialised
                         gradient tensors of rasterize_points.cu:157-166 are allocated per call like there;
  * `ref_render`       = `render()` (gaussian_renderer/__init__.py:18-115) with the GaussianModel getters
                         (scene/gaussian_model.py:95-120): sigmoid / exp / normalize / cat as torch kernels;
  * `RefTrainer.step`  = the loop body of train_st_normals.py:107-194 with the bench's synthetic loss and
                         `torch.optim.Adam` over the six groups (scene/gaussian_model.py:149-167).

Nothing in wast3d_b200/ imports this.  Users: tests/ (full-size parity of the fused model-space path) and
bench.py's reference legs (`--impl reference`, `extra.reference_cuda_sm100`)."""
from __future__ import annotations

import math

import torch

from . import ref


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, opacities, scales, rotations, cam_view_depth, cfg, offsets_padded, rr):
        H, W = cfg["H"], cfg["W"]
        out = rr.forward(bg=cfg["bg"], means3D=means3D, opacities=opacities, view=cfg["view"], proj=cfg["proj"],
                         campos=cfg["campos"], W=W, H=H, tan_fovx=cfg["tan_fovx"], tan_fovy=cfg["tan_fovy"],
                         shs=sh, scales=scales, rotations=rotations, D=cfg["D"], scale_modifier=1.0,
                         sampling_offsets=offsets_padded, offsets_prepadded=True)
        ctx.rr = rr
        ctx.mark_non_differentiable(out["radii"])
        return out["color"], out["depth"], out["radii"]

    @staticmethod
    def backward(ctx, g_color, g_depth, _g_radii):
        rr = ctx.rr
        H, W = rr.args["H"], rr.args["W"]
        dev = rr.radii.device
        if g_color is None:
            g_color = torch.zeros(3, H, W, device=dev)
        if g_depth is None:
            g_depth = torch.zeros(H, W, device=dev)
        g = rr.backward(g_color.contiguous(), g_depth.contiguous())
        # (means3D, means2D, sh, opacities, scales, rotations, cam_view_depth, cfg, offsets, rr)
        return (g["dL_dmean3D"], g["dL_dmean2D"], g["dL_dsh"], g["dL_dopacity"], g["dL_dscale"], g["dL_drot"],
                g["dL_dviewdepth"], None, None, None)


def pad_offsets(offsets: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """The reference reads sampling_offsets of out-of-image threads (forward.cu:287): give it
    an allocation in which those reads stay in bounds (done once per offsets tensor, outside any timing)."""
    pad = torch.zeros(((H + 16) * (W + 16), 2), device=offsets.device)
    pad[: H * W] = offsets.reshape(-1, 2)
    return pad


def ref_render(cam, leaves, active_sh_degree, bg, offsets_padded, rr):
    """gaussian_renderer/__init__.py:18-115 on the six leaf tensors (xyz, f_dc, f_rest, opacity, scaling, rotation)."""
    xyz, f_dc, f_rest, opacity, scaling, rotation = leaves
    screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
    cam_view_depth = torch.zeros(xyz.shape[:-1] + (1,), dtype=xyz.dtype, device=xyz.device, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
        cam_view_depth.retain_grad()
    except Exception:  # noqa: BLE001
        pass
    cfg = dict(H=int(cam.image_height), W=int(cam.image_width), tan_fovx=math.tan(cam.FoVx * 0.5),
               tan_fovy=math.tan(cam.FoVy * 0.5), bg=bg, view=cam.world_view_transform,
               proj=cam.full_proj_transform, campos=cam.camera_center, D=int(active_sh_degree))
    # the GaussianModel getters (scene/gaussian_model.py:95-120)
    shs = torch.cat((f_dc, f_rest), dim=1)
    image, depth, radii = _RefRasterize.apply(xyz, screenspace_points, shs, torch.sigmoid(opacity), torch.exp(scaling),
                                              torch.nn.functional.normalize(rotation), cam_view_depth, cfg,
                                              offsets_padded, rr)
    return {"render": image, "depth": depth, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}


class RefTrainer:
    """The reference's optimisation loop state: six leaf tensors + torch.optim.Adam(eps=1e-15), six groups with
    the learning rates of arguments/__init__.py (scene/gaussian_model.py:149-167)."""

    def __init__(self, arrs: dict, spatial_lr_scale: float, device, active_sh_degree: int = 3, asynchronous=True):
        t = lambda k: torch.nn.Parameter(torch.as_tensor(arrs[k]).float().to(device).contiguous())
        self.leaves = [t("xyz"), t("f_dc"), t("f_rest"), t("opacity_logits"), t("log_scales"), t("rotations")]
        lrs = [0.00016 * spatial_lr_scale, 0.0025, 0.0025 / 20.0, 0.05, 0.005, 0.001]
        self.optimizer = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(self.leaves, lrs)],
                                          lr=0.0, eps=1e-15)
        self.D = active_sh_degree
        self.rr = ref.RefRasterizer(asynchronous=asynchronous)

    def render(self, cam, bg, offsets_padded):
        return ref_render(cam, self.leaves, self.D, bg, offsets_padded, self.rr)

    def step(self, cam, bg, offsets_padded, loss_fn):
        out = self.render(cam, bg, offsets_padded)
        loss = loss_fn(out)
        loss.backward()
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return loss
