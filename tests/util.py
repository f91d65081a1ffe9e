"""Shared input builders for the parity tests (seeded, sizes the oracle finishes in seconds)."""
from __future__ import annotations

import math

import numpy as np
import torch

from wast3d_b200.scene import orbit_cameras, synthetic_gaussians, fov_y_from_x


def raster_case(P=20000, W=320, H=240, seed=0, cam_index=1, degree=3, jitter=True, bg=(0.0, 0.0, 0.0),
                log_scale_mu=-3.6, radius=4.03, fovx=0.6911, garden=False, use_precomp_color=False,
                use_precomp_cov=False, scale_modifier=1.0):
    """Activated rasteriser inputs as float32 numpy arrays (what `_C.rasterize_gaussians` gets)."""
    g = synthetic_gaussians(P, seed=seed, garden=garden, log_scale_mu=log_scale_mu)
    cam = orbit_cameras(8, radius, 2.0 if garden else 0.0, fovx, W, H, device="cpu", sphere=not garden)[cam_index]
    q = g["rotations"] / np.linalg.norm(g["rotations"], axis=1, keepdims=True)
    rng = np.random.default_rng(seed + 1000)
    case = dict(
        W=W, H=H, tan_fovx=math.tan(cam.FoVx / 2), tan_fovy=math.tan(cam.FoVy / 2),
        bg=np.asarray(bg, np.float32), means3D=g["xyz"],
        opacities=(1.0 / (1.0 + np.exp(-g["opacity_logits"]))).astype(np.float32),
        view=cam.world_view_transform.numpy().copy(), proj=cam.full_proj_transform.numpy().copy(),
        campos=cam.camera_center.numpy().copy(), D=degree, scale_modifier=scale_modifier,
        sampling_offsets=(-rng.random((H, W, 2))).astype(np.float32) if jitter else None)
    shs = np.concatenate([g["f_dc"], g["f_rest"]], 1).astype(np.float32)
    if use_precomp_color:
        case["colors_precomp"] = rng.random((P, 3)).astype(np.float32)
    else:
        case["shs"] = shs
    scales = np.exp(g["log_scales"]).astype(np.float32)
    if use_precomp_cov:
        from wast3d_b200.scene import covariance_from_scaling_rotation
        case["cov3D_precomp"] = covariance_from_scaling_rotation(
            torch.from_numpy(scales), scale_modifier, torch.from_numpy(q.astype(np.float32))).numpy()
    else:
        case["scales"], case["rotations"] = scales, q.astype(np.float32)
    return case


def to_cuda(case: dict) -> dict:
    return {k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if isinstance(v, np.ndarray) else v)
            for k, v in case.items()}


EMPTY = None


def call_forward(tc: dict, debug=False, colour_wait_event=None):
    """Run our `_C.rasterize_gaussians` on a CUDA case dict."""
    from wast3d_b200.diff_gaussian_rasterization import _C
    e = torch.empty(0)
    g = lambda k: tc.get(k) if tc.get(k) is not None else e
    return _C.rasterize_gaussians(
        tc["bg"], tc["means3D"], g("colors_precomp"), tc["opacities"], g("scales"), g("rotations"),
        tc["scale_modifier"], g("cov3D_precomp"), tc["view"], tc["proj"], tc["tan_fovx"], tc["tan_fovy"],
        tc["H"], tc["W"], g("shs"), tc["D"], tc["campos"], False, debug, g("sampling_offsets"),
        _colour_wait_event=colour_wait_event)


def call_backward(tc: dict, fwd, dL_dpix, dL_ddepth, debug=False, scratch=True):
    from wast3d_b200.diff_gaussian_rasterization import _C
    e = torch.empty(0)
    g = lambda k: tc.get(k) if tc.get(k) is not None else e
    R, color, depth, radii, geom, binning, img = fwd
    return _C.rasterize_gaussians_backward(
        tc["bg"], tc["means3D"], radii, g("colors_precomp"), g("scales"), g("rotations"),
        tc["scale_modifier"], g("cov3D_precomp"), tc["view"], tc["proj"], tc["tan_fovx"], tc["tan_fovy"],
        dL_dpix, dL_ddepth, g("shs"), tc["D"], tc["campos"], geom, R, binning, img, debug,
        g("sampling_offsets"), _return_scratch=scratch)


def export(tc: dict, fwd):
    from wast3d_b200.diff_gaussian_rasterization import _C
    R, color, depth, radii, geom, binning, img = fwd
    kw = dict(P=tc["means3D"].shape[0], D=tc["D"], M=(tc["shs"].shape[1] if tc.get("shs") is not None else 0),
              W=tc["W"], H=tc["H"], tan_fovx=tc["tan_fovx"], tan_fovy=tc["tan_fovy"],
              scale_modifier=tc["scale_modifier"], prefiltered=False, debug=False, bg=tc["bg"],
              means3D=tc["means3D"], sh=tc.get("shs"), colors=tc.get("colors_precomp"),
              opacity=tc["opacities"], scales=tc.get("scales"), rotations=tc.get("rotations"),
              cov3D_precomp=tc.get("cov3D_precomp"), viewmatrix=tc["view"], projmatrix=tc["proj"],
              campos=tc["campos"], sampling_offsets=tc.get("sampling_offsets"))
    return _C.export_state(kw, R, geom, binning, img)


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / max(b.norm().item(), 1e-30))


class CutMapping:
    """Relation between the reference's per-tile blend lists (every tile of the radius rectangle,
    rasterizer_impl.cu:70-113) and ours under the tile cut (wast3d_set_tile_cut(1)): ours must be the
    reference's lists with some instances removed and the order of the rest preserved.  Raises
    AssertionError otherwise.  All arrays are numpy; ranges are [T,2]."""

    def __init__(self, ref_ranges, ref_pl, our_ranges, our_pl, P):
        T = ref_ranges.shape[0]
        assert our_ranges.shape[0] == T
        rsz = (ref_ranges[:, 1].astype(np.int64) - ref_ranges[:, 0].astype(np.int64))
        osz = (our_ranges[:, 1].astype(np.int64) - our_ranges[:, 0].astype(np.int64))
        assert (rsz >= 0).all() and (osz >= 0).all() and (osz <= rsz).all()
        assert rsz.sum() == len(ref_pl) and osz.sum() == len(our_pl)

        def tile_major(ranges, sizes, pl):
            tile = np.repeat(np.arange(T, dtype=np.int64), sizes)
            first = np.cumsum(sizes) - sizes
            pos = np.arange(len(pl), dtype=np.int64) - np.repeat(first, sizes) + np.repeat(ranges[:, 0].astype(np.int64), sizes)
            return tile, pl.astype(np.int64)[pos], first

        rt, rp, self.ref_first = tile_major(ref_ranges, rsz, ref_pl)
        ot, op, _ = tile_major(our_ranges, osz, our_pl)
        rk, ok = rt * int(P) + rp, ot * int(P) + op
        assert len(np.unique(rk)) == len(rk)
        self.kept = np.isin(rk, ok)
        assert int(self.kept.sum()) == len(ok), "ours holds instances the reference does not"
        assert np.array_equal(rk[self.kept], ok), "order of the kept instances differs from the reference's"
        self.kept_prefix = np.concatenate([[0], np.cumsum(self.kept)]).astype(np.int64)
        self.ref_sizes, self.our_sizes = rsz, osz

    def map_n_contrib(self, n_ref, W, H):
        """The reference's last-contributor positions (1-based, in ITS tile list) expressed in our lists.
        Asserts that every last contributor survived the cut."""
        n_ref = np.asarray(n_ref).astype(np.int64).reshape(H, W)
        tx = (W + 15) // 16
        yy, xx = np.mgrid[0:H, 0:W]
        first = self.ref_first[(yy // 16) * tx + (xx // 16)]
        has = n_ref > 0
        assert self.kept[(first + n_ref - 1)[has]].all(), "a contributing instance was cut"
        out = self.kept_prefix[first + n_ref] - self.kept_prefix[first]
        return np.where(has, out, 0)
