"""TEST INFRASTRUCTURE ONLY — ctypes front end of oracle/_ref/libwast3d_ref.so, i.e. the
UNMODIFIED reference CUDA rasteriser / simple-knn (compiled by oracle/build_ref.sh from
/root/reference) behind oracle/ref_wrap.cu.  Needs a GPU.  Used to pin the CPU oracle and as the
"reference recompiled for sm_100" arm of the parity tests and of profiles/."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PATH = Path(__file__).resolve().parent / "_ref" / "libwast3d_ref.so"
_lib = None


def available() -> bool:
    return _PATH.exists()


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_PATH))
        _lib.ref_ctx_create.restype = C.c_void_p
        _lib.ref_ctx_destroy.argtypes = [C.c_void_p]
        _lib.ref_ctx_set_async.argtypes = [C.c_void_p, C.c_int]
    return _lib


def _p(t):
    if t is None or t.numel() == 0:
        return None
    assert t.is_cuda and t.is_contiguous(), "ref wrapper needs contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


class RefRasterizer:
    """One forward (+ optional backward) of the reference rasteriser on raw tensors."""

    def __init__(self, asynchronous=False):
        """asynchronous=True: the wrapper does not cudaDeviceSynchronize after each call (event timing)."""
        self.ctx = C.c_void_p(lib().ref_ctx_create())
        self.asynchronous = bool(asynchronous)
        lib().ref_ctx_set_async(self.ctx, int(self.asynchronous))

    def __del__(self):
        try:
            lib().ref_ctx_destroy(self.ctx)
        except Exception:
            pass

    def forward(self, *, bg, means3D, opacities, view, proj, campos, W, H, tan_fovx, tan_fovy,
                shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                sampling_offsets=None, D=0, scale_modifier=1.0, prefiltered=False, offsets_prepadded=False):
        if not self.asynchronous:
            torch.cuda.synchronize()
        dev = means3D.device
        P = means3D.shape[0]
        M = shs.shape[1] if shs is not None else 0
        if offsets_prepadded:  # caller padded once (ref_step.pad_offsets): nothing extra inside a timed step
            self.offsets = sampling_offsets
        else:
            if sampling_offsets is None:  # the reference dereferences it unconditionally (forward.cu:287)
                sampling_offsets = torch.zeros(H, W, 2, device=dev)
            # the reference reads sampling_offsets / dL_ddepth of out-of-image threads:
            # pad so that those reads stay inside an allocation
            pad = torch.zeros(((H + 16) * (W + 16), 2), device=dev)
            pad[: H * W] = sampling_offsets.reshape(-1, 2)
            self.offsets = pad
        self.args = dict(bg=bg, means3D=means3D, view=view, proj=proj, campos=campos, W=W, H=H,
                         tan_fovx=tan_fovx, tan_fovy=tan_fovy, shs=shs, colors_precomp=colors_precomp,
                         scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, D=D, M=M,
                         scale_modifier=scale_modifier)
        color = torch.zeros(3, H, W, device=dev)
        depth = torch.zeros(H, W, device=dev)
        radii = torch.zeros(P, dtype=torch.int32, device=dev)
        R = lib().ref_raster_forward(
            self.ctx, P, int(D), int(M), _p(bg), int(W), int(H), _p(means3D), _p(shs),
            _p(colors_precomp), _p(opacities), _p(scales), C.c_float(scale_modifier), _p(rotations),
            _p(cov3D_precomp), _p(view), _p(proj), _p(campos), C.c_float(tan_fovx),
            C.c_float(tan_fovy), int(prefiltered), _p(color), _p(depth), _p(self.offsets), _p(radii))
        if R < 0:
            raise RuntimeError("reference forward failed")
        self.R, self.radii, self.P = R, radii, P
        return {"R": R, "color": color, "depth": depth, "radii": radii}

    def state(self):
        a, dev, P = self.args, self.radii.device, self.P
        N = a["W"] * a["H"]
        o = {"depths": torch.zeros(P, device=dev), "means2D": torch.zeros(P, 2, device=dev),
             "cov3D": torch.zeros(P, 6, device=dev), "conic_opacity": torch.zeros(P, 4, device=dev),
             "rgb": torch.zeros(P, 3, device=dev),
             "tiles_touched": torch.zeros(P, dtype=torch.int32, device=dev),
             "clamped": torch.zeros(P, 3, dtype=torch.uint8, device=dev),
             "final_T": torch.zeros(a["H"], a["W"], device=dev),
             "n_contrib": torch.zeros(a["H"], a["W"], dtype=torch.int32, device=dev),
             "point_list": torch.zeros(max(self.R, 1), dtype=torch.int32, device=dev),
             "ranges": torch.zeros(((a["W"] + 15) // 16) * ((a["H"] + 15) // 16), 2, dtype=torch.int32, device=dev)}
        rc = lib().ref_raster_export_state(
            self.ctx, _p(o["depths"]), _p(o["means2D"]), _p(o["cov3D"]), _p(o["conic_opacity"]),
            _p(o["rgb"]), _p(o["tiles_touched"]), _p(o["clamped"]), _p(o["final_T"]),
            _p(o["n_contrib"]), _p(o["point_list"]), _p(o["ranges"]))
        assert rc == 0
        o["point_list"] = o["point_list"][: self.R]
        # untouched entries of culled Gaussians are uninitialised in the reference: mask them
        vis = self.radii > 0
        for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped"):
            o[k] = torch.where(vis.view(-1, *[1] * (o[k].dim() - 1)), o[k], torch.zeros_like(o[k]))
        return o

    def backward(self, dL_dpix, dL_ddepth):
        a, dev, P = self.args, self.radii.device, self.P
        M = a["M"]
        z = lambda *s: torch.zeros(*s, device=dev)
        g = {"dL_dmean2D": z(P, 3), "dL_dconic": z(P, 2, 2), "dL_dopacity": z(P, 1),
             "dL_dcolor": z(P, 3), "dL_dmean3D": z(P, 3), "dL_dcov3D": z(P, 6), "dL_dsh": z(P, M, 3),
             "dL_dscale": z(P, 3), "dL_drot": z(P, 4), "dL_dviewdepth": z(P, 1)}
        H, W = a["H"], a["W"]
        dpad = torch.zeros((H + 16) * (W + 16), device=dev)
        dpad[: H * W] = dL_ddepth.reshape(-1)
        rc = lib().ref_raster_backward(
            self.ctx, P, int(a["D"]), int(M), int(self.R), _p(a["bg"]), int(W), int(H),
            _p(a["means3D"]), _p(a["shs"]), _p(a["colors_precomp"]), _p(a["scales"]),
            C.c_float(a["scale_modifier"]), _p(a["rotations"]), _p(a["cov3D_precomp"]), _p(a["view"]),
            _p(a["proj"]), _p(a["campos"]), C.c_float(a["tan_fovx"]), C.c_float(a["tan_fovy"]),
            _p(self.radii), _p(dL_dpix.contiguous()), _p(dpad), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]),
            _p(g["dL_dopacity"]), _p(g["dL_dcolor"]), _p(g["dL_dmean3D"]), _p(g["dL_dcov3D"]),
            _p(g["dL_dsh"]), _p(g["dL_dscale"]), _p(g["dL_drot"]), _p(g["dL_dviewdepth"]),
            _p(self.offsets))
        if rc != 0:
            raise RuntimeError("reference backward failed")
        return g


def knn_dist2(points: torch.Tensor) -> torch.Tensor:
    torch.cuda.synchronize()
    pts = points.contiguous().float()
    out = torch.zeros(pts.shape[0], device=pts.device)
    rc = lib().ref_knn_dist2(int(pts.shape[0]), _p(pts), _p(out))
    if rc != 0:
        raise RuntimeError("reference knn failed")
    return out


def mark_visible(means3D, view, proj):
    out = torch.zeros(means3D.shape[0], dtype=torch.bool, device=means3D.device)
    lib().ref_mark_visible(int(means3D.shape[0]), _p(means3D.contiguous()), _p(view.contiguous()),
                           _p(proj.contiguous()), _p(out))
    return out
