"""Host shim with the three entry points of the reference's pybind module
(submodules/diff-gaussian-rasterization/ext.cpp:15-19), argument for argument:

    rasterize_gaussians           <- RasterizeGaussiansCUDA          rasterize_points.cu:35-119
    rasterize_gaussians_backward  <- RasterizeGaussiansBackwardCUDA  rasterize_points.cu:121-206
    mark_visible                  <- markVisible                     rasterize_points.cu:208-227

It only allocates the torch outputs and forwards raw pointers to the C ABI
(include/wast3d_b200.h); all arithmetic happens in libwast3d_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

NUM_CHANNELS = 3  # cuda_rasterizer/config.h:15


def _params(keep, *, P, D, M, W, H, tan_fovx, tan_fovy, scale_modifier, prefiltered, debug, bg,
            means3D, sh, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix, projmatrix,
            campos, sampling_offsets, raw_params=False, sh_rest=None, colour_wait_event=None, preprojected=False):
    f = _lib.fptr
    return _lib.RasterParams(
        P=P, D=int(D), M=M, width=int(W), height=int(H), tan_fovx=float(tan_fovx),
        tan_fovy=float(tan_fovy), scale_modifier=float(scale_modifier),
        prefiltered=int(bool(prefiltered)), debug=int(bool(debug)),
        background=f(bg, keep), means3D=f(means3D, keep), shs=f(sh, keep),
        colors_precomp=f(colors, keep), opacities=f(opacity, keep), scales=f(scales, keep),
        rotations=f(rotations, keep), cov3D_precomp=f(cov3D_precomp, keep),
        viewmatrix=f(viewmatrix, keep), projmatrix=f(projmatrix, keep), campos=f(campos, keep),
        sampling_offsets=f(sampling_offsets, keep), raw_params=int(bool(raw_params)),
        shs_rest=f(sh_rest, keep), colour_wait_event=colour_wait_event, preprojected=int(bool(preprojected)))


def _sh_coeffs(sh) -> int:
    # rasterize_points.cu:85-89: M = sh.size(1) unless sh is the empty placeholder
    return int(sh.size(1)) if sh.numel() != 0 and sh.dim() >= 2 else 0


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                        image_width, sh, degree, campos, prefiltered, debug, sampling_offsets,
                        *, _colour_wait_event=None):
    """`_colour_wait_event` (extension, torch.cuda.Event): `sh` is read only behind this event (ABI v5)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    _lib.require_device(means3D)
    lib = _lib.load()
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    dev = means3D.device
    out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    out_depth = torch.empty((H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    geom, binning, img = _lib.GrowBuffer(dev, "geom"), _lib.GrowBuffer(dev, "binning"), _lib.GrowBuffer(dev, "img")
    keep: list = []
    prm = _params(keep, P=P, D=degree, M=_sh_coeffs(sh), W=W, H=H, tan_fovx=tan_fovx,
                  tan_fovy=tan_fovy, scale_modifier=scale_modifier, prefiltered=prefiltered,
                  debug=debug, bg=background, means3D=means3D, sh=sh, colors=colors,
                  opacity=opacity, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
                  viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos,
                  sampling_offsets=sampling_offsets,
                  colour_wait_event=int(_colour_wait_event.cuda_event) if _colour_wait_event is not None else None)
    rendered = C.c_int(0)
    with torch.cuda.device(dev):
        st = lib.wast3d_raster_forward(
            C.byref(prm), geom.cb, None, binning.cb, None, img.cb, None, out_color.data_ptr(),
            out_depth.data_ptr(), radii.data_ptr() if P else None, C.byref(rendered),
            _lib.stream_ptr())
    for b in (geom, binning, img):
        if b.error is not None:
            err = b.error
            geom.take(), binning.take(), img.take()
            raise err
    _lib.check(st, "rasterize_gaussians")
    return rendered.value, out_color, out_depth, radii, geom.take(), binning.take(), img.take()


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations,
                                 scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                                 tan_fovy, dL_dout_color, dL_dout_depth, sh, degree, campos,
                                 geomBuffer, R, binningBuffer, imageBuffer, debug, sampling_offsets,
                                 _return_scratch=False):
    _lib.require_device(means3D)
    lib = _lib.load()
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = _sh_coeffs(sh)
    opt = dict(dtype=torch.float32, device=means3D.device)
    # every element is written by the kernels -> torch.empty, not the reference's ten torch::zeros
    # (rasterize_points.cu:157-166); P == 0 gives empty tensors either way.
    dL_dmeans3D = torch.empty((P, 3), **opt)
    dL_dmeans2D = torch.empty((P, 3), **opt)
    dL_dcolors = torch.empty((P, NUM_CHANNELS), **opt)
    dL_dopacity = torch.empty((P, 1), **opt)
    dL_dcov3D = torch.empty((P, 6), **opt)
    dL_dsh = torch.empty((P, M, 3), **opt)
    dL_dscales = torch.empty((P, 3), **opt)
    dL_drotations = torch.empty((P, 4), **opt)
    dL_dconic = torch.empty((P, 2, 2), **opt) if _return_scratch else None
    dL_dcamViewDepth = torch.empty((P, 1), **opt) if _return_scratch else None
    if P != 0:
        keep: list = []
        # opacities are not an argument of the reference's backward (they live in its geometry
        # buffer); ours are in the render records as well, so that pointer stays NULL here.
        prm = _params(keep, P=P, D=degree, M=M, W=W, H=H, tan_fovx=tan_fovx, tan_fovy=tan_fovy,
                      scale_modifier=scale_modifier, prefiltered=False, debug=debug, bg=background,
                      means3D=means3D, sh=sh, colors=colors, opacity=None, scales=scales,
                      rotations=rotations, cov3D_precomp=cov3D_precomp, viewmatrix=viewmatrix,
                      projmatrix=projmatrix, campos=campos, sampling_offsets=sampling_offsets)
        p = lambda t: None if t is None or t.numel() == 0 else t.data_ptr()
        with torch.cuda.device(means3D.device):
            st = lib.wast3d_raster_backward(
                C.byref(prm), int(R), _lib.fptr(radii, keep, torch.int32),
                _lib.fptr(geomBuffer, keep, torch.uint8), _lib.fptr(binningBuffer, keep, torch.uint8),
                _lib.fptr(imageBuffer, keep, torch.uint8), _lib.fptr(dL_dout_color, keep),
                _lib.fptr(dL_dout_depth, keep), p(dL_dmeans2D), p(dL_dconic), p(dL_dopacity),
                p(dL_dcolors), p(dL_dmeans3D), p(dL_dcov3D), p(dL_dsh), p(dL_dscales),
                p(dL_drotations), p(dL_dcamViewDepth), _lib.stream_ptr())
        _lib.check(st, "rasterize_gaussians_backward")
    out = (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
           dL_drotations)
    if _return_scratch:
        return out + (dL_dconic, dL_dcamViewDepth)
    return out


def mark_visible(means3D, viewmatrix, projmatrix):
    _lib.require_device(means3D)
    lib = _lib.load()
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        keep: list = []
        with torch.cuda.device(means3D.device):
            st = lib.wast3d_mark_visible(P, _lib.fptr(means3D, keep), _lib.fptr(viewmatrix, keep),
                                         _lib.fptr(projmatrix, keep), present.data_ptr(),
                                         _lib.stream_ptr())
        _lib.check(st, "mark_visible")
    return present


def export_state(prm_kwargs, num_rendered, geomBuffer, binningBuffer, imageBuffer):
    """Test hook: unpack the opaque buffers into the reference's per-Gaussian arrays."""
    lib = _lib.load()
    keep: list = []
    prm = _params(keep, **prm_kwargs)
    P, W, H = prm.P, prm.width, prm.height
    dev = geomBuffer.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out = {
        "depths": torch.zeros(P, device=dev), "means2D": torch.zeros(P, 2, device=dev),
        "conic_opacity": torch.zeros(P, 4, device=dev), "rgb": torch.zeros(P, 3, device=dev),
        "tiles_touched": torch.zeros(P, dtype=torch.int32, device=dev),
        "clamped": torch.zeros(P, 3, dtype=torch.uint8, device=dev),
        "point_list": torch.zeros(max(num_rendered, 0), dtype=torch.int32, device=dev),
        "ranges": torch.zeros(T, 2, dtype=torch.int32, device=dev),
    }
    p = lambda t: t.data_ptr() if t.numel() else None
    st = lib.wast3d_raster_export_state(
        C.byref(prm), int(num_rendered), p(geomBuffer), p(binningBuffer), p(imageBuffer),
        p(out["depths"]), p(out["means2D"]), p(out["conic_opacity"]), p(out["rgb"]),
        p(out["tiles_touched"]), p(out["clamped"]), p(out["point_list"]), p(out["ranges"]),
        _lib.stream_ptr())
    _lib.check(st, "export_state")
    N = W * H
    out["final_T"] = imageBuffer[: 4 * N].view(torch.float32).view(H, W)
    off = (4 * N + 127) // 128 * 128
    out["n_contrib"] = imageBuffer[off: off + 4 * N].view(torch.int32).view(H, W)
    return out
