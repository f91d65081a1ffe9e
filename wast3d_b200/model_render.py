"""Fused "model-space" rasterisation: the autograd op takes the six LEAF
parameters of GaussianModel (scene/gaussian_model.py:149-167) instead of their activated
versions, so the four activation kernels of the reference's render()
(gaussian_renderer/__init__.py:61-90: sigmoid / exp / normalize / cat through the getters of
scene/gaussian_model.py:95-120), the 576 MB `get_features` concatenation and their autograd
backward never run as separate torch kernels: K1 applies the activations while it reads the
parameters, K8+K9 writes the gradients of the raw parameters directly
(`wast3d_raster_params::raw_params`, `wast3d_raster_backward_raw`).

Results equal the unfused path up to the rounding of the activations (tests/test_raster_gpu.py::
test_model_render_matches_unfused).  There is no fallback: without the library it raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .diff_gaussian_rasterization import GaussianRasterizationSettings
from .diff_gaussian_rasterization import _C as dgr_C


def _prm(keep, rs: GaussianRasterizationSettings, xyz, f_dc, f_rest, opacity, scaling, rotation, offsets,
         colour_wait_event=None):
    P = int(xyz.size(0))
    M = 1 + (int(f_rest.size(1)) if f_rest.numel() else 0)
    return dgr_C._params(
        keep, P=P, D=rs.sh_degree, M=M, W=rs.image_width, H=rs.image_height, tan_fovx=rs.tanfovx,
        tan_fovy=rs.tanfovy, scale_modifier=rs.scale_modifier, prefiltered=rs.prefiltered, debug=rs.debug,
        bg=rs.bg, means3D=xyz, sh=f_dc, colors=None, opacity=opacity, scales=scaling, rotations=rotation,
        cov3D_precomp=None, viewmatrix=rs.viewmatrix, projmatrix=rs.projmatrix, campos=rs.campos,
        sampling_offsets=offsets, raw_params=True, sh_rest=f_rest, colour_wait_event=colour_wait_event)


# ---- graph-safe forward (wast3d_raster_forward_async): no host synchronisation in the step -------------------
# The reference reads num_rendered back on the host in every forward to size the binning buffer
# (rasterizer_impl.cu:283-289), which stalls the stream once per step and rules out CUDA-graph capture.  In async
# mode the buffer is sized from the largest instance count SEEN so far for this (device, image size) plus headroom;
# the count of every call lands in pinned host memory through an asynchronous copy and is looked at one call later:
# no wait in steady state.  The first call for a size uses the synchronous protocol (there is no estimate yet).  An
# overflow (the view needed more instances than the capacity: the image of that call was incomplete) raises
# RuntimeError at the next call / at async_forward_check(), after the capacity has been raised.
_ASYNC = {"on": bool(int(__import__("os").environ.get("WAST3D_ASYNC_FORWARD", "0")))}
# CUDA-graph capture (graphed.GraphedStep): a fixed instance capacity and a pinned word block that receives the status of
# every replay; no events, no host reads while the stream is capturing
_CAPTURE: dict = {"capacity": None, "status_host": None}
_LAST: dict = {"num_rendered": 0}


def last_num_rendered() -> int:
    """Instance count of the last synchronous rasterize_model() forward (the capacity in graph-safe calls)."""
    return int(_LAST["num_rendered"])
_ASYNC_STATE: dict = {}   # (device index, W, H) -> {"seen": max R, "pending": [(event, pinned, capacity)]}


def set_async_forward(on: bool) -> bool:
    """Enable / disable the graph-safe forward for rasterize_model(); returns the previous setting."""
    prev, _ASYNC["on"] = _ASYNC["on"], bool(on)
    return prev


def _async_poll(st, wait=False):
    keep = []
    err = None
    for ev, pinned, cap in st["pending"]:
        if wait:
            ev.synchronize()
        if not ev.query():
            keep.append((ev, pinned, cap))
            continue
        r, prefilt, timeout, overflow = (int(v) for v in pinned.tolist())
        st["seen"] = max(st["seen"], r)
        if overflow:
            err = RuntimeError(f"wast3d_b200: a graph-safe forward needed {r} tile instances but its binning buffer held "
                               f"{cap}; that step's image and gradients are incomplete (capacity raised for the next call)")
        elif prefilt & 2:
            err = RuntimeError("wast3d_b200: preprojected forward: the sampling offsets exceed the bounds the projection "
                               "assumed (BackwardFusedAdam.prefetch_view offset_bounds); that step's image is incomplete")
        elif prefilt & 1:
            err = RuntimeError("wast3d_b200: rasterize_model: invalid argument (prefiltered set but a culled point was seen)")
        elif timeout:
            err = RuntimeError("wast3d_b200: look-back time-out in the binning stage")
    st["pending"] = keep
    if err is not None:
        raise err


def async_forward_check(device=None):
    """Wait for the status words of all graph-safe forwards issued so far and raise if one of them overflowed its
    instance capacity (call at a point where synchronising is acceptable: end of training, checkpoints, tests)."""
    for key, st in _ASYNC_STATE.items():
        if device is None or key[0] == torch.device(device).index:
            _async_poll(st, wait=True)


def _projection_key(rs, leaves, P):
    """What a pre-projected geometry buffer is valid for: the camera tensors (identity + version), the image
    geometry, and the six parameter tensors exactly as the optimizer-in-backward left them."""
    cam = tuple((t.data_ptr(), t._version) for t in (rs.viewmatrix, rs.projmatrix, rs.campos))
    return (cam, int(rs.image_width), int(rs.image_height), float(rs.tanfovx), float(rs.tanfovy),
            float(rs.scale_modifier), int(rs.sh_degree), int(P), _lib.set_tile_cut(-1),
            tuple((t.data_ptr(), t._version) for t in leaves))


class _RasterizeModel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, means2D, f_dc, f_rest, opacity, scaling, rotation, raster_settings,
                sampling_offsets, grad_sink=None, colour_wait_event=None):
        rs = raster_settings
        # peer.GradSink: the leaf gradients go straight into the optimizer's peer-visible arena
        ctx.grad_sink = grad_sink
        ctx.sink_params = (xyz, f_dc, f_rest, opacity, scaling, rotation) if grad_sink is not None else None
        _lib.require_device(xyz)
        if xyz.dim() != 2 or xyz.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        lib = _lib.load()
        dev = xyz.device
        P, H, W = int(xyz.size(0)), int(rs.image_height), int(rs.image_width)
        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((H, W), dtype=torch.float32, device=dev)
        # a geometry buffer the previous step's backward already projected into (optim.BackwardFusedAdam.prefetch_view)?
        fused_opt = getattr(grad_sink, "fused_adam", None) if grad_sink is not None else None
        proj = getattr(fused_opt, "projection", None) if fused_opt is not None else None
        pre = False
        if proj is not None:
            fused_opt.projection = None   # one shot
            pre = colour_wait_event is None and P > 0 and \
                proj["key"] == _projection_key(rs, (xyz, f_dc, f_rest, opacity, scaling, rotation), P)
        if pre:
            radii = proj["radii"]
            geom = _lib.FixedBuffer(proj["geom"])
        else:
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            geom = _lib.GrowBuffer(dev, "geom")
        binning, img = _lib.GrowBuffer(dev, "binning"), _lib.GrowBuffer(dev, "img")
        keep: list = []
        # colour_wait_event (torch.cuda.Event): f_dc / f_rest are only read behind it (ABI v5)
        ev = int(colour_wait_event.cuda_event) if colour_wait_event is not None else None
        prm = _prm(keep, rs, xyz, f_dc, f_rest, opacity, scaling, rotation, sampling_offsets, ev)
        prm.preprojected = int(pre)
        rendered = C.c_int(0)
        ast = None
        capturing = _CAPTURE["capacity"] is not None and P > 0
        if capturing:
            with torch.cuda.device(dev):
                status = torch.empty(4, dtype=torch.int32, device=dev)
                st = lib.wast3d_raster_forward_async(
                    C.byref(prm), geom.cb, None, binning.cb, None, img.cb, None, color.data_ptr(), depth.data_ptr(),
                    radii.data_ptr(), int(_CAPTURE["capacity"]), status.data_ptr(), _lib.stream_ptr())
                rendered.value = int(_CAPTURE["capacity"])
                if st == 0:
                    _CAPTURE["status_host"].copy_(status, non_blocking=True)
        elif _ASYNC["on"] and P:
            ast = _ASYNC_STATE.setdefault((dev.index, W, H), {"seen": 0, "pending": []})
            _async_poll(ast)   # raises if an earlier call of this size overflowed
        with torch.cuda.device(dev):
            if capturing:
                pass
            elif ast is not None and ast["seen"] > 0:
                # capacity: 25% + 64k above the largest count seen, bucketed so that it (and the buffer size) is stable
                cap = min(_lib.bucket_bytes(ast["seen"] + ast["seen"] // 4 + 65536), 0x7FFFFFFF)
                status = torch.empty(4, dtype=torch.int32, device=dev)
                st = lib.wast3d_raster_forward_async(
                    C.byref(prm), geom.cb, None, binning.cb, None, img.cb, None, color.data_ptr(), depth.data_ptr(),
                    radii.data_ptr(), int(cap), status.data_ptr(), _lib.stream_ptr())
                rendered.value = int(cap)
                if st == 0:
                    pinned = torch.empty(4, dtype=torch.int32, pin_memory=True)
                    pinned.copy_(status, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    ast["pending"].append((ev, pinned, int(cap)))
            else:
                st = lib.wast3d_raster_forward(
                    C.byref(prm), geom.cb, None, binning.cb, None, img.cb, None, color.data_ptr(),
                    depth.data_ptr(), radii.data_ptr() if P else None, C.byref(rendered), _lib.stream_ptr())
                if ast is not None and st == 0:
                    ast["seen"] = max(ast["seen"], int(rendered.value), 1)
        geom_t, binning_t, img_t = geom.take(), binning.take(), img.take()
        for b in (geom, binning, img):
            if b.error is not None:
                raise b.error
        _lib.check(st, "rasterize_model")
        ctx.raster_settings = rs
        ctx.num_rendered = rendered.value
        _LAST["num_rendered"] = int(rendered.value)
        ctx.save_for_backward(xyz, f_dc, f_rest, opacity, scaling, rotation, radii, geom_t, binning_t, img_t,
                              sampling_offsets if sampling_offsets is not None else torch.empty(0))
        ctx.mark_non_differentiable(radii)
        return color, depth, radii

    @staticmethod
    def backward(ctx, g_color, g_depth, _g_radii):
        rs = ctx.raster_settings
        (xyz, f_dc, f_rest, opacity, scaling, rotation, radii, geom, binning, img, offsets) = ctx.saved_tensors
        lib = _lib.load()
        dev = xyz.device
        P, H, W = int(xyz.size(0)), int(rs.image_height), int(rs.image_width)
        if g_color is None:
            g_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        if g_depth is None:
            g_depth = torch.zeros((H, W), dtype=torch.float32, device=dev)
        opt = dict(dtype=torch.float32, device=dev)
        sink, sunk = ctx.grad_sink, None
        fused_opt = getattr(sink, "fused_adam", None) if sink is not None else None
        if fused_opt is not None:
            # optimizer-in-backward (optim.BackwardFusedAdam): K8+K9 applies the Adam update in place
            if not sink.fresh:
                raise RuntimeError("BackwardFusedAdam: a second backward through render() before optimizer.step() / "
                                   "zero_grad() is not supported (the update was already applied)")
            d_m2d = torch.empty((P, 3), **opt) if ctx.needs_input_grad[1] else None
            if P:
                keep: list = []
                prm = _prm(keep, rs, xyz, f_dc, f_rest, opacity, scaling, rotation,
                           offsets if offsets.numel() else None)
                groups = fused_opt.adam_groups(list(ctx.sink_params))
                grads_out = None
                if fused_opt.capture_grads:  # test hook: also write the leaf gradients
                    fused_opt.last_grads = [torch.empty_like(t) for t in ctx.sink_params]
                    grads_out = (C.c_void_p * 6)(*[t.data_ptr() if t.numel() else None for t in fused_opt.last_grads])
                # prefetch_view(): also project every Gaussian for the next camera from the updated values
                nv, nview, nrs = fused_opt._next_view, None, None
                fused_opt._next_view = None
                fused_opt.projection = None
                if nv is not None:
                    import math
                    cam = nv["cam"]
                    nrs = GaussianRasterizationSettings(
                        image_height=int(cam.image_height), image_width=int(cam.image_width),
                        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=rs.bg,
                        scale_modifier=nv["scale"], viewmatrix=cam.world_view_transform,
                        projmatrix=cam.full_proj_transform, sh_degree=rs.sh_degree if nv["D"] is None else int(nv["D"]),
                        campos=cam.camera_center, prefiltered=False, debug=False)
                    n_geom = _lib.geom_buffer(dev, P)
                    n_radii = torch.empty((P,), dtype=torch.int32, device=dev)
                    b = nv["bounds"]
                    nview = _lib.NextView(
                        width=nrs.image_width, height=nrs.image_height, tan_fovx=nrs.tanfovx, tan_fovy=nrs.tanfovy,
                        scale_modifier=nrs.scale_modifier, D=nrs.sh_degree, viewmatrix=_lib.fptr(nrs.viewmatrix, keep),
                        projmatrix=_lib.fptr(nrs.projmatrix, keep), campos=_lib.fptr(nrs.campos, keep),
                        geom_buffer=n_geom.data_ptr(), radii=n_radii.data_ptr(), offset_min_x=b[0], offset_max_x=b[1],
                        offset_min_y=b[2], offset_max_y=b[3])
                with torch.cuda.device(dev):
                    st = lib.wast3d_raster_backward_raw_adam_next(
                        C.byref(prm), int(ctx.num_rendered), _lib.fptr(radii, keep, torch.int32),
                        _lib.fptr(geom, keep, torch.uint8), _lib.fptr(binning, keep, torch.uint8),
                        _lib.fptr(img, keep, torch.uint8), _lib.fptr(g_color, keep), _lib.fptr(g_depth, keep),
                        groups, grads_out, d_m2d.data_ptr() if d_m2d is not None else None,
                        C.byref(nview) if nview is not None else None, _lib.stream_ptr())
                _lib.check(st, "rasterize_model_backward_adam")
            fused_opt._applied = True
            sink.fresh = False
            # the parameters changed in place outside autograd's view: any other graph branch that saved them
            # must fail loudly instead of silently using post-update values
            for p_ in ctx.sink_params:
                if p_.numel():
                    torch.autograd.graph.increment_version(p_)
            if P and nview is not None:
                fused_opt.projection = {"geom": n_geom, "radii": n_radii,
                                        "key": _projection_key(nrs, ctx.sink_params, P)}
            return None, d_m2d, None, None, None, None, None, None, None, None, None
        # arena path only when nothing else has produced a gradient for these leaves in this step: re-pointing
        # .grad at the arena view would drop an earlier regulariser's gradient (it is accumulated by autograd
        # instead, and the optimizer copies .grad into the arena)
        if sink is not None and sink.fresh and all(p_.grad is None for p_ in ctx.sink_params):
            sunk = [sink.view_for(p) for p in ctx.sink_params]
            if any(v is None for v in sunk):
                sunk = None
        if sunk is not None:
            # first backward since zero_grad(): K8+K9 writes every element of the arena's gradient
            # views; autograd gets None for these leaves and .grad is pointed at the views
            d_xyz, d_dc, d_rest, d_op, d_sc, d_rot = sunk
        else:
            d_xyz = torch.empty_like(xyz)
            d_dc = torch.empty_like(f_dc)
            d_rest = torch.empty_like(f_rest)
            d_op = torch.empty_like(opacity)
            d_sc = torch.empty_like(scaling)
            d_rot = torch.empty_like(rotation)
        d_m2d = torch.empty((P, 3), **opt) if ctx.needs_input_grad[1] else None
        if sunk is not None and getattr(sink, "record_out", None) is not None:
            d_dc = d_rest = None   # colour-record exchange: the SH gradients are rebuilt from the records, never written
        if P:
            keep: list = []
            prm = _prm(keep, rs, xyz, f_dc, f_rest, opacity, scaling, rotation,
                       offsets if offsets.numel() else None)
            p = lambda t: None if t is None or t.numel() == 0 else t.data_ptr()
            with torch.cuda.device(dev):
                st = lib.wast3d_raster_backward_raw(
                    C.byref(prm), int(ctx.num_rendered), _lib.fptr(radii, keep, torch.int32),
                    _lib.fptr(geom, keep, torch.uint8), _lib.fptr(binning, keep, torch.uint8),
                    _lib.fptr(img, keep, torch.uint8), _lib.fptr(g_color, keep), _lib.fptr(g_depth, keep),
                    p(d_xyz), p(d_dc), p(d_rest), p(d_op), p(d_sc), p(d_rot), p(d_m2d), _lib.stream_ptr())
            _lib.check(st, "rasterize_model_backward")
            if sink is not None and getattr(sink, "record_out", None) is not None:
                # staged (peer_records.PeerRecordAdam): this view's 16-byte colour records for the feature exchange
                from .peer_records import write_colour_records
                with torch.cuda.device(dev):
                    write_colour_records(sink, P, radii, geom, _lib.stream_ptr())
        elif sunk is not None:
            for v in sunk:
                v.zero_()
        if sunk is not None:
            if getattr(sink, "record_out", None) is not None:
                sunk = [None if p_ is f_dc or p_ is f_rest else v for p_, v in zip(ctx.sink_params, sunk)]
            for p_, v in zip(ctx.sink_params, sunk):
                p_.grad = v
            sink.fresh = False
            return None, d_m2d, None, None, None, None, None, None, None, None, None
        return d_xyz, d_m2d, d_dc, d_rest, d_op, d_sc, d_rot, None, None, None, None


def rasterize_model(xyz, means2D, features_dc, features_rest, opacity_logits, log_scales, rotations,
                    raster_settings: GaussianRasterizationSettings, sampling_offsets=None, grad_sink=None,
                    colour_wait_event=None):
    """(color[3,H,W], depth[H,W], radii[P]) from the RAW GaussianModel parameters.
    `grad_sink` (peer.GradSink, optional): destination of the leaf gradients (see peer.py).
    `colour_wait_event` (torch.cuda.Event, optional): the features are read only by a colour kernel that
    waits for this event, so their all-gather may overlap projection, sorting and binning."""
    for name, t in (("features_dc", features_dc), ("features_rest", features_rest),
                    ("opacity", opacity_logits), ("scaling", log_scales), ("rotation", rotations)):
        if not t.is_contiguous() or t.dtype != torch.float32:
            raise RuntimeError(f"rasterize_model: {name} must be contiguous float32")
    return _RasterizeModel.apply(xyz, means2D, features_dc, features_rest, opacity_logits, log_scales,
                                 rotations, raster_settings, sampling_offsets, grad_sink, colour_wait_event)


def model_supports_fusion(pc) -> bool:
    """True when `pc` exposes the six raw parameter tensors and uses the reference's activations
    (scene/gaussian_model.py:26-41), so folding them into the kernels does not change results."""
    names = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")
    if not all(isinstance(getattr(pc, n, None), torch.Tensor) for n in names):
        return False
    if getattr(pc, "scaling_activation", torch.exp) is not torch.exp:
        return False
    if getattr(pc, "opacity_activation", torch.sigmoid) is not torch.sigmoid:
        return False
    if getattr(pc, "rotation_activation", torch.nn.functional.normalize) is not torch.nn.functional.normalize:
        return False
    return all(getattr(pc, n).is_contiguous() and getattr(pc, n).dtype == torch.float32 for n in names)
