"""One optimisation step of the style-transfer loop as a replayable CUDA graph.

The reference's step (train_st.py:271 render, :325 loss.backward(), :342-343 optimizer.step() / zero_grad()) stalls once per
forward on the host read of num_rendered (rasterizer_impl.cu:283) and bakes the optimizer's step count into its Adam
launches.  Here nothing in the step depends on a host value that changes from step to step: the forward sizes its
binning buffer from a capacity (wast3d_raster_forward_async), the Adam step count and bias corrections live on the
device (wast3d_adam_schedule_step), camera and targets are static device buffers refreshed before each replay.  The
whole launch sequence (~35 kernels) is captured once with torch.cuda.graph and replayed with one launch.

    gs = GraphedStep(model, pipe, background, camera, loss_fn)      # model.training_setup(in_backward=True)
    for cam, tgt, dtgt in views:
        gs.set_view(cam); gs.set_targets(tgt, dtgt)
        loss = gs.step()          # device scalar (static buffer), valid until the next step()
    gs.check()                    # raises if a replay needed more tile instances than the captured capacity

All cameras must share the image size and field of view of the capture (they are launch constants); anything else
takes the eager path (render() + loss.backward() + optimizer.step()).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

from . import _lib, model_render
from .gaussian_renderer import render
from .optim import BackwardFusedAdam


class GraphedStep:
    def __init__(self, model, pipe, background, camera, loss_fn, target=None, depth_target=None,
                 sampling_offsets=None, warmup_cameras=None, capacity=None):
        """`camera`: any camera of the set (fixes H, W, FoV).  `loss_fn(out, target, depth_target) -> scalar`.
        `sampling_offsets`: optional static [H, W, 2] tensor (default: drawn inside the graph like the reference does).
        `warmup_cameras`: cameras rendered eagerly before the capture to learn the instance count (default: `camera`);
        `capacity`: instance capacity of the captured forward (default: 1.5 x the largest count seen + 64 k)."""
        opt = getattr(model, "optimizer", None)
        if not isinstance(opt, BackwardFusedAdam):
            raise RuntimeError("GraphedStep needs the optimizer-in-backward (training_setup(..., in_backward=True))")
        if getattr(pipe, "debug", False) or getattr(pipe, "convert_SHs_python", False) or getattr(pipe, "compute_cov3D_python", False):
            raise RuntimeError("GraphedStep: pipe.debug (synchronises after every launch) and the Python SH / covariance "
                               "paths cannot be captured")
        self.model, self.pipe, self.bg, self.loss_fn, self.opt = model, pipe, background, loss_fn, opt
        dev = model.get_xyz.device
        self.H, self.W = int(camera.image_height), int(camera.image_width)
        self.cam = SimpleNamespace(
            FoVx=float(camera.FoVx), FoVy=float(camera.FoVy), image_height=self.H, image_width=self.W,
            world_view_transform=camera.world_view_transform.detach().to(dev).clone(),
            full_proj_transform=camera.full_proj_transform.detach().to(dev).clone(),
            camera_center=camera.camera_center.detach().to(dev).clone())
        self.target = torch.zeros(3, self.H, self.W, device=dev) if target is None else target.detach().clone()
        self.depth_target = torch.zeros(self.H, self.W, device=dev) if depth_target is None else depth_target.detach().clone()
        self.offsets = sampling_offsets
        self.loss = torch.zeros((), device=dev)
        self.status_host = torch.zeros(4, dtype=torch.int32, pin_memory=True)
        self.graph = None
        self.replays = 0
        self._capture(warmup_cameras or [camera], capacity)

    # ---- per-step inputs (static buffers) ------------------------------------------------------------------------
    def set_view(self, camera):
        if int(camera.image_height) != self.H or int(camera.image_width) != self.W or \
                not math.isclose(float(camera.FoVx), self.cam.FoVx) or not math.isclose(float(camera.FoVy), self.cam.FoVy):
            raise RuntimeError("GraphedStep: the camera's image size / field of view differ from the captured ones")
        self.cam.world_view_transform.copy_(camera.world_view_transform, non_blocking=True)
        self.cam.full_proj_transform.copy_(camera.full_proj_transform, non_blocking=True)
        self.cam.camera_center.copy_(camera.camera_center, non_blocking=True)

    def set_targets(self, target=None, depth_target=None):
        if target is not None:
            self.target.copy_(target, non_blocking=True)
        if depth_target is not None:
            self.depth_target.copy_(depth_target, non_blocking=True)

    # ---- capture / replay ------------------------------------------------------------------------------------------
    def _eager_step(self, cam):
        out = render(cam, self.model, self.pipe, self.bg, sampling_offsets=self.offsets)
        loss = self.loss_fn(out, self.target, self.depth_target)
        loss.backward()
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss

    def _capture(self, warm_cams, capacity):
        dev = self.loss.device
        self.opt.device_schedule(True)
        # eager steps first: lazy allocations (scratch, optimizer state) happen outside the graph and the instance
        # count of these views sizes the capacity
        seen = 0
        prev = model_render.set_async_forward(False)
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for cam in warm_cams:
                    self.set_view(cam)
                    self._eager_step(self.cam)
                    seen = max(seen, int(model_render.last_num_rendered()))
            torch.cuda.current_stream(dev).wait_stream(side)
        finally:
            model_render.set_async_forward(prev)
        if capacity is None:
            capacity = min(_lib.bucket_bytes(seen + seen // 2 + 65536), 0x7FFFFFFF)
        self.capacity = int(capacity)
        self.graph = torch.cuda.CUDAGraph()
        model_render._CAPTURE.update(capacity=self.capacity, status_host=self.status_host)
        launches0 = _lib.launch_count()
        try:
            with torch.cuda.graph(self.graph):
                loss = self._eager_step(self.cam)
                self.loss.copy_(loss.detach())
        finally:
            model_render._CAPTURE.update(capacity=None, status_host=None)
        self.launches_per_step = _lib.launch_count() - launches0
        # the capture ran the host book-keeping of one step without executing it
        self.opt.advance_host_steps(-1)

    def step(self):
        self.opt.sync_hyper()
        self.graph.replay()
        self.opt.advance_host_steps(1)
        self.replays += 1
        return self.loss

    def check(self):
        """Synchronise and raise if the last replay's forward reported a problem (capacity overflow: that step's image
        and update were incomplete — re-create the GraphedStep with a larger `capacity`)."""
        torch.cuda.current_stream(self.loss.device).synchronize()
        r, flags, timeout, overflow = (int(v) for v in self.status_host.tolist())
        if overflow:
            raise RuntimeError(f"wast3d_b200: a graphed step needed {r} tile instances but its binning buffer held "
                               f"{self.capacity}; that step's image and update are incomplete")
        if flags & 1:
            raise RuntimeError("wast3d_b200: rasterize_model: invalid argument (prefiltered set but a culled point was seen)")
        if timeout:
            raise RuntimeError("wast3d_b200: look-back time-out in the binning stage")
        return r
