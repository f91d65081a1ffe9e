"""GPU tests of the colour-record optimizer's kernels (csrc/sh_adam.cu, include/wast3d_b200_staged.h; first run on
hardware in round 2, profiles/r02_scaling.md).

Three views of one scene: [per-view backward -> dL/dsh summed over the views -> dense fused Adam on _features_dc /
_features_rest] against [16-byte colour record per view -> wast3d_staged_sh_adam_from_records]."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.util import call_backward, call_forward, raster_case, to_cuda
from wast3d_b200._lib import AdamGroup   # struct wast3d_adam_group (ABI v7 layout)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("degree,P", [(3, 6000), (1, 4999), (0, 777)])
def test_sh_adam_from_records_equals_summed_gradients_then_adam(built, degree, P):
    from wast3d_b200 import _lib
    from wast3d_b200.optim import FusedAdam
    lib = _lib.load()
    lib.wast3d_staged_colour_records.restype = C.c_int
    lib.wast3d_staged_colour_records.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wast3d_staged_sh_adam_from_records.restype = C.c_int
    lib.wast3d_staged_sh_adam_from_records.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                       C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    views = [1, 3, 6]
    recs, campos, grads = [], [], None
    gen = torch.Generator(device="cuda").manual_seed(2)
    for ci in views:
        case = raster_case(P=P, W=128, H=96, seed=4, cam_index=ci, degree=degree, log_scale_mu=-3.0)
        tc = to_cuda(case)
        fwd = call_forward(tc)
        dpix = torch.randn(3, case["H"], case["W"], device="cuda", generator=gen)
        ddep = torch.randn(case["H"], case["W"], device="cuda", generator=gen)
        g = call_backward(tc, fwd, dpix, ddep, scratch=False)
        dsh = g[5]
        grads = dsh.clone() if grads is None else grads + dsh
        rec = torch.empty(P, 4, device="cuda")
        st = lib.wast3d_staged_colour_records(P, fwd[3].data_ptr(), fwd[4].data_ptr(), rec.data_ptr(), _lib.stream_ptr())
        assert st == 0
        recs.append(rec)
        campos.append(case["campos"].astype(np.float32))
    shs = tc["shs"]
    M = shs.shape[1]
    scale = 1.0 / len(views)
    lrs = (2.5e-3, 1.25e-4)
    # expected: dense fused Adam (bit-exact with torch's arithmetic order, tests/test_knn_match_gpu.py) on the averaged sum
    e_dc = torch.nn.Parameter(shs[:, :1].contiguous().clone())
    e_rest = torch.nn.Parameter(shs[:, 1:].contiguous().clone())
    opt = FusedAdam([{"params": [e_dc], "lr": lrs[0]}, {"params": [e_rest], "lr": lrs[1]}], lr=0.0, eps=1e-15)
    # staged path state
    p_dc, p_rest = e_dc.detach().clone(), e_rest.detach().clone()
    m_dc, v_dc, m_rest, v_rest = (torch.zeros_like(p_dc), torch.zeros_like(p_dc), torch.zeros_like(p_rest),
                                  torch.zeros_like(p_rest))
    ptrs = (C.c_void_p * len(views))(*[r.data_ptr() for r in recs])
    cam = np.ascontiguousarray(np.stack(campos), np.float32)
    for step in (1, 2):
        e_dc.grad = (grads[:, :1] * scale).contiguous()
        e_rest.grad = (grads[:, 1:] * scale).contiguous()
        opt.step()
        gd = AdamGroup(p_dc.data_ptr(), m_dc.data_ptr(), v_dc.data_ptr(), lrs[0], 0.9, 0.999, 1e-15, step, 0)
        gr = AdamGroup(p_rest.data_ptr() if M > 1 else None, m_rest.data_ptr() if M > 1 else None,
                       v_rest.data_ptr() if M > 1 else None, lrs[1], 0.9, 0.999, 1e-15, step, 0)
        st = lib.wast3d_staged_sh_adam_from_records(
            P, degree, M, len(views), ptrs, cam.ctypes.data, tc["means3D"].data_ptr(), scale, C.byref(gd),
            C.byref(gr) if M > 1 else None, _lib.stream_ptr())
        assert st == 0
        torch.cuda.synchronize()
    for a, b in ((p_dc, e_dc.detach()), (p_rest, e_rest.detach())):
        if a.numel():
            # per-view products are the same bits; the views are summed in the same order; Adam's first steps are
            # +-lr wherever the gradient is non-zero: differences can only come from last-bit sums that cancel
            d = (a - b).abs()
            assert (d > 1e-6).float().mean().item() <= 1e-4, d.max().item()
    never = torch.stack([r[:, 3] == 0 for r in recs]).all(0)
    if never.any():  # never visible: zero gradient, zero moments -> parameters untouched
        assert torch.equal(p_dc[never], shs[:, :1][never])


def test_training_loop_with_colour_records_equals_plain_peer(built):
    """World size 1: peer_records.PeerRecordAdam (features rebuilt from this view's colour records, side stream)
    against the single-launch peer optimizer over three optimisation steps through render()."""
    from wast3d_b200.gaussian_renderer import render
    from wast3d_b200.scene import GaussianModel, PipelineParams, orbit_cameras, synthetic_gaussians
    arrs = synthetic_gaussians(20000, seed=5, log_scale_mu=-3.2)
    cams = orbit_cameras(3, 4.03, 0.0, 0.6911, 160, 112, device="cuda", sphere=True)
    bg = torch.zeros(3, device="cuda")
    offs = -torch.rand(112, 160, 2, device="cuda")

    def run(records):
        m = GaussianModel.from_arrays(arrs, device="cuda")
        m.spatial_lr_scale = 1.0
        m.active_sh_degree = 3
        opt = m.training_setup(peer=True, feature_records=records)
        imgs = []
        for cam in cams:
            out = render(cam, m, PipelineParams(), bg, sampling_offsets=offs)
            imgs.append(out["render"].detach().clone())
            (out["render"].square().mean() + 0.1 * out["depth"].mean()).backward()
            if records:
                opt.set_view_centres(cam.camera_center.detach().cpu().reshape(1, 3))
            opt.step(); opt.zero_grad()
        if records:
            opt.sync()
        torch.cuda.synchronize()
        res = imgs, [p.detach().clone() for p in m.parameters()]
        opt.close()
        return res

    ia, pa = run(False)
    ib, pb = run(True)
    assert torch.equal(ia[0], ib[0])
    for a, b in zip(ia[1:], ib[1:]):
        d = (a - b).abs()
        assert d.mean().item() <= 1e-5 and (d > 2e-5).float().mean().item() <= 1e-2
    for a, b in zip(pa, pb):
        d = (a - b).abs() / max(1.0, b.abs().max().item())
        assert (d > 1e-5).float().mean().item() <= 1e-3
