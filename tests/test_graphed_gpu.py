"""CUDA-graph replay of the whole optimisation step (wast3d_b200.graphed.GraphedStep) against the eager step:
same parameters after the same sequence of views, learning-rate changes picked up, capacity overflow reported."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=5, P=50000, W=256, H=192):
    from wast3d_b200.scene import GaussianModel, OptimizationParams, PipelineParams, orbit_cameras, synthetic_gaussians
    arrs = synthetic_gaussians(P, seed=seed, log_scale_mu=-3.4)
    cams = orbit_cameras(3, 4.03, 0.0, 0.6911, W, H, device="cuda", sphere=True)
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")

    def model():
        m = GaussianModel.from_arrays(arrs, device="cuda")
        m.spatial_lr_scale = 1.0
        m.training_setup(OptimizationParams(), in_backward=True)
        return m
    g = torch.Generator(device="cuda").manual_seed(3)
    offs = -torch.rand(H, W, 2, device="cuda", generator=g)
    tgts = [torch.rand(3, H, W, device="cuda", generator=g) for _ in cams]
    dtgts = [3.0 + torch.rand(H, W, device="cuda", generator=g) for _ in cams]
    return model, cams, bg, offs, tgts, dtgts, PipelineParams()


def _loss(out, tgt, dtgt):
    from wast3d_b200.losses import pixel_loss
    return pixel_loss(out["render"], tgt, out["depth"], dtgt, w_l1=1.0, w_tv=1.0, w_depth=0.1)


def test_graphed_step_equals_eager_step(built):
    from wast3d_b200 import _lib
    from wast3d_b200.gaussian_renderer import render
    from wast3d_b200.graphed import GraphedStep
    model, cams, bg, offs, tgts, dtgts, pipe = _setup()
    order = [0, 1, 2, 1, 0, 2, 2, 1]
    lr_at = {3: 2.5e-4, 6: 1.0e-4}    # xyz learning-rate schedule (scene/gaussian_model.py:169-176)
    prev = _lib.set_deterministic(1)
    try:
        # eager: the warm-up step GraphedStep takes on cams[0], then the sequence
        a = model()
        losses_a = []
        for i, k in enumerate([0] + order):
            if i - 1 in lr_at:
                a.optimizer.param_groups[0]["lr"] = lr_at[i - 1]
            out = render(cams[k], a, pipe, bg, sampling_offsets=offs)
            loss = _loss(out, tgts[k] if i else tgts[0], dtgts[k] if i else dtgts[0])
            loss.backward()
            a.optimizer.step()
            a.optimizer.zero_grad(set_to_none=True)
            losses_a.append(float(loss.detach()))
        # graphed
        b = model()
        gs = GraphedStep(b, pipe, bg, cams[0], _loss, target=tgts[0], depth_target=dtgts[0], sampling_offsets=offs)
        assert gs.launches_per_step >= 10
        losses_b = []
        for i, k in enumerate(order):
            if i in lr_at:
                b.optimizer.param_groups[0]["lr"] = lr_at[i]
            gs.set_view(cams[k])
            gs.set_targets(tgts[k], dtgts[k])
            losses_b.append(float(gs.step()))
        assert gs.check() > 1000
        assert b.optimizer.state[b._xyz]["step"] == a.optimizer.state[a._xyz]["step"] == len(order) + 1
        for la, lb in zip(losses_a[1:], losses_b):
            assert abs(la - lb) <= 1e-6 * abs(la)
        for pa, pb in zip(a.optimizer.param_groups, b.optimizer.param_groups):
            ta, tb = pa["params"][0], pb["params"][0]
            # device-side bias corrections (double pow on the GPU) may round differently from the host's in the last
            # bit of a float: equal up to a few ulp of the update
            assert torch.allclose(ta, tb, rtol=0, atol=1e-7 * float(ta.abs().max()) + 1e-9), pa["name"]
            sa, sb = a.optimizer.state[ta], b.optimizer.state[tb]
            assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-6, atol=1e-12)
    finally:
        _lib.set_deterministic(prev)


def test_graphed_step_draws_offsets_and_reports_overflow(built):
    from wast3d_b200.graphed import GraphedStep
    model, cams, bg, offs, tgts, dtgts, pipe = _setup(seed=6)
    m = model()
    gs = GraphedStep(m, pipe, bg, cams[0], _loss, target=tgts[0], depth_target=dtgts[0])   # offsets drawn in the graph
    first = None
    for i in range(12):
        gs.set_view(cams[i % 3])
        gs.set_targets(tgts[i % 3], dtgts[i % 3])
        l = float(gs.step())
        first = l if first is None else first
    gs.check()
    assert l < first                      # the loop optimises
    m2 = model()
    small = GraphedStep(m2, pipe, bg, cams[0], _loss, target=tgts[0], depth_target=dtgts[0], capacity=4096)
    small.step()
    with pytest.raises(RuntimeError, match="binning buffer held"):
        small.check()
