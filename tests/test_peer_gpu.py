"""GPU tests of the peer-memory optimizer at world size 1 (the multi-rank run is tests/peer_check.py under
torchrun; `test_peer_two_ranks` launches it when the box has two GPUs)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _groups(ps):
    return [{"params": [p], "lr": lr} for p, lr in zip(ps, (1.6e-4, 2.5e-3, 0.05, 1e-3, 5e-3))]


def test_peer_adam_world1_matches_fused_and_torch(built):
    from wast3d_b200.optim import FusedAdam
    from wast3d_b200.peer import PeerShardedAdam
    torch.manual_seed(0)
    shapes = [(100003, 3), (5000, 15, 3), (777, 1), (5,), (64, 4)]
    p0 = [torch.randn(s, device="cuda") for s in shapes]
    pa = [torch.nn.Parameter(p.clone()) for p in p0]
    pb = [torch.nn.Parameter(p.clone()) for p in p0]
    pc = [torch.nn.Parameter(p.clone()) for p in p0]
    oa = PeerShardedAdam(_groups(pa), lr=0.0, eps=1e-15)
    ob, oc = FusedAdam(_groups(pb), lr=0.0, eps=1e-15), torch.optim.Adam(_groups(pc), lr=0.0, eps=1e-15)
    for a, p in zip(pa, p0):  # parameters moved into the arena keep their values and shapes
        assert torch.equal(a.detach(), p) and a.shape == p.shape
    for it in range(6):
        for a, b, c in zip(pa, pb, pc):
            g = torch.randn_like(b) * (10.0 ** (it - 3))
            a.grad, b.grad, c.grad = g.clone(), g.clone(), g.clone()
        oa.step(); ob.step(); oc.step()
        oa.zero_grad()
    oa.check_peers()
    for a, b, c in zip(pa, pb, pc):
        assert torch.equal(a.detach(), b.detach())  # same arithmetic as adam.cu: bit-exact
        assert (a - c).abs().max().item() <= 1e-6 * max(1.0, c.abs().max().item())
    oa.close()


def test_peer_adam_late_class_matches_fused(built):
    """late_params: second launch on a side stream, own flag set, own shard + moment slice; same results."""
    from wast3d_b200.optim import FusedAdam
    from wast3d_b200.peer import PeerShardedAdam
    torch.manual_seed(1)
    shapes = [(100003, 3), (5000, 15, 3), (777, 1), (5,), (64, 4)]
    p0 = [torch.randn(s, device="cuda") for s in shapes]
    pa = [torch.nn.Parameter(p.clone()) for p in p0]
    pb = [torch.nn.Parameter(p.clone()) for p in p0]
    oa = PeerShardedAdam(_groups(pa), lr=0.0, eps=1e-15, late_params=[pa[1], pa[3]])
    ob = FusedAdam(_groups(pb), lr=0.0, eps=1e-15)
    assert oa.overlap_late and oa.take_late_event() is None
    for it in range(5):
        for a, b in zip(pa, pb):
            g = torch.randn_like(b) * (10.0 ** (it - 2))
            a.grad, b.grad = g.clone(), g.clone()
        oa.step(); ob.step()
        oa.zero_grad()
        if it % 2 == 0:  # consumer orders itself; otherwise the next step() does
            ev = oa.take_late_event()
            assert ev is not None and oa.take_late_event() is None
            torch.cuda.current_stream().wait_event(ev)
    oa.sync()
    oa.check_peers()
    for a, b, p in zip(pa, pb, p0):
        assert a.shape == p.shape and torch.equal(a.detach(), b.detach())
    oa.close()


def test_training_loop_with_overlapped_features_equals_plain_peer(built):
    """Three optimisation steps through render(): overlap_features (colour kernel behind the side-stream
    launch) against the single-launch peer optimizer: bit-identical images and parameters with the deterministic
    tile backward; agreement to the rounding of K7's unordered float atomics (~1e-7 in the image) without it."""
    from wast3d_b200.gaussian_renderer import render
    from wast3d_b200.scene import GaussianModel, PipelineParams, orbit_cameras, synthetic_gaussians
    arrs = synthetic_gaussians(20000, seed=5, log_scale_mu=-3.2)
    cams = orbit_cameras(3, 4.03, 0.0, 0.6911, 160, 112, device="cuda", sphere=True)
    bg = torch.zeros(3, device="cuda")
    offs = -torch.rand(112, 160, 2, device="cuda")

    def run(overlap):
        m = GaussianModel.from_arrays(arrs, device="cuda")
        m.spatial_lr_scale = 1.0
        opt = m.training_setup(peer=True, overlap_features=overlap)
        imgs = []
        for cam in cams:
            out = render(cam, m, PipelineParams(), bg, sampling_offsets=offs)
            imgs.append(out["render"].detach().clone())
            (out["render"].square().mean() + 0.1 * out["depth"].mean()).backward()
            opt.step(); opt.zero_grad()
        if overlap:
            assert opt.take_late_event() is not None  # the last step's launch nobody rendered after
            opt._late_pending = True
            opt.sync()
        torch.cuda.synchronize()
        res = imgs, [p.detach().clone() for p in m.parameters()]
        opt.close()
        return res

    # deterministic tile backward (wast3d_set_deterministic: fixed summation order, no float atomics): the two
    # schedules are the same computation, so EVERY image and EVERY parameter must match bit for bit — a single
    # stale feature read (one step too early) would show
    from wast3d_b200 import _lib
    prev = _lib.set_deterministic(1)
    try:
        ia, pa = run(False)
        ib, pb = run(True)
    finally:
        _lib.set_deterministic(prev)
    for a, b in zip(ia, ib):
        assert torch.equal(a, b)
    assert (ia[0] - ia[1]).abs().max().item() > 1e-3  # the views do differ
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
    # default mode (unordered float atomics in K7, like the reference): same loop, agreement to rounding
    ia, pa = run(False)
    ib, pb = run(True)
    assert torch.equal(ia[0], ib[0])
    for a, b in zip(ia[1:], ib[1:]):
        d = (a - b).abs()
        assert d.mean().item() <= 1e-5 and (d > 2e-5).float().mean().item() <= 1e-2


def test_render_writes_gradients_into_the_arena(built):
    """render() with a grad sink: .grad are the arena views, equal to the plain autograd gradients, and a
    second backward before the step accumulates like autograd does."""
    from wast3d_b200.gaussian_renderer import render
    from wast3d_b200.scene import GaussianModel, PipelineParams, orbit_cameras, synthetic_gaussians
    arrs = synthetic_gaussians(6000, seed=2, log_scale_mu=-3.2)
    cam = orbit_cameras(2, 4.03, 0.0, 0.6911, 128, 96, device="cuda", sphere=True)[0]
    bg = torch.zeros(3, device="cuda")
    offs = -torch.rand(96, 128, 2, device="cuda")

    def grads(peer, twice=False):
        m = GaussianModel.from_arrays(arrs, device="cuda")
        m.spatial_lr_scale = 1.0
        opt = m.training_setup(peer=peer, fused=not peer)
        for _ in range(2 if twice else 1):
            out = render(cam, m, PipelineParams(), bg, sampling_offsets=offs)
            (out["render"].square().mean() + 0.1 * out["depth"].mean()).backward()
        return m, opt, [p.grad.detach().clone() for p in m.parameters()]

    m1, o1, g1 = grads(True)
    for p in m1.parameters():
        assert p.grad.data_ptr() == o1.grad_sink.view_for(p).data_ptr()
    _, _, g0 = grads(False)
    for a, b in zip(g1, g0):
        assert (a - b).norm().item() <= 1e-5 * max(b.norm().item(), 1e-12)  # float atomics: unordered
    _, _, g2 = grads(True, twice=True)
    for a, b in zip(g2, g0):
        assert (a - 2 * b).norm().item() <= 1e-4 * max(b.norm().item(), 1e-12)
    before = [p.detach().clone() for p in m1.parameters()]
    o1.step(); o1.zero_grad()
    assert all(p.grad is None for p in m1.parameters()) and o1.grad_sink.fresh
    assert all(not torch.equal(a, p.detach()) for a, p in zip(before, m1.parameters()))
    o1.close()


def test_peer_two_ranks(built):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run tests/peer_check.py under torchrun on a multi-GPU box)")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "tests" / "peer_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "peer_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
