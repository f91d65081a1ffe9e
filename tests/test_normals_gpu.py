"""GPU parity of the fused depth -> normals kernels (csrc/normals.cu, wast3d_b200/normals.py) against the torch
restatement of the reference's kornia expression (oracle/normals.py; train_st_normals.py:113-123), evaluated in
float64 (ground truth) and in float32 (what the reference's own torch kernels would give): our float32 result must be
at least as close to the float64 value as a small multiple of the float32 torch chain's own error, forward and backward.
Sizes: small odd shapes, the reference's 800x800 with its hard-coded K, and 1920x1080 (BASELINE.json configs[4])."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _depth(H, W, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 3, H, device="cuda"), torch.linspace(0, 4, W, device="cuda"), indexing="ij")
    d = 4.0 + torch.sin(xx * 2.1) * torch.cos(yy * 1.7) + 0.3 * torch.sin(xx * 9.0 + yy * 5.0)
    d = d + 0.02 * torch.rand(H, W, device="cuda", generator=g)          # rendered depth is noisy at the pixel level
    d[: H // 5, : W // 4] = 0.0                                            # background: depth 0 -> degenerate normals
    return d.float().contiguous()


@pytest.mark.parametrize("H,W,K", [(7, 5, (30.0, 33.0, 2.5, 3.5)), (1, 9, (10.0, 10.0, 4.0, 0.0)), (33, 1, (10.0, 10.0, 0.0, 16.0)),
                                   (97, 131, (120.0, 110.0, 65.0, 48.0)), (800, 800, (1111.0, 1111.0, 400.0, 400.0)),
                                   (1080, 1920, None)])
def test_normals_forward_backward_match_oracle(built, H, W, K):
    from oracle import normals as on
    from wast3d_b200.normals import depth_to_normals01, intrinsics_for
    from wast3d_b200.scene import CONFIGS, scene_cameras
    if K is None:
        K = intrinsics_for(scene_cameras(CONFIGS["c5"], 8, device="cuda")[0])
    d = _depth(H, W, seed=H * 7 + W)
    gen = torch.Generator(device="cuda").manual_seed(5)
    g = torch.randn(3, H, W, device="cuda", generator=gen)

    d1 = d.clone().requires_grad_(True)
    out = depth_to_normals01(d1, *K)
    (out * g).sum().backward()

    d64 = d.double().requires_grad_(True)
    o64 = on.depth_to_normals01(d64, *K)
    (o64 * g.double()).sum().backward()
    d32 = d.clone().requires_grad_(True)
    o32 = on.depth_to_normals01(d32, *K)
    (o32 * g).sum().backward()

    # forward: max abs error against float64, bounded by the float32 torch chain's own error (x4) + 2e-6
    e_ours = (out.double() - o64).abs()
    e_t32 = (o32.double() - o64).abs()
    # degenerate pixels (a x b == 0 exactly or nearly: flat zero-depth background) have an arbitrary direction in every
    # float32 evaluation; compare where the float64 cross product is well away from zero
    n64 = on.depth_to_normals(d64.detach(), *K)
    ok = (n64.norm(dim=0) > 0.5)[None].expand_as(o64)
    if H == 1 or W == 1:
        # one of the two Sobel gradients vanishes identically, so a x b is 0 up to rounding residue: the direction of
        # the "normal" is noise in EVERY evaluation (the float64 torch chain returns non-zero junk here as well, from
        # the summation order of its convolution) and nothing downstream is comparable.  Ours is exactly zero (the
        # kernel takes differences of identical points) and must stay finite.
        assert out.abs().max().item() == 0.0
        assert torch.isfinite(d1.grad).all()
        return
    assert ok.float().mean().item() > 0.7
    assert e_ours[ok].max().item() <= 4.0 * e_t32[ok].max().item() + 2e-6, (e_ours[ok].max().item(), e_t32[ok].max().item())
    assert out.min().item() >= 0.0 and out.max().item() <= 1.0 + 1e-6
    # backward: relative L2 against float64, bounded the same way
    def rel(a, b):
        return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
    r_ours, r_t32 = rel(d1.grad, d64.grad), rel(d32.grad, d64.grad)
    assert r_ours <= 4.0 * r_t32 + 1e-5, (r_ours, r_t32)
    assert torch.isfinite(d1.grad).all()


def test_normals_are_deterministic(built):
    from wast3d_b200.normals import depth_to_normals01
    d = _depth(240, 320, seed=3)
    g = torch.randn(3, 240, 320, device="cuda")
    res = []
    for _ in range(2):
        x = d.clone().requires_grad_(True)
        o = depth_to_normals01(x, 300.0, 300.0, 160.0, 120.0)
        (o * g).sum().backward()
        res.append((o.detach().clone(), x.grad.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
