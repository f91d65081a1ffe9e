"""CPU checks of the depth -> normals oracle (oracle/normals.py = kornia's depth_to_normals restated in torch,
train_st_normals.py:113-123): closed-form cases that pin the conventions (axes, sign, replicate padding, min/max
rescale).  kornia itself is absent and unpinned in the reference: parity unpinned, stated in oracle/normals.py."""
import math

import torch

from oracle import normals as on


def test_fronto_parallel_plane_has_normal_minus_or_plus_z():
    d = torch.full((12, 17), 3.0, dtype=torch.float64)
    n = on.depth_to_normals(d, 100.0, 100.0, 8.0, 6.0)
    # a = d/dx xyz = (z/fx, 0, 0), b = (0, z/fy, 0): a x b = +z
    assert torch.allclose(n[2], torch.ones_like(n[2])) and n[:2].abs().max() < 1e-12


def test_tilted_plane_matches_analytic_normal():
    # plane z = z0 + s * X in camera space: points (X, Y, z); with X = (u - cx)/fx * z -> z = z0 / (1 - s (u-cx)/fx)
    H, W, fx, fy, cx, cy, s, z0 = 20, 30, 50.0, 50.0, 15.0, 10.0, 0.2, 4.0
    u = torch.arange(W, dtype=torch.float64)[None].expand(H, W)
    z = z0 / (1.0 - s * (u - cx) / fx)
    n = on.depth_to_normals(z, fx, fy, cx, cy)
    want = torch.tensor([-s, 0.0, 1.0], dtype=torch.float64)
    want = want / want.norm()
    inner = n[:, 2:-2, 2:-2]                    # Sobel of a plane is exact away from the replicate border
    assert (inner - want[:, None, None]).abs().max() < 1e-9


def test_rescale_and_gradient_through_min_max():
    torch.manual_seed(0)
    d = (torch.rand(9, 11, dtype=torch.float64) + 2.0).requires_grad_(True)
    out = on.depth_to_normals01(d, 40.0, 45.0, 5.0, 4.0)
    assert out.min().item() == 0.0 and abs(out.max().item() - 1.0) < 1e-5
    g = torch.randn_like(out)
    (out * g).sum().backward()
    # finite differences on a few pixels (interior, border, corner)
    for (y, x) in ((4, 5), (0, 3), (8, 10), (0, 0)):
        eps = 1e-6
        dp = d.detach().clone(); dp[y, x] += eps
        dm = d.detach().clone(); dm[y, x] -= eps
        fd = ((on.depth_to_normals01(dp, 40.0, 45.0, 5.0, 4.0) * g).sum() -
              (on.depth_to_normals01(dm, 40.0, 45.0, 5.0, 4.0) * g).sum()) / (2 * eps)
        assert math.isclose(fd.item(), d.grad[y, x].item(), rel_tol=1e-4, abs_tol=1e-6)
