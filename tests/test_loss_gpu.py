"""Fused pixel losses (csrc/loss.cu) against the reference's torch expressions (utils/loss_utils.py:18-19
l1_loss, :213-215 tv_loss) in fp32 on the same inputs.  Tolerances: loss 1e-5 relative (the kernel sums in
double, torch in fp32), gradients 1e-6 absolute relative to the largest gradient entry (identical formula,
only the constant factors are rounded differently)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_l1(a, b):
    return torch.abs(a - b).mean()


def ref_tv(img):
    return 0.5 * (torch.abs(img[..., 1:, :] - img[..., :-1, :]).mean() + torch.abs(img[..., :, 1:] - img[..., :, :-1]).mean())


@pytest.mark.parametrize("shape", [(3, 64, 96), (3, 840, 1297), (1, 1, 7), (3, 5, 1), (4, 33, 65)])
def test_pixel_loss_matches_torch(built, shape):
    from wast3d_b200.losses import pixel_loss
    torch.manual_seed(1)
    C, H, W = shape
    img = torch.rand(C, H, W, device="cuda", requires_grad=True)
    dep = (torch.rand(H, W, device="cuda") * 10).requires_grad_(True)
    gt, dgt = torch.rand(C, H, W, device="cuda"), torch.rand(H, W, device="cuda") * 10
    with torch.no_grad():  # exact ties exercise sign(0) = 0
        img[:, : H // 2, : W // 2] = gt[:, : H // 2, : W // 2]
    tv = ref_tv(img) if H > 1 and W > 1 else (0.5 * torch.abs(img[..., :, 1:] - img[..., :, :-1]).mean() if W > 1 else
                                               0.5 * torch.abs(img[..., 1:, :] - img[..., :-1, :]).mean())
    want = 0.7 * ref_l1(img, gt) + 1.3 * tv + 0.1 * ((dep - dgt) ** 2).mean()
    gi, gd = torch.autograd.grad(want * 2.5, (img, dep))
    got = pixel_loss(img, gt, dep, dgt, w_l1=0.7, w_tv=1.3, w_depth=0.1)
    hi, hd = torch.autograd.grad(got * 2.5, (img, dep))
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    assert (hi - gi).abs().max().item() <= 1e-6 * gi.abs().max().item() + 1e-12
    assert (hd - gd).abs().max().item() <= 1e-6 * gd.abs().max().item() + 1e-12
    again = pixel_loss(img, gt, dep, dgt, w_l1=0.7, w_tv=1.3, w_depth=0.1)
    assert again.item() == got.item()  # fixed reduction order


def test_reference_named_losses(built):
    from wast3d_b200.losses import l1_loss, tv_loss
    torch.manual_seed(2)
    a, b = torch.rand(3, 50, 70, device="cuda", requires_grad=True), torch.rand(3, 50, 70, device="cuda")
    assert abs(l1_loss(a, b).item() - ref_l1(a, b).item()) <= 1e-6
    assert abs(tv_loss(a).item() - ref_tv(a).item()) <= 1e-6
    g1, = torch.autograd.grad(tv_loss(a), a)
    g2, = torch.autograd.grad(ref_tv(a), a)
    assert (g1 - g2).abs().max().item() <= 1e-6 * g2.abs().max().item()


def test_pixel_loss_rejects_cpu_tensors(built):
    from wast3d_b200.losses import pixel_loss
    with pytest.raises(RuntimeError):
        pixel_loss(torch.rand(3, 4, 4), torch.rand(3, 4, 4))
