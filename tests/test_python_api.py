"""CPU: the Python surface mirrors the reference's (names, fields, argument checks)."""
import inspect

import pytest
import torch

from wast3d_b200.diff_gaussian_rasterization import (GaussianRasterizationSettings, GaussianRasterizer,
                                                     rasterize_gaussians)


def _settings(H=16, W=16):
    return GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3), scale_modifier=1.0,
        viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3),
        prefiltered=False, debug=False)


def test_settings_fields_match_reference():
    # diff_gaussian_rasterization/__init__.py:173-185
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
        "projmatrix", "sh_degree", "campos", "prefiltered", "debug")


def test_forward_signature_matches_reference():
    sig = inspect.signature(GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                    "rotations", "cov3D_precomp", "cam_view_depth", "sampling_offsets"]
    assert list(inspect.signature(rasterize_gaussians).parameters) == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings", "cam_view_depth", "sampling_offsets"]


def test_invalid_combinations_raise_like_reference():
    r = GaussianRasterizer(_settings())
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(x, x, torch.zeros(4, 1), scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(x, x, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.zeros(4, 1), colors_precomp=x)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.zeros(4, 1), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.zeros(4, 1), colors_precomp=x, scales=x)  # scales without rotations


def test_dropin_aliases():
    import sys
    import wast3d_b200
    wast3d_b200.install_dropin()
    from diff_gaussian_rasterization import GaussianRasterizer as G2  # noqa
    from simple_knn._C import distCUDA2  # noqa
    assert G2 is GaussianRasterizer
    for k in ("diff_gaussian_rasterization", "diff_gaussian_rasterization._C", "simple_knn", "simple_knn._C"):
        sys.modules.pop(k, None)


def test_camera_conventions():
    """scene/cameras.py:54-57: stored matrices are transposed; campos = inverse(view)[3,:3]."""
    import numpy as np
    from wast3d_b200.scene import Camera, look_at
    R, T = look_at((3.0, 1.0, 2.0))
    cam = Camera(R, T, 0.9, 0.7, 64, 48, device="cpu")
    wv = cam.world_view_transform
    p = torch.tensor([0.0, 0.0, 0.0, 1.0])
    v = p @ wv  # row-vector convention
    assert abs(v[2].item() - np.linalg.norm([3.0, 1.0, 2.0])) < 1e-5  # origin straight ahead at distance |eye|
    assert abs(v[0].item()) < 1e-5 and abs(v[1].item()) < 1e-5
    assert torch.allclose(cam.camera_center, torch.tensor([3.0, 1.0, 2.0]), atol=1e-5)
    assert torch.allclose(cam.full_proj_transform, wv @ cam.projection_matrix)


def test_update_learning_rate_follows_the_reference_schedule():
    """scene/gaussian_model.py:169-176 + utils/general_utils.py:29-62 (lr_delay_steps = 0): log-linear from
    position_lr_init to position_lr_final (both times spatial_lr_scale) over position_lr_max_steps, xyz group only."""
    import numpy as np
    from wast3d_b200.scene import GaussianModel, OptimizationParams, synthetic_gaussians
    m = GaussianModel.from_arrays(synthetic_gaussians(64, seed=0), device="cpu")
    m.spatial_lr_scale = 5.0
    m.training_setup(OptimizationParams())
    others = [g["lr"] for g in m.optimizer.param_groups[1:]]
    for it in (0, 1, 777, 29_999, 30_000, 45_000):
        t = np.clip(it / 30_000, 0, 1)
        want = float(np.exp(np.log(0.00016 * 5.0) * (1 - t) + np.log(0.0000016 * 5.0) * t))
        got = m.update_learning_rate(it)
        assert abs(got - want) <= 1e-12 * want
        assert m.optimizer.param_groups[0]["lr"] == got
    assert [g["lr"] for g in m.optimizer.param_groups[1:]] == others
    assert m.update_learning_rate(-1) == 0.0


def test_graphed_step_needs_the_optimizer_in_backward():
    """GraphedStep captures render -> loss -> backward (+ Adam): without optim.BackwardFusedAdam the step has
    step-dependent host values and a separate optimizer launch sequence; it refuses instead of capturing a wrong graph."""
    from types import SimpleNamespace
    from wast3d_b200.graphed import GraphedStep
    from wast3d_b200.scene import GaussianModel, OptimizationParams, PipelineParams, synthetic_gaussians
    m = GaussianModel.from_arrays(synthetic_gaussians(32, seed=0), device="cpu")
    m.training_setup(OptimizationParams())     # torch.optim.Adam
    cam = SimpleNamespace(image_height=8, image_width=8, FoVx=0.7, FoVy=0.7)
    with pytest.raises(RuntimeError, match="optimizer-in-backward"):
        GraphedStep(m, PipelineParams(), torch.zeros(3), cam, lambda out, t, d: out["render"].mean())
