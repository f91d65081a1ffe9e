"""CPU test of oracle/sh_records.py (round-2 design): the SH gradient summed over several views,
rebuilt from 16-byte colour records + xyz + the camera centres, equals the sum of the per-view SH gradients of the
oracle's preprocess backward (which is pinned on the reference library, tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from tests.util import raster_case


@pytest.mark.parametrize("degree", [3, 1, 0])
def test_sh_gradient_from_colour_records_equals_sum_of_views(built, degree):
    from oracle import cpu, sh_records
    views = []
    for cam_index in (1, 3, 6):
        case = raster_case(P=3000, W=96, H=64, seed=4, cam_index=cam_index, degree=degree, log_scale_mu=-3.0)
        inp = cpu.RasterInputs(**case)
        fwd = cpu.forward_all(inp)
        rng = np.random.default_rng(10 + cam_index)
        dpix = rng.normal(size=(3, case["H"], case["W"])).astype(np.float32)
        ddep = rng.normal(size=(case["H"], case["W"])).astype(np.float32)
        g = cpu.backward_all(inp, fwd, dpix, ddep)
        views.append((case, fwd["pre"], g))
    xyz = views[0][0]["means3D"]
    M = views[0][0]["shs"].shape[1]
    want = np.zeros((xyz.shape[0], M, 3), np.float32)
    for _, _, g in views:
        want = (want + g["dL_dsh"]).astype(np.float32)
    recs = [sh_records.colour_records(g["dL_dcolor"], pre["clamped"], pre["radii"]) for _, pre, g in views]
    got = sh_records.sh_grad_from_records(xyz, [c["campos"] for c, _, _ in views], recs, degree, M)
    assert np.abs(want).max() > 0
    # same products, same view order; the C oracle contracts nothing, numpy rounds every product: last-bit room
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-6 * scale
    # culled in every view -> exactly zero; coefficients above the active degree -> exactly zero
    never = np.all([pre["radii"] == 0 for _, pre, _ in views], axis=0)
    assert never.any() and np.all(got[never] == 0) and np.all(want[never] == 0)
    assert np.all(got[:, (degree + 1) ** 2:] == 0) and np.all(want[:, (degree + 1) ** 2:] == 0)
    # records are 16 bytes per Gaussian and view
    assert recs[0].dtype == np.float32 and recs[0].shape == (xyz.shape[0], 4)
