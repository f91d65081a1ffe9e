"""CPU: the oracle (oracle/*.c) reproduces fixtures that were produced by the UNMODIFIED reference
CUDA sources on a B200 (tests/golden/make_golden.py).  This is what pins the oracle."""
from pathlib import Path

import numpy as np
import pytest

from oracle import cpu

GOLD = Path(__file__).resolve().parent / "golden"


def load_case(name):
    z = np.load(GOLD / f"{name}.npz")
    case = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    for k in ("W", "H", "D"):
        case[k] = int(case[k])
    for k in ("tan_fovx", "tan_fovy", "scale_modifier"):
        case[k] = float(case[k])
    return z, case


@pytest.mark.parametrize("name", ["raster_sh_jitter", "raster_precomp"])
def test_preprocess_and_binning(built, name):
    z, case = load_case(name)
    inp = cpu.RasterInputs(**case)
    pre = cpu.preprocess(inp)
    ok = pre["fragile"] == 0
    assert (pre["radii"][ok] == z["radii"][ok]).all()
    assert (pre["radii"] != z["radii"]).sum() <= (~ok).sum()
    vis = z["radii"] > 0
    np.testing.assert_allclose(pre["means2D"][vis], z["st_means2D"][vis], rtol=0, atol=2e-4)
    np.testing.assert_allclose(pre["depths"][vis], z["st_depths"][vis], rtol=2e-6, atol=0)
    # conic = inverse of the 2x2 covariance: elements of O(1) agree to ~1e-7 relative; the off-diagonal term can be
    # 1e-4 of the diagonal (a cancellation, -cov.y / det), so its error is bounded in ABSOLUTE terms (observed: max abs
    # 2.4e-6, median relative 7.5e-8)
    np.testing.assert_allclose(pre["conic_opacity"][vis], z["st_conic_opacity"][vis], rtol=2e-5, atol=4e-6)
    if "colors_precomp" not in case:
        np.testing.assert_allclose(pre["rgb"][vis], z["st_rgb"][vis], rtol=0, atol=2e-6)
        assert (pre["clamped"][vis] == z["st_clamped"][vis]).mean() > 0.999
    # binning on the REFERENCE's K1 state must reproduce the reference's list exactly
    b = cpu.bin_instances(case["W"], case["H"], z["radii"], z["st_means2D"], z["st_depths"])
    assert b["R"] == int(z["R"])
    assert (b["point_list"] == z["st_point_list"].astype(np.uint32)).all()


@pytest.mark.parametrize("name", ["raster_sh_jitter", "raster_precomp"])
def test_render_forward_and_backward(built, name):
    z, case = load_case(name)
    W, H = case["W"], case["H"]
    colors = case["colors_precomp"] if "colors_precomp" in case else z["st_rgb"]
    b = cpu.bin_instances(W, H, z["radii"], z["st_means2D"], z["st_depths"])
    img = cpu.render_forward(W, H, case["bg"], case.get("sampling_offsets"), b["ranges"], b["point_list"],
                             z["st_means2D"], colors, z["st_depths"], z["st_conic_opacity"])
    solid = img["fragile"] == 0
    assert solid.mean() > 0.99
    # north star: colour / depth / alpha within 1e-4 max abs
    assert np.abs(img["color"] - z["color"])[:, solid].max() <= 1e-4
    assert np.abs(img["depth"] - z["depth"])[solid].max() <= 1e-4
    assert np.abs(img["final_T"] - z["st_final_T"])[solid].max() <= 1e-4
    assert (img["n_contrib"][solid] == z["st_n_contrib"].astype(np.uint32)[solid]).all()
    # K7 on the reference's own forward state, then K8+K9
    g7 = cpu.render_backward(len(z["radii"]), W, H, case["bg"], case.get("sampling_offsets"), b["ranges"],
                             b["point_list"], z["st_means2D"], z["st_conic_opacity"], colors, z["st_final_T"],
                             z["st_n_contrib"], z["dL_dpix"], z["dL_ddepth"])
    inp = cpu.RasterInputs(**case)
    g9 = cpu.gaussian_backward(inp, z["radii"], z["st_clamped"], g7["dL_dmean2D"], g7["dL_dconic"],
                               g7["dL_dcolor"], g7["dL_dviewdepth"])
    got = {**g7, **g9}
    rel = lambda a, r: np.linalg.norm(a.ravel().astype(np.float64) - r.ravel()) / max(np.linalg.norm(r.ravel()), 1e-30)
    for k in ("dL_dmean2D", "dL_dopacity", "dL_dcolor", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot", "dL_dsh"):
        refv = z["g_" + k]
        if refv.size == 0 or np.linalg.norm(refv) == 0:
            assert np.abs(got[k]).max(initial=0) == 0
            continue
        assert rel(got[k], refv) <= 1e-3, k   # north star: per-parameter gradients within 1e-3 rel L2
    c = z["g_dL_dconic"].reshape(-1, 4)
    assert rel(got["dL_dconic"][:, [0, 1, 3]], c[:, [0, 1, 3]]) <= 1e-3
    assert rel(got["dL_dviewdepth"], z["g_dL_dviewdepth"].ravel()) <= 1e-3


def test_knn_bit_exact_with_reference(built):
    z = np.load(GOLD / "knn_3000.npz")
    d, _ = cpu.knn(z["points"])
    assert (d == z["mean_dist2"]).all()          # simple_knn.cu semantics, bit for bit
    assert (d[50:56] == d[50]).all()             # duplicates: distance 0 neighbours, self excluded by index
