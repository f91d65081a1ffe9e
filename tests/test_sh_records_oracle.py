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


def test_staged_kernel_statements_on_the_cpu(built):
    """csrc/sh_adam.cu compiles its per-Gaussian accumulation for host and device; the host emulation entry point
    runs those statements on the CPU: gradient rebuilt from the records (checked through two Adam steps from zero
    moments, where the update is -lr * g / (|g| + eps) and then depends on g's magnitude) against the numpy
    restatement + a numpy Adam."""
    import ctypes as C
    from oracle import cpu, sh_records

    from wast3d_b200._lib import AdamGroup   # struct wast3d_adam_group (ABI v7 layout)

    lib = C.CDLL(str(built["cuda"]))
    f = lib.wast3d_staged_sh_adam_host_emulation
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    degree, P = 3, 2001
    views = []
    for cam_index in (0, 2, 5):
        case = raster_case(P=P, W=96, H=64, seed=8, cam_index=cam_index, degree=degree, log_scale_mu=-3.0)
        inp = cpu.RasterInputs(**case)
        fwd = cpu.forward_all(inp)
        rng = np.random.default_rng(cam_index)
        g = cpu.backward_all(inp, fwd, rng.normal(size=(3, 64, 96)).astype(np.float32),
                             rng.normal(size=(64, 96)).astype(np.float32))
        views.append((case, sh_records.colour_records(g["dL_dcolor"], fwd["pre"]["clamped"], fwd["pre"]["radii"])))
    xyz = np.ascontiguousarray(views[0][0]["means3D"], np.float32)
    shs = views[0][0]["shs"]
    M = shs.shape[1]
    recs = [np.ascontiguousarray(r, np.float32) for _, r in views]
    campos = np.ascontiguousarray(np.stack([c["campos"] for c, _ in views]), np.float32)
    scale = np.float32(1.0 / len(views))
    grad = (sh_records.sh_grad_from_records(xyz, list(campos), recs, degree, M) * scale).astype(np.float32)

    def np_adam(p, g, m, v, lr, t, b1=0.9, b2=0.999, eps=1e-15):
        m[:] = m + np.float32(1 - b1) * (g - m)
        v[:] = v * np.float32(b2) + np.float32(1 - b2) * g * g
        step = np.float32(lr / (1 - b1 ** t))
        inv = np.float32(1.0 / np.sqrt(1 - b2 ** t))
        p[:] = p - step * (m / (np.sqrt(v) * inv + np.float32(eps)))

    lrs = (2.5e-3, 1.25e-4)
    e_dc, e_rest = shs[:, :1].copy(), shs[:, 1:].copy()
    em = [np.zeros_like(e_dc), np.zeros_like(e_dc), np.zeros_like(e_rest), np.zeros_like(e_rest)]
    p_dc, p_rest = np.ascontiguousarray(e_dc.copy()), np.ascontiguousarray(e_rest.copy())
    sm = [np.zeros_like(p_dc), np.zeros_like(p_dc), np.zeros_like(p_rest), np.zeros_like(p_rest)]
    ptrs = (C.c_void_p * len(recs))(*[r.ctypes.data for r in recs])
    for t in (1, 2):
        np_adam(e_dc, grad[:, :1], em[0], em[1], lrs[0], t)
        np_adam(e_rest, grad[:, 1:], em[2], em[3], lrs[1], t)
        gd = AdamGroup(p_dc.ctypes.data, sm[0].ctypes.data, sm[1].ctypes.data, lrs[0], 0.9, 0.999, 1e-15, t, 0, None)
        gr = AdamGroup(p_rest.ctypes.data, sm[2].ctypes.data, sm[3].ctypes.data, lrs[1], 0.9, 0.999, 1e-15, t, 0, None)
        assert f(P, degree, M, len(recs), ptrs, campos.ctypes.data, xyz.ctypes.data, float(scale), C.byref(gd), C.byref(gr)) == 0
    assert np.abs(p_dc - shs[:, :1]).max() > 1e-3  # something moved
    # moments carry the gradient itself: first moment after two steps = (1-b1)(1 + b1) g
    want_m = (np.float32(0.1) * np.float32(1.9)) * grad[:, 1:]
    assert np.abs(sm[2] - want_m).max() <= 1e-5 * max(1e-12, np.abs(want_m).max())
    assert np.abs(p_dc - e_dc).max() <= 1e-6 and np.abs(p_rest - e_rest).max() <= 1e-6


def _record_exchange_job(rank, world):
    """One view-parallel step at world size 2 over gloo with REAL per-view gradients (CPU oracle): the features'
    update rebuilt on every rank from all-gathered 16-byte colour records against all-reduce(dL/dsh) + Adam."""
    import torch
    import torch.distributed as dist
    from oracle import cpu, sh_records
    cpu.set_threads(2)
    degree, P = 2, 1500
    case = raster_case(P=P, W=64, H=48, seed=6, cam_index=1 + 3 * rank, degree=degree, log_scale_mu=-3.0)
    inp = cpu.RasterInputs(**case)
    fwd = cpu.forward_all(inp)
    rng = np.random.default_rng(50 + rank)
    g = cpu.backward_all(inp, fwd, rng.normal(size=(3, 48, 64)).astype(np.float32),
                         rng.normal(size=(48, 64)).astype(np.float32))
    xyz, shs = case["means3D"], case["shs"]
    M = shs.shape[1]

    def adam1(p, grad, lr):  # first step from zero moments
        m = np.float32(0.1) * grad
        v = np.float32(1 - 0.999) * grad * grad
        return p - np.float32(lr / (1 - 0.9)) * (m / (np.sqrt(v) * np.float32(1 / np.sqrt(1 - 0.999)) + np.float32(1e-15)))

    # (a) textbook: all-reduce (average) of the SH gradients
    gsum = torch.from_numpy(g["dL_dsh"].copy())
    dist.all_reduce(gsum)
    base = adam1(shs, (gsum.numpy() / np.float32(world)).astype(np.float32), 2.5e-3)
    # (b) all-gather of the colour records and camera centres; every rank rebuilds the summed gradient
    rec = torch.from_numpy(sh_records.colour_records(g["dL_dcolor"], fwd["pre"]["clamped"], fwd["pre"]["radii"]))
    recs = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(recs, rec)
    cam = torch.from_numpy(case["campos"].astype(np.float32))
    cams = [torch.empty_like(cam) for _ in range(world)]
    dist.all_gather(cams, cam)
    grad = sh_records.sh_grad_from_records(xyz, [c.numpy() for c in cams], [r.numpy() for r in recs], degree, M)
    mine = adam1(shs, (grad / np.float32(world)).astype(np.float32), 2.5e-3)
    # replicas: every rank computed the same bits
    flat = torch.from_numpy(mine.copy())
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    return {"identical": bool(torch.equal(flat, ref)), "max_diff": float(np.abs(mine - base).max()),
            "moved": float(np.abs(mine - shs).max()), "bytes_per_gaussian": rec.element_size() * rec.shape[1]}


def test_colour_record_exchange_world2_gloo(built):
    from tests.test_distributed_gloo import _run
    out = _run(_record_exchange_job, world=2)
    for o in out:
        assert o["identical"] and o["bytes_per_gaussian"] == 16
        assert o["moved"] > 1e-3 and o["max_diff"] <= 2e-6
