"""TEST INFRASTRUCTURE ONLY — numpy/ctypes front end of oracle/liboracle.so (raster_oracle.c,
knn_oracle.c, match_oracle.c).  See the headers of those files for the reference citations and
the parity-pinning status."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


class OracleParams(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float)] + \
               [(n, C.c_void_p) for n in ("bg", "means3D", "shs", "colors_precomp", "opacities",
                                          "scales", "rotations", "cov3D_precomp", "view", "proj",
                                          "campos", "sampling_offsets")]


def lib():
    global _lib
    if _lib is None:
        path = _HERE / "liboracle.so"
        if not path.exists():
            from wast3d_b200._build import build_oracle
            build_oracle()
        _lib = C.CDLL(str(path))
        _lib.oracle_bin.restype = C.c_int64
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def set_threads(n: int):
    lib().oracle_set_threads(int(n))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RasterInputs:
    """Host copy of one rasteriser call's inputs (same meaning as wast3d_raster_params)."""

    def __init__(self, *, W, H, tan_fovx, tan_fovy, bg, means3D, opacities, view, proj, campos,
                 shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                 sampling_offsets=None, D=0, scale_modifier=1.0):
        self.W, self.H, self.D = int(W), int(H), int(D)
        self.tan_fovx, self.tan_fovy, self.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
        self.bg, self.means3D, self.opacities = _f32(bg), _f32(means3D), _f32(opacities)
        self.view, self.proj, self.campos = _f32(view), _f32(proj), _f32(campos)
        self.shs, self.colors_precomp = _f32(shs), _f32(colors_precomp)
        self.scales, self.rotations, self.cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
        self.sampling_offsets = _f32(sampling_offsets)
        self.P = int(self.means3D.shape[0])
        self.M = int(self.shs.shape[1]) if self.shs is not None else 0

    def cstruct(self):
        return OracleParams(
            P=self.P, D=self.D, M=self.M, W=self.W, H=self.H, tan_fovx=self.tan_fovx,
            tan_fovy=self.tan_fovy, scale_modifier=self.scale_modifier, bg=_p(self.bg),
            means3D=_p(self.means3D), shs=_p(self.shs), colors_precomp=_p(self.colors_precomp),
            opacities=_p(self.opacities), scales=_p(self.scales), rotations=_p(self.rotations),
            cov3D_precomp=_p(self.cov3D_precomp), view=_p(self.view), proj=_p(self.proj),
            campos=_p(self.campos), sampling_offsets=_p(self.sampling_offsets))


def preprocess(inp: RasterInputs) -> dict:
    P = inp.P
    o = {"radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32),
         "depths": np.zeros(P, np.float32), "cov3D": np.zeros((P, 6), np.float32),
         "conic_opacity": np.zeros((P, 4), np.float32), "rgb": np.zeros((P, 3), np.float32),
         "clamped": np.zeros((P, 3), np.uint8), "tiles_touched": np.zeros(P, np.uint32),
         "fragile": np.zeros(P, np.uint8)}
    s = inp.cstruct()
    lib().oracle_preprocess(C.byref(s), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]),
                            _p(o["cov3D"]), _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["clamped"]),
                            _p(o["tiles_touched"]), _p(o["fragile"]))
    return o


def bin_instances(W, H, radii, means2D, depths, tiles_touched=None) -> dict:
    radii = np.ascontiguousarray(radii, np.int32)
    means2D, depths = _f32(means2D), _f32(depths)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    # upper bound on R: recount exactly like the oracle does (rect area per visible Gaussian)
    cap = int(np.asarray(tiles_touched, dtype=np.int64).sum()) if tiles_touched is not None else None
    if cap is None:
        r = radii.astype(np.float32)
        x0 = np.clip(((means2D[:, 0] - r) / 16).astype(np.int32), 0, gx)
        x1 = np.clip(((means2D[:, 0] + r + 15) / 16).astype(np.int32), 0, gx)
        y0 = np.clip(((means2D[:, 1] - r) / 16).astype(np.int32), 0, gy)
        y1 = np.clip(((means2D[:, 1] + r + 15) / 16).astype(np.int32), 0, gy)
        cap = int((((x1 - x0) * (y1 - y0)).astype(np.int64) * (radii > 0)).sum()) + 16
    point_list = np.zeros(max(cap, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    R = lib().oracle_bin(int(radii.shape[0]), int(W), int(H), _p(radii), _p(means2D), _p(depths),
                         _p(point_list), _p(ranges))
    assert R <= cap, (R, cap)
    return {"R": int(R), "point_list": point_list[:R], "ranges": ranges}


def render_forward(W, H, bg, sampling_offsets, ranges, point_list, means2D, colors, depths,
                   conic_opacity) -> dict:
    N = W * H
    o = {"final_T": np.zeros((H, W), np.float32), "n_contrib": np.zeros((H, W), np.uint32),
         "color": np.zeros((3, H, W), np.float32), "depth": np.zeros((H, W), np.float32),
         "fragile": np.zeros((H, W), np.uint8)}
    a = [_f32(bg), _f32(sampling_offsets), np.ascontiguousarray(ranges, np.uint32),
         np.ascontiguousarray(point_list, np.uint32), _f32(means2D), _f32(colors), _f32(depths),
         _f32(conic_opacity)]
    lib().oracle_render_forward(int(W), int(H), *[_p(x) for x in a], _p(o["final_T"]),
                                _p(o["n_contrib"]), _p(o["color"]), _p(o["depth"]), _p(o["fragile"]))
    return o


def render_backward(P, W, H, bg, sampling_offsets, ranges, point_list, means2D, conic_opacity,
                    colors, final_T, n_contrib, dL_dpix, dL_ddepth) -> dict:
    o = {"dL_dmean2D": np.zeros((P, 3), np.float32), "dL_dconic": np.zeros((P, 4), np.float32),
         "dL_dopacity": np.zeros(P, np.float32), "dL_dcolor": np.zeros((P, 3), np.float32),
         "dL_dviewdepth": np.zeros(P, np.float32)}
    a = [_f32(bg), _f32(sampling_offsets), np.ascontiguousarray(ranges, np.uint32),
         np.ascontiguousarray(point_list, np.uint32), _f32(means2D), _f32(conic_opacity),
         _f32(colors), _f32(final_T), np.ascontiguousarray(n_contrib, np.uint32), _f32(dL_dpix),
         _f32(dL_ddepth)]
    lib().oracle_render_backward(int(P), int(W), int(H), *[_p(x) for x in a], _p(o["dL_dmean2D"]),
                                 _p(o["dL_dconic"]), _p(o["dL_dopacity"]), _p(o["dL_dcolor"]),
                                 _p(o["dL_dviewdepth"]))
    return o


def gaussian_backward(inp: RasterInputs, radii, clamped, dL_dmean2D, dL_dconic, dL_dcolor,
                      dL_dviewdepth) -> dict:
    P, M = inp.P, inp.M
    o = {"dL_dmean3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
         "dL_dsh": np.zeros((P, M, 3), np.float32), "dL_dscale": np.zeros((P, 3), np.float32),
         "dL_drot": np.zeros((P, 4), np.float32)}
    s = inp.cstruct()
    a = [np.ascontiguousarray(radii, np.int32), np.ascontiguousarray(clamped, np.uint8),
         _f32(dL_dmean2D), _f32(dL_dconic), _f32(dL_dcolor), _f32(dL_dviewdepth)]
    lib().oracle_gaussian_backward(C.byref(s), *[_p(x) for x in a], _p(o["dL_dmean3D"]),
                                   _p(o["dL_dcov3D"]), _p(o["dL_dsh"]) if M else None,
                                   _p(o["dL_dscale"]), _p(o["dL_drot"]))
    return o


def forward_all(inp: RasterInputs) -> dict:
    """K1 -> binning -> K6 entirely on the CPU (what bench.py times as the CPU baseline)."""
    pre = preprocess(inp)
    b = bin_instances(inp.W, inp.H, pre["radii"], pre["means2D"], pre["depths"], pre["tiles_touched"])
    colors = inp.colors_precomp if inp.colors_precomp is not None else pre["rgb"]
    img = render_forward(inp.W, inp.H, inp.bg, inp.sampling_offsets, b["ranges"], b["point_list"],
                         pre["means2D"], colors, pre["depths"], pre["conic_opacity"])
    return {"pre": pre, "bin": b, "img": img, "colors": colors}


def backward_all(inp: RasterInputs, fwd: dict, dL_dpix, dL_ddepth) -> dict:
    pre, b, img = fwd["pre"], fwd["bin"], fwd["img"]
    g7 = render_backward(inp.P, inp.W, inp.H, inp.bg, inp.sampling_offsets, b["ranges"],
                         b["point_list"], pre["means2D"], pre["conic_opacity"], fwd["colors"],
                         img["final_T"], img["n_contrib"], dL_dpix, dL_ddepth)
    g9 = gaussian_backward(inp, pre["radii"], pre["clamped"], g7["dL_dmean2D"], g7["dL_dconic"],
                           g7["dL_dcolor"], g7["dL_dviewdepth"])
    return {**g7, **g9}


def mark_visible(means3D, view):
    m = _f32(means3D)
    out = np.zeros(m.shape[0], np.uint8)
    lib().oracle_mark_visible(int(m.shape[0]), _p(m), _p(_f32(view)), _p(out))
    return out.astype(bool)


# ----------------------------------------------------------------------------- matching / kNN
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def cdist(a, b, sqrt=True):
    a, b = _f32(a), _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    lib().oracle_cdist(a.shape[0], b.shape[0], _p(a), _p(b), _p(out), int(bool(sqrt)))
    return out


def nn_match(a, b):
    a, b = _f32(a), _f32(b)
    idx = np.zeros(a.shape[0], np.int32)
    dist = np.zeros(a.shape[0], np.float32)
    lib().oracle_nn_match(a.shape[0], b.shape[0], _p(a), _p(b), _p(idx), _p(dist))
    return idx, dist


def cdist_topk(a, b, k):
    a, b = _f32(a), _f32(b)
    dist = np.zeros((a.shape[0], k), np.float32)
    idx = np.zeros((a.shape[0], k), np.int32)
    lib().oracle_cdist_topk(a.shape[0], b.shape[0], _p(a), _p(b), int(k), _p(dist), _p(idx))
    return dist, idx


def emd2_uniform(xa, xb, want_matrix=False):
    xa, xb = _f32(xa), _f32(xb)
    n = xa.shape[0]
    perm = np.zeros(n, np.int32)
    M = np.zeros((n, n), np.float32) if want_matrix else None
    f = lib().oracle_emd2_uniform
    f.restype = C.c_float
    cost = float(f(n, _p(xa), _p(xb), _p(perm), _p(M)))
    return (cost, perm, M) if want_matrix else (cost, perm)


def w2_match(mean_c, cov_c, mean_s, cov_s, want_matrix=False):
    mean_c, cov_c, mean_s, cov_s = _f32(mean_c), _f32(cov_c), _f32(mean_s), _f32(cov_s)
    Kc, Ks = mean_c.shape[0], mean_s.shape[0]
    idx = np.zeros(Kc, np.int32)
    cost = np.zeros(Kc, np.float32)
    mat = np.zeros((Kc, Ks), np.float32) if want_matrix else None
    lib().oracle_w2_match(Kc, Ks, _p(mean_c), _p(cov_c), _p(mean_s), _p(cov_s), _p(idx), _p(cost), _p(mat))
    return (idx, cost, mat) if want_matrix else (idx, cost)


def cluster_stats(points, labels, K):
    points, labels = _f32(points), _i32(labels)
    mean = np.zeros((K, 3), np.float32)
    cov = np.zeros((K, 6), np.float32)
    count = np.zeros(K, np.int32)
    lib().oracle_cluster_stats(points.shape[0], int(K), _p(points), _p(labels), _p(mean), _p(cov), _p(count))
    return mean, cov, count


def knn(points, want_index=True):
    points = _f32(points)
    P = points.shape[0]
    out = np.zeros(P, np.float32)
    idx = np.zeros((P, 3), np.int32) if want_index else None
    lib().oracle_knn(P, _p(points), _p(out), _p(idx))
    return out, idx
