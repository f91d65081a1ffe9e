"""ctypes binding of libwast3d_b200.so (the C ABI declared in include/wast3d_b200.h).

There is deliberately no fallback: if the library is missing, or a call is made without an
sm_100 device, the caller gets a RuntimeError.  Nothing here imports or executes `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libwast3d_b200.so"

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)


class RasterParams(C.Structure):
    """struct wast3d_raster_params (include/wast3d_b200.h)."""

    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int), ("debug", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p), ("campos", C.c_void_p), ("sampling_offsets", C.c_void_p),
        ("raw_params", C.c_int), ("shs_rest", C.c_void_p), ("colour_wait_event", C.c_void_p),
        ("preprojected", C.c_int),
    ]


class NextView(C.Structure):
    """struct wast3d_next_view (include/wast3d_b200.h)."""

    _fields_ = [("width", C.c_int), ("height", C.c_int), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("scale_modifier", C.c_float), ("D", C.c_int), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
                ("campos", C.c_void_p), ("geom_buffer", C.c_void_p), ("radii", C.c_void_p),
                ("offset_min_x", C.c_float), ("offset_max_x", C.c_float), ("offset_min_y", C.c_float),
                ("offset_max_y", C.c_float)]


class AdamSegment(C.Structure):
    """struct wast3d_adam_segment (include/wast3d_b200.h)."""

    _fields_ = [("begin4", C.c_ulonglong), ("end4", C.c_ulonglong), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int), ("reserved", C.c_int)]


class AdamGroup(C.Structure):
    """struct wast3d_adam_group (include/wast3d_b200.h)."""

    _fields_ = [("param", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int),
                ("reserved", C.c_int), ("schedule_dev", C.c_void_p)]


class PairArgs(C.Structure):
    """struct wast3d_pair_args (include/wast3d_b200.h)."""

    _fields_ = [("n", C.c_longlong), ("k", C.c_int), ("formula", C.c_int), ("a", C.c_void_p), ("lda", C.c_int),
                ("b", C.c_void_p), ("ldb", C.c_int), ("center", C.c_void_p), ("idx", C.c_void_p),
                ("row_scale", C.c_void_p), ("a2", C.c_void_p), ("lda2", C.c_int)]


# name -> (restype, argtypes); every symbol of include/wast3d_b200.h
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    "wast3d_strerror": (C.c_char_p, [_i]),
    "wast3d_abi_version": (_i, []),
    "wast3d_adam_schedule_step": (_i, [_i, _vp, _vp, _vp, _vp]),
    "wast3d_device_check": (_i, [_i]),
    "wast3d_raster_forward": (_i, [C.POINTER(RasterParams), ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp,
                                   _vp, _vp, _vp, C.POINTER(_i), _vp]),
    "wast3d_raster_forward_async": (_i, [C.POINTER(RasterParams), ALLOC_FN, _vp, ALLOC_FN, _vp, ALLOC_FN, _vp,
                                         _vp, _vp, _vp, _i, _vp, _vp]),
    "wast3d_raster_backward": (_i, [C.POINTER(RasterParams), _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_raster_backward_raw": (_i, [C.POINTER(RasterParams), _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_raster_backward_raw_adam": (_i, [C.POINTER(RasterParams), _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                             C.POINTER(AdamGroup), C.POINTER(_vp), _vp, _vp]),
    "wast3d_raster_backward_raw_adam_next": (_i, [C.POINTER(RasterParams), _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                                  C.POINTER(AdamGroup), C.POINTER(_vp), _vp, _vp, _vp]),
    "wast3d_raster_geom_bytes": (_sz, [_i]),
    "wast3d_raster_export_state": (_i, [C.POINTER(RasterParams), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp]),
    "wast3d_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_set_tile_cut": (_i, [_i]),
    "wast3d_set_deterministic": (_i, [_i]),
    "wast3d_knn_scratch_bytes": (_sz, [_i]),
    "wast3d_knn_dist2": (_i, [_i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wast3d_cluster_stats": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_cluster_sums": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_cluster_scatter": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_match_scratch_bytes": (_sz, [_i, _i]),
    "wast3d_nn_match": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wast3d_cdist_topk": (_i, [_i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "wast3d_emd2_uniform": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_w2_match": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wast3d_w2_match_debug": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_test_sort_pairs": (_i, [_sz, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "wast3d_test_scan": (_i, [_sz, _vp, _vp, _vp, _vp, _i, _vp]),
    "wast3d_profile_set": (_i, [C.c_uint]),
    "wast3d_profile_slots": (_i, []),
    "wast3d_profile_slot_name": (C.c_char_p, [_i]),
    "wast3d_profile_read": (_i, [_vp, _vp]),
    "wast3d_launch_count": (C.c_ulonglong, [_i]),
    "wast3d_adam_step": (_i, [_sz, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _vp]),
    "wast3d_pixel_loss_scratch_bytes": (_sz, []),
    "wast3d_pixel_loss_forward": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp]),
    "wast3d_pixel_loss_backward": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "wast3d_depth_normals_scratch_bytes": (_sz, []),
    "wast3d_depth_normals_forward": (_i, [_i, _i, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_depth_normals_backward": (_i, [_i, _i, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_kmeans_lloyd": (_i, [_i, _i, _vp, _vp, _vp, _i, C.c_double, C.POINTER(C.c_double), C.POINTER(_i),
                                 C.POINTER(C.c_double), _vp]),
    "wast3d_pair_dist_forward": (_i, [C.POINTER(PairArgs), _vp, _vp]),
    "wast3d_pair_dist_backward": (_i, [C.POINTER(PairArgs), _vp, _vp, _vp, _vp, _vp]),
    "wast3d_pair_loss_scratch_bytes": (_sz, []),
    "wast3d_pair_loss_forward": (_i, [C.POINTER(PairArgs), _vp, _vp, _i, C.c_double, _vp, _vp, _vp]),
    "wast3d_pair_loss_backward": (_i, [C.POINTER(PairArgs), _vp, _vp, _i, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "wast3d_peer_flag_bytes": (_sz, []),
    "wast3d_peer_adam_step": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _vp, _i, _f, C.c_uint,
                                   C.c_double, _i, _vp]),
    "wast3d_peer_error": (_i, [_i]),
    "wast3d_peer_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "wast3d_peer_export": (_i, [_vp, _vp]),
    "wast3d_peer_import": (_i, [_vp, C.POINTER(_vp)]),
    "wast3d_peer_release": (_i, [_vp, _i]),
}

_lib = None
MISSING: list = []  # symbols of the header the loaded library does not export


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("WAST3D_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -m wast3d_b200._build` "
            "(there is no CPU or PyTorch fallback for these kernels)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:  # header and library disagree: calling it raises, tests flag it
            MISSING.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "wast3d_b200"):
    if status != 0:
        msg = load().wast3d_strerror(status).decode()
        raise RuntimeError(f"{what}: {msg} (status {status})")


def require_device(t: torch.Tensor | None = None):
    """The product path must fail loudly without the GPU (no silent eager fallback)."""
    if not torch.cuda.is_available():
        raise RuntimeError("wast3d_b200 needs a CUDA sm_100 device; there is no CPU fallback")
    if t is not None and not t.is_cuda:
        raise RuntimeError("wast3d_b200: expected a CUDA tensor, got one on " + str(t.device))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def fptr(t: torch.Tensor | None, keep: list, dtype=torch.float32) -> int | None:
    """data_ptr of a contiguous `dtype` CUDA view of t; None for empty/absent tensors
    (the reference's "not provided" = size-0 tensor with null data_ptr)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != dtype:
        raise RuntimeError(f"wast3d_b200: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError("wast3d_b200: expected a CUDA tensor, got one on " + str(t.device))
    c = t.contiguous()
    keep.append(c)
    return c.data_ptr()


def bucket_bytes(n: int) -> int:
    """Scratch sizes are rounded up to 1/16..1/8 of their magnitude (>= 1 MiB granules): the
    binning buffer scales with the per-view instance count, and a slightly different size every
    call would make torch's caching allocator cudaMalloc a fresh segment for every new camera
    (tens of milliseconds each, and tens of GB of reserved memory after a few hundred steps)."""
    n = int(n)
    if n <= (1 << 20):
        return n
    g = 1 << max(20, n.bit_length() - 4)
    return (n + g - 1) // g * g


# (device index, kind) -> largest size handed out so far.  The binning buffer scales with the per-view
# instance count; asking torch's caching allocator for a slightly different size per camera makes it
# cudaMalloc a fresh ~0.4 GB segment for every new size (tens to hundreds of milliseconds each, reserved
# memory growing to 2-3x the live set over the first ~60 steps: tools/diag_phases.py).  Requesting the
# high-water size every time makes every call after the first hit the allocator's cache.
_HIGH_WATER: dict = {}


class GrowBuffer:
    """One of the three opaque byte buffers; plays resizeFunctional (rasterize_points.cu:27-33)."""

    def __init__(self, device, kind: str = "scratch"):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.error = None
        key = (torch.device(device).index, kind)

        def _alloc(nbytes, _user):
            try:
                want = bucket_bytes(nbytes)
                have = _HIGH_WATER.get(key, 0)
                if want > have:
                    # a new maximum: leave 1/8 headroom so the next slightly larger view fits as well
                    _HIGH_WATER[key] = have = bucket_bytes(want + want // 8) if have else want
                self.tensor = torch.empty(have, dtype=torch.uint8, device=self.device)
                return self.tensor.data_ptr()
            except Exception as e:  # never let an exception cross the C boundary
                self.error = e
                return None

        self.cb = ALLOC_FN(_alloc)

    def take(self):
        """Hand the buffer to the caller and drop the callback.  The callback's closure refers back to
        this object, a reference cycle: left alone it keeps the (0.4 GB) buffer alive until Python's
        cyclic collector happens to run, the caching allocator then has to cudaMalloc another one
        (a ~10 ms stall every few dozen steps)."""
        t, self.tensor, self.cb = self.tensor, None, None
        return t


class FixedBuffer:
    """GrowBuffer stand-in that hands the library one pre-filled tensor (the geometry buffer a previous
    wast3d_raster_backward_raw_adam_next already wrote K1's outputs into)."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor
        self.error = None

        def _alloc(nbytes, _user):
            if self.tensor is None or nbytes > self.tensor.numel():
                self.error = RuntimeError("wast3d_b200: pre-projected geometry buffer is too small")
                return None
            return self.tensor.data_ptr()

        self.cb = ALLOC_FN(_alloc)

    def take(self):
        t, self.tensor, self.cb = self.tensor, None, None
        return t


def geom_buffer(device, P: int) -> torch.Tensor:
    """A geometry buffer for P Gaussians (wast3d_raster_geom_bytes), sized like GrowBuffer(device, 'geom') would."""
    gb = GrowBuffer(device, "geom")
    ptr = gb.cb(int(load().wast3d_raster_geom_bytes(int(P))), None)
    if gb.error is not None or not ptr:
        raise gb.error or RuntimeError("geometry buffer allocation failed")
    return gb.take()


def set_tile_cut(mode: int) -> int:
    """1 (default) = Gaussians are instantiated only in tiles that can see alpha >= 1/255; 0 = the
    reference's radius rectangles (reproduces its num_rendered / point list).  Returns the old mode."""
    return int(load().wast3d_set_tile_cut(int(mode)))


def set_deterministic(mode: int) -> int:
    """1 = the tile backward adds in a fixed order (no float atomics; bit-reproducible gradients, tests);
    0 (default) = unordered float atomics like the reference.  Returns the old mode."""
    return int(load().wast3d_set_deterministic(int(mode)))


def profile_slots() -> list:
    lib = load()
    return [lib.wast3d_profile_slot_name(i).decode() for i in range(lib.wast3d_profile_slots())]


def profile_enable(names=None):
    """Enable event timing for the named stages (None = all, [] = off)."""
    slots = profile_slots()
    mask = 0
    for i, n in enumerate(slots):
        if names is None or n in names:
            mask |= 1 << i
    load().wast3d_profile_set(mask)


def profile_read() -> dict:
    """{stage: (total_ms, scopes)} accumulated since the last read."""
    lib = load()
    n = lib.wast3d_profile_slots()
    ms = (C.c_double * n)()
    cnt = (C.c_ulonglong * n)()
    check(lib.wast3d_profile_read(ms, cnt), "profile_read")
    return {name: (ms[i], int(cnt[i])) for i, name in enumerate(profile_slots()) if cnt[i]}


def launch_count(reset: bool = False) -> int:
    return int(load().wast3d_launch_count(int(reset)))
