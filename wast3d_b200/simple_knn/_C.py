"""`from simple_knn._C import distCUDA2` (scene/gaussian_model.py:20) — host shim over the C ABI.

Replaces distCUDA2 -> SimpleKNN::knn (submodules/simple-knn/spatial.cu:15-26,
simple_knn.cu:185-221): mean squared distance to the 3 nearest OTHER points.
"""
from __future__ import annotations

import torch

from .. import _lib


def _run(points: torch.Tensor, want_index: bool):
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    _lib.require_device(points)
    lib = _lib.load()
    P = int(points.size(0))
    means = torch.empty((P,), dtype=torch.float32, device=points.device)
    index = torch.empty((P, 3), dtype=torch.int32, device=points.device) if want_index else None
    if P:
        keep: list = []
        nbytes = lib.wast3d_knn_scratch_bytes(P)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
        with torch.cuda.device(points.device):
            st = lib.wast3d_knn_dist2(P, _lib.fptr(points, keep), means.data_ptr(),
                                      index.data_ptr() if want_index else None,
                                      scratch.data_ptr(), nbytes, _lib.stream_ptr())
        _lib.check(st, "distCUDA2")
    return means, index


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    return _run(points, False)[0]


def knn3(points: torch.Tensor):
    """Extension: (mean_dist2 [P], index [P,3] int32, nearest first, ties -> lowest index)."""
    return _run(points, True)
