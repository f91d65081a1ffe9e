"""wast3d_b200 — B200-native (sm_100a) hot path of WaSt3D's style-transfer optimisation step.

Drop-in modules (same names and signatures as the reference):
    wast3d_b200.diff_gaussian_rasterization   GaussianRasterizationSettings, GaussianRasterizer
    wast3d_b200.simple_knn._C                 distCUDA2
    wast3d_b200.gaussian_renderer             render()
plus wast3d_b200.matching (cluster statistics, nearest-cluster / Gaussian-W2 matching),
wast3d_b200.optim.FusedAdam, wast3d_b200.scene (cameras, synthetic scenes, GaussianModel slice)
and wast3d_b200.distributed (view-parallel gradients, sharded matching).

`install_dropin()` aliases the first two under their reference import names so unmodified
reference scripts (`from diff_gaussian_rasterization import ...`,
`from simple_knn._C import distCUDA2`) pick up this implementation.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_dropin():
    from . import diff_gaussian_rasterization as dgr
    from . import simple_knn as sk
    from .simple_knn import _C as sk_C
    sys.modules["diff_gaussian_rasterization"] = dgr
    sys.modules["diff_gaussian_rasterization._C"] = dgr._C
    sys.modules["simple_knn"] = sk
    sys.modules["simple_knn._C"] = sk_C
    return dgr, sk
