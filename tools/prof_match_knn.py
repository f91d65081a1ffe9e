"""ncu target: a few steady-state calls of the W2 match (16384 x 4096, BASELINE.json configs[3] shape), the
nearest-point match (50k x 10k, configs[0]) and distCUDA2 at 3M points.
    ncu --set full -k regex:"match_kernel|knn_" ... python tools/prof_match_knn.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wast3d_b200 import matching  # noqa: E402
from wast3d_b200.scene import synthetic_gaussians  # noqa: E402
from wast3d_b200.simple_knn._C import distCUDA2  # noqa: E402

dev = "cuda"
rng = np.random.default_rng(0)


def clusters(K):
    m = rng.normal(size=(K, 3)) * 8.0
    A = rng.normal(size=(K, 3, 3)) * rng.uniform(0.05, 0.6, size=(K, 1, 3))
    S = A @ A.transpose(0, 2, 1)
    c6 = np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)
    return torch.from_numpy(m.astype(np.float32)).to(dev), torch.from_numpy(c6.astype(np.float32)).to(dev)


mc, cc = clusters(16384)
ms, cs = clusters(4096)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    matching.w2_match(mc, cc, ms, cs)
a = torch.randn(50000, 3, device=dev) * 1.3
b = torch.randn(10000, 3, device=dev) * 1.3
for _ in range(n):
    matching.nn_match(a, b)
pts = torch.from_numpy(synthetic_gaussians(3_000_000, seed=0, garden=True, log_scale_mu=-4.0)["xyz"]).to(dev)
for _ in range(n):
    distCUDA2(pts)
torch.cuda.synchronize()
# timing outside the profiler: CUDA events
if len(sys.argv) > 2:
    for name, fn in (("w2_match 16384x4096", lambda: matching.w2_match(mc, cc, ms, cs)),
                     ("nn_match 50000x10000", lambda: matching.nn_match(a, b)),
                     ("distCUDA2 3M", lambda: distCUDA2(pts))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 20:.4f} ms per call")
