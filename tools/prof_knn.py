"""Profiling driver (not a test): distCUDA2 on a scene-shaped cloud, ours and the reference's."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from wast3d_b200.scene import synthetic_gaussians
from wast3d_b200.simple_knn._C import distCUDA2

P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
garden = P > 1_000_000
pts = torch.from_numpy(synthetic_gaussians(P, seed=1, garden=garden)["xyz"]).cuda()
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r
ms, d = t(lambda: distCUDA2(pts))
print(f"ours  P={P}: {ms:.3f} ms")
if "--ref" in sys.argv:
    from oracle import ref
    ms2, d2 = t(lambda: ref.knn_dist2(pts))
    print(f"ref   P={P}: {ms2:.3f} ms   bit-equal {(d == d2).float().mean().item():.6f}")
