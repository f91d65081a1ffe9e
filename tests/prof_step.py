"""Profiling driver (not a pytest): N forward+backward calls of one configuration through the
C ABI, printing the library's per-stage CUDA-event times (wast3d_profile_*)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests.util import *
from wast3d_b200 import _lib

def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    stages = len(sys.argv) > 3 and sys.argv[3] == "stages"
    if cfg == "c2":
        case = raster_case(P=300000, W=800, H=800, seed=0, log_scale_mu=-4.6)
    elif cfg == "c5":
        case = raster_case(P=3000000, W=1920, H=1080, seed=0, log_scale_mu=-4.0, garden=True, radius=6.0, fovx=1.19)
    else:
        case = raster_case(P=3000000, W=1297, H=840, seed=0, log_scale_mu=-4.0, garden=True, radius=6.0, fovx=1.19)
    tc = to_cuda(case)
    H, W = case["H"], case["W"]
    dpix = torch.randn(3, H, W, device="cuda"); ddep = torch.randn(H, W, device="cuda")
    if stages:
        for _ in range(3):
            fwd = call_forward(tc); call_backward(tc, fwd, dpix, ddep, scratch=False)
        _lib.profile_enable(None); _lib.profile_read()
    for i in range(iters):
        torch.cuda.synchronize(); t0 = time.time()
        fwd = call_forward(tc)
        torch.cuda.synchronize(); t1 = time.time()
        call_backward(tc, fwd, dpix, ddep, scratch=False)
        torch.cuda.synchronize(); t2 = time.time()
        if not stages:
            print(f"iter {i}: fwd {1e3*(t1-t0):.3f} ms  bwd {1e3*(t2-t1):.3f} ms  R={fwd[0]} visible={(fwd[3]>0).sum().item()}")
    if stages:
        pr = _lib.profile_read()
        print(cfg, "R=%d" % fwd[0], " ".join(f"{k}={v[0]/v[1]:.4f}" for k, v in pr.items()),
              "sum=%.4f" % sum(v[0] / v[1] for v in pr.values()))

if __name__ == "__main__":
    main()
