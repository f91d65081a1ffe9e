"""Diagnostic (not a test): per-step wall time of the phases of the bench step (render / loss / backward /
optimizer, a synchronize after each) and allocator activity, step by step from a cold process, to
explain why a fresh process needs many steps to reach its steady-state step time.
usage: python tools/diag_phases.py [steps] [expandable|native] [peer|fused]"""
import os
import sys
import time
from pathlib import Path

mode = sys.argv[2] if len(sys.argv) > 2 else "expandable"
if mode == "expandable":
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = "expandable_segments:True"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import bench  # noqa: E402
from wast3d_b200.gaussian_renderer import render  # noqa: E402
from wast3d_b200.scene import CONFIGS, GaussianModel, PipelineParams, scene_cameras, synthetic_gaussians  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
optk = sys.argv[3] if len(sys.argv) > 3 else "peer"
spec = CONFIGS["c3"]
dev = torch.device("cuda", 0)
pc = GaussianModel.from_arrays(synthetic_gaussians(spec.P, seed=0, garden=spec.garden, log_scale_mu=spec.log_scale_mu),
                               sh_degree=3, device=dev)
pc.spatial_lr_scale = 5.0
opt = pc.training_setup(peer=True) if optk == "peer" else pc.training_setup(fused=True)
cams = scene_cameras(spec, 8, device=dev)
pipe, bg = PipelineParams(), torch.zeros(3, device=dev)
H, W = spec.height, spec.width
tgt, dtgt = torch.rand(3, H, W, device=dev), torch.rand(H, W, device=dev) * 10
sync = torch.cuda.synchronize
rows = []
print(f"allocator={mode} optimizer={optk}")
for i in range(steps):
    st0 = torch.cuda.memory_stats()
    sync(); t0 = time.perf_counter()
    out = render(cams[i % 8], pc, pipe, bg)
    sync(); t1 = time.perf_counter()
    loss = bench.style_loss(out, tgt, dtgt)
    sync(); t2 = time.perf_counter()
    loss.backward()
    sync(); t3 = time.perf_counter()
    opt.step(); opt.zero_grad(set_to_none=True)
    del out, loss
    sync(); t4 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3,
                 st1["num_device_alloc"] - st0["num_device_alloc"], st1["num_device_free"] - st0["num_device_free"],
                 st1["reserved_bytes.all.current"] / 2**30))
print("step  render  loss  backward  optim | dev_alloc dev_free reservedGiB")
for i, r in enumerate(rows):
    if i < 24 or i % 8 == 0 or i >= steps - 8:
        print(f"{i:4d} {r[0]:7.2f} {r[1]:6.2f} {r[2]:8.2f} {r[3]:6.2f} | {r[4]:3d} {r[5]:3d} {r[6]:6.2f}")
