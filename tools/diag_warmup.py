"""Diagnostic: round-by-round (8 steps) device time of the bench step over a long run, with
allocator statistics, to see how long a fresh process takes to reach steady state."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from wast3d_b200.gaussian_renderer import render
from wast3d_b200.scene import CONFIGS, GaussianModel, PipelineParams, scene_cameras, synthetic_gaussians

spec = CONFIGS["c3"]
dev = torch.device("cuda", 0)
pc = GaussianModel.from_arrays(synthetic_gaussians(spec.P, seed=0, garden=spec.garden, log_scale_mu=spec.log_scale_mu), sh_degree=3, device=dev)
pc.spatial_lr_scale = 5.0
opt = pc.training_setup(fused=True)
cams = scene_cameras(spec, 8, device=dev)
pipe, bg = PipelineParams(), torch.zeros(3, device=dev)
H, W = spec.height, spec.width
tgt, dtgt = torch.rand(3, H, W, device=dev), torch.rand(H, W, device=dev) * 10
def step(i):
    out = render(cams[i % 8], pc, pipe, bg)
    bench.style_loss(out, tgt, dtgt).backward()
    opt.step(); opt.zero_grad(set_to_none=True)
rows = []
for r in range(40):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for i in range(8): step(r * 8 + i)
    b.record(); torch.cuda.synchronize()
    st = torch.cuda.memory_stats()
    rows.append((a.elapsed_time(b) / 8, st["reserved_bytes.all.current"] / 2**30, st["segment.all.current"], st["num_device_alloc"]))
print("round ms/step:", " ".join(f"{x[0]:.2f}" for x in rows))
print("reserved GiB :", " ".join(f"{x[1]:.1f}" for x in rows))
print("cudaMallocs  :", " ".join(f"{x[3]}" for x in rows))
# CPU issue time without the GPU in the way is not separable (forward syncs), so time single steps
ts = []
for i in range(24):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(i); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("single steps :", " ".join(f"{t:.2f}" for t in ts))
