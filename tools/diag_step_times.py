"""Diagnostic (not a test): per-step wall/device times of the bench step with and without the
nvidia-smi sampler, and allocator statistics, to explain step-time outliers."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import bench  # noqa: E402
from wast3d_b200.gaussian_renderer import render  # noqa: E402
from wast3d_b200.scene import CONFIGS, GaussianModel, PipelineParams, scene_cameras, synthetic_gaussians  # noqa: E402


def main():
    spec = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
    dev = torch.device("cuda", 0)
    pc = GaussianModel.from_arrays(synthetic_gaussians(spec.P, seed=0, garden=spec.garden, log_scale_mu=spec.log_scale_mu),
                                   sh_degree=3, device=dev)
    pc.spatial_lr_scale = 5.0
    opt = pc.training_setup(fused=True)
    cams = scene_cameras(spec, 8, device=dev)
    pipe, bg = PipelineParams(), torch.zeros(3, device=dev)
    H, W = spec.height, spec.width
    tgt, dtgt = torch.rand(3, H, W, device=dev), torch.rand(H, W, device=dev) * 10

    def step(i):
        out = render(cams[i % 8], pc, pipe, bg)
        bench.style_loss(out, tgt, dtgt).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    def run(n, tag, first=0):
        ts = []
        for i in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step(first + i)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        st = torch.cuda.memory_stats()
        print(f"{tag}: " + " ".join(f"{t:.1f}" for t in ts))
        print(f"   reserved {st['reserved_bytes.all.current'] / 2**30:.2f} GiB, segments {st['segment.all.current']}, "
              f"cudaMalloc retries {st['num_alloc_retries']}, allocs {st['allocation.all.allocated']}")

    run(10, "cold steps 0-9")
    run(16, "warm steps (sync each)")
    # unsynchronised throughput
    for tag, sampler in (("no sampler", False), ("nvidia-smi -lms 100 sampler", True), ("no sampler again", False)):
        s = bench.ClockSampler(0) if sampler else None
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(32):
            step(i)
        b.record()
        torch.cuda.synchronize()
        print(f"{tag}: {a.elapsed_time(b) / 32:.3f} ms/step", s.stop() if s else "")


if __name__ == "__main__":
    main()
