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
    import threading
    for tag, period in (("no sampler", None), ("nvml 20 ms", 0.02), ("no sampler", None), ("nvml 100 ms", 0.1),
                        ("nvml 20 ms clocks only", -0.02), ("no sampler", None)):
        smp = None
        if period is not None:
            smp = bench.ClockSampler(0, period_s=abs(period))
            if period < 0:  # clocks only: drop the power / reasons queries
                nv, h = smp.nv, smp.h
                smp._stop.set(); smp.th.join()
                smp._stop = threading.Event()
                def _run(smp=smp, nv=nv, h=h):
                    while not smp._stop.is_set():
                        smp.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), 0.0, 0))
                        smp._stop.wait(0.02)
                smp.th = threading.Thread(target=_run, daemon=True); smp.th.start()
            time.sleep(0.3)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if smp: smp.begin()
        a.record()
        for i in range(30):
            step(i)
        b.record()
        torch.cuda.synchronize()
        if smp: smp.end()
        print(f"{tag}: {a.elapsed_time(b) / 30:.3f} ms/step", smp.stop() if smp else "")


if __name__ == "__main__":
    main()
