"""torchrun script (N >= 2 GPUs) for the STAGED colour-record optimizer (wast3d_b200/peer_records.py): a few
optimisation steps of a small scene with `feature_records=True` against the peer optimizer with overlapped feature
exchange, replicas compared across ranks.  Not collected by pytest; run it first thing in the next round:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/peer_records_check.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from wast3d_b200 import distributed as wd
    from wast3d_b200.gaussian_renderer import render
    from wast3d_b200.scene import GaussianModel, PipelineParams, orbit_cameras, synthetic_gaussians
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
    arrs = synthetic_gaussians(P, seed=5, log_scale_mu=-3.2)
    cams = orbit_cameras(8, 4.03, 0.0, 0.6911, 320, 240, device=dev, sphere=True)
    centres = [c.camera_center.detach().cpu() for c in cams]
    bg = torch.zeros(3, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    offs = -torch.rand(240, 320, 2, device=dev, generator=g)

    def run(records, steps=4):
        m = GaussianModel.from_arrays(arrs, device=dev)
        m.spatial_lr_scale = 1.0
        m.active_sh_degree = 3
        opt = m.training_setup(peer=True, average=True, overlap_features=not records, feature_records=records)
        for i in range(steps):
            cam = wd.view_for_rank(cams, i, rank, world)
            out = render(cam, m, PipelineParams(), bg, sampling_offsets=offs)
            (out["render"].square().mean() + 0.1 * out["depth"].mean()).backward()
            if records:
                opt.set_view_centres(torch.stack([centres[(i * world + q) % len(cams)] for q in range(world)]))
            opt.step(); opt.zero_grad()
        opt.sync()
        torch.cuda.synchronize()
        opt.check_peers()
        ps = [p.detach().clone() for p in m.parameters()]
        flat = torch.cat([p.reshape(-1) for p in ps])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(flat, ref), "replicas diverged"
        dist.barrier()
        opt.close()
        return ps

    a = run(False)
    b = run(True)
    worst = 0.0
    for x, y in zip(a, b):
        d = (x - y).abs() / max(1.0, y.abs().max().item())
        worst = max(worst, (d > 1e-5).float().mean().item())
    if rank == 0:
        print(f"colour-record optimizer vs peer optimizer on {world} GPUs: fraction of elements off by > 1e-5: {worst:.2e}",
              flush=True)
    assert worst <= 1e-3, worst
    if rank == 0:
        print("peer_records_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
