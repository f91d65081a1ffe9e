"""torchrun diagnostic: where the side-stream feature exchange sits relative to the next forward.
Per step (ms from the step's start on the main stream): forward done, backward done, early launch done,
late launch done (side stream).  python -m torch.distributed.run --nproc-per-node N tools/diag_overlap.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch, torch.distributed as dist

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from bench import style_loss
    from wast3d_b200 import distributed as wd
    from wast3d_b200.gaussian_renderer import render
    from wast3d_b200.scene import CONFIGS, GaussianModel, PipelineParams, scene_cameras, synthetic_gaussians
    spec = CONFIGS["c3"]
    pc = GaussianModel.from_arrays(synthetic_gaussians(spec.P, seed=0, garden=spec.garden, log_scale_mu=spec.log_scale_mu), sh_degree=3, device=dev)
    pc.spatial_lr_scale = 5.0
    opt = pc.training_setup(peer=True, average=True, overlap_features=os.environ.get("OVERLAP", "1") == "1")
    cams = scene_cameras(spec, 8, device=dev); pipe = PipelineParams(); bg = torch.zeros(3, device=dev)
    H, W = spec.height, spec.width
    tgt = torch.rand(3, H, W, device=dev); dtgt = torch.rand(H, W, device=dev) * 10
    E = lambda: torch.cuda.Event(enable_timing=True)
    rows = []
    n = 40
    for i in range(n):
        cam = wd.view_for_rank(cams, i, rank, world)
        e0, e1, e2, e3, e4 = E(), E(), E(), E(), E()
        e0.record()
        out = render(cam, pc, pipe, bg); e1.record()
        loss = style_loss(out, tgt, dtgt, fused=True); loss.backward(); e2.record()
        opt.step(); e3.record()
        if opt._side is not None:
            e4.record(opt._side)
        else:
            e4.record()
        opt.zero_grad(set_to_none=True)
        rows.append((e0, e1, e2, e3, e4))
    opt.sync(); torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        for i in range(n - 8, n - 1):
            e0, e1, e2, e3, e4 = rows[i]
            nxt = rows[i + 1]
            print(f"step {i}: fwd {e0.elapsed_time(e1):.3f} bwd {e0.elapsed_time(e2):.3f} early {e0.elapsed_time(e3):.3f} "
                  f"late {e0.elapsed_time(e4):.3f} | next fwd done {e0.elapsed_time(nxt[1]):.3f} (late->next fwd {e4.elapsed_time(nxt[1]):.3f})", flush=True)
        tot = rows[n - 9][0].elapsed_time(rows[n - 1][0]) / 8
        print(f"mean step {tot:.3f} ms", flush=True)
    opt.close(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
