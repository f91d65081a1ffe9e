"""Stage times of the optimizer-in-backward step (K8+K9 with the Adam epilogue) at one configuration.
    WAST3D_ADAM_UNROLL=1|2|4 python tools/prof_adam_bwd.py [c3] [iters]
Also the target of `ncu -k regex:gaussian_backward` captures of that variant."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch
from wast3d_b200 import _lib
from wast3d_b200.gaussian_renderer import render
from wast3d_b200.scene import CONFIGS, GaussianModel, PipelineParams, scene_cameras, synthetic_gaussians


def main():
    spec = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    mode = sys.argv[3] if len(sys.argv) > 3 else "backward"
    dev = torch.device("cuda", 0)
    pc = GaussianModel.from_arrays(synthetic_gaussians(spec.P, seed=0, garden=spec.garden, log_scale_mu=spec.log_scale_mu),
                                   sh_degree=3, device=dev)
    pc.spatial_lr_scale = 5.0
    opt = pc.training_setup(in_backward=True) if mode == "backward" else pc.training_setup(fused=True)
    cams = scene_cameras(spec, 8, device=dev)
    bg = torch.zeros(3, device=dev)
    tgt = torch.rand(3, spec.height, spec.width, device=dev)

    def step(i):
        out = render(cams[i % 8], pc, PipelineParams(), bg)
        ((out["render"] - tgt).abs().mean() + 0.1 * out["depth"].mean()).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    _lib.profile_enable(None)
    _lib.profile_read()
    for i in range(iters):
        step(i)
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    print(mode, "unroll=" + os.environ.get("WAST3D_ADAM_UNROLL", "default"),
          " ".join(f"{k}={v[0] / max(v[1], 1):.4f}" for k, v in pr.items() if v[1]))


if __name__ == "__main__":
    main()
