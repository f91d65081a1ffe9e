"""Generates tests/golden/*.npz ON THE GPU BOX from the UNMODIFIED reference CUDA sources
(oracle/_ref/libwast3d_ref.so, built by oracle/build_ref.sh from /root/reference).

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy into tests/golden/

The fixtures pin the CPU oracle (tests/test_oracle_golden.py, runs without a GPU) and are an
independent check of the CUDA path (tests/test_raster_gpu.py::test_against_golden).
Inputs come from tests/util.raster_case (seeded numpy), so they are regenerated rather than stored
where possible; they are stored anyway so that the fixture is self-contained."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref  # noqa: E402
from tests.util import raster_case, to_cuda  # noqa: E402

CASES = {
    "raster_sh_jitter": dict(P=600, W=80, H=56, seed=11, log_scale_mu=-3.0, bg=(0.2, 0.1, 0.4)),
    "raster_precomp": dict(P=600, W=64, H=48, seed=12, log_scale_mu=-3.0, use_precomp_color=True,
                           use_precomp_cov=True, jitter=False, degree=0),
}
KEYS = ("bg", "means3D", "opacities", "view", "proj", "campos", "W", "H", "tan_fovx", "tan_fovy", "shs",
        "colors_precomp", "scales", "rotations", "cov3D_precomp", "sampling_offsets", "D", "scale_modifier")


def main(out_dir):
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    for name, kw in CASES.items():
        case = raster_case(**kw)
        tc = to_cuda(case)
        rr = ref.RefRasterizer()
        f = rr.forward(**{k: tc.get(k) for k in KEYS})
        st = rr.state()
        rng = np.random.default_rng(kw["seed"] + 7)
        dpix = rng.normal(size=(3, case["H"], case["W"])).astype(np.float32)
        ddep = rng.normal(size=(case["H"], case["W"])).astype(np.float32)
        g = rr.backward(torch.from_numpy(dpix).cuda(), torch.from_numpy(ddep).cuda())
        blob = {"in_" + k: np.asarray(v) for k, v in case.items() if v is not None}
        blob.update({"dL_dpix": dpix, "dL_ddepth": ddep, "R": np.int64(f["R"]),
                     "color": f["color"].cpu().numpy(), "depth": f["depth"].cpu().numpy(),
                     "radii": f["radii"].cpu().numpy()})
        blob.update({"st_" + k: v.cpu().numpy() for k, v in st.items()})
        blob.update({"g_" + k: v.cpu().numpy() for k, v in g.items()})
        np.savez_compressed(out / f"{name}.npz", **blob)
        print(name, "R =", f["R"], "visible =", int((f["radii"] > 0).sum()))
    from wast3d_b200.scene import synthetic_gaussians
    pts = synthetic_gaussians(3000, seed=21)["xyz"]
    pts[50:56] = pts[49]  # duplicates
    d = ref.knn_dist2(torch.from_numpy(pts).cuda()).cpu().numpy()
    np.savez_compressed(out / "knn_3000.npz", points=pts, mean_dist2=d)
    print("knn ok")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
