"""Diagnostic script (not a pytest): our CUDA path vs the real reference lib vs the CPU oracle."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from tests.util import *
from oracle import cpu, ref

def main(P=20000, W=320, H=240, seed=0, **kw):
    case = raster_case(P=P, W=W, H=H, seed=seed, **kw)
    tc = to_cuda(case)
    fwd = call_forward(tc, debug=True)
    R, color, depth, radii, geom, binning, img = fwd
    torch.cuda.synchronize()
    st = export(tc, fwd)
    print(f"ours: R={R} visible={(radii>0).sum().item()} color mean={color.mean().item():.6f}")
    rr = ref.RefRasterizer()
    rf = rr.forward(**{k: tc.get(k) for k in ("bg","means3D","opacities","view","proj","campos","W","H","tan_fovx","tan_fovy","shs","colors_precomp","scales","rotations","cov3D_precomp","sampling_offsets","D","scale_modifier")})
    rs = rr.state()
    print(f"ref : R={rf['R']} visible={(rf['radii']>0).sum().item()} color mean={rf['color'].mean().item():.6f}")
    print("radii mismatches:", (radii != rf["radii"]).sum().item())
    for k in ("depths","means2D","conic_opacity","rgb","tiles_touched","clamped"):
        a, b = st[k], rs[k]
        neq = (a != b).sum().item()
        md = (a.float()-b.float()).abs().max().item()
        print(f"  K1 {k:14s} not-bit-equal={neq:8d} maxabs={md:.3e}")
    if R == rf["R"]:
        print("point_list mismatches:", (st["point_list"] != rs["point_list"]).sum().item())
    print("color maxabs:", (color-rf["color"]).abs().max().item(), "depth maxabs:", (depth-rf["depth"]).abs().max().item(),
          "final_T maxabs:", (st["final_T"]-rs["final_T"]).abs().max().item(), "n_contrib mism:", (st["n_contrib"]!=rs["n_contrib"]).sum().item())
    print("color bit-equal frac:", (color==rf["color"]).float().mean().item())
    # backward
    gen = torch.Generator(device="cuda").manual_seed(seed)
    dpix = torch.randn(3, H, W, device="cuda", generator=gen)
    ddep = torch.randn(H, W, device="cuda", generator=gen)
    ours = call_backward(tc, fwd, dpix, ddep, debug=True)
    names = ("dL_dmean2D","dL_dcolor","dL_dopacity","dL_dmean3D","dL_dcov3D","dL_dsh","dL_dscale","dL_drot","dL_dconic","dL_dviewdepth")
    rg = rr.backward(dpix, ddep)
    for n, t in zip(names, ours):
        b = rg[n]
        print(f"  bwd {n:14s} rel_l2 vs ref={rel_l2(t.reshape(b.shape), b):.3e}  |ref|={b.norm().item():.3e}")
    # CPU oracle staged on OUR forward state
    o6 = cpu.render_forward(W, H, case["bg"], case.get("sampling_offsets"), st["ranges"].cpu().numpy(), st["point_list"].cpu().numpy(),
                            st["means2D"].cpu().numpy(), st["rgb"].cpu().numpy() if case.get("colors_precomp") is None else case["colors_precomp"],
                            st["depths"].cpu().numpy(), st["conic_opacity"].cpu().numpy())
    dc = np.abs(o6["color"] - color.cpu().numpy()).max(0)
    frag = o6["fragile"].astype(bool)
    print("oracle K6: color maxabs all=%.3e nonfragile=%.3e fragile px=%d  n_contrib mism=%d" % (dc.max(), dc[~frag].max(), frag.sum(), (o6["n_contrib"] != st["n_contrib"].cpu().numpy().astype(np.uint32)).sum()))
    inp = cpu.RasterInputs(**case)
    opre = cpu.preprocess(inp)
    print("oracle K1: radii mism=%d (fragile=%d) means2D maxabs=%.3e conic maxrel=%.3e rgb maxabs=%.3e" % (
        (opre["radii"] != radii.cpu().numpy()).sum(), opre["fragile"].sum(), np.abs(opre["means2D"]-st["means2D"].cpu().numpy()).max(),
        (np.abs(opre["conic_opacity"]-st["conic_opacity"].cpu().numpy())/(np.abs(opre["conic_opacity"])+1e-6)).max(), np.abs(opre["rgb"]-st["rgb"].cpu().numpy()).max()))
    ofull = cpu.forward_all(inp)
    ob = cpu.backward_all(inp, ofull, dpix.cpu().numpy(), ddep.cpu().numpy())
    for n, t in zip(names, ours):
        key = n
        b = torch.from_numpy(ob[key])
        print(f"  bwd {n:14s} rel_l2 vs oracle={rel_l2(t.reshape(-1).cpu(), b.reshape(-1)):.3e}")

if __name__ == "__main__":
    main()
    main(P=50000, W=801, H=437, seed=3, bg=(0.3, 0.5, 0.1))
    main(P=30000, W=400, H=300, seed=5, use_precomp_color=True, use_precomp_cov=True, jitter=False)
    # timing sanity
    case = raster_case(P=300000, W=800, H=800, seed=0, log_scale_mu=-4.6)
    tc = to_cuda(case)
    for _ in range(3): fwd = call_forward(tc)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(10): fwd = call_forward(tc)
    torch.cuda.synchronize(); print("fwd 300k 800x800 ms:", (time.time()-t)*100, "R=", fwd[0])
    dpix = torch.randn(3, 800, 800, device="cuda"); ddep = torch.randn(800, 800, device="cuda")
    for _ in range(3): call_backward(tc, fwd, dpix, ddep, scratch=False)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(10): call_backward(tc, fwd, dpix, ddep, scratch=False)
    torch.cuda.synchronize(); print("bwd ms:", (time.time()-t)*100)
    rr = ref.RefRasterizer()
    kw = {k: tc.get(k) for k in ("bg","means3D","opacities","view","proj","campos","W","H","tan_fovx","tan_fovy","shs","colors_precomp","scales","rotations","cov3D_precomp","sampling_offsets","D","scale_modifier")}
    for _ in range(3): rr.forward(**kw)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(10): rr.forward(**kw)
    torch.cuda.synchronize(); print("ref fwd ms:", (time.time()-t)*100)
    for _ in range(3): rr.backward(dpix, ddep)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(10): rr.backward(dpix, ddep)
    torch.cuda.synchronize(); print("ref bwd ms (incl. zeros alloc):", (time.time()-t)*100)
