"""Diagnostic (not a pytest): kNN, matching (incl. tcgen05 lower bound) and Adam vs oracles."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from oracle import cpu, ref
from wast3d_b200.simple_knn._C import knn3, distCUDA2
from wast3d_b200 import matching
from wast3d_b200.scene import synthetic_gaussians

def t_knn():
    for P, garden in [(3, False), (5, False), (5000, False), (20000, True)]:
        g = synthetic_gaussians(max(P, 4), seed=P, garden=garden)["xyz"][:P]
        if P == 5000: g[100:110] = g[99]  # duplicates
        pts = torch.from_numpy(g).cuda()
        d, idx = knn3(pts); torch.cuda.synchronize()
        od, oi = cpu.knn(g)
        rd = ref.knn_dist2(pts).cpu().numpy() if P >= 4 else od
        print(f"knn P={P}: dist bit-equal oracle={(d.cpu().numpy()==od).mean():.4f} ref={(d.cpu().numpy()==rd).mean():.4f} oracle-vs-ref={(od==rd).mean():.4f} idx equal={(idx.cpu().numpy()==oi).mean():.4f}")
    for P, garden in [(300000, False), (3000000, True)]:
        g = synthetic_gaussians(P, seed=1, garden=garden)["xyz"]
        pts = torch.from_numpy(g).cuda()
        for _ in range(2): d = distCUDA2(pts)
        torch.cuda.synchronize(); t = time.time(); d = distCUDA2(pts); torch.cuda.synchronize(); t1 = time.time() - t
        for _ in range(1): rd = ref.knn_dist2(pts)
        torch.cuda.synchronize(); t = time.time(); rd = ref.knn_dist2(pts); torch.cuda.synchronize(); t2 = time.time() - t
        print(f"knn P={P}: ours {t1*1e3:.2f} ms ref {t2*1e3:.2f} ms bit-equal={(d==rd).float().mean().item():.6f}")

def rand_clusters(K, rng, spread=3.0):
    m = rng.normal(size=(K, 3)) * spread
    A = rng.normal(size=(K, 3, 3)) * rng.uniform(0.05, 0.6, size=(K, 1, 3))
    S = A @ A.transpose(0, 2, 1)
    c6 = np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)
    return m.astype(np.float32), c6.astype(np.float32)

def t_match():
    rng = np.random.default_rng(0)
    for Kc, Ks in [(5, 3), (128, 128), (512, 512), (1000, 777), (2048, 4096)]:
        mc, cc = rand_clusters(Kc, rng); ms, cs = rand_clusters(Ks, rng)
        oi, oc, omat = cpu.w2_match(mc, cc, ms, cs, want_matrix=True)
        T = lambda x: torch.from_numpy(x).cuda()
        idx, cost, stats, lb = matching.w2_match(T(mc), T(cc), T(ms), T(cs), return_stats=True, _lb_dump=True)
        torch.cuda.synchronize()
        lbn = lb.cpu().numpy()
        viol = (lbn > omat * (1 + 1e-6) + 1e-6)
        # true (float64) lower bound for reference
        u = np.concatenate([mc, np.sqrt(cc[:, [0]] + cc[:, [3]] + cc[:, [5]])], 1).astype(np.float64)
        v = np.concatenate([ms, np.sqrt(cs[:, [0]] + cs[:, [3]] + cs[:, [5]])], 1).astype(np.float64)
        LB = ((u[:, None, :] - v[None, :, :]) ** 2).sum(-1)
        nrm = (u ** 2).sum(1)[:, None] + (v ** 2).sum(1)[None, :]
        gap = (LB - nrm / 2048.0 - lbn)
        print(f"w2 Kc={Kc} Ks={Ks}: idx equal={(idx.cpu().numpy()==oi).mean():.4f} cost bit-equal={(cost.cpu().numpy()==oc).mean():.4f} "
              f"LB>cost violations={viol.sum()} nan={np.isnan(lbn).sum()} gemm err max={np.abs(gap).max():.3e} (rel {np.abs(gap/nrm).max():.2e}) stats={stats}")
        a = rng.normal(size=(Kc, 3)).astype(np.float32); b = rng.normal(size=(Ks, 3)).astype(np.float32)
        ni, nd = cpu.nn_match(a, b)
        gi, gd = matching.nn_match(T(a), T(b))
        ti = torch.cdist(T(a), T(b)).argmin(1)
        print(f"nn  Na={Kc} Nb={Ks}: idx equal oracle={(gi.cpu().numpy()==ni).mean():.4f} dist bit-equal={(gd.cpu().numpy()==nd).mean():.4f} vs torch-gpu argmin={(gi==ti).float().mean().item():.4f}")
    # stats + speed at C4
    Kc, Ks = 16384, 4096
    mc, cc = rand_clusters(Kc, rng, 8.0); ms, cs = rand_clusters(Ks, rng, 8.0)
    T = lambda x: torch.from_numpy(x).cuda()
    a = [T(mc), T(cc), T(ms), T(cs)]
    for _ in range(3): matching.w2_match(*a)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10): o = matching.w2_match(*a, return_stats=True)
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print(f"w2 C4 {Kc}x{Ks}: {dt*1e3:.3f} ms/call -> {Kc*Ks/dt/1e9:.2f} Gpairs/s stats={o[2]}")
    pts = T(rng.normal(size=(200000, 3)).astype(np.float32)); lab = torch.randint(0, 512, (200000,), device="cuda")
    m, c, n = matching.cluster_stats(pts, lab, 512)
    om, oc_, on = cpu.cluster_stats(pts.cpu().numpy(), lab.cpu().numpy(), 512)
    print("cluster_stats: mean maxabs", np.abs(m.cpu().numpy() - om).max(), "cov maxabs", np.abs(c.cpu().numpy() - oc_).max(), "count eq", (n.cpu().numpy() == on).all())

def t_adam():
    from wast3d_b200.optim import FusedAdam
    torch.manual_seed(0)
    p1 = torch.nn.Parameter(torch.randn(100003, 3, device="cuda")); p2 = torch.nn.Parameter(p1.detach().clone())
    o1 = FusedAdam([{"params": [p1], "lr": 1e-2}], lr=0.0, eps=1e-15); o2 = torch.optim.Adam([{"params": [p2], "lr": 1e-2}], lr=0.0, eps=1e-15)
    for it in range(5):
        g = torch.randn_like(p1) * (10.0 ** (it - 2))
        p1.grad = g.clone(); p2.grad = g.clone(); o1.step(); o2.step()
    print("adam: max abs diff", (p1 - p2).abs().max().item(), "rel", ((p1 - p2).norm() / p2.norm()).item())

if __name__ == "__main__":
    which = sys.argv[1:] or ["knn", "match", "adam"]
    if "adam" in which: t_adam()
    if "knn" in which: t_knn()
    if "match" in which: t_match()
