"""Device-resident throughput (CUDA events) of the hot path for every BASELINE config + kernel variant -> stdout table."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def dev(b):
    return [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]


rows = []
N = 1_000_000
a2 = dev(wl.c2_batch(N))
for name, cfg in (("C2 A n12 l32 m64 (specialised)", AloConfig(**wl.C2_CFG)),
                  ("C2 A n12 l32 m64 (generic kernel)", AloConfig(kernel=1, **wl.C2_CFG)),
                  ("C1-style A n12 l24 m48 tol1e-5 max20 (specialised)", AloConfig("A", 12, 24, 48, 1e-5, 20)),
                  ("A+derivative n12 l32 m64 (specialised)", AloConfig(**{**wl.C2_CFG, "use_derivative": True})),
                  ("B n12 l32 m64 max_iters 200 (specialised, matrix kernel)", AloConfig(**{**wl.C2_CFG, "solver": "B", "max_iters": 200}))):
    ms, out = timeit(lambda: batch.price_batch(*a2, cfg=cfg, want_boundary=False))
    rows.append((name, N, ms, N / ms * 1e3, float(out["iters"].double().mean()), int(torch.isnan(out["price"]).sum())))
N3 = 200_000
a3 = dev(wl.c3_batch(N3))
ms, out = timeit(lambda: batch.price_batch(*a3, cfg=AloConfig(**wl.C3_CFG), want_boundary=False), reps=2)
rows.append(("C3 B n24 l64 m128 max_iters 200 (Clenshaw + L2 scratch)", N3, ms, N3 / ms * 1e3, float(out["iters"].double().mean()), int(torch.isnan(out["price"]).sum())))
surf = dict(K0=50.0, K1=150.0, T0=1.0 / 52, T1=3.0, V0=0.10, V1=0.60, S=100.0, r=0.04, q=0.01, nK=1000, nT=1000, nV=100)
N5 = 4_000_000
ms, out = timeit(lambda: batch.surface_batch(surf, 48_000_000, N5, AloConfig(**wl.C5_CFG), want_iters=True), reps=2)
rows.append(("C5 surface slice [48M, 52M) decoded on device", N5, ms, N5 / ms * 1e3, float(out[1].double().mean()), int(torch.isnan(out[0]).sum())))
N5d = 100_000_000
buf = torch.empty(N5d, dtype=torch.float64, device="cuda")
ms, out = timeit(lambda: batch.surface_batch(surf, 0, N5d, AloConfig(**wl.C5_CFG), dedup=True, out=buf), reps=3)
rows.append(("C5 full 100M surface, boundary solves deduplicated", N5d, ms, N5d / ms * 1e3, float("nan"), int(torch.isnan(out).sum())))
# leaf kernels (thread per element) on 10M elements
NL = 10_000_000
bl = wl.c2_batch(NL)
al = {k: torch.as_tensor(bl[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")}
for name, fn, bytes_per in (("leaf: European closed form (64 B/option)", lambda: batch.european_batch(al["T"], al["S"], al["r"], al["q"], al["sigma"], al["K"], al["is_put"]), 57),
                            ("leaf: QD+ price incl. boundary root (72 B/option)", lambda: batch.qd_price_batch(al["T"], al["S"], al["r"], al["q"], al["sigma"], al["K"], al["is_put"])[0], 73),
                            ("leaf: Bjerksund-Stensland price (64 B/option)", lambda: batch.bjerksund_stensland_batch(al["S"], al["r"], al["q"], al["sigma"], al["K"], al["T"])[0], 64)):
    ms, out = timeit(fn, reps=3)
    rows.append((name + " %.0f GB/s" % (NL * bytes_per / ms / 1e6), NL, ms, NL / ms * 1e3, float("nan"), int(torch.isnan(out).sum())))
print("%-62s %10s %10s %14s %8s %6s" % ("config", "options", "ms", "options/s", "mean_it", "nan"))
for r in rows:
    print("%-62s %10d %10.1f %14.0f %8.2f %6d" % r)
