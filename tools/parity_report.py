"""Measured parity of the CUDA path against the golden outputs of the unmodified reference (tests/golden): max |price - ref|,
max relative boundary-node difference, sweep-count mismatches, per fixture and kernel family."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
print("%-34s %-12s %8s %8s %12s %12s %10s %12s" % ("fixture", "kernel", "options", "parity", "max|dprice|", "max rel dB", "iters !=", "max rel dB0"))
for fn, cfgd in (("ref_c1_testA.npz", dict(solver="A", n=12, l=24, m=48, iter_tol=1e-5, max_iters=20)), ("ref_c2_A.npz", wl.C2_CFG),
                 ("ref_c3_B.npz", wl.C3_CFG)):
    g = np.load(os.path.join(G, fn))
    a = [g["in_" + k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
    ps = np.isfinite(g["price"]) & (g["err"] < 1e-2)
    for kname, k in (("specialised", 0), ("generic", 1)):
        for seeds in ("own", "borrowed"):
            out = {kk: v.cpu().numpy() for kk, v in batch.price_batch(*a, cfg=AloConfig(kernel=k, **cfgd),
                                                                       B0=g["B0"] if seeds == "borrowed" else None).items()}
            dp = np.abs(out["price"] - g["price"])[ps].max()
            dB = np.nanmax(np.abs(out["B"] / g["B"] - 1)[ps])
            dB0 = np.nanmax(np.abs(out["B0"] / g["B0"] - 1)[ps])
            print("%-34s %-12s %8d %8d %12.2e %12.2e %10d %12.2e" % (fn + " / " + seeds + " seeds", kname, len(ps), ps.sum(), dp, dB,
                                                                (out["iters"][ps] != g["iters"][ps]).sum(), dB0))
print("bars: |dprice| <= 1e-9, rel dB <= 1e-10 (BASELINE.json north_star); reference = antdvid/FastAmericanOptionPricing run unmodified")
