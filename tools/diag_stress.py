"""Diagnostic: which options of the stress batch break the boundary bar (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
from oracle import c_oracle as co
N = 6000
b = {k: v[:N].copy() for k, v in wl.c3_batch(N).items()}
g = np.random.default_rng(123)
b["S"][0:500] = b["K"][0:500] * g.uniform(0.02, 0.3, 500)
b["S"][500:1000] = b["K"][500:1000] * g.uniform(3.0, 20.0, 500)
b["T"][1000:1500] = 10 ** g.uniform(-4, -2, 500)
b["q"][1500:1800] = b["r"][1500:1800]
b["q"][1800:2000] = 0.0
b["r"][2000:2100] = 1e-6
a = [b[k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
np.set_printoptions(linewidth=200, precision=6)
for solver, n, l, m, mi in (("A", 12, 32, 64, 40), ("B", 24, 64, 128, 60), ("B", 12, 24, 48, 60), ("A", 16, 40, 50, 40)):
    ref = co.alo_price_batch(solver, *a, n=n, l=l, m=m, iter_tol=1e-5, max_iters=mi)
    for kernel in (0, 1):
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in batch.price_batch(*a, cfg=AloConfig(solver, n, l, m, 1e-5, mi, kernel=kernel)).items()}
        sane = ((ref["B"] > 1e-3 * b["K"][:, None]) & (ref["B"] < 1e3 * b["K"][:, None])).all(axis=1)
        ps = np.isfinite(ref["price"]) & (ref["err"] < 1e-2) & sane
        dB = np.abs(out["B"] / ref["B"] - 1)
        dB[~ps] = 0
        dP = np.abs(out["price"] - ref["price"]); dP[~ps] = 0
        rowmax = np.nanmax(dB, axis=1)
        order = np.argsort(-rowmax)[:8]
        print("==", solver, n, l, "kernel", kernel, "ps", ps.sum(), "max dB", rowmax.max(), "max dP", dP.max(),
              "iters mismatch", (out["iters"][ps] != ref["iters"][ps]).sum(), "n>1e-10:", (rowmax > 1e-10).sum())
        for i in order:
            print("  i=%d dB=%.2e dP=%.2e put=%d r=%.5f q=%.5f sg=%.4f K=%.2f T=%.5g S=%.2f it=%d/%d err=%.2e/%.2e node=%d Bref/K=%s" % (
                i, rowmax[i], dP[i], b["is_put"][i], b["r"][i], b["q"][i], b["sigma"][i], b["K"][i], b["T"][i], b["S"][i],
                out["iters"][i], ref["iters"][i], out["err"][i], ref["err"][i], int(np.nanargmax(dB[i])), (ref["B"][i, ::4] / b["K"][i])))
