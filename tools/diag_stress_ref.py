import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fastamericanoptionpricing_b200 import batch
from fastamericanoptionpricing_b200.batch import AloConfig
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_stress_A.npz"))
a = [g["in_" + k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
ps = np.isfinite(g["price"]) & (g["err"] < 1e-2) & ~((g["in_q"] == 0) & ~g["in_is_put"])
for kernel in (0, 1):
    out = {k: v.cpu().numpy() for k, v in batch.price_batch(*a, cfg=AloConfig("A", 12, 32, 64, 1e-5, 40, kernel=kernel), B0=np.nan_to_num(g["B0"], nan=1.0)).items()}
    dB = np.nanmax(np.abs(out["B"] / g["B"] - 1), axis=1); dB[~ps] = 0
    order = np.argsort(-dB)[:6]
    print("kernel", kernel, "max dB %.2e" % dB.max())
    for i in order:
        print("   i=%d dB=%.2e dprice=%.2e put=%d r=%.4f q=%.4f sg=%.4f K=%.1f T=%.4g S=%.1f iters %d/%d err %.2e/%.2e Bref/K=%s" % (
            g["index"][i], dB[i], abs(out["price"][i] - g["price"][i]), g["in_is_put"][i], g["in_r"][i], g["in_q"][i], g["in_sigma"][i], g["in_K"][i],
            g["in_T"][i], g["in_S"][i], out["iters"][i], g["iters"][i], out["err"][i], g["err"][i], np.round(g["B"][i][::4] / g["in_K"][i], 4)))
