"""Stress batch, own seeds: options over the conditioned bars — seed distance, conditioning probe, sweeps."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
from oracle import c_oracle as co
N = 6000
b = wl.stress_batch(N)
a = [b[k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
for solver, n, l, m, mi in (("A", 12, 32, 64, 40), ("B", 24, 64, 128, 60), ("B", 12, 24, 48, 60), ("A", 16, 40, 50, 40)):
    ref = co.alo_price_batch(solver, *a, n=n, l=l, m=m, iter_tol=1e-5, max_iters=mi)
    sane = ((ref["B"] > 1e-3 * b["K"][:, None]) & (ref["B"] < 1e3 * b["K"][:, None])).all(axis=1)
    ps = np.isfinite(ref["price"]) & (ref["err"] < 1e-2) & sane
    cond = co.seed_conditioning(b["T"], b["r"], b["q"], b["sigma"], b["K"], b["is_put"], n)
    allow = 4 * cond * 0.5 ** np.minimum(ref["iters"], 60)
    out = {k: v.cpu().numpy() for k, v in batch.price_batch(*a, cfg=AloConfig(solver, n, l, m, 1e-5, mi, kernel=1)).items()}
    sd = np.nanmax(np.abs(out["B0"] / ref["B0"] - 1), axis=1)
    dB = np.nanmax(np.abs(out["B"] / ref["B"] - 1), axis=1)
    bad = ps & (dB > 1e-10 + allow)
    print(solver, n, "parity set", ps.sum(), "over plain bar", (ps & (dB > 1e-10)).sum(), "over conditioned bar", bad.sum(),
          "allow>1e-10:", (ps & (allow > 1e-10)).sum())
    for i in np.nonzero(bad)[0][:10]:
        j = np.nanargmax(np.abs(out["B0"][i] / ref["B0"][i] - 1))
        print("   opt %d put %d r %.5f q %.5f sg %.4f T %.3e K %.1f: seed %.2e (node %d: %.10g vs %.10g) cond %.2e iters %d/%d dB %.2e err %.2e" % (
            i, b["is_put"][i], b["r"][i], b["q"][i], b["sigma"][i], b["T"][i], b["K"][i], sd[i], j, out["B0"][i][j], ref["B0"][i][j], cond[i],
            out["iters"][i], ref["iters"][i], dB[i], ref["err"][i]))
