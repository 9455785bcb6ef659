"""Full-size differential run: the whole 1M-option C2 batch of the bench and 200k options of C3, CUDA vs the plain-C oracle
on the host cores (about two minutes of CPU).  Prints the measured maxima; the bars are 1e-9 (price) and 1e-10 (boundary)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
from oracle import c_oracle as co
for name, N, gen, cfgk in (("C2 (1M, Solver A n=12 l=32 m=64)", 1_000_000, wl.c2_batch, wl.C2_CFG),
                           ("C3 (200k, Solver B n=24 l=64 m=128)", 200_000, wl.c3_batch, wl.C3_CFG)):
    b = {k: v[:N] for k, v in gen(N).items()}
    a = [b[k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
    t0 = time.perf_counter()
    ref = co.alo_price_batch(cfgk["solver"], *a, **{k: cfgk[k] for k in ("n", "l", "m", "iter_tol", "max_iters")})
    tc = time.perf_counter() - t0
    out = {k: v.cpu().numpy() for k, v in batch.price_batch(*a, cfg=AloConfig(**cfgk)).items()}
    ps = np.isfinite(ref["price"]) & (ref["err"] < 1e-2)
    dp = np.abs(out["price"] - ref["price"])[ps]
    dB = np.abs(out["B"] / ref["B"] - 1)[ps]
    print("%s: parity set %d of %d; max |dprice| %.3e (99.9%%: %.1e); max rel dB %.3e; sweep-count mismatches %d; NaN/finite pattern "
          "differs on %d options; outside the parity set: %d (oracle %.0f s on %d threads)"
          % (name, ps.sum(), N, dp.max(), np.quantile(dp, 0.999), np.nanmax(dB), (out["iters"][ps] != ref["iters"][ps]).sum(),
             (np.isfinite(out["price"]) != np.isfinite(ref["price"])).sum(), (~ps).sum(), tc, len(os.sched_getaffinity(0))))
