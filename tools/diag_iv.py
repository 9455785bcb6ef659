import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
N = 4000
b = {k: v[:N].copy() for k, v in wl.c2_batch(N).items()}
b["q"] = np.maximum(b["q"], 0.002)
cfg = AloConfig("A", 12, 32, 64, 1e-9, 60)
a = [b[k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
out = batch.price_batch(*a, cfg=cfg, want_boundary=False)
target = out["price"]
sig, trials = batch.implied_vol_batch(target, b["r"], b["q"], b["K"], b["T"], b["S"], b["is_put"], cfg=cfg, sigma_lo=0.08, sigma_hi=1.5)
sig = sig.cpu().numpy()
d = np.abs(sig - b["sigma"])
bad = np.where((d > 1e-7) | np.isnan(d))[0]
re = batch.price_batch(b["r"], b["q"], np.nan_to_num(sig, nan=0.3), b["K"], b["T"], b["S"], b["is_put"], cfg=cfg, want_boundary=False)
lo = batch.price_batch(b["r"], b["q"], np.full(N, 0.08), b["K"], b["T"], b["S"], b["is_put"], cfg=cfg, want_boundary=False)
for i in bad[:40]:
    print("i=%d put=%d r=%.4f q=%.4f sg=%.4f K=%.1f T=%.3f S=%.1f target=%.8f intrinsic=%.4f sig=%.6f price(sig)=%.8f err=%.1e it=%d | price(0.08)=%.6f err %.1e" % (
        i, b["is_put"][i], b["r"][i], b["q"][i], b["sigma"][i], b["K"][i], b["T"][i], b["S"][i], target[i].item(),
        max(b["K"][i] - b["S"][i], 0) if b["is_put"][i] else max(b["S"][i] - b["K"][i], 0), sig[i], re["price"][i].item(), re["err"][i].item(), re["iters"][i].item(),
        lo["price"][i].item(), lo["err"][i].item()))
