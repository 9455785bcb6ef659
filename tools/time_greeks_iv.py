"""Throughput of the layers on top of the price kernel: Greeks (delta/gamma/theta on the resident boundary + vega/rho by
re-solves) and implied volatility (bracketed root search, one full batch solve per trial) on the C2 batch."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
N = 1_000_000
b = wl.c2_batch(N)
a = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
cfg = AloConfig(**wl.C2_CFG)


def timed(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    return time.perf_counter() - t0, out


dt, g = timed(lambda: batch.greeks_batch(*a, cfg=cfg, want_vega=False, want_rho=False))
print("price + delta + gamma + theta (one solve + four price integrals on its boundary): %d options in %.1f ms -> %.2f M options/s"
      % (N, dt * 1e3, N / dt / 1e6))
dt, g = timed(lambda: batch.greeks_batch(*a, cfg=cfg))
print("... + vega + rho (four more solves): %.1f ms -> %.2f M options/s" % (dt * 1e3, N / dt / 1e6))
target = batch.price_batch(*a, cfg=cfg, want_boundary=False)["price"]
dt, (iv, trials) = timed(lambda: batch.implied_vol_batch(target, a[0], a[1], a[3], a[4], a[5], a[6], cfg=cfg, sigma_lo=0.08, sigma_hi=1.5))
ok = torch.isfinite(iv)
print("implied volatility: %d options, at most %d trials per option (finished options leave the batch) in %.1f ms -> %.2f M implied vols/s; %.1f%% bracketed"
      % (N, trials, dt * 1e3, N / dt / 1e6, 100 * ok.float().mean().item()))
