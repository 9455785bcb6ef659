"""Times the C3 configuration (Solver B, n=24, l=64, m=128) on 200k options: kernel variants with / without the L2 scratch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl, _lib
from fastamericanoptionpricing_b200.batch import AloConfig
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
b = wl.c3_batch(N)
a = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
outs = {}
for name, flag in (("scratch", _lib.FLAG_WS_SCRATCH), ("recompute", 0)):
    saved = _lib.FLAG_WS_SCRATCH
    _lib.FLAG_WS_SCRATCH = flag
    cfg = AloConfig(**wl.C3_CFG)
    batch.price_batch(*a, cfg=cfg, want_boundary=False); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = batch.price_batch(*a, cfg=cfg, want_boundary=False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    _lib.FLAG_WS_SCRATCH = saved
    outs[name] = out
    print("%-10s %d options %.1f ms -> %.0f options/s, mean iters %.2f" % (name, N, np.median(ts), N / np.median(ts) * 1e3, out["iters"].double().mean().item()))
d = (outs["scratch"]["price"] - outs["recompute"]["price"]).abs()
print("max |price diff| between variants %.3e (nan-free %d)" % (d[torch.isfinite(d)].max().item(), torch.isfinite(d).sum().item()),
      "iters equal:", bool((outs["scratch"]["iters"] == outs["recompute"]["iters"]).all()))
