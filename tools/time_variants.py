import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
N = 1_000_000
b = wl.c2_batch(N)
a = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
for k in (0, 2, 3):
    cfg = AloConfig(kernel=k, **wl.C2_CFG)
    batch.price_batch(*a, cfg=cfg, want_boundary=False); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.price_batch(*a, cfg=cfg, want_boundary=False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("kernel variant %d: %.2f ms -> %.2f M options/s" % (k, np.median(ts), N / np.median(ts) / 1e3))
