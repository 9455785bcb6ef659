"""Times the seed kernels alone (CUDA events) on the 1M-option C2 batch: hybr (default) and Newton (opt-in)."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
b = wl.c2_batch()
d = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")] + [torch.as_tensor(b["is_put"]).cuda().to(torch.uint8)]
for seed in ("hybr", "newton"):
    # max_iters = 0: the boundary kernel runs no sweep, so the step is the seed kernel + the price integral on the seed
    for mi in (0, 20):
        cfg = AloConfig(seed=seed, **{**wl.C2_CFG, "max_iters": mi})
        for _ in range(2):
            batch.price_batch(*d, cfg=cfg, want_boundary=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            batch.price_batch(*d, cfg=cfg, want_boundary=False)
        e1.record(); torch.cuda.synchronize()
        print("seed %-6s max_iters %2d: %.2f ms per 1M options" % (seed, mi, e0.elapsed_time(e1) / 5))
