import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
b = wl.c2_batch(262144)
d = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")] + [torch.as_tensor(b["is_put"]).cuda().to(torch.uint8)]
cfg = AloConfig(**{**wl.C2_CFG, "max_iters": 0})
batch.price_batch(*d, cfg=cfg, want_boundary=False)
torch.cuda.synchronize()
