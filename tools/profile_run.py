"""Small driver for ncu: prices a prefix of the C2 batch a few times with one kernel variant."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=262144)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--config", default="c2")
a = ap.parse_args()
if a.config == "c2":
    b, cfgd = wl.c2_batch(a.n), wl.C2_CFG
else:
    b, cfgd = wl.c3_batch(a.n), wl.C3_CFG
cfg = batch.AloConfig(kernel=a.kernel, **cfgd)
dev = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
for _ in range(a.reps):
    out = batch.price_batch(*dev, cfg=cfg, want_boundary=False)
torch.cuda.synchronize()
print("ok", float(out["price"].nansum()), float(out["iters"].double().mean()))
