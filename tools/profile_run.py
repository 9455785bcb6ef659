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
if a.config == "c5d":     # deduplicated surface sweep: --n options from the start of the C5 surface
    surf = dict(K0=50.0, K1=150.0, T0=1.0 / 52, T1=3.0, V0=0.10, V1=0.60, S=100.0, r=0.04, q=0.01, nK=1000, nT=1000, nV=100)
    buf = torch.empty(a.n, dtype=torch.float64, device="cuda")
    for _ in range(a.reps):
        out = batch.surface_batch(surf, 0, a.n, batch.AloConfig(**wl.C5_CFG), dedup=True, out=buf)
    torch.cuda.synchronize()
    print("ok", float(out.sum()))
    sys.exit(0)
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
