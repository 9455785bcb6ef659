"""Single-process multi-GPU host path (batch.MultiGpuHost): equality with the one-GPU result and wall-clock scaling."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
cfg = AloConfig(**wl.C2_CFG)
G = torch.cuda.device_count()
N = 1_000_000 * G
b = wl.c2_batch(N)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
a = [pin(b[k]) for k in ("r", "q", "sigma", "K", "T", "S")] + [pin(b["is_put"].astype(np.uint8))]
out = {"price": pin(np.empty(N)), "iters": pin(np.empty(N, dtype=np.int32))}
one = batch.HostContext(0)
ref = one.price_batch(*[x[:200_000] for x in a], cfg=cfg)
mg = batch.MultiGpuHost()
mg.price_batch(*a, cfg=cfg, out=out)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); mg.price_batch(*a, cfg=cfg, out=out); ts.append(time.perf_counter() - t0)
ok = np.array_equal(out["price"][:200_000], ref["price"], equal_nan=True) and np.array_equal(out["iters"][:200_000], ref["iters"])
print("devices %d: %d options in %.1f ms -> %.1f M options/s end to end from one host process; equal to the 1-GPU result: %s"
      % (G, N, 1e3 * min(ts), N / min(ts) / 1e6, ok))
