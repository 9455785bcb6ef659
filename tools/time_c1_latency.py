"""C1 (BASELINE.json configs[0]): the reference's testA.py case through the drop-in class — wall-clock latency of one solve()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fastamericanoptionpricing_b200.FastAmericanOptionSolverA import FastAmericanOptionSolverA, qd
from fastamericanoptionpricing_b200.FastAmericanOptionSolverB import FastAmericanOptionSolverB
from fastamericanoptionpricing_b200 import batch
from fastamericanoptionpricing_b200.batch import AloConfig
for cls, name in ((FastAmericanOptionSolverA, "FastAmericanOptionSolverA"), (FastAmericanOptionSolverB, "FastAmericanOptionSolverB")):
    s = cls(0.04, 0.01, 0.2, 100, 3, qd.OptionType.Put)
    s.DEBUG = False; s.max_iters = 20 if cls is FastAmericanOptionSolverA else 200; s.iter_tol = 1e-5
    v = s.solve(0.0, 80)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); v = s.solve(0.0, 80); ts.append(time.perf_counter() - t0)
    print("%s(0.04, 0.01, 0.2, 100, 3, Put).solve(0, 80) = %.12f in %d sweeps: median %.3f ms, min %.3f ms per call (reference: ~3.8 s)"
          % (name, v, s.num_iters, 1e3 * np.median(ts), 1e3 * min(ts)))
# the same option as a batch of one through the batch API (specialised kernel, no per-iteration record)
cfg = AloConfig("A", 12, 24, 48, 1e-5, 20)
a = [torch.tensor([x], dtype=torch.float64, device="cuda") for x in (0.04, 0.01, 0.2, 100.0, 3.0, 80.0)] + [torch.ones(1, dtype=torch.uint8, device="cuda")]
batch.price_batch(*a, cfg=cfg, want_boundary=False); torch.cuda.synchronize()
ts = []
for _ in range(50):
    t0 = time.perf_counter(); p = batch.price_batch(*a, cfg=cfg, want_boundary=False)["price"].item(); ts.append(time.perf_counter() - t0)
print("batch of one through price_batch (device tensors in, .item() out): median %.3f ms, min %.3f ms" % (1e3 * np.median(ts), 1e3 * min(ts)))
