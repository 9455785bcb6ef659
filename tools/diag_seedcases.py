import os, sys, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
warnings.simplefilter("ignore")
import numpy as np
from fastamericanoptionpricing_b200 import batch
from fastamericanoptionpricing_b200.batch import AloConfig
g = np.load("/root/repo/tests/golden/ref_seedcases_A.npz")
a = [g["in_" + k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
for kernel in (0, 1):
    out = {k: v.cpu().numpy() for k, v in batch.price_batch(*a, cfg=AloConfig("A", 12, 32, 64, 1e-5, 40, kernel=kernel)).items()}
    nanm = np.nonzero(np.isnan(out["price"]) != np.isnan(g["price"]))[0]
    print("kernel", kernel, "nan mismatches", nanm)
    for i in nanm[:8]:
        print("  opt", i, "grp", g["group"][i], "ref price", g["price"][i], "iters", g["iters"][i], "err", g["err"][i], "| got", out["price"][i], out["iters"][i], out["err"][i])
        print("     ref B", g["B"][i][:4], "got B", out["B"][i][:4], "B0 ref", g["B0"][i][:3], "got", out["B0"][i][:3])
    isn = np.isnan(g["price"]) & np.isnan(out["price"])
    print("  iters mismatch among common NaN:", (out["iters"][isn] != g["iters"][isn]).sum())
