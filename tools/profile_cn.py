"""Small driver for ncu: Crank-Nicolson kernel on a few hundred options of the C4 batch (2000 x 2000 grid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = wl.c4_batch(n)
a = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")]
p = batch.cn_put_batch(*a, n_space=2000, max_dt=b["T"] / 2000, mode=mode, Smax=wl.c4_smax(b))
torch.cuda.synchronize()
print("ok", float(p.sum()))
