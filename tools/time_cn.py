"""Times the Crank-Nicolson kernel alone (CUDA events) at the C4 grid (2000 x 2000)."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = wl.c4_batch(n)
a = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")]
for mode in (2, 0):
    batch.cn_put_batch(*a, n_space=2000, max_dt=b["T"] / 2000, mode=mode)
    torch.cuda.synchronize()
    lib = _lib.load()
    dts, ns = batch.cn_dt_schedule_batch(b["T"], b["T"] / 2000)
    d = torch.from_numpy(dts).cuda(); nsd = torch.from_numpy(ns).cuda()
    price = torch.empty(n, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.faop_cn_put_batch(n, 2000, mode, *[ctypes.c_void_p(x.data_ptr()) for x in a[:4]], ctypes.c_void_p(a[5].data_ptr()),
                               ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(nsd.data_ptr()), dts.shape[1], 1.2, 1e-5, 200,
                               ctypes.c_void_p(price.data_ptr()), None, None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("mode %d: %d options, 2000x%d grid: kernel %.1f ms -> %.0f options/s, %.2f ns per node-step" % (
        mode, n, dts.shape[1], ms, n / ms * 1e3, ms * 1e6 / (n * 2000.0 * dts.shape[1])), "rc", rc)
