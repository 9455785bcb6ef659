"""Times the Crank-Nicolson kernel (CUDA events) at the C4 grid (2000 x 2000): kernel alone on a sample, then the whole
100k-put C4 batch through batch.cn_put_batch, cross-checked against spectral prices (BASELINE.json configs[3])."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl, _lib
from fastamericanoptionpricing_b200.batch import AloConfig
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
full = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
b = wl.c4_batch(max(n, full))
lib = _lib.load()
a = [torch.as_tensor(b[k][:n]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")]
mdt = torch.as_tensor(b["T"][:n] / 2000).cuda()
for mode in (2, 0):
    batch.cn_put_batch(*a, n_space=2000, max_dt=mdt, mode=mode)
    torch.cuda.synchronize()
    price = torch.empty(n, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.faop_cn_put_countdown_batch(n, 2000, mode, *[ctypes.c_void_p(x.data_ptr()) for x in a], ctypes.c_void_p(mdt.data_ptr()),
                                         1.2, 1e-5, 200, None, ctypes.c_void_p(price.data_ptr()), None, None,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("mode %d: %d options, 2000 nodes x ~2000 steps: kernel %.1f ms -> %.0f options/s, %.2f ps per node-step, rc %d" % (
        mode, n, ms, n / ms * 1e3, ms * 1e9 / (n * 2000.0 * 2000.0), rc))
if full:
    af = [torch.as_tensor(b[k][:full]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bf = {k: v[:full] for k, v in b.items()}
    cn = batch.cn_put_batch(*af, n_space=2000, max_dt=torch.as_tensor(b["T"][:full] / 2000), mode=2, Smax=wl.c4_smax(bf))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sp = batch.price_batch(*af[:4], af[4], af[5], torch.ones(full, dtype=torch.uint8), cfg=AloConfig(**{**wl.C2_CFG, "iter_tol": 1e-8, "max_iters": 40}),
                           want_boundary=False)["price"]
    d = (cn - sp).abs().cpu().numpy()
    spread = b["sigma"][:full] * np.sqrt(b["T"][:full])
    print("C4 full: %d puts, 2000 x 2000, projected, Smax = K max(3, exp(3 sigma sqrt T)): %.1f ms -> %.0f options/s; |CN - spectral|: max %.2e "
          "over ALL options (bar 5e-3), median %.2e, max where sigma sqrt(T) > 0.45: %.2e" % (
              full, ms, full / ms * 1e3, d.max(), np.median(d), d[spread > 0.45].max()))
