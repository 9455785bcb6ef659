import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from fastamericanoptionpricing_b200 import batch
    mode, key, val = int(sys.argv[1]), sys.argv[2], float(sys.argv[3])
    d = dict(r=0.04, q=0.01, sigma=0.2, K=100.0, T=3.0, S=80.0)
    d[key] = val
    t0 = time.perf_counter()
    p = batch.cn_put_batch(d["r"], d["q"], d["sigma"], d["K"], d["T"], d["S"], n_space=64, max_dt=0.05, mode=mode, max_iter=20)
    torch.cuda.synchronize()
    print("mode %d %s=%g -> %r in %.3f s" % (mode, key, val, p.item(), time.perf_counter() - t0), flush=True)
    sys.exit(0)
for mode in (0, 1, 2):
    for key in ("r", "q", "sigma", "K", "T", "S"):
        for val in ("nan", "inf", "-inf", "0.0", "-1.0", "1e-320", "1e300"):
            try:
                out = subprocess.run([sys.executable, __file__, str(mode), key, val], capture_output=True, text=True, timeout=25)
                line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr.strip()[-200:]
                if "in 0." not in line and "in 1." not in line:
                    print("SLOW/ODD:", line, flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT: mode %d %s=%s" % (mode, key, val), flush=True)
print("done")
