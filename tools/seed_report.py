"""QD seed report: the hybr seed (reference's root finder, default) and the Newton seed (opt-in) against the live-reference
fixtures (seed distance, final boundary / price distance, sweep-count mismatches, violations of the conditioned bar), and the
cost of each on the 1M-option C2 batch."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
warnings.simplefilter("ignore")
from fastamericanoptionpricing_b200 import batch, workloads as wl
from fastamericanoptionpricing_b200.batch import AloConfig
from oracle import c_oracle as co
from conftest import load_golden, parity_set

for name, cfgkw in (("ref_c2_A.npz", wl.C2_CFG), ("ref_c3_B.npz", wl.C3_CFG),
                    ("ref_stress_A.npz", dict(solver="A", n=12, l=32, m=64, iter_tol=1e-5, max_iters=40))):
    g = load_golden(name)
    a = [g["in_" + k] for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
    ps = parity_set(g) & ~((g["in_q"] == 0) & ~g["in_is_put"])
    if "stress" in name:
        ps &= (g["B"] > 1e-3 * g["in_K"][:, None]).all(axis=1)
    cond = co.seed_conditioning(g["in_T"], g["in_r"], g["in_q"], g["in_sigma"], g["in_K"], g["in_is_put"], cfgkw["n"])
    allow = 4 * cond * 0.5 ** g["iters"]
    for seed in ("hybr", "newton"):
        for kernel in (0, 1):
            out = {k: v.cpu().numpy() for k, v in batch.price_batch(*a, cfg=AloConfig(kernel=kernel, seed=seed, **cfgkw)).items()}
            sd = np.nanmax(np.abs(out["B0"] / g["B0"] - 1), axis=1)
            dB = np.nanmax(np.abs(out["B"] / g["B"] - 1), axis=1)
            dp = np.abs(out["price"] - g["price"])
            mism = (out["iters"] != g["iters"]) & ps
            plain = ps & ((dB > 1e-10) | (dp > 1e-9))
            condv = ps & ((dB > 1e-10 + allow) | (dp > 1e-9 + allow * g["in_K"]))
            print("%-18s seed %-6s kernel %d: parity set %4d  seed max %.2e med %.1e | dB max %.2e dprice max %.2e | sweep mismatches %d | "
                  "over plain bars %d, over conditioned bars %d (options with allowance > 1e-10: %d)" % (
                      name, seed, kernel, ps.sum(), sd[ps].max(), np.median(sd[ps]), dB[ps].max(), dp[ps].max(), mism.sum(),
                      plain.sum(), condv.sum(), (ps & (allow > 1e-10)).sum()))
            for i in np.nonzero(plain | mism)[0][:6]:
                print("      opt %d T %.2e r %.4f q %.4f sg %.3f put %d: seed %.1e cond %.1e dB %.1e dp %.1e iters %d/%d" % (
                    i, g["in_T"][i], g["in_r"][i], g["in_q"][i], g["in_sigma"][i], g["in_is_put"][i], sd[i], cond[i], dB[i], dp[i],
                    out["iters"][i], g["iters"][i]))

# cost on the 1M-option C2 batch
b = wl.c2_batch()
d = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S")] + [torch.as_tensor(b["is_put"]).cuda().to(torch.uint8)]
for seed in ("hybr", "newton"):
    cfg = AloConfig(seed=seed, **wl.C2_CFG)
    for _ in range(2):
        batch.price_batch(*d, cfg=cfg, want_boundary=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        batch.price_batch(*d, cfg=cfg, want_boundary=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tau = batch.collocation_tau(d[4], 12).reshape(-1)
    rep = [x.repeat_interleave(13) for x in (d[0], d[1], d[2], d[3])]
    e0.record()
    if seed == "hybr":
        batch.qd_boundary_batch(tau, *rep, d[6].repeat_interleave(13))
    e1.record(); torch.cuda.synchronize()
    print("C2 1M, seed %-6s: %.2f ms per step -> %.2f M options/s%s" % (
        seed, ms, 1e-3 * len(b["r"]) / ms, ("; 13M hybr roots as a leaf call: %.2f ms" % e0.elapsed_time(e1)) if seed == "hybr" else ""))
