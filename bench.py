#!/usr/bin/env python
"""bench.py — American options priced per second (FP64) on N B200s, next to the CPU path timed on the host cores.

Contract:  python bench.py --gpus N --steps K --warmup W   prints ONE JSON line.
  step      one pass of the hot path (QD seed kernel + boundary/price kernel) over one batch of synthetic options:
            BASELINE.json configs[1] = 1M random puts/calls, FastAmericanOptionSolverA n=12 l=32 (m=64), tol 1e-5,
            max_iters 20 — per GPU (weak scaling: every rank prices its own 1M-option slice, no collective).
  value     options/s over all ranks, inputs resident in HBM, CUDA-event time per step, max over ranks
  e2e       the same through the host-buffer C-ABI call (faop_alo_price_batch_host) with pinned host arrays:
            H2D of the inputs and D2H of prices inside the timed region
  roofline  nominal FP64 FLOPs (SURVEY §8d model) / time of the dominant kernel alone, over the DFMA-pipe peak
            measured live by faop_measure_fp64_peak (MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline  the plain-C oracle port (oracle/alo_oracle.c) on all host cores over a bounded prefix of the batch
--impl reference  times that CPU port alone (rank 0 only) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from fastamericanoptionpricing_b200 import workloads as wl  # noqa: E402

METRIC = "American options priced/sec (FP64)"
UNIT = "options/s"
N_PER_GPU = 1_000_000
kSMSP = 148 * 4          # SM sub-partitions (issue ports) of one B200


def flops_per_option(solver, n, l, m, J):
    """Nominal algorithmic FLOP model of SURVEY.md §8d (add/mul=1, FMA=2, div=sqrt=16, exp=40, log=56, Phi=120, phi=42)."""
    P = (234 if solver == "A" else 268) + 2 * (n + 1)
    return J * (n * l * P + 300 * n) + 2880 * n + 80 * n * l + m * (468 + 2 * (n + 1)) + 500


def cpu_port_rate(cfg, b, target_seconds, threads):
    """Options/s of the plain-C oracle port on `threads` host threads over a prefix sized for ~target_seconds."""
    from oracle import c_oracle as co          # cpu_baseline leg: the one place bench.py may execute oracle/
    kw = dict(n=cfg["n"], l=cfg["l"], m=cfg["m"], iter_tol=cfg["iter_tol"], max_iters=cfg["max_iters"],
              use_derivative=cfg["use_derivative"], nthreads=threads)
    keys = ("r", "q", "sigma", "K", "T", "S", "is_put")
    probe = 64 * threads
    t0 = time.perf_counter()
    co.alo_price_batch(cfg["solver"], *[b[k][:probe] for k in keys], **kw)
    rate = probe / (time.perf_counter() - t0)
    count = int(min(len(b["r"]), max(probe, rate * target_seconds)))
    t0 = time.perf_counter()
    out = co.alo_price_batch(cfg["solver"], *[b[k][:count] for k in keys], **kw)
    dt = time.perf_counter() - t0
    return count / dt, count, dt, out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (the oracle's C port of the reference algorithm) on all host
    threads; each step = a bounded prefix of the same workload.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = len(os.sched_getaffinity(0))
    cfg = dict(wl.C2_CFG)
    b = wl.c2_batch(N_PER_GPU)
    rate0, count, _, _ = cpu_port_rate(cfg, b, 2.0, threads)
    per_step = int(max(64 * threads, min(N_PER_GPU, rate0 * 8.0)))     # ~8 s of CPU work per step
    from oracle import c_oracle as co
    keys = ("r", "q", "sigma", "K", "T", "S", "is_put")
    kw = dict(n=cfg["n"], l=cfg["l"], m=cfg["m"], iter_tol=cfg["iter_tol"], max_iters=cfg["max_iters"], nthreads=threads)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        co.alo_price_batch(cfg["solver"], *[b[k][:per_step] for k in keys], **kw)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = per_step / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: random puts/calls, FastAmericanOptionSolverA n=12 l=32 m=64 tol=1e-5 max_iters=20",
                       "options_per_step": per_step, "seed": wl.C2_SEED},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "first %d options of the 1M C2 batch per step, plain-C port of the reference "
                                       "algorithm (oracle/alo_oracle.c) on %d pthreads" % (per_step, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_c5(args):
    """BASELINE.json configs[4]: 100M puts = 1000 strikes x 1000 maturities x 100 vols (S=100, r=.04, q=.01), Solver A n=12
    l=32 m=64.  The (T, sigma) scenarios are cut across the ranks; each rank solves its scenarios' boundaries once (K = 1)
    and prices all 1000 strikes of each (faop_alo_ladder_batch).  Total work is fixed: strong scaling."""
    import torch
    import torch.distributed as dist
    from fastamericanoptionpricing_b200 import batch
    from fastamericanoptionpricing_b200.sharding import shard_bounds
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    cfg = batch.AloConfig(**wl.C5_CFG)
    nK, nT, nV = 1000, 1000, 100
    Kl = np.linspace(50.0, 150.0, nK)
    Tg, Vg = np.meshgrid(np.linspace(1.0 / 52, 3.0, nT), np.linspace(0.10, 0.60, nV), indexing="ij")
    G = nT * nV
    lo, hi = shard_bounds(G, world, rank)
    g = hi - lo
    T = torch.from_numpy(Tg.reshape(-1)[lo:hi].copy()).to(dev)
    V = torch.from_numpy(Vg.reshape(-1)[lo:hi].copy()).to(dev)
    r = torch.full((g,), 0.04, dtype=torch.float64, device=dev)
    q = torch.full((g,), 0.01, dtype=torch.float64, device=dev)
    S = torch.full((g,), 100.0, dtype=torch.float64, device=dev)
    put = torch.ones(g, dtype=torch.uint8, device=dev)
    Kd = torch.from_numpy(Kl).to(dev)
    out = torch.empty((g, nK), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        batch.ladder_batch(r, q, V, T, S, put, Kd, cfg=cfg, out=out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = batch.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):       # the 800 MB/world output per step is itself larger than L2: nothing survives between steps
        ev[s][0].record()
        batch.ladder_batch(r, q, V, T, S, put, Kd, cfg=cfg, out=out)
        ev[s][1].record()
    barrier()
    launches = batch.launch_count() - l0
    clocks = sampler.stop()
    ms = float(np.mean([a.elapsed_time(b_) for a, b_ in ev]))
    red = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    ms = red.item()
    if rank == 0:
        line = {"metric": METRIC, "value": G * nK / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "C5: 100M-put surface (1000 strikes x 1000 maturities x 100 vols), FastAmericanOptionSolverA "
                                       "n=12 l=32 m=64 tol=1e-5 max_iters=20; boundary solves shared per (T, sigma) scenario "
                                       "(BASELINE.json configs[4], SURVEY 8f row 1)",
                           "scenarios": G, "strikes": nK, "parallelism": "scenarios cut across %d ranks, no collective" % world,
                           "l2": "output (%d MB per rank) exceeds L2" % (g * nK * 8 >> 20)},
                "gpu_launches": int(launches), "clocks": clocks,
                "checksum": float(out.sum().item())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--options", type=int, default=N_PER_GPU, help="options per GPU per step (default: the 1M of configs[1])")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 (default, the headline line) or c5: the 100M-option surface as 100k (T, sigma) scenarios x "
                         "1000 strikes through the deduplicated ladder path, scenarios cut across the ranks (strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_c5(args)

    import torch
    import torch.distributed as dist
    from fastamericanoptionpricing_b200 import batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    Nopt = args.options
    cfgd = dict(wl.C2_CFG)
    cfg = batch.AloConfig(kernel=args.kernel, **cfgd)

    # every rank prices its own slice of a world*Nopt batch: rank r gets the r-th independent 1M draw
    b = wl.c2_batch(Nopt, seed=wl.C2_SEED + 7919 * rank)
    keys = ("r", "q", "sigma", "K", "T", "S")
    host = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in keys}
    host["is_put"] = torch.from_numpy(b["is_put"].astype(np.uint8)).pin_memory()
    d = {k: v.to(dev) for k, v in host.items()}
    args_dev = [d[k] for k in keys] + [d["is_put"]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return batch.price_batch(*args_dev, cfg=cfg, want_boundary=False)

    for _ in range(W):
        out = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = batch.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record()
        out = step()
        ev[s][1].record()
    barrier()
    launches = batch.launch_count() - l0
    ms_steps = [a.elapsed_time(b_) for a, b_ in ev]
    ms_step = float(np.mean(ms_steps))

    # dominant kernel alone: borrowed seeds -> only the boundary/price kernel is launched
    seeds = batch.price_batch(*args_dev, cfg=cfg, want_boundary=True)
    B0 = seeds["B0"]
    iters_mean = float(seeds["iters"].double().mean().item())
    del seeds
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.zero_()
        kev[s][0].record()
        batch.price_batch(*args_dev, cfg=cfg, B0=B0, want_boundary=False)
        kev[s][1].record()
    torch.cuda.synchronize()
    ms_kernel = float(np.mean([a.elapsed_time(b_) for a, b_ in kev]))
    del B0

    # end to end through the host-buffer C-ABI entry point (pinned host arrays in, host arrays out)
    ctx = batch.HostContext(local)
    hout = {"price": torch.empty(Nopt, dtype=torch.float64).pin_memory().numpy()}
    hargs = [host[k].numpy() for k in keys] + [host["is_put"].numpy()]
    for _ in range(2):
        ctx.price_batch(*hargs, cfg=cfg, out=hout)
    barrier()
    e2e_t = []
    l1 = batch.launch_count()
    for s in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.price_batch(*hargs, cfg=cfg, out=hout)          # returns after the D2H copy has completed
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    ms_e2e = 1e3 * float(np.mean(e2e_t))
    e2e_launches = batch.launch_count() - l1
    same = bool(np.array_equal(hout["price"], out["price"].cpu().numpy(), equal_nan=True))

    tf_peak, _ = batch.measure_fp64_peak(20)
    mix_pps, _ = batch.measure_point_mix_peak(10)

    red = torch.tensor([ms_step, ms_e2e, ms_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e, ms_kernel = red.tolist()
    value = world * Nopt / (ms_step / 1e3)
    e2e_value = world * Nopt / (ms_e2e / 1e3)

    if rank == 0:
        fl = flops_per_option(cfgd["solver"], cfgd["n"], cfgd["l"], cfgd["m"], iters_mean)
        achieved = fl * Nopt / (ms_kernel / 1e3) / 1e12
        threads = len(os.sched_getaffinity(0))
        cpu_rate, cpu_count, cpu_dt, cpu_out = cpu_port_rate(cfgd, b, args.cpu_seconds, threads)
        # the CPU leg doubles as a live parity check of this very run
        p = out["price"].cpu().numpy()[:cpu_count]
        ok = np.isfinite(cpu_out["price"]) & (cpu_out["err"] < 1e-2)
        max_abs = float(np.max(np.abs(p - cpu_out["price"])[ok]))
        h2d = Nopt * (6 * 8 + 1)
        d2h = Nopt * 8
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "C2: 1M random puts/calls per GPU, FastAmericanOptionSolverA n=12 l=32 m=64 tol=1e-5 "
                                       "max_iters=20 (BASELINE.json configs[1])",
                           "options_per_gpu": Nopt, "seed": wl.C2_SEED, "parallelism": "dp%d independent slices, no collective" % world,
                           "l2": "flushed (256 MB write) between timed iterations; inputs 49 MB/GPU",
                           "mean_iterations": iters_mean, "kernel": "auto" if args.kernel == 0 else "generic"},
                "max_abs_err_vs_cpu_port": max_abs, "parity_sample": int(ok.sum()),
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": achieved / tf_peak,
                             # dram__bytes_read+write of the dominant kernel, per launch: 153.5 B/option measured by
                             # ncu --set full (profiles/r1_alo_fast_A12_ncu_full.csv) == the algorithmic 49 B in + 104 B seeds
                             "traffic": 153.5 * Nopt if args.kernel == 0 else None,
                             "kernel": "alo_fast_A12_kernel" if args.kernel == 0 else "alo_generic_kernel",
                             "kernel_ms": ms_kernel, "flops_per_option": fl,
                             "flop_model": "SURVEY 8d nominal: add/mul 1, FMA 2, div/sqrt 16, exp 40, log 56, Phi 120, phi 42 "
                                           "at the realised mean iteration count",
                             # hardware-side view of the same launch: DFMA x2 + DMUL + DADD thread instructions per option
                             # counted by ncu (profiles/r1_alo_fast_A12_ncu_full.csv, 1.131 MFLOP at 19.32 sweeps), scaled
                             # to this run's sweep count; the nominal model prices exp/Phi/sqrt as 40/120/16 flops where
                             # the kernel spends 24/50/12, which is why `frac` can exceed 1
                             "executed": None if args.kernel != 0 else {
                                 "flops_per_option": 1.131e6 * iters_mean / 19.316,
                                 "tflops": 1.131e6 * iters_mean / 19.316 * Nopt / (ms_kernel / 1e3) / 1e12,
                                 "frac": 1.131e6 * iters_mean / 19.316 * Nopt / (ms_kernel / 1e3) / 1e12 / tf_peak,
                                 "fp64_pipe_active_pct_ncu": 70.7,
                                 # what actually bounds the kernel: an FP64 warp instruction holds the SMSP issue
                                 # port for two cycles, any other for one.  21.34 k FP64 + 15.12 k other warp instructions per
                                 # option at 19.32 sweeps (ncu source page) over this run's cycles per option and SMSP
                                 "issue_port_frac": (2 * 21335.4 + 15116.2) * iters_mean / 19.316 /
                                                    (ms_kernel * 1e-3 * (clocks["sm_mhz"] or 1965.0) * 1e6 * kSMSP / Nopt),
                                 # quadrature points/s of this launch (options * sweeps * n * l) over the rate of the same
                                 # per-point math on register operands alone, measured in this run (faop_measure_point_mix_peak):
                                 # what the loads, reductions, node closures, seeds and the price integral cost on top
                                 "points_per_s": Nopt * iters_mean * cfgd["n"] * cfgd["l"] / (ms_kernel / 1e3),
                                 "point_math_peak_points_per_s": mix_pps,
                                 "point_math_frac": Nopt * iters_mean * cfgd["n"] * cfgd["l"] / (ms_kernel / 1e3) / mix_pps},
                             "peak_source": "register-operand DFMA chains measured in this run (faop_measure_fp64_peak); "
                                            "MEASURED_PEAKS.json has no FP64 figure"},
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": "first %d options of rank 0's batch, %.1f s, plain-C port of the reference "
                                           "algorithm on %d pthreads" % (cpu_count, cpu_dt, threads)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e, "bitwise_equal_to_device_path": same},
                "gpu_launches": int(launches), "gpu_launches_e2e": e2e_launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
