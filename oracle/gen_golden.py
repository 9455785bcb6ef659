#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (test infrastructure).

Runs only in the development container: it imports the reference's Python
modules from /root/reference (read-only) behind a no-op `matplotlib` stub
(the reference imports pyplot at module scope, FastAmericanOptionSolverBase.py:8,
CrankNicolsonOptionSolver.py:2; nothing numerical touches it).  The reference
has no golden vectors or asserting tests of its own, so these
outputs ARE the pin for the oracle and for the CUDA path.  Outputs go to
tests/golden/*.npz and are committed; /root/reference is never read by tests,
smoke() or bench.py.

Usage:  python oracle/gen_golden.py [--only NAME[,NAME...]] [--procs P]
Recorded library versions are stored in every fixture (`versions`).
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FAOP_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")

_STUB = '''
import sys, types
class _N:
    def __getattr__(self, k): return _N()
    def __call__(self, *a, **k): return _N()
for _sub in ("pyplot", "cm", "colors"):
    _m = types.ModuleType("matplotlib." + _sub)
    def _ga(k):
        if k.startswith("__"):      # inspect / importlib probe dunders (e.g. __file__) on every module in sys.modules
            raise AttributeError(k)
        return _N()
    _m.__getattr__ = _ga
    sys.modules["matplotlib." + _sub] = _m
    globals()[_sub] = _m
'''


def _install_reference_on_path():
    d = tempfile.mkdtemp(prefix="faop_mplstub_")
    os.makedirs(os.path.join(d, "matplotlib"))
    with open(os.path.join(d, "matplotlib", "__init__.py"), "w") as f:
        f.write(_STUB)
    sys.path[:0] = [d, REF]


_install_reference_on_path()
sys.path.insert(0, REPO)
warnings.simplefilter("ignore")
np.seterr(all="ignore")

import scipy  # noqa: E402
import FastAmericanOptionSolverA as refA  # noqa: E402
import FastAmericanOptionSolverB as refB  # noqa: E402
import QDplusAmericanOptionSolver as refqd  # noqa: E402
import EuropeanOptionSolver as refeu  # noqa: E402
import ChebyshevInterpolation as refcheb  # noqa: E402
import CrankNicolsonOptionSolver as refcn  # noqa: E402
from fastamericanoptionpricing_b200 import workloads as wl  # noqa: E402

VERSIONS = "numpy %s scipy %s python %s" % (np.__version__, scipy.__version__, sys.version.split()[0])


def _otype(is_put):
    return refqd.OptionType.Put if is_put else refqd.OptionType.Call


def solve_one(args):
    """One reference solve. args = (solver, r,q,sigma,K,T,S,t,is_put, n,l,m,tol,max_iters,deriv)."""
    tau_max = tau_min = None
    if len(args) == 17:            # gen_domain: the tau_max / tau_min attributes (Base.py:28-29)
        args, (tau_max, tau_min) = args[:15], args[15:]
    (solver, r, q, sigma, K, T, S, t, is_put, n, l, m, tol, max_iters, deriv) = args
    cls = refA.FastAmericanOptionSolverA if solver == "A" else refB.FastAmericanOptionSolverB
    s = cls(float(r), float(q), float(sigma), float(K), float(T), _otype(is_put))
    s.DEBUG = False
    s.collocation_num = int(n)
    s.quadrature_num = int(l)
    s.integration_num = int(m)
    s.iter_tol = float(tol)
    s.max_iters = int(max_iters)
    s.use_derivative = bool(deriv)
    if tau_max is not None:
        s.tau_max = float(tau_max)
    if tau_min is not None:
        s.tau_min = float(tau_min)
    t0 = time.perf_counter()
    price = float(s.solve(float(t), float(S)))
    dt = time.perf_counter() - t0
    nn = int(n) + 1
    B = np.full(nn, np.nan)
    B0 = np.full(nn, np.nan)
    tau = np.full(nn, np.nan)
    if len(s.shared_B) == nn:
        B[:] = np.asarray(s.shared_B, dtype=float)
        B0[:] = np.asarray(s.shared_B0, dtype=float)
        tau[:] = np.asarray(s.shared_tau, dtype=float)
    errs = np.full(int(max_iters), np.nan)
    for k, e in s.iter_records:
        errs[k - 1] = e
    return dict(price=price, european=float(s.european_price), B=B, B0=B0, tau=tau,
                iters=int(s.num_iters), err=float(s.error), iter_errs=errs, seconds=dt)


def run_batch(pool, solver, b, cfg, idx):
    jobs = [(solver, b["r"][i], b["q"][i], b["sigma"][i], b["K"][i], b["T"][i], b["S"][i], b["t"][i],
             bool(b["is_put"][i]), cfg["n"], cfg["l"], cfg["m"], cfg["iter_tol"], cfg["max_iters"],
             cfg["use_derivative"]) for i in idx]
    res = pool.map(solve_one, jobs, chunksize=1)
    out = {k: np.array([r_[k] for r_ in res]) for k in res[0]}
    for k in ("r", "q", "sigma", "K", "T", "S", "t", "is_put"):
        out["in_" + k] = np.asarray(b[k])[idx]
    for k, v in cfg.items():
        out["cfg_" + k] = np.array(v)
    out["versions"] = np.array(VERSIONS)
    return out


def gen_c1(pool):
    c = wl.c1_testA()
    b = {k: np.array([c[k]]) for k in ("r", "q", "sigma", "K", "T", "S", "t", "is_put")}
    cfg = {k: c[k] for k in ("solver", "n", "l", "m", "iter_tol", "max_iters", "use_derivative")}
    out = run_batch(pool, "A", b, cfg, [0])
    np.savez(os.path.join(OUT, "ref_c1_testA.npz"), **out)
    print("c1:", repr(out["price"][0]), out["iters"][0], repr(out["err"][0]))


def gen_c2(pool, count=None):
    count = count or int(os.environ.get("FAOP_GOLDEN_C2", "4096"))   # first 4,096 options through the live reference
    b = wl.c2_batch()
    out = run_batch(pool, "A", b, wl.C2_CFG, list(range(count)))
    np.savez(os.path.join(OUT, "ref_c2_A.npz"), **out)
    print("c2: %d options, mean %.2f s/option, iters min/med/max %d/%d/%d, nan %d" % (
        count, out["seconds"].mean(), out["iters"].min(), np.median(out["iters"]), out["iters"].max(),
        np.isnan(out["price"]).sum()))


def gen_c3(pool, count=None):
    count = count or int(os.environ.get("FAOP_GOLDEN_C3", "96"))
    b = wl.c3_batch()
    out = run_batch(pool, "B", b, wl.C3_CFG, list(range(count)))
    out["in_category"] = b["category"][:count]
    np.savez(os.path.join(OUT, "ref_c3_B.npz"), **out)
    print("c3: %d options, mean %.2f s/option, iters min/med/max %d/%d/%d, nan %d" % (
        count, out["seconds"].mean(), out["iters"].min(), np.median(out["iters"]), out["iters"].max(),
        np.isnan(out["price"]).sum()))


def gen_misc(pool):
    """Mixed knobs at the class defaults n=12,l=24,m=48,max_iters=200: A/B x derivative x t!=0 x n,l,m."""
    g = np.random.default_rng(777)
    base = [
        # r, q, sigma, K, T, S, t, is_put
        (0.04, 0.01, 0.2, 100.0, 3.0, 80.0, 0.0, True),     # testA
        (0.04, 0.06, 0.2, 100.0, 3.0, 120.0, 0.0, False),   # call r<q (App. C)
        (0.04, 0.06, 0.2, 100.0, 3.0, 80.0, 0.0, True),     # put r<q -> X=rK/q
        (0.03, 0.07, 0.3, 90.0, 0.5, 85.0, 0.0, True),      # App. C derivative case
        (0.06, 0.02, 0.25, 100.0, 1.0, 110.0, 0.0, False),  # call r>q -> X=rK/q
        (0.05, 0.05, 0.3, 100.0, 2.0, 100.0, 0.0, True),    # r == q
        (0.05, 0.05, 0.3, 100.0, 2.0, 100.0, 0.0, False),   # r == q call
        (0.04, 0.01, 0.2, 100.0, 3.0, 80.0, 1.0, True),     # t = 1 (tau = 2, boundary still on [0,T])
        (0.04, 0.01, 0.2, 100.0, 3.0, 80.0, 3.0, True),     # t = T -> tau = 0
        (0.04, 0.01, 0.2, 100.0, 3.0, 50.0, 0.0, True),     # deep in the exercise region (no clamp)
        (0.04, 0.0, 0.2, 100.0, 3.0, 110.0, 0.0, False),    # q == 0 call -> European shortcut
        (0.04, 0.0, 0.2, 100.0, 3.0, 90.0, 0.0, True),      # q == 0 put
        (0.02, 0.03, 0.45, 120.0, 0.1, 118.0, 0.0, False),  # short maturity call
        (0.08, 0.0, 0.15, 60.0, 5.0, 55.0, 0.0, True),      # long maturity put, q=0
    ]
    for _ in range(6):
        base.append((g.uniform(.001, .1), g.uniform(0, .1), g.uniform(.1, .6), g.uniform(50, 150),
                     g.uniform(.05, 3), 0.0, 0.0, bool(g.random() < .5)))
        lst = list(base[-1])
        lst[5] = lst[3] * g.uniform(.5, 1.5)
        base[-1] = tuple(lst)
    variants = [
        ("A", 12, 24, 48, 1e-5, 200, False),
        ("B", 12, 24, 48, 1e-5, 200, False),
        ("A", 12, 24, 48, 1e-5, 200, True),
        ("B", 12, 24, 48, 1e-5, 200, True),     # reference -> NaN for most (App. B#7)
        ("A", 8, 16, 20, 1e-6, 50, False),
        ("A", 16, 40, 50, 1e-7, 60, False),
        ("B", 6, 10, 30, 1e-4, 40, False),
        ("A", 12, 32, 64, 1e-8, 30, False),
        ("A", 5, 7, 9, 1e-5, 10, False),
    ]
    jobs, meta = [], []
    for vi, (solver, n, l, m, tol, mi, deriv) in enumerate(variants):
        for bi, (r, q, sig, K, T, S, t, is_put) in enumerate(base):
            jobs.append((solver, r, q, sig, K, T, S, t, is_put, n, l, m, tol, mi, deriv))
            meta.append((vi, bi))
    res = pool.map(solve_one, jobs, chunksize=1)
    nmax = max(v[1] for v in variants) + 1
    imax = max(v[5] for v in variants)

    def pad(a, n_):
        o = np.full(n_, np.nan)
        o[:len(a)] = a
        return o
    out = dict(
        solver=np.array([j[0] for j in jobs]),
        in_r=np.array([j[1] for j in jobs]), in_q=np.array([j[2] for j in jobs]),
        in_sigma=np.array([j[3] for j in jobs]), in_K=np.array([j[4] for j in jobs]),
        in_T=np.array([j[5] for j in jobs]), in_S=np.array([j[6] for j in jobs]),
        in_t=np.array([j[7] for j in jobs]), in_is_put=np.array([j[8] for j in jobs]),
        n=np.array([j[9] for j in jobs]), l=np.array([j[10] for j in jobs]), m=np.array([j[11] for j in jobs]),
        iter_tol=np.array([j[12] for j in jobs]), max_iters=np.array([j[13] for j in jobs]),
        use_derivative=np.array([j[14] for j in jobs]),
        price=np.array([r_["price"] for r_ in res]), european=np.array([r_["european"] for r_ in res]),
        iters=np.array([r_["iters"] for r_ in res]), err=np.array([r_["err"] for r_ in res]),
        B=np.array([pad(r_["B"], nmax) for r_ in res]), B0=np.array([pad(r_["B0"], nmax) for r_ in res]),
        tau=np.array([pad(r_["tau"], nmax) for r_ in res]),
        iter_errs=np.array([pad(r_["iter_errs"], imax) for r_ in res]),
        versions=np.array(VERSIONS))
    np.savez(os.path.join(OUT, "ref_misc.npz"), **out)
    print("misc: %d solves, nan prices %d" % (len(jobs), np.isnan(out["price"]).sum()))


def _qd_root(args):
    r, q, sig, K, tau, is_put = args
    s = refqd.QDplus(r, q, sig, K, _otype(is_put))
    return float(s.compute_exercise_boundary(tau))


def _qd_price(args):
    r, q, sig, K, tau, S, is_put = args
    s = refqd.QDplus(r, q, sig, K, _otype(is_put))
    with contextlib.redirect_stdout(io.StringIO()):
        p = float(s.price(tau, S))
    return p, float(s.exercise_boundary)


def gen_qd(pool, count=2048):
    # reference test_QDplus_curve.py:7-19 parameters and its 13 tau
    r0, q0, K0, sig0 = 0.0975729097939295, 0.011804520625954162, 105.61782314803582, 0.2
    s = refqd.QDplus(r0, q0, sig0, K0, refqd.OptionType.Put)
    # the 13 tau literals of reference test_QDplus_curve.py:16-19
    tau_curve = np.array([3.00000000e+00, 2.89864827e+00, 2.61153811e+00, 2.18566017e+00, 1.68750000e+00, 1.18846904e+00,
                          7.50000000e-01, 4.12011906e-01, 1.87500000e-01, 6.43398282e-02, 1.34618943e-02, 8.70786986e-04,
                          0.00000000e+00])
    curve = np.array([s.compute_exercise_boundary(t_) for t_ in tau_curve])
    Sgrid = np.linspace(50.0, 150.0, 41)
    resid_put = np.asarray(s.exercise_boundary_func(Sgrid, 0.75))
    sc = refqd.QDplus(0.04, 0.06, 0.25, 100.0, refqd.OptionType.Call)
    resid_call = np.asarray(sc.exercise_boundary_func(Sgrid, 0.75))
    g = np.random.default_rng(4242)
    is_put = g.random(count) < 0.5
    r = g.uniform(0.001, 0.10, count)
    q = g.uniform(0.0, 0.10, count)
    sig = g.uniform(0.05, 0.60, count)
    K = g.uniform(50.0, 150.0, count)
    # mix of collocation-like taus (down to ~1e-4) and plain uniforms
    tau = np.where(g.random(count) < 0.5, g.uniform(0.01, 5.0, count),
                   3.0 * np.square(1 + np.cos(np.pi * g.integers(0, 12, count) / 12)) / 4 * g.uniform(0.02, 1, count))
    roots = np.array(pool.map(_qd_root, [(r[i], q[i], sig[i], K[i], tau[i], bool(is_put[i])) for i in range(count)]))
    npx = 64
    Sx = K[:npx] * g.uniform(0.6, 1.4, npx)
    pr = pool.map(_qd_price, [(r[i], q[i], sig[i], K[i], tau[i], Sx[i], bool(is_put[i])) for i in range(npx)])
    np.savez(os.path.join(OUT, "ref_qd.npz"), curve_r=r0, curve_q=q0, curve_K=K0, curve_sigma=sig0,
             curve_tau=tau_curve, curve_B=curve, resid_S=Sgrid, resid_tau=0.75, resid_put=resid_put,
             resid_call=resid_call, resid_call_params=np.array([0.04, 0.06, 0.25, 100.0]),
             in_r=r, in_q=q, in_sigma=sig, in_K=K, in_tau=tau, in_is_put=is_put, B=roots,
             price_S=Sx, price=np.array([p[0] for p in pr]), price_boundary=np.array([p[1] for p in pr]),
             test_QDplus_price=_qd_price((r0, q0, sig0, K0, 0.000870786986, 30.543317986992072, True))[0],   # test_QDplus.py:8-18
             versions=np.array(VERSIONS))
    print("qd: curve[0]=%r nan roots=%d" % (curve[0], np.isnan(roots).sum()))


def gen_european(pool, count=2048):
    g = np.random.default_rng(99)
    tau = g.uniform(0.0, 5.0, count)
    tau[:16] = 0.0
    S = g.uniform(20.0, 250.0, count)
    r = g.uniform(0.0, 0.1, count)
    q = g.uniform(0.0, 0.1, count)
    sig = g.uniform(0.02, 0.8, count)
    K = g.uniform(50.0, 150.0, count)
    E = refeu.EuropeanOption
    put = np.array([E.european_put_value(tau[i], S[i], r[i], q[i], sig[i], K[i]) for i in range(count)], dtype=float)
    call = np.array([E.european_call_value(tau[i], S[i], r[i], q[i], sig[i], K[i]) for i in range(count)], dtype=float)
    theta = np.array([E.european_option_theta(tau[i], S[i], r[i], q[i], sig[i], K[i]) for i in range(count)], dtype=float)
    with np.errstate(all="ignore"):
        d1 = np.array([E.d1(tau[i], S[i], r[i], q[i], sig[i], K[i]) for i in range(count)], dtype=float)
        d2 = np.array([E.d2(tau[i], S[i], r[i], q[i], sig[i], K[i]) for i in range(count)], dtype=float)
    np.savez(os.path.join(OUT, "ref_european.npz"), tau=tau, S=S, r=r, q=q, sigma=sig, K=K, put=put, call=call,
             theta=theta, d1=d1, d2=d2, versions=np.array(VERSIONS))
    print("european: %d" % count)


def gen_cheb(pool):
    g = np.random.default_rng(5)
    out = dict(versions=np.array(VERSIONS))
    for n in (2, 5, 12, 24, 40):
        nodes = refcheb.ChebyshevInterpolation.get_std_cheby_points(n)
        y = g.normal(size=n + 1)
        z = g.uniform(-1, 1, 50)
        c = refcheb.ChebyshevInterpolation(y, None, 0, 0)
        out["n%d_nodes" % n] = nodes
        out["n%d_y" % n] = y
        out["n%d_a" % n] = np.array(c.a)
        out["n%d_z" % n] = z
        out["n%d_val" % n] = np.array(c.value(z))
        # with the solver's sqrt map on [0, 2.5]  (FastAmericanOptionSolverBase.py:235-237)
        x = g.uniform(0, 2.5, 50)
        c2 = refcheb.ChebyshevInterpolation(y, lambda x_, a, b: np.sqrt(4 * (x_ - a) / (b - a)) - 1, 0.0, 2.5)
        out["n%d_x" % n] = x
        out["n%d_valx" % n] = np.array(c2.value(x))
    # test_cheb_cvg.py: interpolation error of x*exp(x) at n=2..20 on 200 pts (reference script's table)
    errs = []
    xs = np.linspace(-1, 1, 200)
    for n in range(2, 21):
        nodes = refcheb.ChebyshevInterpolation.get_std_cheby_points(n)
        c = refcheb.ChebyshevInterpolation(nodes * np.exp(nodes), None, 0, 0)
        errs.append(np.max(np.abs(np.array(c.value(xs)) - xs * np.exp(xs))))
    out["cvg_n"] = np.arange(2, 21)
    out["cvg_err"] = np.array(errs)
    np.savez(os.path.join(OUT, "ref_cheb.npz"), **out)
    print("cheb ok")


def gen_domain(pool):
    """The collocation-domain knobs tau_max / tau_min (FastAmericanOptionSolverBase.py:28-29, consumed at :185 and :223):
    tau_max above and below the maturity, with t != 0, Solver A and B, puts and calls; tau_min != 0 (upstream: NaN, 1 sweep)."""
    base = [
        (0.04, 0.01, 0.2, 100.0, 3.0, 80.0, 0.0, True),
        (0.04, 0.01, 0.2, 100.0, 3.0, 80.0, 1.0, True),
        (0.04, 0.06, 0.2, 100.0, 3.0, 120.0, 0.0, False),
        (0.03, 0.07, 0.3, 90.0, 0.5, 85.0, 0.1, True),
        (0.06, 0.02, 0.25, 100.0, 1.0, 110.0, 0.0, False),
        (0.05, 0.05, 0.3, 100.0, 2.0, 100.0, 0.5, True),
    ]
    jobs = []
    for (r, q, sig, K, T, S, t, put) in base:
        for solver, n, l, m in (("A", 12, 32, 64), ("B", 12, 24, 48), ("A", 16, 40, 50)):
            for mult in (1.5, 0.7, 1.0000001):
                jobs.append((solver, r, q, sig, K, T, S, t, put, n, l, m, 1e-5, 60, False, mult * T, 0.0))
    for (r, q, sig, K, T, S, t, put) in base[:3]:
        jobs.append(("A", r, q, sig, K, T, S, t, put, 12, 24, 48, 1e-5, 60, False, T, 0.1))
        jobs.append(("A", r, q, sig, K, T, S, t, put, 12, 24, 48, 1e-5, 60, False, T, -0.1))
    res = pool.map(solve_one, jobs, chunksize=1)
    nmax = max(j[9] for j in jobs) + 1

    def pad(a):
        o = np.full(nmax, np.nan)
        o[:len(a)] = a
        return o
    names = ("solver", "in_r", "in_q", "in_sigma", "in_K", "in_T", "in_S", "in_t", "in_is_put", "n", "l", "m", "iter_tol",
             "max_iters", "use_derivative", "tau_max", "tau_min")
    out = {k: np.array([j[i] for j in jobs]) for i, k in enumerate(names)}
    for k in ("price", "european", "iters", "err"):
        out[k] = np.array([r_[k] for r_ in res])
    for k in ("B", "B0", "tau"):
        out[k] = np.array([pad(r_[k]) for r_ in res])
    out["versions"] = np.array(VERSIONS)
    np.savez(os.path.join(OUT, "ref_domain.npz"), **out)
    print("domain: %d solves, nan prices %d (expected %d)" % (len(jobs), np.isnan(out["price"]).sum(), 6))


def _cn_one(args):
    r, q, sig, K, T, S0, N, max_dt, psor, tol, max_iter, omega = args
    s = refcn.CNOptionSolver(r, q, sig, K, T)
    s.N = N
    s.max_dt = max_dt
    s.USE_PSOR = psor
    s.tol = tol
    s.max_iter = max_iter
    s.omega = omega
    with contextlib.redirect_stdout(io.StringIO()):
        p = float(s.solve(S0))
    return p, np.asarray(s.X, dtype=float).flatten(), np.asarray(s.S, dtype=float).flatten()


def gen_cn(pool):
    g = np.random.default_rng(31)
    jobs = []
    base = [(0.04, 0.04, 0.2, 100.0, 3.0, 80.0), (0.04, 0.01, 0.2, 100.0, 3.0, 80.0)]
    for _ in range(6):
        base.append((g.uniform(.01, .08), g.uniform(0, .06), g.uniform(.15, .5), 100.0, g.uniform(.25, 3), g.uniform(70, 130)))
    for (r, q, sig, K, T, S0) in base:
        jobs.append((r, q, sig, K, T, S0, 200, 1 / 12.0, False, 1e-5, 200, 1.2))   # class defaults
        jobs.append((r, q, sig, K, T, S0, 200, 1 / 12.0, True, 1e-5, 200, 1.2))    # PSOR
        jobs.append((r, q, sig, K, T, S0, 64, T / 40, True, 1e-7, 400, 1.2))       # small grid, tighter SOR
        jobs.append((r, q, sig, K, T, S0, 301, T / 300, False, 1e-5, 200, 1.2))    # 301-step float countdown (App. B#14)
    res = pool.map(_cn_one, jobs, chunksize=1)
    Nmax = max(j[6] for j in jobs)
    X = np.full((len(jobs), Nmax), np.nan)
    for i, r_ in enumerate(res):
        X[i, :len(r_[1])] = r_[1]
    cols = list(zip(*jobs))
    np.savez(os.path.join(OUT, "ref_cn.npz"), in_r=np.array(cols[0]), in_q=np.array(cols[1]), in_sigma=np.array(cols[2]),
             in_K=np.array(cols[3]), in_T=np.array(cols[4]), in_S0=np.array(cols[5]), N=np.array(cols[6]),
             max_dt=np.array(cols[7]), use_psor=np.array(cols[8]), tol=np.array(cols[9]), max_iter=np.array(cols[10]),
             omega=np.array(cols[11]), price=np.array([r_[0] for r_ in res]), X=X, versions=np.array(VERSIONS))
    print("cn: %d solves; first two %r %r" % (len(jobs), res[0][0], res[1][0]))


def _cn_state_one(args):
    """Non-default Smax (CrankNicolsonOptionSolver.py:12 -> :48), err / iter after a PSOR solve (:103-104), and the per-step
    surface: setInitialCondition -> setCoeff(dt) (A's three diagonals + closure row, b) -> solveLinearSystem / solvePSOR."""
    r, q, sig, K, T, S0, N, max_dt, psor, smax_mult, step_dt = args
    s = refcn.CNOptionSolver(r, q, sig, K, T)
    s.N = N
    s.max_dt = max_dt
    s.USE_PSOR = psor
    s.Smax = smax_mult * K
    with contextlib.redirect_stdout(io.StringIO()):
        p = float(s.solve(S0))
    X = np.asarray(s.X, dtype=float).flatten()
    err, it = float(s.err), int(s.iter)
    # per-step surface on a fresh object
    u = refcn.CNOptionSolver(r, q, sig, K, T)
    u.N = N
    u.Smax = smax_mult * K
    u.setInitialCondition()
    u.setCoeff(step_dt)
    A = np.array(u.A, dtype=float)
    i = np.arange(N)
    di = A[i, i].copy()
    lo = np.zeros(N); lo[1:] = A[i[1:], i[1:] - 1]
    up = np.zeros(N); up[:-1] = A[i[:-1], i[:-1] + 1]
    last = A[-1, N - 4:].copy()
    b = np.array(u.b, dtype=float).flatten()
    if psor:
        u.solvePSOR()
    else:
        u.solveLinearSystem()
    X1 = np.asarray(u.X, dtype=float).flatten()
    # a second step from the new state (the PSOR path keeps a 1-D X; the dense path needs it flattened like solve() never does)
    return dict(price=p, X=X, err=err, iter=it, di=di, lo=lo, up=up, last=last, b=b, X1=X1, err1=float(u.err), iter1=int(u.iter))


def gen_cn_state(pool):
    g = np.random.default_rng(32)
    jobs = []
    base = [(0.04, 0.04, 0.2, 100.0, 3.0, 80.0), (0.04, 0.01, 0.2, 100.0, 3.0, 80.0)]
    for _ in range(4):
        base.append((g.uniform(.01, .08), g.uniform(0, .06), g.uniform(.15, .5), g.uniform(60, 140), g.uniform(.25, 3), g.uniform(70, 130)))
    for (r, q, sig, K, T, S0) in base:
        for psor in (False, True):
            for mult in (3.0, 4.5, 2.0, 7.25):
                jobs.append((r, q, sig, K, T, S0, 160, 1 / 12.0, psor, mult, 0.05))
    res = pool.map(_cn_state_one, jobs, chunksize=1)
    cols = list(zip(*jobs))
    out = dict(in_r=np.array(cols[0]), in_q=np.array(cols[1]), in_sigma=np.array(cols[2]), in_K=np.array(cols[3]),
               in_T=np.array(cols[4]), in_S0=np.array(cols[5]), N=np.array(cols[6]), max_dt=np.array(cols[7]),
               use_psor=np.array(cols[8]), smax_mult=np.array(cols[9]), step_dt=np.array(cols[10]), versions=np.array(VERSIONS))
    for k in res[0]:
        out[k] = np.array([r_[k] for r_ in res])
    np.savez(os.path.join(OUT, "ref_cn_state.npz"), **out)
    print("cn_state: %d solves; price[0..3] %r" % (len(jobs), [r_["price"] for r_ in res[:4]]))


def _stage_one(args):
    solver, r, q, sig, K, T, is_put, n, l, m = args
    cls = refA.FastAmericanOptionSolverA if solver == "A" else refB.FastAmericanOptionSolverB
    s = cls(r, q, sig, K, T, _otype(is_put))
    s.DEBUG = False
    s.collocation_num, s.quadrature_num, s.integration_num = n, l, m
    s.use_derivative = True          # so that compute_f_and_fprime returns the analytic f'
    s.set_collocation_points()
    s.set_initial_guess()
    tau = list(s.shared_tau)
    B = list(s.shared_B)
    Nv, Dv, f, fp, Np, Dp = [], [], [], [], [], []
    for ti, Bi in zip(tau, B):
        if ti == 0:
            Nv.append(np.nan); Dv.append(np.nan); Np.append(np.nan); Dp.append(np.nan)
        else:
            Nv.append(s.N_func(ti, Bi)); Dv.append(s.D_func(ti, Bi))
            Np.append(s.Nprime_func(ti, Bi)); Dp.append(s.Dprime_func(ti, Bi))
        ff = s.compute_f_and_fprime(ti, Bi)
        f.append(ff[0]); fp.append(ff[1])
    # price integral on the seed boundary at three spots and tau = T and T/2
    v = []
    for S0 in (0.8 * K, K, 1.2 * K):
        for tt in (T, 0.5 * T):
            v.append(s.american_value_with_known_boundary(tt, S0, r, q, sig, K))
    match2 = s.check_value_match_condition2()
    return dict(tau=np.array(tau), B=np.array(B), N=np.array(Nv, dtype=float), D=np.array(Dv, dtype=float),
                Np=np.array(Np, dtype=float), Dp=np.array(Dp, dtype=float), f=np.array(f, dtype=float),
                fp=np.array(fp, dtype=float), v=np.array(v, dtype=float), match2=float(match2))


def gen_stage(pool):
    """Stage-wise values on the QD seed boundary: N, D, N', D', f, f' per node, price integral, match condition."""
    g = np.random.default_rng(2718)
    jobs = []
    for solver in ("A", "B"):
        for i in range(12):
            is_put = bool(i % 2)
            r = g.uniform(.005, .1); q = g.uniform(.005, .1); sig = g.uniform(.1, .5)
            jobs.append((solver, r, q, sig, g.uniform(50, 150), g.uniform(.1, 3), is_put, 12, 24, 48))
    res = pool.map(_stage_one, jobs, chunksize=1)
    cols = list(zip(*jobs))
    out = {k: np.array([r_[k] for r_ in res]) for k in res[0]}
    out.update(solver=np.array(cols[0]), in_r=np.array(cols[1]), in_q=np.array(cols[2]), in_sigma=np.array(cols[3]),
               in_K=np.array(cols[4]), in_T=np.array(cols[5]), in_is_put=np.array(cols[6]), n=np.array(cols[7]),
               l=np.array(cols[8]), m=np.array(cols[9]), versions=np.array(VERSIONS))
    np.savez(os.path.join(OUT, "ref_stage.npz"), **out)
    print("stage: %d" % len(jobs))


def _bs_module(Kg, Tg):
    """The reference module executed with the globals `K`, `T` that computeExerciseBoundary reads
    (BjerksundStenslandSolver.py:40,42); its __main__ block is not run."""
    import types
    src = open(os.path.join(REF, "BjerksundStenslandSolver.py")).read()
    mod = types.ModuleType("BjerksundStenslandSolver_ref")
    mod.__dict__.update(K=Kg, T=Tg)
    exec(compile(src, "BjerksundStenslandSolver.py", "exec"), mod.__dict__)
    return mod


def gen_bs(pool, count=1024):
    g = np.random.default_rng(77)
    r = g.uniform(0.005, 0.10, count)
    q = g.uniform(0.005, 0.10, count)
    sig = g.uniform(0.08, 0.6, count)
    K = g.uniform(50.0, 150.0, count)
    T = g.uniform(0.05, 5.0, count)
    S = K * g.uniform(0.5, 1.4, count)
    # globals as the object's own strike / maturity (what the class means) ...
    price = np.empty(count); X = np.empty(count); alpha = np.empty(count); phis = np.empty((count, 6))
    for i in range(count):
        m = _bs_module(float(K[i]), float(T[i]))
        o = m.BksdStld(float(r[i]), float(q[i]), float(sig[i]), float(K[i]), float(T[i]))
        price[i] = o.price(float(S[i]))
        X[i] = o.computeExerciseBoundary()
        alpha[i] = o.computeAlpha(X[i])
        phis[i] = [o.phi(S[i], T[i], o.beta, X[i], X[i]), o.phi(S[i], T[i], 1, X[i], X[i]), o.phi(S[i], T[i], 1, K[i], X[i]),
                   o.phi(S[i], T[i], 0, X[i], X[i]), o.phi(S[i], T[i], 0, K[i], X[i]), o.beta]
    # ... and foreign globals (the __main__ put-by-symmetry line has K_global = 100 != self.K = 80, :52-53)
    Kg = g.uniform(50.0, 150.0, 128); Tg = g.uniform(0.05, 5.0, 128)
    price_g = np.array([_bs_module(float(Kg[i]), float(Tg[i])).BksdStld(float(r[i]), float(q[i]), float(sig[i]), float(K[i]),
                                                                         float(T[i])).price(float(S[i])) for i in range(128)])
    # the module's own __main__ numbers (:46-57)
    m = _bs_module(100, 3.0)
    main_call = m.BksdStld(0.04, 0.04, 0.2, 100, 3.0).price(80)
    main_put = m.BksdStld(0.04, 0.04, 0.2, 80, 3.0).price(100)
    np.savez(os.path.join(OUT, "ref_bs.npz"), r=r, q=q, sigma=sig, K=K, T=T, S=S, price=price, X=X, alpha=alpha, phis=phis,
             Kg=Kg, Tg=Tg, price_g=price_g, main_call=main_call, main_put=main_put, versions=np.array(VERSIONS))
    print("bs: %d, main call %r put %r, nan %d" % (count, main_call, main_put, np.isnan(price).sum()))


def gen_c5(pool, count=None):
    """Config 5 (surface sweep): a seeded sample of flat indices of the 1000 x 1000 x 100 grid through the live reference,
    plus whole strike ladders of a few (T, sigma) scenarios (the deduplicated path's unit of work)."""
    count = count or int(os.environ.get("FAOP_GOLDEN_C5", "256"))
    g = np.random.default_rng(55)
    flat = np.sort(g.choice(100_000_000, size=count, replace=False))
    ladders = []
    for iT, iV in ((3, 5), (500, 50), (999, 99)):
        for iK in range(0, 1000, 37):
            ladders.append((iK * 1000 + iT) * 100 + iV)
    idx = np.concatenate([flat, np.array(ladders, dtype=np.int64)])
    parts = [wl.c5_surface(int(i), 1) for i in idx]
    b = {k: np.concatenate([p_[k] for p_ in parts]) for k in parts[0]}
    out = run_batch(pool, "A", b, wl.C5_CFG, list(range(len(idx))))
    out["flat_index"] = idx
    np.savez(os.path.join(OUT, "ref_c5_surface.npz"), **out)
    print("c5: %d options, iters min/med/max %d/%d/%d, nan %d" % (len(idx), out["iters"].min(), np.median(out["iters"]),
                                                                 out["iters"].max(), np.isnan(out["price"]).sum()))


def gen_stress(pool):
    """Hard corners through the live reference (Solver A, n=12, l=32, m=64, tol 1e-5, 40 sweeps): a slice of every block of
    workloads.stress_batch — deep ITM/OTM, tiny maturities (the ill-conditioned QD seeds), r == q, q == 0, r ~ 0."""
    b = wl.stress_batch(6000)
    b["t"] = np.zeros(6000)
    idx = [i for lo in (0, 500, 1000, 1500, 1800, 2000, 2100) for i in range(lo, lo + 12)] + list(range(1400, 1436))
    cfg = dict(wl.C2_CFG, max_iters=40)
    out = run_batch(pool, "A", b, cfg, idx)
    out["index"] = np.array(idx)
    np.savez(os.path.join(OUT, "ref_stress_A.npz"), **out)
    print("stress: %d options, nan %d, err>=1e-2 %d" % (len(idx), np.isnan(out["price"]).sum(), (out["err"] >= 1e-2).sum()))


def gen_seedcases(pool):
    """Where the seed's root finder matters (round 2): options picked by the ORACLE's own measures, run through the live reference.
    (a) C2 options whose hybrd iterate is far from the Newton root at some node (hybrd exits on slow progress; upstream then turns to
    NaN in the second sweep or converges from a different seed), (b) the worst-conditioned seeds of the stress batch, (c) a random
    slice of the stress batch.  Solver A, n=12, l=32, m=64, tol 1e-5, 40 sweeps."""
    from oracle import c_oracle as co
    n = 12
    c = np.cos(np.arange(n + 1) * np.pi / n)
    picks = []
    b2 = wl.c2_batch(200_000)
    tau = np.square(c + 1)[None, :] * b2["T"][:, None] / 4
    a2 = (tau, b2["r"][:, None], b2["q"][:, None], b2["sigma"][:, None], b2["K"][:, None], b2["is_put"][:, None])
    with np.errstate(all="ignore"):
        far = np.nanmax(np.abs(co.qd_boundary_batch(*a2, seed="hybr") / co.qd_boundary_batch(*a2, seed="newton") - 1), axis=1)
    ia = np.argsort(-np.nan_to_num(far))[:48]
    bs = wl.stress_batch(6000)
    cond = co.seed_conditioning(bs["T"], bs["r"], bs["q"], bs["sigma"], bs["K"], bs["is_put"], n)
    ib = np.argsort(-np.nan_to_num(cond))[:72]
    ic = np.random.default_rng(99).choice(6000, 120, replace=False)
    keys = ("r", "q", "sigma", "K", "T", "S", "is_put")
    b = {k: np.concatenate([b2[k][ia], bs[k][ib], bs[k][ic]]) for k in keys}
    b["t"] = np.zeros(len(b["r"]))
    cfg = dict(wl.C2_CFG, max_iters=40)
    out = run_batch(pool, "A", b, cfg, list(range(len(b["r"]))))
    out["group"] = np.concatenate([np.zeros(len(ia), dtype=int), np.ones(len(ib), dtype=int), np.full(len(ic), 2)])
    out["source_index"] = np.concatenate([ia, ib, ic])
    np.savez(os.path.join(OUT, "ref_seedcases_A.npz"), **out)
    print("seedcases: %d options, nan %d, err>=1e-2 %d, hybr-vs-newton seed distance of group a: min %.1e" % (
        len(b["r"]), np.isnan(out["price"]).sum(), (out["err"] >= 1e-2).sum(), far[ia].min()))


GENS = dict(c1=gen_c1, european=gen_european, cheb=gen_cheb, qd=gen_qd, stage=gen_stage, misc=gen_misc, cn=gen_cn, cn_state=gen_cn_state, domain=gen_domain,
            c3=gen_c3, c2=gen_c2, bs=gen_bs, c5=gen_c5, stress=gen_stress, seedcases=gen_seedcases)


def main():
    import multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    names = [x for x in a.only.split(",") if x] or list(GENS)
    with mp.Pool(a.procs) as pool:
        for nme in names:
            t0 = time.time()
            GENS[nme](pool)
            print("  [%s done in %.1f s]" % (nme, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
