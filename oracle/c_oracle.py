"""ctypes binding of the plain-C oracle (oracle/alo_oracle.c).  TEST INFRASTRUCTURE ONLY — see alo_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np
from numpy.polynomial.legendre import leggauss

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "alo_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def alo_price_batch(solver, r, q, sigma, K, T, S, is_put, t=None, n=12, l=24, m=48, iter_tol=1e-5, max_iters=200,
                    use_derivative=False, eta=0.5, B0=None, nthreads=None, seed="hybr", rule_l=None, rule_m=None, tau_max=None):
    """seed: "hybr" (the reference's root finder, default) or "newton"; rule_l / rule_m: (y, w) on [-1, 1] replacing
    leggauss(l) / leggauss(m) (sizes l, m); tau_max: per-option end of the collocation domain (Base.py:28), default T."""
    r, q, sigma, K, T, S = (_d(np.atleast_1d(a)) for a in (r, q, sigma, K, T, S))
    N = r.shape[0]
    ip = np.ascontiguousarray(np.atleast_1d(is_put), dtype=np.uint8)
    t = np.zeros(N) if t is None else _d(np.broadcast_to(np.asarray(t, dtype=float), (N,)))
    yl, wl = leggauss(l) if rule_l is None else rule_l
    ym, wm = leggauss(m) if rule_m is None else rule_m
    assert len(yl) == l and len(ym) == m
    tm = None if tau_max is None else _d(np.broadcast_to(np.asarray(tau_max, dtype=float), (N,)).copy())
    out = dict(price=np.empty(N), european=np.empty(N), B=np.empty((N, n + 1)), B0=np.empty((N, n + 1)),
               iters=np.empty(N, dtype=np.int32), err=np.empty(N))
    B0c = None if B0 is None else _d(B0)
    if nthreads is None:
        nthreads = len(os.sched_getaffinity(0))
    rc = lib().oracle_alo_price_batch(
        ctypes.c_int(0 if solver == "A" else 1), ctypes.c_int(int(use_derivative)), ctypes.c_int(n), ctypes.c_int(l),
        ctypes.c_int(m), ctypes.c_int(max_iters), ctypes.c_double(iter_tol), ctypes.c_double(eta), ctypes.c_int64(N),
        _p(r), _p(q), _p(sigma), _p(K), _p(T), _p(t), _p(S), _p(ip), _p(_d(yl)), _p(_d(wl)), _p(_d(ym)), _p(_d(wm)),
        _p(B0c), _p(out["price"]), _p(out["european"]), _p(out["B"]), _p(out["B0"]), _p(out["iters"]), _p(out["err"]),
        ctypes.c_int(nthreads), ctypes.c_int(1 if seed == "hybr" else 0), _p(tm))
    assert rc == 0
    return out


def qd_boundary_batch(tau, r, q, sigma, K, is_put, seed="hybr", return_nfev=False):
    tau, r, q, sigma, K = (_d(a) for a in np.broadcast_arrays(tau, r, q, sigma, K))
    ip = np.ascontiguousarray(np.broadcast_to(is_put, tau.shape), dtype=np.uint8)
    B = np.empty(tau.shape)
    nf = np.zeros(tau.shape, dtype=np.int32)
    lib().oracle_qd_boundary_batch(ctypes.c_int64(tau.size), _p(r), _p(q), _p(sigma), _p(K), _p(tau), _p(ip), _p(B),
                                   ctypes.c_int(1 if seed == "hybr" else 0), _p(nf))
    return (B, nf) if return_nfev else B


def european_batch(tau, S, r, q, sigma, K, is_put):
    tau, S, r, q, sigma, K = (_d(a) for a in np.broadcast_arrays(tau, S, r, q, sigma, K))
    ip = np.ascontiguousarray(np.broadcast_to(is_put, tau.shape), dtype=np.uint8)
    out = np.empty(tau.shape)
    lib().oracle_european_batch(ctypes.c_int64(tau.size), _p(tau), _p(S), _p(r), _p(q), _p(sigma), _p(K), _p(ip), _p(out))
    return out


def cn_put_batch(r, q, sigma, K, T, S0, n_space, max_dt, mode, omega=1.2, tol=1e-5, max_iter=200, return_X=False,
                 Smax=None, return_state=False):
    """mode 0 = unprojected direct solve, 1 = reference SOR+project, 2 = projected Thomas.  Smax (default 3K) is the
    CNOptionSolver.Smax attribute (CrankNicolsonOptionSolver.py:12,48); return_state adds (err, iter) of the last SOR solve."""
    from .alo_oracle import cn_dt_schedule
    r, q, sigma, K, T, S0, max_dt = (_d(np.atleast_1d(a)) for a in np.broadcast_arrays(r, q, sigma, K, T, S0, max_dt))
    N = r.shape[0]
    sched = [cn_dt_schedule(T[i], max_dt[i]) for i in range(N)]
    ms = max(len(s) for s in sched)
    dts = np.zeros((N, ms))
    ns = np.zeros(N, dtype=np.int32)
    for i, s in enumerate(sched):
        dts[i, :len(s)] = s
        ns[i] = len(s)
    price = np.empty(N)
    X = np.empty((N, n_space)) if return_X else None
    sm = None if Smax is None else _d(np.broadcast_to(np.asarray(Smax, dtype=np.float64), (N,)).copy())
    err = np.zeros(N)
    it = np.zeros(N, dtype=np.int32)
    lib().oracle_cn_put_batch(ctypes.c_int64(N), ctypes.c_int(n_space), ctypes.c_int(mode), _p(r), _p(q), _p(sigma), _p(K),
                              _p(sm), _p(S0), _p(dts), _p(ns), ctypes.c_int(ms), ctypes.c_double(omega), ctypes.c_double(tol),
                              ctypes.c_int(max_iter), _p(price), _p(X), _p(err), _p(it))
    out = (price, X) if return_X else price
    if return_state:
        return (out, err, it)
    return out


def seed_conditioning(T, r, q, sigma, K, is_put, n, tau_max=None):
    """How far the reference's own QD seed (hybr from x0 = K at every collocation node) moves when the normal cdf inside its
    residual changes by 1-2 ulp: max over nodes and over probes k = +-1, +-2 of |B0_k / B0_0 - 1|, per option.  A few 1e-16
    where the residual is well conditioned; up to ~1e-6 where K - S - p cancels catastrophically (tiny maturities with
    X = rK/q << K) — there the reference's seed is an accident of its own rounding and no other implementation (scipy on another
    CPU included) can reproduce it better than this number."""
    T, r, q, sigma, K = (_d(np.atleast_1d(a)) for a in (T if tau_max is None else tau_max, r, q, sigma, K))
    c = np.cos(np.arange(n + 1) * np.pi / n)
    tau = np.square(c + 1)[None, :] * T[:, None] / 4
    a = (tau, r[:, None], q[:, None], sigma[:, None], K[:, None], np.atleast_1d(is_put).astype(bool)[:, None])
    try:
        base = qd_boundary_batch(*a)
        spread = np.zeros(T.shape[0])
        for k in (1, -1, 2, -2):
            lib().oracle_set_seed_probe(ctypes.c_int(k))
            with np.errstate(all="ignore"):
                spread = np.fmax(spread, np.nanmax(np.abs(qd_boundary_batch(*a) / base - 1), axis=1))
    finally:
        lib().oracle_set_seed_probe(ctypes.c_int(0))
    return spread


SEED_PROBES = (1, -1, 2, -2, 3, -3, 4, -4, 6, -6, 8, -8)


def seed_sensitivity(solver, r, q, sigma, K, T, S, is_put, probes=SEED_PROBES, **kw):
    """What the reference's own seed noise is worth in its FINAL answer: the whole solve (hybrd seed, damped sweeps, price
    integral) is repeated with the normal cdf inside the QD residual scaled by (1 + k ulp), k in `probes`; returns per option
    (max_k |B_k / B_0 - 1| over the nodes, max_k |price_k - price_0|, any sweep-count change).  This is the band inside which two
    correct implementations of the reference (scipy's ndtr vs libm's or libdevice's erfc — or scipy on another CPU) legitimately
    differ: zero-ish for well-conditioned options, and up to 1e-8 where hybrd stops on a flat residual after few sweeps or
    bifurcates between two slow-progress exits.  Meant for the handful of options that exceed the plain bars, not for whole batches."""
    base = alo_price_batch(solver, r, q, sigma, K, T, S, is_put, **kw)
    dB = np.zeros(len(base["price"]))
    dP = np.zeros(len(base["price"]))
    flip = np.zeros(len(base["price"]), dtype=bool)
    try:
        for k in probes:
            lib().oracle_set_seed_probe(ctypes.c_int(k))
            o = alo_price_batch(solver, r, q, sigma, K, T, S, is_put, **kw)
            with np.errstate(all="ignore"):
                dB = np.fmax(dB, np.nan_to_num(np.nanmax(np.abs(o["B"] / base["B"] - 1), axis=1), nan=np.inf))
                dP = np.fmax(dP, np.nan_to_num(np.abs(o["price"] - base["price"]), nan=np.inf))
            flip |= o["iters"] != base["iters"]
    finally:
        lib().oracle_set_seed_probe(ctypes.c_int(0))
    return dB, dP, flip
