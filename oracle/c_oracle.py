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
                    use_derivative=False, eta=0.5, B0=None, nthreads=None):
    r, q, sigma, K, T, S = (_d(np.atleast_1d(a)) for a in (r, q, sigma, K, T, S))
    N = r.shape[0]
    ip = np.ascontiguousarray(np.atleast_1d(is_put), dtype=np.uint8)
    t = np.zeros(N) if t is None else _d(np.broadcast_to(np.asarray(t, dtype=float), (N,)))
    yl, wl = leggauss(l)
    ym, wm = leggauss(m)
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
        ctypes.c_int(nthreads))
    assert rc == 0
    return out


def qd_boundary_batch(tau, r, q, sigma, K, is_put):
    tau, r, q, sigma, K = (_d(a) for a in np.broadcast_arrays(tau, r, q, sigma, K))
    ip = np.ascontiguousarray(np.broadcast_to(is_put, tau.shape), dtype=np.uint8)
    B = np.empty(tau.shape)
    lib().oracle_qd_boundary_batch(ctypes.c_int64(tau.size), _p(r), _p(q), _p(sigma), _p(K), _p(tau), _p(ip), _p(B))
    return B


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
