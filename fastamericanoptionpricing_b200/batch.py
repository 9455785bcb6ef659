"""Batch API: PyTorch tensors in, PyTorch tensors out, CUDA through the C ABI (include/faop.h).

PyTorch is plumbing here — device memory, streams, torch.distributed — and nothing else: every
number is produced by a hand-written sm_100a kernel in csrc/.  All functions take 1-D structure-of-
arrays inputs (anything torch.as_tensor accepts; moved to the current CUDA device as float64) and
return CUDA tensors.  They raise if CUDA or libfaop.so is unavailable; there is no CPU path.
"""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch
from numpy.polynomial.legendre import leggauss

from . import _lib
from ._lib import FaopCfg, FaopSurface, check


@dataclass(frozen=True)
class AloConfig:
    """Solver knobs; field names follow the reference's attributes
    (FastAmericanOptionSolverBase.py:19-23, 46-47): collocation_num, quadrature_num, integration_num."""
    solver: str = "A"
    n: int = 12            # collocation_num
    l: int = 24            # quadrature_num
    m: int = 48            # integration_num (frozen at 2*24 in the reference constructor)
    iter_tol: float = 1e-5
    max_iters: int = 200
    use_derivative: bool = False
    eta: float = 0.5
    kernel: int = 0        # 0 = auto (specialised kernel when available), 1 = generic kernel
    # opt-in accuracy modes named in the north-star but absent from the reference (they change results):
    rule_l: str = "gauss-legendre"   # or "tanh-sinh": rule of the boundary-equation integrals (l points)
    rule_m: str = "gauss-legendre"   # or "tanh-sinh": rule of the price integral (m points)
    # root finder of the QD seed: "hybr" = the reference's (MINPACK hybrd as scipy.optimize.root drives it,
    # QDplusAmericanOptionSolver.py:73), "newton" = safeguarded Newton, the faster opt-in (FAOP_FLAG_SEED_NEWTON)
    seed: str = "hybr"

    def c(self, flags=0, tau_max=None):
        if self.seed not in ("hybr", "newton"):
            raise ValueError("seed must be 'hybr' or 'newton'")
        if self.seed == "newton":
            flags |= _lib.FLAG_SEED_NEWTON
        return FaopCfg(0 if self.solver == "A" else 1, int(bool(self.use_derivative)), int(self.n), int(self.l),
                       int(self.m), int(self.max_iters), float(self.iter_tol), float(self.eta), int(self.kernel), int(flags),
                       None if tau_max is None else tau_max.data_ptr())


def _device(device=None):
    if not torch.cuda.is_available():
        raise _lib.FaopError("CUDA device required: fastamericanoptionpricing_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _f64(x, dev, N=None):
    t = torch.as_tensor(x, dtype=torch.float64)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    t = t.reshape(-1) if t.dim() != 1 else t
    if N is not None and t.numel() == 1 and N != 1:
        t = t.expand(N)
    return t.contiguous()


def _u8(x, dev, N=None):
    t = torch.as_tensor(x)
    t = t.to(device=dev, dtype=torch.uint8, non_blocking=True).reshape(-1)
    if N is not None and t.numel() == 1 and N != 1:
        t = t.expand(N)
    return t.contiguous()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_TABLES = {}


_GL_CACHE = {}


def gauss_legendre(num):
    """The reference's quadrature rule: numpy.polynomial.legendre.leggauss (FastAmericanOptionSolverBase.py:175).
    Cached per size (an eigenvalue solve each time otherwise); the returned arrays are read-only."""
    num = int(num)
    hit = _GL_CACHE.get(num)
    if hit is None:
        y, w = leggauss(num)
        y, w = np.ascontiguousarray(y, dtype=np.float64), np.ascontiguousarray(w, dtype=np.float64)
        y.setflags(write=False)
        w.setflags(write=False)
        hit = _GL_CACHE[num] = (y, w)
    return hit


def tanh_sinh(num, t_max=3.16):
    """Double-exponential (tanh-sinh) rule on [-1, 1] with `num` points: y = tanh(pi/2 sinh t), w = h pi/2 cosh t /
    cosh^2(pi/2 sinh t) on the uniform grid t = (k - (num-1)/2) h, h = 2 t_max / (num - 1); t_max = 3.16 puts the outermost
    nodes 1e-16 from the end points, where the weights have decayed below 1e-15.  Not in the reference."""
    num = int(num)
    if num < 2:
        return np.zeros(1), np.full(1, 2.0)
    h = 2.0 * t_max / (num - 1)
    t = (np.arange(num) - 0.5 * (num - 1)) * h
    u = 0.5 * np.pi * np.sinh(t)
    y = np.tanh(u)
    w = h * 0.5 * np.pi * np.cosh(t) / np.cosh(u) ** 2
    y = np.clip(y, -1.0 + 2.0 ** -52, 1.0 - 2.0 ** -53)     # keep (1 + y)/2 > 0: the kernels divide by it
    return np.ascontiguousarray(y), np.ascontiguousarray(w)


def quadrature_rule(name, num):
    if name == "gauss-legendre":
        return gauss_legendre(num)
    if name == "tanh-sinh":
        return tanh_sinh(num)
    if name.startswith("tanh-sinh:"):          # "tanh-sinh:2.0" = truncation t_max = 2.0 (nodes stay 2e-5 away from +-1)
        return tanh_sinh(num, t_max=float(name.split(":", 1)[1]))
    raise ValueError("unknown quadrature rule %r" % (name,))


def tables_for(cfg: AloConfig, dev):
    """Device table blob for (n, l, m) and the two rules, built once per device by faop_tables_build."""
    key = (cfg.n, cfg.l, cfg.m, cfg.rule_l, cfg.rule_m, str(dev))
    t = _TABLES.get(key)
    if t is None:
        lib = _lib.load()
        c = cfg.c()
        nbytes = ctypes.c_size_t()
        check(lib.faop_tables_bytes(ctypes.byref(c), ctypes.byref(nbytes)), "faop_tables_bytes")
        host = np.empty(nbytes.value // 8, dtype=np.float64)
        yl, wl = quadrature_rule(cfg.rule_l, cfg.l)
        ym, wm = quadrature_rule(cfg.rule_m, cfg.m)
        check(lib.faop_tables_build(ctypes.byref(c), yl.ctypes.data, wl.ctypes.data, ym.ctypes.data, wm.ctypes.data,
                                    host.ctypes.data), "faop_tables_build")
        t = torch.from_numpy(host).to(dev)
        _TABLES[key] = t
    return t


def price_batch(r, q, sigma, K, T, S, is_put, t=None, cfg: AloConfig = AloConfig(), B0=None,
                want_boundary=True, want_iter_errs=False, device=None, tau_max=None):
    """FastAmericanOptionSolver{A,B}.solve(t, s0) for a batch (FastAmericanOptionSolverBase.py:49-76).

    Returns dict(price, european, iters, err[, B, B0, tau]); B/B0/tau are N x (n+1) (node i at
    tau_i = T (1 + cos(i pi/n))^2 / 4).  B0 (optional, N x (n+1)) borrows a seed boundary instead of the QD seed.
    want_iter_errs adds iter_errs (N x max_iters, the errors of the reference's iter_records, NaN padded).
    tau_max (optional, per option): the reference's tau_max attribute (Base.py:28), the end of the collocation domain; default T.
    """
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        r = _f64(r, dev)
        N = r.numel()
        q, sigma, K, T, S = (_f64(x, dev, N) for x in (q, sigma, K, T, S))
        is_put = _u8(is_put, dev, N)
        tt = None if t is None else _f64(t, dev, N)
        tmax = None if tau_max is None else _f64(tau_max, dev, N)
        for x in (q, sigma, K, T, S, is_put) + (() if tmax is None else (tmax,)):
            if x.numel() != N:
                raise ValueError("all option arrays must have the same length")
        nn = cfg.n + 1
        tab = tables_for(cfg, dev)
        price = torch.empty(N, dtype=torch.float64, device=dev)
        eur = torch.empty(N, dtype=torch.float64, device=dev)
        iters = torch.empty(N, dtype=torch.int32, device=dev)
        err = torch.empty(N, dtype=torch.float64, device=dev)
        B = torch.empty((N, nn), dtype=torch.float64, device=dev) if want_boundary else None
        B0_in = None
        B0_out = None
        if B0 is not None:
            B0_in = torch.as_tensor(B0, dtype=torch.float64).to(dev).reshape(N, nn).contiguous()
        elif want_boundary:
            B0_out = torch.empty((N, nn), dtype=torch.float64, device=dev)
        # workspace: QD seeds (when no B0 tensor takes them) + the kernels' invariant-table scratch
        c = cfg.c(_lib.FLAG_WS_SCRATCH, tau_max=tmax)
        nb = ctypes.c_size_t()
        seeds_elsewhere = B0_in is not None or B0_out is not None       # then the workspace holds the scratch only
        check(lib.faop_workspace_bytes(ctypes.byref(c), 0 if seeds_elsewhere else N, ctypes.byref(nb)), "faop_workspace_bytes")
        ws = torch.empty(max(nb.value, 256), dtype=torch.uint8, device=dev)
        ierr = torch.empty((N, cfg.max_iters), dtype=torch.float64, device=dev) if want_iter_errs else None
        check(lib.faop_alo_price_batch(ctypes.byref(c), N, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(T), _ptr(tt),
                                       _ptr(S), _ptr(is_put), _ptr(tab), _ptr(B0_in), _ptr(price), _ptr(eur), _ptr(B),
                                       _ptr(B0_out), _ptr(iters), _ptr(err), _ptr(ierr), _ptr(ws), _stream(dev)),
              "faop_alo_price_batch")
        out = dict(price=price, european=eur, iters=iters, err=err)
        if want_iter_errs:
            out["iter_errs"] = ierr
        if want_boundary:
            out["B"] = B
            out["B0"] = B0_in if B0_in is not None else B0_out
            out["tau"] = collocation_tau(T if tmax is None else tmax, cfg.n)
        return out


_ONE = {}


def price_one(r, q, sigma, K, T, S, is_put, t=0.0, cfg: AloConfig = AloConfig(), tau_max=None, device=None):
    """solve(t, s0) of ONE option (the drop-in classes' scalar path) with one host-to-device copy, the two kernel launches and
    one device-to-host copy: the seven scalars travel as one pinned 80-byte block, every output (price, european, err, iters,
    B, B0, the per-sweep errors) comes back as one block.  Buffers are cached per (device, n, max_iters).
    Returns a dict of numpy values / arrays."""
    lib = _lib.load()
    dev = _device(device)
    nn, mi = cfg.n + 1, int(cfg.max_iters)
    key = (str(dev), nn, mi)
    st = _ONE.get(key)
    with torch.cuda.device(dev):
        if st is None:
            nout = 4 + 2 * nn + mi                  # price, european, err, iters(int32 in an 8-byte slot), B, B0, iter_errs
            st = dict(hin=torch.empty(10, dtype=torch.float64).pin_memory(), din=torch.empty(10, dtype=torch.float64, device=dev),
                      hout=torch.empty(nout, dtype=torch.float64).pin_memory(), dout=torch.empty(nout, dtype=torch.float64, device=dev))
            _ONE[key] = st
        h = st["hin"].numpy()
        h[:8] = (r, q, sigma, K, T, t, S, T if tau_max is None else tau_max)
        h[8:9].view(np.uint8)[0] = 1 if is_put else 0
        din, dout = st["din"], st["dout"]
        din.copy_(st["hin"], non_blocking=True)
        base, ob = din.data_ptr(), dout.data_ptr()

        def ip(i):
            return ctypes.c_void_p(base + 8 * i)

        def op(i):
            return ctypes.c_void_p(ob + 8 * i)
        c = cfg.c()
        if tau_max is not None:
            c.tau_max = base + 8 * 7
        check(lib.faop_alo_price_batch(ctypes.byref(c), 1, ip(0), ip(1), ip(2), ip(3), ip(4), ip(5), ip(6), ip(8),
                                       _ptr(tables_for(cfg, dev)), None, op(0), op(1), op(4), op(4 + nn), op(3), op(2),
                                       op(4 + 2 * nn) if mi > 0 else None, None, _stream(dev)), "faop_alo_price_batch")
        st["hout"].copy_(dout, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        o = st["hout"].numpy()
        return dict(price=np.float64(o[0]), european=np.float64(o[1]), err=float(o[2]), iters=int(o[3:4].view(np.int32)[0]),
                    B=o[4:4 + nn].copy(), B0=o[4 + nn:4 + 2 * nn].copy(), iter_errs=o[4 + 2 * nn:4 + 2 * nn + mi].copy())


def collocation_tau(T, n):
    """set_collocation_points (FastAmericanOptionSolverBase.py:221-223): tau_i = (cos(i pi/n) + 1)^2 T / 4."""
    c = torch.as_tensor(np.cos(np.arange(n + 1) * np.pi / n), dtype=torch.float64, device=T.device)
    return (c + 1).square()[None, :] * T[:, None] / 4


def eval_points_batch(B, tau_pts, B_pts, r, q, sigma, K, T, is_put, cfg: AloConfig = AloConfig(), device=None):
    """N_func, D_func, Nprime_func, Dprime_func, compute_f_and_fprime and the damped update at arbitrary
    (tau, Q) points (N x p) on a given boundary B (N x (n+1)); FastAmericanOptionSolverBase.py:144-165, 267-279.
    Returns dict(N, D, Nprime, Dprime, f, fprime, Bnew, I1, I2), each N x p (I1, I2: the two quadrature integrals,
    K12/K3 for Solver A, the D-/N-integrals for Solver B)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        r = _f64(r, dev)
        N = r.numel()
        q, sigma, K, T = (_f64(x, dev, N) for x in (q, sigma, K, T))
        is_put = _u8(is_put, dev, N)
        Bd = torch.as_tensor(B, dtype=torch.float64).to(dev).reshape(N, cfg.n + 1).contiguous()
        tp = torch.as_tensor(tau_pts, dtype=torch.float64).to(dev).reshape(N, -1).contiguous()
        bp = torch.as_tensor(B_pts, dtype=torch.float64).to(dev).reshape(N, -1).contiguous()
        if tp.shape != bp.shape:
            raise ValueError("tau_pts and B_pts must have the same shape")
        p = tp.shape[1]
        outs = [torch.empty((N, p), dtype=torch.float64, device=dev) for _ in range(9)]
        tab = tables_for(cfg, dev)
        c = cfg.c()
        check(lib.faop_alo_eval_points_batch(ctypes.byref(c), N, p, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(T),
                                             _ptr(is_put), _ptr(tab), _ptr(Bd), _ptr(tp), _ptr(bp),
                                             *[_ptr(o) for o in outs], _stream(dev)), "faop_alo_eval_points_batch")
        return dict(zip(("N", "D", "Nprime", "Dprime", "f", "fprime", "Bnew", "I1", "I2"), outs))


def eval_nodes_batch(B, r, q, sigma, K, T, is_put, cfg: AloConfig = AloConfig(), device=None):
    """eval_points_batch at the n active collocation nodes (tau_i, B_i) of the boundary itself."""
    dev = _device(device)
    Bd = torch.as_tensor(B, dtype=torch.float64).to(dev).reshape(-1, cfg.n + 1)
    Td = _f64(T, dev, Bd.shape[0])
    tau = collocation_tau(Td, cfg.n)
    return eval_points_batch(Bd, tau[:, :cfg.n], Bd[:, :cfg.n], r, q, sigma, K, Td, is_put, cfg=cfg, device=device)


def interp_boundary_batch(B, u, r, q, K, T, is_put, cfg: AloConfig = AloConfig(), device=None):
    """B(u) at arbitrary u (N x p) from the Chebyshev interpolant of ln(B/X)^2 on [0, T]
    (compute_integration_terms, FastAmericanOptionSolverBase.py:167-195)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        r = _f64(r, dev)
        N = r.numel()
        q, K, T = (_f64(x, dev, N) for x in (q, K, T))
        is_put = _u8(is_put, dev, N)
        Bd = torch.as_tensor(B, dtype=torch.float64).to(dev).reshape(N, cfg.n + 1).contiguous()
        ud = torch.as_tensor(u, dtype=torch.float64).to(dev).reshape(N, -1).contiguous()
        out = torch.empty_like(ud)
        tab = tables_for(cfg, dev)
        c = cfg.c()
        check(lib.faop_alo_interp_boundary_batch(ctypes.byref(c), N, ud.shape[1], _ptr(r), _ptr(q), _ptr(K), _ptr(T),
                                                 _ptr(is_put), _ptr(tab), _ptr(Bd), _ptr(ud), _ptr(out), _stream(dev)),
              "faop_alo_interp_boundary_batch")
        return out


def price_known_boundary_batch(B, r, q, sigma, K, T, S, is_put, t=None, cfg: AloConfig = AloConfig(), device=None,
                               tau_max=None):
    """american_value_with_known_boundary (FastAmericanOptionSolverBase.py:92-103). Returns (price, european).
    tau_max: end of the boundary's collocation domain (Base.py:28,185), default T."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        r = _f64(r, dev)
        N = r.numel()
        q, sigma, K, T, S = (_f64(x, dev, N) for x in (q, sigma, K, T, S))
        is_put = _u8(is_put, dev, N)
        tt = None if t is None else _f64(t, dev, N)
        Bd = torch.as_tensor(B, dtype=torch.float64).to(dev).reshape(N, cfg.n + 1).contiguous()
        price = torch.empty(N, dtype=torch.float64, device=dev)
        eur = torch.empty(N, dtype=torch.float64, device=dev)
        tab = tables_for(cfg, dev)
        tmax = None if tau_max is None else _f64(tau_max, dev, N)
        c = cfg.c(tau_max=tmax)
        check(lib.faop_alo_price_known_boundary_batch(ctypes.byref(c), N, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(T),
                                                      _ptr(tt), _ptr(S), _ptr(is_put), _ptr(tab), _ptr(Bd), _ptr(price),
                                                      _ptr(eur), _stream(dev)), "faop_alo_price_known_boundary_batch")
        return price, eur


def surface_batch(surface: dict, first, count, cfg: AloConfig, want_iters=False, device=None, workspace_mb=512,
                  dedup=False, out=None):
    """Config-5 implied-surface sweep; options are decoded from the flat index on the device
    (workloads.c5_surface defines the index order).  Returns price (and iters).
    dedup=True: one boundary solve per (T, sigma) group shared by all strikes (faop_alo_surface_dedup_batch);
    prices agree with the plain sweep to rounding.  out: optional preallocated price tensor (count,)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        s = FaopSurface(surface["K0"], surface["K1"], surface["T0"], surface["T1"], surface["V0"], surface["V1"],
                        surface["S"], surface["r"], surface["q"], surface["nK"], surface["nT"], surface["nV"],
                        int(surface.get("is_put", 1)))
        price = torch.empty(count, dtype=torch.float64, device=dev) if out is None else out
        iters = torch.empty(count, dtype=torch.int32, device=dev) if want_iters else None
        if dedup:
            tab = tables_for(cfg, dev)
            c = cfg.c()
            nb = ctypes.c_size_t()
            check(lib.faop_alo_surface_dedup_workspace_bytes(ctypes.byref(c), ctypes.byref(s), int(first), int(count),
                                                             ctypes.byref(nb)), "faop_alo_surface_dedup_workspace_bytes")
            ws = torch.empty(max(nb.value, 256), dtype=torch.uint8, device=dev)
            check(lib.faop_alo_surface_dedup_batch(ctypes.byref(c), ctypes.byref(s), int(first), int(count), _ptr(tab),
                                                   _ptr(price), _ptr(iters), _ptr(ws), ws.numel(), _stream(dev)),
                  "faop_alo_surface_dedup_batch")
            return (price, iters) if want_iters else price
        ws = torch.empty(workspace_mb << 20, dtype=torch.uint8, device=dev)
        tab = tables_for(cfg, dev)
        c = cfg.c()
        check(lib.faop_alo_surface_batch(ctypes.byref(c), ctypes.byref(s), int(first), int(count), _ptr(tab), _ptr(price),
                                         _ptr(iters), _ptr(ws), ws.numel(), _stream(dev)), "faop_alo_surface_batch")
        return (price, iters) if want_iters else price


def greeks_batch(r, q, sigma, K, T, S, is_put, t=None, cfg: AloConfig = AloConfig(), bump_S_rel=1e-3, bump_t=1e-4,
                 bump_sigma=1e-4, bump_r=1e-5, want_vega=True, want_rho=True, device=None):
    """Price and Greeks of a batch.  delta, gamma, theta (dV/dt) reuse the solved boundary — it depends neither on S nor on t —
    through faop_alo_greeks_known_boundary_batch; vega and rho are central differences of two re-solved batches each
    (the boundary does depend on sigma and r).  Returns dict(price, delta, gamma, theta[, vega][, rho], iters, err)."""
    lib = _lib.load()
    base = price_batch(r, q, sigma, K, T, S, is_put, t=t, cfg=cfg, want_boundary=True, device=device)
    dev = base["price"].device
    with torch.cuda.device(dev):
        N = base["price"].numel()
        r_, q_, sg_, K_, T_, S_ = (_f64(x, dev, N) for x in (r, q, sigma, K, T, S))
        ip = _u8(is_put, dev, N)
        tt = None if t is None else _f64(t, dev, N)
        tab = tables_for(cfg, dev)
        price, delta, gamma, theta = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(4))
        c = cfg.c()
        check(lib.faop_alo_greeks_known_boundary_batch(ctypes.byref(c), N, _ptr(r_), _ptr(q_), _ptr(sg_), _ptr(K_), _ptr(T_),
                                                       _ptr(tt), _ptr(S_), _ptr(ip), _ptr(tab), _ptr(base["B"]),
                                                       float(bump_S_rel), float(bump_t), _ptr(price), _ptr(delta), _ptr(gamma),
                                                       _ptr(theta), _stream(dev)), "faop_alo_greeks_known_boundary_batch")
        out = dict(price=base["price"], delta=delta, gamma=gamma, theta=theta, iters=base["iters"], err=base["err"])
        # the finite differences of re-solved prices need a converged boundary (at iter_tol = 1e-5 the up and down solves may
        # stop after different sweep counts, a 1e-5-sized kink against a 1e-4-sized bump): the bumped solves run with the
        # stopping rule tightened 1000x and room for the extra sweeps; the price itself keeps the caller's configuration
        import dataclasses
        fd = dataclasses.replace(cfg, iter_tol=cfg.iter_tol * 1e-3, max_iters=max(cfg.max_iters, 64), seed="newton")
        if want_vega:
            up = price_batch(r_, q_, sg_ + bump_sigma, K_, T_, S_, ip, t=tt, cfg=fd, want_boundary=False)["price"]
            dn = price_batch(r_, q_, sg_ - bump_sigma, K_, T_, S_, ip, t=tt, cfg=fd, want_boundary=False)["price"]
            out["vega"] = (up - dn) / (2 * bump_sigma)
        if want_rho:
            up = price_batch(r_ + bump_r, q_, sg_, K_, T_, S_, ip, t=tt, cfg=fd, want_boundary=False)["price"]
            dn = price_batch(r_ - bump_r, q_, sg_, K_, T_, S_, ip, t=tt, cfg=fd, want_boundary=False)["price"]
            out["rho"] = (up - dn) / (2 * bump_r)
        return out


def _iv_answer(sig, lo, hi, flo, fhi, side):
    """The bracket end with the smaller price residual (both ends have been priced); NaN where the target was not bracketed."""
    best = torch.where(flo.abs() <= fhi.abs(), lo, hi)
    return torch.where(side >= 0, best, sig)


def implied_vol_batch(target, r, q, K, T, S, is_put, cfg: AloConfig = AloConfig(), sigma_lo=0.05, sigma_hi=3.0, iters=48,
                      tol=1e-9, f_tol=1e-9, check_every=2, device=None):
    """Implied volatility of American prices: per option the root of price_batch(sigma) - target in [sigma_lo, sigma_hi]
    by a bracketed Illinois search (bisection when the secant stalls) whose state lives on the device
    (faop_iv_update_batch); every trial is one batch solve of the options still searching.  An option is finished when its
    bracket is narrower than tol or the price residual at one end of it is below f_tol; finished options leave the batch
    (checked every `check_every` trials), so a few slow ones do not keep a million converged ones re-solving.  Stops after
    `iters` trials at the latest.  Targets outside [price(sigma_lo), price(sigma_hi)] give NaN.
    Returns (sigma, trials) with trials = the number of batch solves of the longest-searching option."""
    lib = _lib.load()
    dev, N, (target, r, q, K, T, S) = _common([target, r, q, K, T, S], device)
    with torch.cuda.device(dev):
        ip = _u8(is_put, dev, N)
        result = torch.full((N,), float("nan"), dtype=torch.float64, device=dev)
        idx = torch.arange(N, device=dev)
        cur = [target, r, q, K, T, S, ip]
        n = N
        lo = torch.full((n,), float(sigma_lo), dtype=torch.float64, device=dev)
        hi = torch.full((n,), float(sigma_hi), dtype=torch.float64, device=dev)
        flo = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
        fhi = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
        side = torch.full((n,), -2, dtype=torch.int32, device=dev)
        sig = lo.clone()
        trials = 0
        for it in range(int(iters) + 2):
            tg, r_, q_, K_, T_, S_, ip_ = cur
            ok = torch.isfinite(sig)
            p = price_batch(r_, q_, torch.where(ok, sig, lo), K_, T_, S_, ip_, cfg=cfg, want_boundary=False)["price"]
            trials += 1
            check(lib.faop_iv_update_batch(n, _ptr(tg), _ptr(p), _ptr(sig), _ptr(lo), _ptr(hi), _ptr(flo), _ptr(fhi),
                                           _ptr(side), _stream(dev)), "faop_iv_update_batch")
            if it >= 2 and (it - 2) % check_every == check_every - 1:
                live = (side >= 0) & ((hi - lo) > tol) & (torch.minimum(flo.abs(), fhi.abs()) > f_tol)
                n_live = int(live.sum().item())
                if n_live == 0:
                    break
                if n_live <= n // 2:          # compact: the finished options keep their answer, the rest go on alone
                    result[idx] = _iv_answer(sig, lo, hi, flo, fhi, side)
                    keep = live.nonzero(as_tuple=True)[0]
                    idx = idx[keep]
                    cur = [x[keep].contiguous() for x in cur]
                    lo, hi, flo, fhi, side, sig = (x[keep].contiguous() for x in (lo, hi, flo, fhi, side, sig))
                    n = n_live
        result[idx] = _iv_answer(sig, lo, hi, flo, fhi, side)
        return result, trials


def ladder_batch(r, q, sigma, T, S, is_put, K, cfg: AloConfig = AloConfig(), want_iters=False, device=None, out=None):
    """Strike ladder: G scenarios (r, q, sigma, T, S, is_put) x the strikes K[0..nK) -> price[G, nK] (and iters).
    One boundary solve per scenario is shared by all of its strikes (the boundary scales with K); each price is what
    FastAmericanOptionSolver(r, q, sigma, K_k, T, type).solve(0, S) returns, to rounding (faop_alo_ladder_batch)."""
    lib = _lib.load()
    dev, G, (r, q, sigma, T, S) = _common([r, q, sigma, T, S], device)
    with torch.cuda.device(dev):
        ip = _u8(is_put, dev, G)
        Kh = np.ascontiguousarray(np.atleast_1d(K.detach().cpu().numpy() if isinstance(K, torch.Tensor) else np.asarray(K, dtype=np.float64)),
                                  dtype=np.float64).reshape(-1)
        nK = Kh.shape[0]
        Kd = _f64(K, dev)
        price = torch.empty((G, nK), dtype=torch.float64, device=dev) if out is None else out
        iters = torch.empty((G, nK), dtype=torch.int32, device=dev) if want_iters else None
        if G == 0 or nK == 0:
            return (price, iters) if want_iters else price
        tab = tables_for(cfg, dev)
        c = cfg.c()
        nb = ctypes.c_size_t()
        check(lib.faop_alo_ladder_workspace_bytes(ctypes.byref(c), G, ctypes.byref(nb)), "faop_alo_ladder_workspace_bytes")
        ws = torch.empty(max(nb.value, 256), dtype=torch.uint8, device=dev)
        check(lib.faop_alo_ladder_batch(ctypes.byref(c), G, _ptr(r), _ptr(q), _ptr(sigma), _ptr(T), _ptr(S), _ptr(ip), nK,
                                        _ptr(Kd), float(Kh.min()), float(Kh.max()), _ptr(tab), _ptr(price), _ptr(iters),
                                        _ptr(ws), ws.numel(), _stream(dev)), "faop_alo_ladder_batch")
        return (price, iters) if want_iters else price


# ------------------------------------------------------------------------------------------ leaf kernels
def _common(arrs, device):
    dev = _device(device)
    first = _f64(arrs[0], dev)
    N = first.numel()
    rest = [_f64(x, dev, N) for x in arrs[1:]]
    N = max([N] + [x.numel() for x in rest])
    allv = [first.expand(N).contiguous() if first.numel() == 1 and N != 1 else first] + \
           [x.expand(N).contiguous() if x.numel() == 1 and N != 1 else x for x in rest]
    return dev, N, allv


def qd_boundary_batch(tau, r, q, sigma, K, is_put, device=None):
    """QDplus.compute_exercise_boundary (QDplusAmericanOptionSolver.py:69-76) per element."""
    lib = _lib.load()
    dev, N, (tau, r, q, sigma, K) = _common([tau, r, q, sigma, K], device)
    with torch.cuda.device(dev):
        ip = _u8(is_put, dev, N)
        B = torch.empty(N, dtype=torch.float64, device=dev)
        check(lib.faop_qd_boundary_batch(N, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(tau), _ptr(ip), _ptr(B), _stream(dev)),
              "faop_qd_boundary_batch")
        return B


def qd_residual_batch(S, tau, r, q, sigma, K, is_put, device=None):
    """QDplus.exercise_boundary_func(S, tau) (QDplusAmericanOptionSolver.py:110-125) per element."""
    lib = _lib.load()
    dev, N, (S, tau, r, q, sigma, K) = _common([S, tau, r, q, sigma, K], device)
    with torch.cuda.device(dev):
        ip = _u8(is_put, dev, N)
        F = torch.empty(N, dtype=torch.float64, device=dev)
        check(lib.faop_qd_residual_batch(N, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(tau), _ptr(S), _ptr(ip), _ptr(F),
                                         _stream(dev)), "faop_qd_residual_batch")
        return F


def qd_price_batch(tau, S, r, q, sigma, K, is_put, device=None):
    """QDplus.price(tau, S) (QDplusAmericanOptionSolver.py:44-67). Returns (price, exercise_boundary)."""
    lib = _lib.load()
    dev, N, (tau, S, r, q, sigma, K) = _common([tau, S, r, q, sigma, K], device)
    with torch.cuda.device(dev):
        ip = _u8(is_put, dev, N)
        P = torch.empty(N, dtype=torch.float64, device=dev)
        Bd = torch.empty(N, dtype=torch.float64, device=dev)
        check(lib.faop_qd_price_batch(N, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(tau), _ptr(S), _ptr(ip), _ptr(P),
                                      _ptr(Bd), _stream(dev)), "faop_qd_price_batch")
        return P, Bd


def bjerksund_stensland_batch(S, r, q, sigma, K, T, is_put=None, Kb=None, Tb=None, device=None):
    """BksdStld(r, q, sigma, K, T).price(S) (BjerksundStenslandSolver.py:15-21).  Returns (price, boundary X).
    Kb / Tb: the module globals the reference's computeExerciseBoundary reads (:40, :42); default K, T."""
    lib = _lib.load()
    arrs = [S, r, q, sigma, K, T] + ([Kb] if Kb is not None else []) + ([Tb] if Tb is not None else [])
    dev, N, ts = _common(arrs, device)
    S, r, q, sigma, K, T = ts[:6]
    Kb_t = ts[6] if Kb is not None else None
    Tb_t = ts[6 + (Kb is not None)] if Tb is not None else None
    with torch.cuda.device(dev):
        ip = None if is_put is None else _u8(is_put, dev, N)
        P, X = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(2))
        check(lib.faop_bjerksund_stensland_batch(N, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(T), _ptr(S), _ptr(ip),
                                                 _ptr(Kb_t), _ptr(Tb_t), _ptr(P), _ptr(X), _stream(dev)),
              "faop_bjerksund_stensland_batch")
        return P, X


def bs_phi_batch(S, T, gamma, H, X, r, q, sigma, device=None):
    """BksdStld.phi (BjerksundStenslandSolver.py:26-33)."""
    lib = _lib.load()
    dev, N, (S, T, gamma, H, X, r, q, sigma) = _common([S, T, gamma, H, X, r, q, sigma], device)
    with torch.cuda.device(dev):
        out = torch.empty(N, dtype=torch.float64, device=dev)
        check(lib.faop_bs_phi_batch(N, _ptr(S), _ptr(T), _ptr(gamma), _ptr(H), _ptr(X), _ptr(r), _ptr(q), _ptr(sigma),
                                    _ptr(out), _stream(dev)), "faop_bs_phi_batch")
        return out


def european_batch(tau, S, r, q, sigma, K, is_put, device=None):
    """EuropeanOption.european_put_value / european_call_value (EuropeanOptionSolver.py:7-22)."""
    lib = _lib.load()
    dev, N, (tau, S, r, q, sigma, K) = _common([tau, S, r, q, sigma, K], device)
    with torch.cuda.device(dev):
        ip = _u8(is_put, dev, N)
        P = torch.empty(N, dtype=torch.float64, device=dev)
        check(lib.faop_european_batch(N, _ptr(tau), _ptr(S), _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(ip), _ptr(P),
                                      _stream(dev)), "faop_european_batch")
        return P


def european_greeks_batch(tau, S, r, q, sigma, K, device=None):
    """european_option_theta, d1, d2 (EuropeanOptionSolver.py:25-40). Returns (theta, d1, d2)."""
    lib = _lib.load()
    dev, N, (tau, S, r, q, sigma, K) = _common([tau, S, r, q, sigma, K], device)
    with torch.cuda.device(dev):
        th, d1, d2 = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3))
        check(lib.faop_european_greeks_batch(N, _ptr(tau), _ptr(S), _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(th),
                                             _ptr(d1), _ptr(d2), _stream(dev)), "faop_european_greeks_batch")
        return th, d1, d2


def cheb_fit_batch(y, device=None):
    """ChebyshevInterpolation coefficient fit (ChebyshevInterpolation.py:43-51); y is N x (n+1)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        y = torch.as_tensor(y, dtype=torch.float64).to(dev)
        y = y.reshape(1, -1) if y.dim() == 1 else y.contiguous()
        N, nn = y.shape
        a = torch.empty_like(y)
        check(lib.faop_cheb_fit_batch(N, nn - 1, _ptr(y), _ptr(a), _stream(dev)), "faop_cheb_fit_batch")
        return a


def cheb_eval_batch(a, z, device=None):
    """Clenshaw evaluation (ChebyshevInterpolation.py:31-41); a is N x (n+1), z is N x npts in [-1, 1]."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        a = torch.as_tensor(a, dtype=torch.float64).to(dev)
        a = a.reshape(1, -1) if a.dim() == 1 else a.contiguous()
        z = torch.as_tensor(z, dtype=torch.float64).to(dev)
        z = z.reshape(1, -1) if z.dim() == 1 else z.contiguous()
        N, nn = a.shape
        out = torch.empty_like(z)
        check(lib.faop_cheb_eval_batch(N, nn - 1, z.shape[1], _ptr(a), _ptr(z), _ptr(out), _stream(dev)), "faop_cheb_eval_batch")
        return out


# ------------------------------------------------------------------------------------------ Crank-Nicolson
def cn_dt_schedule(T, max_dt):
    """CNOptionSolver.solvePDE's float countdown (CrankNicolsonOptionSolver.py:36-45), host-side control logic."""
    out = []
    t = float(T)
    max_dt = float(max_dt)
    while t > 0:
        dt = min(t, max_dt)
        out.append(dt)
        t -= dt
    return out


def cn_dt_schedule_batch(T, max_dt):
    """cn_dt_schedule for many options at once: the same float countdown, vectorised over options (one numpy pass per
    time step).  Returns (dts [N x max_steps], n_steps [N])."""
    t = np.array(T, dtype=np.float64, copy=True)
    max_dt = np.asarray(max_dt, dtype=np.float64)
    cols = []
    ns = np.zeros(t.shape[0], dtype=np.int32)
    live = t > 0
    while live.any():
        dt = np.where(live, np.minimum(t, max_dt), 0.0)
        cols.append(dt)
        ns += live
        t = np.where(live, t - dt, t)
        live = t > 0
    dts = np.stack(cols, axis=1) if cols else np.zeros((t.shape[0], 1))
    return np.ascontiguousarray(dts), ns


CN_MODE_DIRECT, CN_MODE_SOR, CN_MODE_PROJECTED = 0, 1, 2


def cn_put_batch(r, q, sigma, K, T, S0, n_space=200, max_dt=1 / 12.0, mode=CN_MODE_PROJECTED, omega=1.2, tol=1e-5,
                 max_iter=200, want_grid=False, device=None, schedule=None, Smax=None, X_in=None, want_state=False):
    """CNOptionSolver(r,q,sigma,K,T).solve(S0) for a batch of American puts (CrankNicolsonOptionSolver.py:27-107).
    mode 0 = the reference default (dense solve, no projection), 1 = the reference's SOR+project loop,
    2 = projected tridiagonal sweep.  Returns price (and the final grid values X, N x n_space).
    The time steps are the reference's float countdown from T in steps of max_dt, run on the device
    (faop_cn_put_countdown_batch); schedule=(dts [N x steps], n_steps [N]) supplies explicit steps instead.
    Smax = the CNOptionSolver.Smax attribute per option (:12,48; default 3K); X_in = N x n_space state to start from
    instead of the payoff; want_state adds (err, iter), the attributes of the last solvePSOR (:103-104), to the result."""
    lib = _lib.load()
    dev, N, (r, q, sigma, K, T, S0) = _common([r, q, sigma, K, T, S0], device)
    with torch.cuda.device(dev):
        price = torch.empty(N, dtype=torch.float64, device=dev)
        X = torch.empty((N, n_space), dtype=torch.float64, device=dev) if want_grid else None
        nb = ctypes.c_size_t()
        check(lib.faop_cn_workspace_bytes(N, n_space, mode, ctypes.byref(nb)), "faop_cn_workspace_bytes")
        ws = torch.empty(max(nb.value, 1), dtype=torch.uint8, device=dev)
        sm = None if Smax is None else _f64(Smax, dev, N)
        if sm is not None and sm.numel() != N:
            raise ValueError("Smax must be a scalar or have one entry per option")
        xin = None
        if X_in is not None:
            xin = torch.as_tensor(X_in, dtype=torch.float64).to(dev).reshape(N, -1).contiguous()
            if xin.shape[1] != n_space:
                raise ValueError("X_in must be N x n_space")
        err = torch.empty(N, dtype=torch.float64, device=dev) if want_state else None
        it = torch.empty(N, dtype=torch.int32, device=dev) if want_state else None
        ex = _lib.FaopCnExtra(_ptr(sm), _ptr(xin), _ptr(err), _ptr(it))
        if schedule is None:
            mdt = _f64(np.broadcast_to(np.asarray(max_dt.cpu() if isinstance(max_dt, torch.Tensor) else max_dt, dtype=np.float64),
                                       (N,)).copy(), dev, N)
            check(lib.faop_cn_put_countdown_batch(N, n_space, mode, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(T), _ptr(S0),
                                                  _ptr(mdt), float(omega), float(tol), int(max_iter), ctypes.byref(ex),
                                                  _ptr(price), _ptr(X), _ptr(ws), _stream(dev)), "faop_cn_put_countdown_batch")
        else:
            dts, ns = schedule
            dts = np.ascontiguousarray(dts, dtype=np.float64)
            dts_d = torch.from_numpy(dts).to(dev)
            ns_d = torch.from_numpy(np.ascontiguousarray(ns, dtype=np.int32)).to(dev)
            check(lib.faop_cn_put_batch(N, n_space, mode, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(S0), _ptr(dts_d),
                                        _ptr(ns_d), dts.shape[1], float(omega), float(tol), int(max_iter), ctypes.byref(ex),
                                        _ptr(price), _ptr(X), _ptr(ws), _stream(dev)), "faop_cn_put_batch")
        out = (price, X) if want_grid else price
        return (out, err, it) if want_state else out


def cn_rows_batch(r, q, sigma, K, dt, X, Smax=None, device=None):
    """CNOptionSolver.setCoeff(dt) (CrankNicolsonOptionSolver.py:55-79) as data: the three diagonals lo, di, up and the
    right-hand side b (each N x n_space) of one time step from the state X (faop_cn_rows_batch)."""
    lib = _lib.load()
    dev, N, (r, q, sigma, K, dt) = _common([r, q, sigma, K, dt], device)
    with torch.cuda.device(dev):
        X = torch.as_tensor(X, dtype=torch.float64).to(dev).reshape(N, -1).contiguous()
        n_space = X.shape[1]
        sm = None if Smax is None else _f64(Smax, dev, N)
        out = [torch.empty((N, n_space), dtype=torch.float64, device=dev) for _ in range(4)]
        check(lib.faop_cn_rows_batch(N, n_space, _ptr(r), _ptr(q), _ptr(sigma), _ptr(K), _ptr(sm), _ptr(dt), _ptr(X),
                                     *[_ptr(o) for o in out], _stream(dev)), "faop_cn_rows_batch")
        return out


def launch_count():
    return int(_lib.load().faop_launch_count())


def dev_math_batch(kind, x, device=None):
    """Unit-test hook: the device math of csrc/faop_fastmath.cuh applied elementwise (see faop_dev_math_batch)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        x = _f64(x, dev)
        out = torch.empty_like(x)
        check(lib.faop_dev_math_batch(int(kind), x.numel(), _ptr(x), _ptr(out), _stream(dev)), "faop_dev_math_batch")
        return out


def measure_fp64_peak(launches=20, device=None):
    """DFMA-pipe peak of this GPU in TFLOP/s (the roofline denominator): returns (tflops, ms_per_launch)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        tf, ms = ctypes.c_double(), ctypes.c_double()
        check(lib.faop_measure_fp64_peak(int(launches), ctypes.byref(tf), ctypes.byref(ms), _stream(dev)),
              "faop_measure_fp64_peak")
        return tf.value, ms.value


def measure_point_mix_peak(launches=10, device=None):
    """Quadrature points per second of the flagship kernel's point math alone (register operands, no memory traffic):
    the speed of light that options/s * sweeps * n * l of the real kernel is compared with.  Returns (points_per_s, ms)."""
    lib = _lib.load()
    dev = _device(device)
    with torch.cuda.device(dev):
        pps, ms = ctypes.c_double(), ctypes.c_double()
        check(lib.faop_measure_point_mix_peak(int(launches), ctypes.byref(pps), ctypes.byref(ms), _stream(dev)),
              "faop_measure_point_mix_peak")
        return pps.value, ms.value


class HostContext:
    """faop_ctx wrapper: the host-buffer entry point (numpy / pinned host arrays in and out), used by the
    e2e leg of bench.py and by callers that do not hold device memory."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        check(self._lib.faop_ctx_create(int(device), ctypes.byref(self._h)), "faop_ctx_create")

    def close(self):
        if self._h:
            self._lib.faop_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _host_in(x, N, dtype, name):
        """An input as a C-contiguous host array of `dtype` and length N: numpy arrays of the right layout and pinned torch
        CPU tensors pass through untouched (no copy); anything else numpy can convert (lists, float32, strided views, bool
        flags) is converted; a wrong length raises.  The C call reads exactly N elements from the returned buffer."""
        if isinstance(x, torch.Tensor):
            if x.device.type != "cpu" or x.dtype != (torch.float64 if dtype == np.float64 else torch.uint8) or not x.is_contiguous():
                x = x.detach().cpu().numpy()
            else:
                if x.numel() != N:
                    raise ValueError("%s has %d elements, expected %d" % (name, x.numel(), N))
                return x
        a = np.asarray(x)
        if dtype == np.uint8 and a.dtype == np.bool_:
            a = a.view(np.uint8)
        a = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
        if a.shape[0] == 1 and N != 1:
            a = np.full(N, a[0], dtype=dtype)
        if a.shape[0] != N:
            raise ValueError("%s has %d elements, expected %d" % (name, a.shape[0], N))
        return a

    @staticmethod
    def _host_out(x, shape, dtype, name):
        """A caller-supplied output buffer must already be what the C call writes: right dtype, C-contiguous, right size."""
        if isinstance(x, torch.Tensor):
            ok = x.device.type == "cpu" and x.is_contiguous() and tuple(x.shape) == shape and \
                x.dtype == {np.float64: torch.float64, np.int32: torch.int32}[dtype]
        else:
            ok = isinstance(x, np.ndarray) and x.dtype == dtype and x.flags.c_contiguous and x.flags.writeable and x.shape == shape
        if not ok:
            raise ValueError("out[%r] must be a writable C-contiguous %s array of shape %r" % (name, np.dtype(dtype).name, shape))
        return x

    def price_batch(self, r, q, sigma, K, T, S, is_put, t=None, cfg: AloConfig = AloConfig(), out=None, want_boundary=False,
                    tau_max=None):
        """Host arrays (numpy float64 / uint8; pinned memory makes the copies asynchronous) -> dict of numpy arrays.
        Inputs are checked and, where needed, converted (see _host_in); arrays in `out` are checked (see _host_out)."""
        def hp(a):
            return None if a is None else ctypes.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())
        N = int(r.numel()) if isinstance(r, torch.Tensor) else int(np.asarray(r).size)
        nn = cfg.n + 1
        r, q, sigma, K, T, S = (self._host_in(x, N, np.float64, nm) for x, nm in
                                ((r, "r"), (q, "q"), (sigma, "sigma"), (K, "K"), (T, "T"), (S, "S")))
        is_put = self._host_in(is_put, N, np.uint8, "is_put")
        t = None if t is None else self._host_in(t, N, np.float64, "t")
        tmax = None if tau_max is None else self._host_in(tau_max, N, np.float64, "tau_max")
        if out is None:
            out = dict(price=np.empty(N), european=np.empty(N), iters=np.empty(N, dtype=np.int32), err=np.empty(N))
            if want_boundary:
                out["B"] = np.empty((N, nn))
                out["B0"] = np.empty((N, nn))
        else:
            if "price" not in out:
                raise ValueError("out needs at least a 'price' array")
            for k, v in out.items():
                if k not in ("price", "european", "err", "iters", "B", "B0"):
                    raise ValueError("unknown output %r" % (k,))
                self._host_out(v, (N, nn) if k in ("B", "B0") else (N,), np.int32 if k == "iters" else np.float64, k)
        yl, wl = quadrature_rule(cfg.rule_l, cfg.l)
        ym, wm = quadrature_rule(cfg.rule_m, cfg.m)
        c = cfg.c()
        if tmax is not None:
            c.tau_max = tmax.ctypes.data if isinstance(tmax, np.ndarray) else tmax.data_ptr()      # a HOST array for the *_host entry
        check(self._lib.faop_alo_price_batch_host(self._h, ctypes.byref(c), N, hp(r), hp(q), hp(sigma), hp(K), hp(T), hp(t),
                                                  hp(S), hp(is_put), hp(yl), hp(wl), hp(ym), hp(wm), None, hp(out["price"]),
                                                  hp(out.get("european")), hp(out.get("B")), hp(out.get("B0")),
                                                  hp(out.get("iters")), hp(out.get("err"))), "faop_alo_price_batch_host")
        return out


class MultiGpuHost:
    """One host process, several GPUs (SURVEY 8e's single-process form): host arrays are cut into contiguous slices, one per
    device; every slice goes through its own faop_ctx (own stream pair, own arena) on its own Python thread — ctypes releases
    the GIL for the duration of faop_alo_price_batch_host, so the devices run concurrently — and the results land directly in
    the caller's output arrays.  No collective, no copy between devices."""

    def __init__(self, devices=None):
        if not torch.cuda.is_available():
            raise _lib.FaopError("CUDA device required: fastamericanoptionpricing_b200 has no CPU fallback")
        self.devices = list(range(torch.cuda.device_count())) if devices is None else [int(d) for d in devices]
        self.ctx = [HostContext(d) for d in self.devices]

    def close(self):
        for c in self.ctx:
            c.close()

    def price_batch(self, r, q, sigma, K, T, S, is_put, t=None, cfg: AloConfig = AloConfig(), out=None):
        from concurrent.futures import ThreadPoolExecutor
        from .sharding import shard_bounds
        N = int(r.numel()) if isinstance(r, torch.Tensor) else int(np.asarray(r).size)
        hin = HostContext._host_in
        r, q, sigma, K, T, S = (hin(x, N, np.float64, nm) for x, nm in
                                ((r, "r"), (q, "q"), (sigma, "sigma"), (K, "K"), (T, "T"), (S, "S")))
        is_put = hin(is_put, N, np.uint8, "is_put")
        t = None if t is None else hin(t, N, np.float64, "t")
        if out is None:
            out = dict(price=np.empty(N), european=np.empty(N), iters=np.empty(N, dtype=np.int32), err=np.empty(N))
        else:
            for k, v in out.items():
                HostContext._host_out(v, (N,), np.int32 if k == "iters" else np.float64, k)
        W = len(self.ctx)

        def work(g):
            lo, hi = shard_bounds(N, W, g)
            if hi > lo:
                sl = slice(lo, hi)
                self.ctx[g].price_batch(r[sl], q[sl], sigma[sl], K[sl], T[sl], S[sl], is_put[sl], None if t is None else t[sl],
                                        cfg=cfg, out={k: v[sl] for k, v in out.items()})
            return hi - lo

        with ThreadPoolExecutor(max_workers=W) as ex:
            done = sum(ex.map(work, range(W)))
        assert done == N
        return out
