"""CPU oracle (numpy): restatement of the reference's spectral-collocation American pricer.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fastamericanoptionpricing_b200/)
imports this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may.  It is a checker, never the thing shipped or measured.

Every function restates one piece of antdvid/FastAmericanOptionPricing (pure Python,
one option per object) and cites the reference file:line it follows.  The restatement
is vectorised over options (leading axis) but keeps the reference's operation order
inside one option: same formulas for d+-, same `tau - u` round trip, Chebyshev fit +
Clenshaw per node, sequential quadrature sums, the same stopping rule.

Pinned against the live reference: tests/golden/ref_*.npz were produced by
oracle/gen_golden.py running the unmodified reference in the dev container
(numpy 2.3.5 / scipy 1.18.1); tests/test_oracle_golden.py checks this file against them.
The QD seed root follows the reference's root finder: `hybr1` below restates MINPACK hybrd for one
unknown exactly as scipy.optimize.root(method="hybr", x0=K) drives it (reference
QDplusAmericanOptionSolver.py:73) and returns the same iterate bit for bit
(tests/test_oracle_golden.py::test_hybr1_equals_scipy_root); a safeguarded Newton iteration on the
same residual is kept as `method="newton"` (the CUDA path's opt-in fast seed).

Third-party arithmetic the reference leans on (versions unpinned upstream):
scipy.stats.norm.cdf -> scipy.special.ndtr (used here directly), norm.pdf ->
exp(-x^2/2)/sqrt(2 pi), numpy.polynomial.legendre.leggauss (used here directly).
"""
import numpy as np
from numpy.polynomial.legendre import leggauss
from scipy.special import ndtr

_INV_SQRT_2PI = 1.0 / np.sqrt(2.0 * np.pi)


def norm_pdf(x):
    """scipy.stats.norm.pdf restated: exp(-x^2/2)/sqrt(2 pi)."""
    return np.exp(-0.5 * x * x) * _INV_SQRT_2PI


# --------------------------------------------------------------------------- European
def euro_d1(tau, s0, r, q, vol, strike):
    """EuropeanOptionSolver.py:35-36."""
    return np.log(s0 * np.exp((r - q) * tau) / strike) / (vol * np.sqrt(tau)) + 0.5 * vol * np.sqrt(tau)


def euro_d2(tau, s0, r, q, vol, strike):
    """EuropeanOptionSolver.py:39-40."""
    return euro_d1(tau, s0, r, q, vol, strike) - vol * np.sqrt(tau)


def european_put_value(tau, s0, r, q, vol, strike):
    """EuropeanOptionSolver.py:7-13 (tau==0 -> intrinsic)."""
    tau, s0, r, q, vol, strike = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (tau, s0, r, q, vol, strike)])
    with np.errstate(all="ignore"):
        d1 = euro_d1(tau, s0, r, q, vol, strike)
        d2 = d1 - vol * np.sqrt(tau)
        v = strike * np.exp(-r * tau) * ndtr(-d2) - s0 * np.exp(-q * tau) * ndtr(-d1)
    return np.where(tau == 0, np.maximum(0.0, strike - s0), v)


def european_call_value(tau, s0, r, q, vol, strike):
    """EuropeanOptionSolver.py:16-22."""
    tau, s0, r, q, vol, strike = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (tau, s0, r, q, vol, strike)])
    with np.errstate(all="ignore"):
        d1 = euro_d1(tau, s0, r, q, vol, strike)
        d2 = d1 - vol * np.sqrt(tau)
        v = s0 * np.exp(-q * tau) * ndtr(d1) - strike * np.exp(-r * tau) * ndtr(d2)
    return np.where(tau == 0, np.maximum(0.0, s0 - strike), v)


def european_option_theta(tau, s0, r, q, vol, strike):
    """EuropeanOptionSolver.py:25-32 (put theta; r and tau floored at 1e-10)."""
    r = np.maximum(r, 1e-10)
    tau = np.maximum(tau, 1e-10)
    d1 = euro_d1(tau, s0, r, q, vol, strike)
    d2 = d1 - vol * np.sqrt(tau)
    return r * strike * np.exp(-r * tau) * ndtr(-d2) - q * s0 * np.exp(-q * tau) * ndtr(-d1) \
        - vol * s0 * np.exp(-q * tau) * norm_pdf(d1) / (2 * np.sqrt(tau))


# --------------------------------------------------------------------------- Bjerksund-Stensland (1993) call
def bs_beta(r, q, sigma):
    """BjerksundStenslandSolver.py:11-13."""
    b = r - q
    s2 = np.power(sigma, 2.0)
    return (0.5 - b / s2) + np.sqrt(np.power(b / s2 - 0.5, 2.0) + 2.0 * r / s2)


def bs_boundary(r, q, sigma, K, T, Kb=None, Tb=None):
    """BjerksundStenslandSolver.py:38-43 (computeExerciseBoundary).  The reference reads the MODULE GLOBALS `K` and `T`
    at :40 and :42 (it only runs as __main__); Kb / Tb are those two values, defaulting to the object's own strike and
    maturity (the evident intent)."""
    Kb = K if Kb is None else Kb
    Tb = T if Tb is None else Tb
    b = r - q
    beta = bs_beta(r, q, sigma)
    with np.errstate(all="ignore"):
        B0 = np.maximum(K, (r / (r - b) * Kb))
        Binf = beta / (beta - 1) * K
        h = -(b * Tb + 2 * sigma * np.sqrt(T)) * (np.power(K, 2.0) / ((Binf - B0) * B0))
        return B0 + (Binf - B0) * (1 - np.exp(h))


def bs_phi(S, T, gamma, H, X, r, q, sigma):
    """BjerksundStenslandSolver.py:26-33."""
    b = r - q
    s2 = np.power(sigma, 2.0)
    with np.errstate(all="ignore"):
        lamb = -r + gamma * b + 0.5 * gamma * (gamma - 1) * s2
        kappa = 2 * b / s2 + (2 * gamma - 1.0)
        d1 = -(np.log(S / H) + (b + (gamma - 0.5) * s2) * T) / (sigma * np.sqrt(T))
        d2 = -(np.log(X * X / (S * H)) + (b + (gamma - 0.5) * s2) * T) / (sigma * np.sqrt(T))
        return np.exp(lamb * T) * np.power(S, gamma) * (ndtr(d1) - np.power(X / S, kappa) * ndtr(d2))


def bs_price(S, r, q, sigma, K, T, Kb=None, Tb=None):
    """BjerksundStenslandSolver.py:15-21 (price) with computeAlpha (:23-24).  No S >= X early-exercise branch, as upstream."""
    S, r, q, sigma, K, T = (np.asarray(a, dtype=float) for a in (S, r, q, sigma, K, T))
    X = bs_boundary(r, q, sigma, K, T, Kb, Tb)
    beta = bs_beta(r, q, sigma)
    with np.errstate(all="ignore"):
        alpha = (X - K) * np.power(X, -beta)
        ph = lambda g, H: bs_phi(S, T, g, H, X, r, q, sigma)  # noqa: E731
        return alpha * np.power(S, beta) - alpha * ph(beta, X) + ph(1, X) - ph(1, K) - K * ph(0, X) + K * ph(0, K)


# --------------------------------------------------------------------------- Chebyshev
def cheb_points(n):
    """ChebyshevInterpolation.py:15-17: cos(i pi / n), i = 0..n."""
    i = np.arange(0, n + 1)
    return np.cos(i * np.pi / n)


def cheb_fit(y):
    """ChebyshevInterpolation.py:43-51: a_k = (2/n) sum'' y_i cos(i k pi / n).  y: (..., n+1)."""
    y = np.asarray(y, dtype=float)
    n = y.shape[-1] - 1
    a = np.zeros_like(y)
    for k in range(n + 1):
        ans = np.zeros(y.shape[:-1])
        for i in range(n + 1):
            term = y[..., i] * np.cos(i * k * np.pi / n)
            if i == 0 or i == n:
                term = term * 0.5
            ans = ans + term
        a[..., k] = ans * (2.0 / n)
    return a


def cheb_eval(a, z):
    """ChebyshevInterpolation.py:31-41 (Clenshaw).  a: (..., n+1), z: (..., npts) already in [-1, 1]."""
    a = np.asarray(a, dtype=float)
    z = np.asarray(z, dtype=float)
    n = a.shape[-1] - 1
    b0 = a[..., n, None] * 0.5 + 0.0 * z
    b1 = np.zeros_like(b0)
    b2 = np.zeros_like(b0)
    for k in range(n - 1, -1, -1):
        b1, b2 = b0, b1
        b0 = a[..., k, None] + 2 * z * b1 - b2
    return 0.5 * (b0 - b2)


def to_cheby_point(x, x_min, x_max):
    """FastAmericanOptionSolverBase.py:235-237."""
    return np.sqrt(4 * (x - x_min) / (x_max - x_min)) - 1


def to_orig_point(c, x_min, x_max):
    """FastAmericanOptionSolverBase.py:239-240."""
    return np.square(c + 1) * (x_max - x_min) / 4 + x_min


# --------------------------------------------------------------------------- QD seed
def B_at_zero(r, q, K, is_put):
    """FastAmericanOptionSolverBase.py:246-256 / QDplusAmericanOptionSolver.py:78-88."""
    r, q, K = (np.asarray(a, dtype=float) for a in (r, q, K))
    is_put = np.asarray(is_put, dtype=bool)
    with np.errstate(all="ignore"):
        rkq = r / q * K
    return np.where(is_put, np.where(r >= q, K, rkq), np.where(r <= q, K, rkq))


def qd_residual(S, tau, r, q, sigma, K, is_put, want_deriv=False):
    """QDplusAmericanOptionSolver.py:110-134: the QD boundary residual (and its analytic dF/dS).

    put : (1 - e^{-q tau} Phi(-d1)) S + qQD (K - S - p)      (:124)
    call: (1 - e^{-q tau} Phi(+d1)) S - qQD (S - K - c)      (:122)
    """
    with np.errstate(all="ignore"):
        Nn = 2 * (r - q) / (sigma * sigma)                           # :217-218
        M = 2 * r / (sigma * sigma)                                  # :214-215
        h = 1 - np.exp(-r * tau)                                     # :211-212
        disc = np.sqrt((Nn - 1) * (Nn - 1) + 4 * M / h)
        qQD = np.where(is_put, -0.5 * (Nn - 1) - 0.5 * disc, -0.5 * (Nn - 1) + 0.5 * disc)   # :127-134
        d1 = euro_d1(tau, S, r, q, sigma, K)
        d2 = d1 - sigma * np.sqrt(tau)
        eq = np.exp(-q * tau)
        put = K * np.exp(-r * tau) * ndtr(-d2) - S * eq * ndtr(-d1)
        call = S * eq * ndtr(d1) - K * np.exp(-r * tau) * ndtr(d2)
        Fp = (1 - eq * ndtr(-d1)) * S + qQD * (K - S - put)
        Fc = (1 - eq * ndtr(d1)) * S - qQD * (S - K - call)
        F = np.where(is_put, Fp, Fc)
        if not want_deriv:
            return F
        svt = sigma * np.sqrt(tau)
        # put: dFdS as written at QDplusAmericanOptionSolver.py:207-208; call: derived the same way
        dp = (1 - qQD) * (1 - eq * ndtr(-d1)) + eq * norm_pdf(d1) / svt
        dc = (1 - qQD) * (1 - eq * ndtr(d1)) - eq * norm_pdf(d1) / svt
        return F, np.where(is_put, dp, dc)


_EPSMCH = float(np.finfo(float).eps)


def hybr1(f, x0, xtol=1.49012e-08, factor=100.0, maxfev=400, epsfcn=_EPSMCH):
    """MINPACK `hybrd` (Powell hybrid: forward-difference Jacobian, dogleg step inside a trust region, Broyden rank-one
    updates) specialised to ONE unknown, with the settings scipy.optimize.root(fun, x0, method="hybr") passes by default:
    xtol = 1.49012e-8, maxfev = 200 (n+1), ml = mu = n-1, epsfcn = machine eps, mode = 1 (diag from the first Jacobian),
    factor = 100.  That call is how the reference finds its QD seed (QDplusAmericanOptionSolver.py:73); MINPACK is a
    third-party dependency absent from /root/reference (scipy 1.18.1 ships a C translation), so this restates the published
    algorithm.  With n = 1 the QR factorisation is r = -J, q = -1 (Householder of a 1 x 1 matrix), every enorm is |.|, r1updt
    and r1mpyq apply no rotation.  Returns (x, nfev, info); x is what `root(...).x[0]` holds whatever `info` says (the
    reference never looks at `success`).  Bit-identical to scipy on 12 000 seed problems
    (tests/test_oracle_golden.py::test_hybr1_equals_scipy_root checks a sample); scipy's own nfev counts two extra shape-probing calls."""
    sqrt = np.sqrt
    x = float(x0)
    fvec = f(x)
    nfev = 1
    fnorm = abs(fvec)
    it = 1
    ncsuc = ncfail = nslow1 = nslow2 = 0
    diag = delta = xnorm = 0.0
    eps = float(sqrt(max(epsfcn, _EPSMCH)))
    while True:
        jeval = True
        h = eps * abs(x)                          # fdjac1
        if h == 0.0:
            h = eps
        J = (f(x + h) - fvec) / h
        nfev += 1
        acn = abs(J)                              # qrfac: rdiag = -J, acnorm = |J|
        if it == 1:
            diag = acn if acn != 0.0 else 1.0
            xnorm = abs(diag * x)
            delta = factor * xnorm
            if delta == 0.0:
                delta = factor
        if J != 0.0:
            qtf, r, Q = -fvec, -J, -1.0           # (q transpose) fvec with q = -1
        else:
            qtf, r, Q = fvec, -0.0, 1.0
        diag = max(diag, acn)
        while True:
            temp = r                              # dogleg: Gauss-Newton direction
            if temp == 0.0:
                temp = _EPSMCH * abs(r)
                if temp == 0.0:
                    temp = _EPSMCH
            xg = qtf / temp
            qnorm = abs(diag * xg)
            if qnorm <= delta:
                p = xg
            else:                                 # scaled gradient direction / dogleg point
                wa1 = (r * qtf) / diag
                gnorm = abs(wa1)
                sgnorm = 0.0
                alpha = delta / qnorm
                if gnorm != 0.0:
                    wa1 = (wa1 / gnorm) / diag
                    t2 = abs(r * wa1)
                    sgnorm = (gnorm / t2) / t2
                    alpha = 0.0
                    if sgnorm < delta:
                        bnorm = abs(qtf)
                        tt = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta)
                        tt = tt - (delta / qnorm) * (sgnorm / delta) ** 2 + float(sqrt(
                            (tt - (delta / qnorm)) ** 2 + (1.0 - (delta / qnorm) ** 2) * (1.0 - (sgnorm / delta) ** 2)))
                        alpha = ((delta / qnorm) * (1.0 - (sgnorm / delta) ** 2)) / tt
                tt = (1.0 - alpha) * min(sgnorm, delta)
                p = tt * wa1 + alpha * xg
            w1 = -p
            w2 = x + w1
            w3 = diag * w1
            pnorm = abs(w3)
            if it == 1:
                delta = min(delta, pnorm)
            w4 = f(w2)
            nfev += 1
            fnorm1 = abs(w4)
            actred = -1.0
            if fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            w3 = qtf + r * w1
            temp = abs(w3)
            prered = 0.0
            if temp < fnorm:
                prered = 1.0 - (temp / fnorm) ** 2
            ratio = 0.0
            if prered > 0.0:
                ratio = actred / prered
            if ratio < 0.1:
                ncsuc = 0
                ncfail += 1
                delta = 0.5 * delta
            else:
                ncfail = 0
                ncsuc += 1
                if ratio >= 0.5 or ncsuc > 1:
                    delta = max(delta, pnorm / 0.5)
                if abs(ratio - 1.0) <= 0.1:
                    delta = pnorm / 0.5
            if ratio >= 1e-4:                     # successful iteration
                x = w2
                w2 = diag * x
                fvec = w4
                xnorm = abs(w2)
                fnorm = fnorm1
                it += 1
            nslow1 += 1
            if actred >= 0.001:
                nslow1 = 0
            if jeval:
                nslow2 += 1
            if actred >= 0.1:
                nslow2 = 0
            if delta <= xtol * xnorm or fnorm == 0.0:
                return x, nfev, 1
            info = 0
            if nfev >= maxfev:
                info = 2
            if 0.1 * max(0.1 * delta, pnorm) <= _EPSMCH * xnorm:
                info = 3
            if nslow2 == 5:
                info = 4
            if nslow1 == 10:
                info = 5
            if info != 0:
                return x, nfev, info
            if ncfail == 2:
                break                             # recompute the Jacobian by forward differences
            sm = Q * w4                           # Broyden rank-one update of r = -J (and of qtf)
            w2 = (sm - w3) / pnorm
            w1 = diag * ((diag * w1) / pnorm)
            if ratio >= 1e-4:
                qtf = sm
            r = r + w2 * w1
            jeval = False


def qd_boundary(tau, r, q, sigma, K, is_put, max_newton=100, method="hybr"):
    """QDplusAmericanOptionSolver.py:69-76: root of the QD residual started from x0 = K; tau == 0 -> B_at_zero.

    method="hybr" (default): the reference's own root finder, scipy.optimize.root(hybr, x0=K), restated by `hybr1`, one
    scalar problem per element.  method="newton": Newton with the analytic derivative, a positivity safeguard (halve towards
    0 / step-limit), a relative stop of 1e-10 (quadratic convergence makes the accepted iterate exact to rounding) and a stall
    test for flat residuals — vectorised, the CUDA path's opt-in fast seed.
    All arguments broadcast; returns an array.
    """
    tau, r, q, sigma, K = np.broadcast_arrays(*[np.asarray(a, dtype=float) for a in (tau, r, q, sigma, K)])
    is_put = np.broadcast_to(np.asarray(is_put, dtype=bool), tau.shape)
    if method == "hybr":
        out = np.empty(tau.shape)
        X0 = B_at_zero(r, q, K, is_put)
        with np.errstate(all="ignore"):
            for ix in np.ndindex(tau.shape):
                if tau[ix] == 0:
                    out[ix] = X0[ix]
                    continue
                a = (tau[ix], r[ix], q[ix], sigma[ix], K[ix], is_put[ix])
                out[ix] = hybr1(lambda S: float(qd_residual(np.float64(S), *a)), K[ix])[0]
        return out
    S = K.astype(float).copy()
    act = tau > 0
    prev = np.full(S.shape, np.inf)
    with np.errstate(all="ignore"):
        for _ in range(max_newton):
            if not act.any():
                break
            F, dF = qd_residual(S, tau, r, q, sigma, K, is_put, want_deriv=True)
            step = F / dF
            Snew = S - step
            bad = ~np.isfinite(Snew) | (Snew <= 0)
            Snew = np.where(bad, 0.5 * S, Snew)
            # limit growth to a factor of 2 per step (keeps the flat small-tau cases on the near root)
            Snew = np.minimum(Snew, 2.0 * S)
            d = np.abs(Snew - S)
            # converged (the next Newton step would be ~d^2), or stalled at the rounding floor of a flat residual
            done = (d <= 1e-10 * np.abs(Snew)) | ((d >= 0.5 * prev) & (prev <= 1e-7 * np.abs(Snew)))
            prev = np.where(act, d, prev)
            S = np.where(act, Snew, S)
            act = act & ~done
    return np.where(tau == 0, B_at_zero(r, q, K, is_put), S)


def qd_price(tau, S, r, q, sigma, K, is_put):
    """QDplusAmericanOptionSolver.py:44-67 (scalars).  Quirks kept: tau==0 -> max(S-K,0) for puts too (:45-47),
    put-form correction (K - Sb - pSb) for calls (:67), c() uses Phi(-d1) for both types (:172).
    Returns (price, exercise_boundary)."""
    if tau == 0:
        return max(S - K, 0.0), K
    Sb = float(qd_boundary(tau, r, q, sigma, K, is_put))
    ind = -1 if is_put else 1
    with np.errstate(all="ignore"):
        Nn = 2 * (r - q) / (sigma * sigma)
        M = 2 * r / (sigma * sigma)
        h = 1 - np.exp(-r * tau)
        disc = np.sqrt((Nn - 1) * (Nn - 1) + 4 * M / h)
        qQD = -0.5 * (Nn - 1) + ind * 0.5 * disc
        qQDdot = M / (h * h * disc)                                                   # :136-140
        d1 = euro_d1(tau, Sb, r, q, sigma, K)
        value = european_put_value if is_put else european_call_value
        p = float(value(tau, Sb, r, q, sigma, K))
        theta = european_option_theta(tau, Sb, r, q, sigma, K)
        dFdh = qQD * theta * np.exp(r * tau) / r + qQDdot * (K - Sb - p) \
            + (Sb * q * np.exp(-q * tau) * ndtr(-d1)) / (r * (1 - h)) \
            - (Sb * np.exp(-q * tau) * norm_pdf(d1)) / (2 * r * tau * (1 - h)) \
            * (2 * np.log(Sb / K) / (sigma * np.sqrt(tau)) - d1)                          # :202-205
        dFdS = (1 - qQD) * (1 - np.exp(-q * tau) * ndtr(-d1)) + (np.exp(-q * tau) * norm_pdf(d1)) / (sigma * np.sqrt(tau))
        dlogSdh = -dFdh / dFdS                                                        # :209
        c0 = -(1 - h) * M / (2 * qQD + Nn - 1) * (1 / h - (theta * np.exp(r * tau)) / (r * (K - Sb - p)) + qQDdot / (2 * qQD + Nn - 1))
        c = c0 - ((1 - h) * M) / (2 * qQD + Nn - 1) * ((1 - np.exp(-q * tau) * ndtr(-d1)) / (K - Sb - p) + qQD / Sb) * dlogSdh
        b = ((1 - h) * M * qQDdot) / (2 * (2 * qQD + Nn - 1))
        pS = float(value(tau, S, r, q, sigma, K))
        if ind * (Sb - S) <= 0:
            return ind * (S - K), Sb
        lg = np.log(S / Sb)
        return pS + (K - Sb - p) / (1 - b * np.square(lg) - c * lg) * np.power(S / Sb, qQD), Sb


# --------------------------------------------------------------------------- solver core
def _dpm(s, z, r, q, sigma):
    """FastAmericanOptionSolverBase.py:306-310, same operation order."""
    lz = np.log(z)
    dm = (lz + (r - q) * s - 0.5 * sigma * sigma * s) / (sigma * np.sqrt(s))
    dp = (lz + (r - q) * s + 0.5 * sigma * sigma * s) / (sigma * np.sqrt(s))
    return dm, dp


def _cdf(sign, d, s, z):
    """CDF_{pos,neg}_{dminus,dplus}, FastAmericanOptionSolverBase.py:312-346, with the s == 0 limits."""
    v = ndtr(sign * d)
    if sign > 0:
        lim = np.where(z > 1, 1.0, 0.0)
    else:
        lim = np.where(z > 1, 0.0, 1.0)
    return np.where(s == 0, lim, v)


def _pdf(d, s):
    """PDF_*, FastAmericanOptionSolverBase.py:348-370 (0 at s == 0; the normal pdf is even)."""
    return np.where(s == 0, 0.0, norm_pdf(d))


def _seq_sum(terms):
    """quadrature_sum's `ans += adding` loop (FastAmericanOptionSolverBase.py:384-388): left-to-right over the last axis."""
    ans = np.zeros(terms.shape[:-1])
    for k in range(terms.shape[-1]):
        ans = ans + terms[..., k]
    return ans


def interp_boundary(H, tau_pts, T, X, is_put, y):
    """compute_integration_terms, FastAmericanOptionSolverBase.py:167-195.

    H: (P, n+1) = ln(B/X)^2, tau_pts: (P, nodes), y: (npts,) Gauss-Legendre nodes.
    Returns u (P, nodes, npts) and B(u) (P, nodes, npts); Chebyshev domain is [0, T] (tau_min=0, tau_max=T).
    """
    a = cheb_fit(H)                                                    # Cheb.py:4-12
    u = tau_pts[..., None] - tau_pts[..., None] * np.square(1 + y) / 4.0           # Base.py:186
    with np.errstate(all="ignore"):
        z = to_cheby_point(u, 0.0, T[:, None, None])                   # Base.py:235-237
        Hu = cheb_eval(a[:, None, :], z)
        root = np.sqrt(np.maximum(0.0, Hu))
        Bu = np.where(is_put[:, None, None], np.exp(-root), np.exp(root)) * X[:, None, None]   # Base.py:190-193
    return u, Bu


def nd_terms(solver, use_derivative, B, tau, T, r, q, sigma, K, X, is_put, y, w):
    """N, D (and N', D') at every collocation node for the current boundary iterate.

    Solver A: FastAmericanOptionSolverA.py:5-83, Solver B: FastAmericanOptionSolverB.py:5-75.
    B, tau: (P, n+1) with the last node tau == 0.  Returns arrays (P, n) for the n active nodes.
    """
    P, nn = B.shape
    n = nn - 1
    with np.errstate(all="ignore"):
        H = np.square(np.log(B / X[:, None]))                                  # Base.py:184
    ta = tau[:, :n]
    u, Bu = interp_boundary(H, ta, T, X, is_put, y)
    r3, q3, s3 = r[:, None, None], q[:, None, None], sigma[:, None, None]
    put3 = is_put[:, None, None]
    put2 = is_put[:, None]
    r2, q2, s2 = r[:, None], q[:, None], sigma[:, None]
    Bi = B[:, :n]
    with np.errstate(all="ignore"):
        s = ta[..., None] - u                                                  # tau - u as the reference forms it
        zz = Bi[..., None] / Bu
        dm, dp = _dpm(s, zz, r3, q3, s3)
        jac = 0.5 * (ta[..., None] - 0) * (1 + y)                              # Base.py:242-244
        W = w * jac
        dmK, dpK = _dpm(ta, Bi / K[:, None], r2, q2, s2)
        svt = s2 * np.sqrt(ta)
        ssu = s3 * np.sqrt(s)
        eq, er = np.exp(q3 * u), np.exp(r3 * u)
        if solver == "A":
            # K12 / K3 integrands, A.py:24-44 (pdf is even, so the call's PDF_neg_* are the same numbers)
            cdf_dp = np.where(put3, _cdf(+1, dp, s, zz), -_cdf(-1, dp, s, zz))
            k12 = eq * cdf_dp + eq / ssu * _pdf(dp, s)
            k3 = er / ssu * _pdf(dm, s)
            K12 = _seq_sum(k12 * w * jac)
            K3 = _seq_sum(k3 * w * jac)
            Nv = _pdf(dmK, ta) / svt + r2 * K3                                                         # A.py:5-10
            Dv = _pdf(dpK, ta) / svt + np.where(put2, _cdf(+1, dpK, ta, Bi / K[:, None]),
                                                -_cdf(-1, dpK, ta, Bi / K[:, None])) + q2 * K12           # A.py:13-19
            Np = Dp = None
            if use_derivative:
                # A.py:46-83; tau = max(1e-10, tau) is a no-op for active nodes
                npi = er * dm / (Bi[..., None] * s3 * s3 * s) * _pdf(dm, s)
                dpi = eq * _pdf(dp, s) / (ssu * Bi[..., None]) * (1 - dp / ssu)
                Np = -dmK * _pdf(dmK, ta) / (Bi * s2 * s2 * ta) - r2 * _seq_sum(npi * w * jac)
                Dp = _pdf(dpK, ta) / (svt * Bi) * (1 - dpK / svt) + q2 * _seq_sum(dpi * w * jac)
        else:
            # B.py:5-39: put uses Phi(+d), call uses Phi(-d) in all four places
            sgn_cdf_dm = np.where(put3, _cdf(+1, dm, s, zz), _cdf(-1, dm, s, zz))
            sgn_cdf_dp = np.where(put3, _cdf(+1, dp, s, zz), _cdf(-1, dp, s, zz))
            Nint = _seq_sum(er * sgn_cdf_dm * w * jac)
            Dint = _seq_sum(eq * sgn_cdf_dp * w * jac)
            zK = Bi / K[:, None]
            Nv = np.where(put2, _cdf(+1, dmK, ta, zK), _cdf(-1, dmK, ta, zK)) + r2 * Nint
            Dv = np.where(put2, _cdf(+1, dpK, ta, zK), _cdf(-1, dpK, ta, zK)) + q2 * Dint
            Np = Dp = None
            if use_derivative:
                # B.py:41-75.  Quirk kept: for calls D' integrates Nprime_integrand (e^{ru}, phi(d-)) (B.py:57).
                npi = er * _pdf(dm, s) / (ssu * Bi[..., None])
                dpi = eq * _pdf(dp, s) / (ssu * Bi[..., None])
                Nsum = _seq_sum(npi * w * jac)
                Dsum = _seq_sum(dpi * w * jac)
                Np = np.where(put2, _pdf(dmK, ta) / (svt * Bi) + r2 * Nsum, -_pdf(dmK, ta) / (svt * Bi) - r2 * Nsum)
                Dp = np.where(put2, _pdf(dpK, ta) / (svt * Bi) + q2 * Dsum, -_pdf(dpK, ta) / (svt * Bi) - q2 * Nsum)
    return Nv, Dv, Np, Dp


def f_and_fprime(solver, use_derivative, B, tau, T, r, q, sigma, K, X, is_put, y, w):
    """compute_f_and_fprime, FastAmericanOptionSolverBase.py:267-279, for the n active nodes."""
    n = B.shape[1] - 1
    Nv, Dv, Np, Dp = nd_terms(solver, use_derivative, B, tau, T, r, q, sigma, K, X, is_put, y, w)
    with np.errstate(all="ignore"):
        disc = K[:, None] * np.exp(-tau[:, :n] * (r - q)[:, None])
        f = disc * Nv / Dv
        fp = np.ones_like(f)
        if use_derivative:
            fp = disc * (Np / Dv - Dp * Nv / (Dv * Dv))
    return f, fp, Nv, Dv


def solve_boundary(solver, r, q, sigma, K, T, is_put, n, l, iter_tol=1e-5, max_iters=200,
                   use_derivative=False, eta=0.5, B0=None, seed="hybr", rule_l=None):
    """compute_exercise_boundary + iterate_once(_foreach_tau), FastAmericanOptionSolverBase.py:105-165.

    Jacobi sweep on all nodes from the previous iterate, B_i += eta (B_i - f)/(f' - 1) with f' = 0 unless
    use_derivative, node tau == 0 pinned to X, stop when ||B_old - B_new||_2 <= tol or count == max_iters.
    T is the end of the collocation domain, i.e. the `tau_max` attribute (= maturity unless changed, Base.py:28, used at :185, :223).
    seed: root finder of the QD seed ("hybr" as upstream, or "newton"); rule_l: (y, w) on [-1, 1] replacing leggauss(l) (opt-in).
    Returns dict(tau, B0, B, iters, err, iter_errs).
    """
    r, q, sigma, K, T = (np.ascontiguousarray(a, dtype=float) for a in (r, q, sigma, K, T))
    is_put = np.ascontiguousarray(is_put, dtype=bool)
    P = r.shape[0]
    y, w = leggauss(l) if rule_l is None else (np.asarray(rule_l[0], dtype=float), np.asarray(rule_l[1], dtype=float))   # Base.py:175
    c = cheb_points(n)
    tau = to_orig_point(c[None, :], 0.0, T[:, None])                  # Base.py:221-223
    X = B_at_zero(r, q, K, is_put)
    if B0 is None:
        B0 = qd_boundary(tau, r[:, None], q[:, None], sigma[:, None], K[:, None], is_put[:, None], method=seed)   # Base.py:258-265
    B = np.array(B0, dtype=float, copy=True)
    iters = np.zeros(P, dtype=np.int64)
    err = np.ones(P)
    iter_errs = np.full((P, max_iters), np.nan)
    act = np.ones(P, dtype=bool) if max_iters > 0 else np.zeros(P, dtype=bool)
    while act.any():
        ix = np.nonzero(act)[0]
        f, fp, _, _ = f_and_fprime(solver, use_derivative, B[ix], tau[ix], T[ix], r[ix], q[ix], sigma[ix], K[ix],
                                   X[ix], is_put[ix], y, w)
        Bold = B[ix]
        Bnew = Bold.copy()
        fprime = fp if use_derivative else np.zeros_like(fp)
        with np.errstate(all="ignore"):
            Bnew[:, :n] = Bold[:, :n] + eta * (Bold[:, :n] - f) / (fprime - 1)
        Bnew[:, n] = X[ix]
        with np.errstate(all="ignore"):
            e = np.sqrt(np.sum(np.square(np.abs(Bold - Bnew)), axis=1))         # norm1_error is the 2-norm, Base.py:230-233
        B[ix] = Bnew
        iters[ix] += 1
        err[ix] = e
        iter_errs[ix, iters[ix] - 1] = e
        # `while iter_err > tol and iter_count < max_iters` — NaN compares False and stops the loop, as in Python
        act[ix] = (e > iter_tol) & (iters[ix] < max_iters)
    return dict(tau=tau, B0=np.asarray(B0, dtype=float), B=B, iters=iters, err=err, iter_errs=iter_errs, X=X)


def price_with_boundary(B, T, t, S, r, q, sigma, K, is_put, m, tau_max=None, rule_m=None):
    """american_value_with_known_boundary + v_integrand_12, FastAmericanOptionSolverBase.py:92-103, 211-219.
    tau_max: end of the boundary's collocation domain (Base.py:28,185), default T; rule_m: (y, w) replacing leggauss(m)."""
    r, q, sigma, K, T, t, S = (np.ascontiguousarray(a, dtype=float) for a in (r, q, sigma, K, T, t, S))
    is_put = np.ascontiguousarray(is_put, dtype=bool)
    X = B_at_zero(r, q, K, is_put)
    tau = T - t
    Tc = T if tau_max is None else np.ascontiguousarray(np.broadcast_to(np.asarray(tau_max, dtype=float), T.shape))
    eur = np.where(is_put, european_put_value(tau, S, r, q, sigma, K), european_call_value(tau, S, r, q, sigma, K))
    y, w = leggauss(m) if rule_m is None else (np.asarray(rule_m[0], dtype=float), np.asarray(rule_m[1], dtype=float))
    with np.errstate(all="ignore"):
        H = np.square(np.log(B / X[:, None]))
        u, Bu = interp_boundary(H, tau[:, None], Tc, X, is_put, y)
        u, Bu = u[:, 0, :], Bu[:, 0, :]
        tt, SS, rr, qq, ss, KK = (a[:, None] for a in (tau, S, r, q, sigma, K))
        s = tt - u
        zz = SS / Bu
        dm, dp = _dpm(s, zz, rr, qq, ss)
        putv = rr * KK * np.exp(-rr * s) * _cdf(-1, dm, s, zz) - qq * SS * np.exp(-qq * s) * _cdf(-1, dp, s, zz)
        callv = qq * SS * np.exp(-qq * s) * _cdf(+1, dp, s, zz) - rr * KK * np.exp(-rr * s) * _cdf(+1, dm, s, zz)
        integrand = np.where(is_put[:, None], putv, callv)
        jac = 0.5 * tt * (1 + y)
        v12 = _seq_sum(integrand * w * jac)
    v12 = np.where(tau == 0, 0.0, v12)                                 # quadrature_sum returns 0 at tau == 0, Base.py:382-383
    return eur + v12, eur


def alo_price_batch(solver, r, q, sigma, K, T, S, is_put, t=None, n=12, l=24, m=48, iter_tol=1e-5, max_iters=200,
                    use_derivative=False, eta=0.5, B0=None, chunk=4096, seed="hybr", rule_l=None, rule_m=None, tau_max=None):
    """FastAmericanOptionSolver.solve(t, s0) for a batch, FastAmericanOptionSolverBase.py:49-76.
    seed / rule_l / rule_m / tau_max: see solve_boundary and price_with_boundary (tau_min stays 0: any other value makes
    upstream return NaN after one sweep, because z = sqrt(4 (u - tau_min)/(tau_max - tau_min)) - 1 meets u < tau_min).

    q == 0 calls take the European shortcut (:50-53): price = european call, iters = 0, err = 1e6 (the
    constructor's initial values, :41-43), boundary arrays NaN.
    """
    r, q, sigma, K, T, S = (np.atleast_1d(np.asarray(a, dtype=float)) for a in (r, q, sigma, K, T, S))
    is_put = np.atleast_1d(np.asarray(is_put, dtype=bool))
    P = r.shape[0]
    t = np.zeros(P) if t is None else np.broadcast_to(np.asarray(t, dtype=float), (P,)).copy()
    out = dict(price=np.empty(P), european=np.empty(P), B=np.full((P, n + 1), np.nan), B0=np.full((P, n + 1), np.nan),
               tau=np.full((P, n + 1), np.nan), iters=np.zeros(P, dtype=np.int64), err=np.full(P, 1e6),
               iter_errs=np.full((P, max_iters), np.nan))
    short = (q == 0) & ~is_put
    if short.any():
        e = european_call_value(T[short] - t[short], S[short], r[short], q[short], sigma[short], K[short])
        out["price"][short] = e
        out["european"][short] = e
    idx = np.nonzero(~short)[0]
    Tc = T if tau_max is None else np.broadcast_to(np.asarray(tau_max, dtype=float), (P,)).copy()
    for c0 in range(0, len(idx), chunk):
        ix = idx[c0:c0 + chunk]
        sb = solve_boundary(solver, r[ix], q[ix], sigma[ix], K[ix], Tc[ix], is_put[ix], n, l, iter_tol, max_iters,
                            use_derivative, eta, None if B0 is None else np.asarray(B0)[ix], seed=seed, rule_l=rule_l)
        price, eur = price_with_boundary(sb["B"], T[ix], t[ix], S[ix], r[ix], q[ix], sigma[ix], K[ix], is_put[ix], m,
                                         tau_max=Tc[ix], rule_m=rule_m)
        out["price"][ix] = price
        out["european"][ix] = eur
        for k in ("B", "B0", "tau", "iters", "err", "iter_errs"):
            out[k][ix] = sb[k]
    return out


def match_condition2(solver, B, tau, T, r, q, sigma, K, is_put, l):
    """check_value_match_condition2, FastAmericanOptionSolverBase.py:401-412 (diagnostic only)."""
    y, w = leggauss(l)
    X = B_at_zero(r, q, K, is_put)
    n = B.shape[1] - 1
    Nv, Dv, _, _ = nd_terms(solver, False, B, tau, T, r, q, sigma, K, X, is_put, y, w)
    left = Nv * K[:, None] * np.exp(-r[:, None] * tau[:, :n])
    right = Dv * B[:, :n] * np.exp(-q[:, None] * tau[:, :n])
    return np.sqrt(np.sum(np.square(left - right), axis=1))


# --------------------------------------------------------------------------- Crank-Nicolson
def cn_dt_schedule(T, max_dt):
    """CrankNicolsonOptionSolver.py:36-45: `while t > 0: dt = min(t, max_dt); ...; t -= dt` in floating point."""
    out = []
    t = float(T)
    while t > 0:
        dt = min(t, max_dt)
        out.append(dt)
        t -= dt
    return out


def cn_solve(r, q, sigma, K, T, S0, N=200, max_dt=1 / 12.0, use_psor=False, tol=1e-5, max_iter=200, omega=1.2,
             projected_thomas=False, Smax=None, state=None):
    """CNOptionSolver.solve, CrankNicolsonOptionSolver.py:27-107, one option (scalars).

    use_psor=False : the reference's default dense `np.linalg.solve` per step (NO early-exercise projection, :81-82),
                     restated as an exact tridiagonal solve (last row closed through three affine back-substitutions).
    use_psor=True  : the reference's "SOR sweep then project the whole vector" loop (:84-107), operation for operation,
                     including the last-row quirk (leaked index i = N-2, j = N-1 term on itself).
    projected_thomas=True : Brennan-Schwartz projected back-substitution max(., K - S_i) (not in the reference; the
                     CUDA path's throughput mode, cross-checked against spectral prices).
    Smax: the CNOptionSolver.Smax attribute (:12, consumed at :48), default 3K.  `state`, if a dict, receives err / iter of the
    last SOR solve (:103-104).
    Returns (price, S grid, X).
    """
    S = np.linspace(0, 3 * K if Smax is None else Smax, N)
    X = np.maximum(K - S, 0)
    dS = S[1] - S[0]
    payoff = K - S
    for dt in cn_dt_schedule(T, max_dt):
        i = np.arange(N - 1)
        alpha = 0.25 * dt * (np.square(sigma * S[i] / dS) - (r - q) * S[i] / dS)          # :65
        beta = 0.5 * dt * (r + np.square(sigma * S[i] / dS))                              # :66
        gamma = 0.25 * dt * (np.square(sigma * S[i] / dS) + (r - q) * S[i] / dS)          # :67
        b = np.zeros(N)
        b[0] = X[0] * (1 - beta[0])
        b[1:N - 1] = alpha[1:] * X[0:N - 2] + (1 - beta[1:]) * X[1:N - 1] + gamma[1:] * X[2:N]
        lo = np.zeros(N)
        di = np.zeros(N)
        up = np.zeros(N)
        di[:N - 1] = 1 + beta
        lo[1:N - 1] = -alpha[1:]
        up[1:N - 1] = -gamma[1:]
        if use_psor and not projected_thomas:
            it = 0
            err = 1000.0
            X = X.copy()
            while err > tol and it < max_iter:
                it += 1
                x_old = X.copy()
                for k in range(N - 1):
                    xk = (1 - omega) * X[k] + omega * b[k] / di[k]
                    xk -= up[k] * X[k + 1] * omega / di[k]
                    xk -= lo[k] * X[k - 1] * omega / di[k]      # k == 0 reads X[-1] with a zero coefficient
                    X[k] = xk
                k = N - 2
                xl = (1 - omega) * X[k] + omega * b[k] / di[k]
                last = {N - 4: -1.0, N - 3: 4.0, N - 2: -5.0, N - 1: 2.0}
                X[N - 1] = xl
                for j in range(N - 4, N):
                    X[N - 1] -= last[j] * X[j] * omega / 2.0
                X = np.maximum(X, payoff)
                err = np.linalg.norm(x_old - X)
            if state is not None:
                state["err"], state["iter"] = float(err), it
        else:
            # rows 0..N-2 are tridiagonal: forward elimination gives X_i = dp_i - cp_i X_{i+1}; the 4-term last row
            # (-1, 4, -5, 2 | 0) is closed by writing X_{N-2}, X_{N-3}, X_{N-4} as affine functions of x = X_{N-1}
            cp = np.zeros(N)
            dp = np.zeros(N)
            cp[0] = up[0] / di[0]
            dp[0] = b[0] / di[0]
            for k in range(1, N - 1):
                den = di[k] - lo[k] * cp[k - 1]
                cp[k] = up[k] / den
                dp[k] = (b[k] - lo[k] * dp[k - 1]) / den
            p2, m2 = dp[N - 2], -cp[N - 2]
            p3, m3 = dp[N - 3] - cp[N - 3] * p2, -cp[N - 3] * m2
            p4, m4 = dp[N - 4] - cp[N - 4] * p3, -cp[N - 4] * m3
            Xn = np.zeros(N)
            Xn[N - 1] = (p4 - 4.0 * p3 + 5.0 * p2) / (2.0 - m4 + 4.0 * m3 - 5.0 * m2)
            if projected_thomas:
                Xn[N - 1] = max(Xn[N - 1], payoff[N - 1])
            for k in range(N - 2, -1, -1):
                Xn[k] = dp[k] - cp[k] * Xn[k + 1]
                if projected_thomas:
                    Xn[k] = max(Xn[k], payoff[k])
            X = Xn
    return float(np.interp(S0, S, X)), S, X
