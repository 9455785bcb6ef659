"""Drop-in for the reference's BjerksundStenslandSolver.py: class BksdStld (Bjerksund-Stensland 1993 flat-boundary
American call; a put through the symmetry of the reference's __main__, :53).  Same constructor, methods and
attributes; the arithmetic runs in the CUDA leaf kernels faop_bjerksund_stensland_batch / faop_bs_phi_batch.

Upstream quirk: computeExerciseBoundary (:38-43) reads the MODULE GLOBALS `K` and `T` (:40, :42), so the class only
works inside its own `__main__`, where they happen to be the call's strike and maturity.  Here the object's own
strike / maturity are used; set `boundary_K` / `boundary_T` on the instance to reproduce a run with foreign globals
(the upstream put line, `BksdStld(q, r, sigma, S0, T).price(K)` with global K = 100, is boundary_K = 100).
As upstream there is no S >= X early-exercise branch and no exception for numerical failure (NaN / inf propagate).
"""
import numpy as np

try:
    from . import batch as _batch
except ImportError:  # flat import (package directory on sys.path, as the reference's scripts expect)
    from fastamericanoptionpricing_b200 import batch as _batch


class BksdStld:
    def __init__(self, riskfree, dividend, volatility, strike, maturity):
        """reference BjerksundStenslandSolver.py:5-13"""
        self.r = riskfree
        self.q = dividend
        self.sigma = volatility
        self.K = strike
        self.T = maturity
        self.b = self.r - self.q
        sig_sqr = np.power(self.sigma, 2.0)
        self.beta = (0.5 - self.b / sig_sqr) + np.sqrt(np.power(self.b / sig_sqr - 0.5, 2.0) + 2.0 * self.r / sig_sqr)
        self.boundary_K = None   # the global `K` of :40 (None: self.K)
        self.boundary_T = None   # the global `T` of :42 (None: self.T)

    def _run(self, S):
        Sa = np.ascontiguousarray(np.atleast_1d(np.asarray(S, dtype=np.float64)).reshape(-1))
        P, X = _batch.bjerksund_stensland_batch(Sa, self.r, self.q, self.sigma, self.K, self.T, None,
                                                self.boundary_K, self.boundary_T)
        return P.cpu().numpy(), X.cpu().numpy()

    def price(self, S):
        """reference :15-21; S may be a scalar or an array (numpy-broadcast upstream)"""
        P, _ = self._run(S)
        return np.float64(P[0]) if np.ndim(S) == 0 else P.reshape(np.shape(S))

    def computeAlpha(self, X):
        """reference :23-24 (two flops; kept on the host like the constructor's beta)"""
        return (X - self.K) * np.power(X, -self.beta)

    def phi(self, S, T, gamma, H, X):
        """reference :26-33"""
        args = (S, T, gamma, H, X)
        bs = np.broadcast_arrays(*[np.asarray(a, dtype=np.float64) for a in args])
        flat = [np.ascontiguousarray(b).reshape(-1) for b in bs]
        out = _batch.bs_phi_batch(*flat, self.r, self.q, self.sigma).cpu().numpy()
        return np.float64(out[0]) if all(np.ndim(a) == 0 for a in args) else out.reshape(bs[0].shape)

    @staticmethod
    def cdf(x):
        """reference :35-37 (scipy.stats.norm.cdf) through the device Phi"""
        xa = np.atleast_1d(np.asarray(x, dtype=np.float64)).reshape(-1)
        out = _batch.dev_math_batch(11, xa).cpu().numpy()
        return np.float64(out[0]) if np.ndim(x) == 0 else out.reshape(np.shape(x))

    def computeExerciseBoundary(self):
        """reference :38-43"""
        return np.float64(self._run(self.K)[1][0])
