"""Drop-in for the reference's FastAmericanOptionSolverBase.py.

Same class name, constructor, knobs (plain attributes mutated after construction), result attributes and helper
methods as the reference (file:line citations below are relative to the reference root).  All numerical work is
done by the CUDA kernels behind the C ABI: a scalar `solve(t, s0)` is a batch of one option through
faop_alo_price_batch (one warp on the GPU).  Conventions kept: no exceptions for numerical failure (NaN
propagates; non-convergence shows only in `error` / `num_iters`), `DEBUG` defaults to True and prints,
`iter_records` accumulates across solves, `integration_num` is frozen at 2*24 in the constructor.
Not kept: the `(tau, num_points)` scratch cache that can serve stale B(u), the result-neutral
diagnostic inside the iteration loop (App. B#3, evaluated only when DEBUG prints it), and `exit()` in
check_f_with_B (App. B#12)."""
import math
from abc import ABC, abstractmethod

import numpy as np
import numpy.linalg as alg
import numpy.polynomial.legendre as legendre

try:
    from . import batch as _batch
    from . import ChebyshevInterpolation as intrp
    from . import EuropeanOptionSolver as europ
    from . import QDplusAmericanOptionSolver as qd
except ImportError:  # flat import: this directory is on sys.path, as the reference's scripts expect
    from fastamericanoptionpricing_b200 import batch as _batch
    from fastamericanoptionpricing_b200 import ChebyshevInterpolation as intrp
    from fastamericanoptionpricing_b200 import EuropeanOptionSolver as europ
    from fastamericanoptionpricing_b200 import QDplusAmericanOptionSolver as qd

try:  # the reference re-exports `plt` through `from FastAmericanOptionSolverBase import *` (testA.py:24-35)
    import matplotlib.pyplot as plt
except Exception:  # matplotlib is optional here; plotting is outside the hot path
    plt = None


def _Phi(x):
    """standard normal cdf of a scalar, evaluated by the device math (faop_dev_math_batch kind 4)"""
    return float(_batch.dev_math_batch(4, [float(x)]).item())


def _phi(x):
    """standard normal pdf of a scalar: exp on the device (kind 0)"""
    return float(_batch.dev_math_batch(0, [-0.5 * float(x) * float(x)]).item()) / math.sqrt(2.0 * math.pi)


class FastAmericanOptionSolver(ABC):
    _SOLVER = "A"

    def __init__(self, riskfree, dividend, volatility, strike, maturity, option_type):
        """reference :13-47"""
        self.r = riskfree
        self.q = dividend
        self.sigma = volatility
        self.K = strike
        self.T = maturity
        self.collocation_num = 12
        self.quadrature_num = 24
        self.integration_num = 2 * self.quadrature_num
        self.max_iters = 200
        self.iter_tol = 1e-5
        self.shared_B0 = []
        self.shared_B = []
        self.shared_B_old = []
        self.shared_tau = []
        self.tau_max = self.T
        self.tau_min = 0
        self.european_price = 0
        self.option_type = option_type
        self.y = [-0.90618, -0.538469, 0, 0.538469, 0.90618]
        self.w = [0.236927, 0.478629, 0.568889, 0.478629, 0.236927]
        self.shared_Bu = [None] * len(self.y)
        self.shared_u = [None] * len(self.y)
        self.iter_records = []
        self.error = 1000000
        self.num_iters = 0
        self.DEBUG = True
        self.use_derivative = False

    # ------------------------------------------------------------------ plumbing
    def _is_put(self):
        return self.option_type == qd.OptionType.Put

    def _cfg(self, **over):
        kw = dict(solver=self._SOLVER, n=int(self.collocation_num), l=int(self.quadrature_num), m=int(self.integration_num),
                  iter_tol=float(self.iter_tol), max_iters=int(self.max_iters), use_derivative=bool(self.use_derivative))
        kw.update(over)
        return _batch.AloConfig(**kw)

    def _opt(self):
        """option scalars for the kernels that only see the collocation domain [0, tau_max] (reference :28, :185, :223)"""
        return (self.r, self.q, self.sigma, self.K, self.tau_max)

    def _tau_max(self):
        return None if self.tau_max == self.T else float(self.tau_max)

    def _nan_solve(self):
        """tau_min != 0 (reference :29): upstream maps u to z = sqrt(4 (u - tau_min)/(tau_max - tau_min)) - 1 (:235-237), which
        meets u < tau_min at the node tau_n = tau_min (and tau_n < 0 when tau_min < 0), so the first sweep produces NaN, the loop
        `while iter_err > tol` stops on it and solve() returns NaN with num_iters = 1 [measured on the reference].  Nothing is
        launched for that; the same state is reported (shared_B all NaN — upstream keeps a few finite nodes when tau_min < 0)."""
        self.set_collocation_points()
        self.set_initial_guess()
        self.shared_B = [np.float64(np.nan)] * (self.collocation_num + 1)
        self.num_iters = 1
        self.error = np.float64(np.nan)
        self.iter_records.append((1, float("nan")))
        return np.float64(np.nan)

    # ------------------------------------------------------------------ the solve
    def solve(self, t, s0):
        """reference :49-76"""
        if self.q == 0 and self.option_type == qd.OptionType.Call:
            self.european_price = europ.EuropeanOption.european_call_value(self.T - t, s0, self.r, self.q, self.sigma, self.K)
            return self.european_price
        if self.tau_min != 0:
            return self._nan_solve()
        self.set_collocation_points()
        self.debug("step 1. checking collocation points ...")
        self.debug("collocation point = {0}".format(self.shared_tau))
        self.debug("step 3. checking numerical integration ...")
        self.test_numerical_integration()
        out = _batch.price_one(self.r, self.q, self.sigma, self.K, self.T, s0, self._is_put(), t=t, cfg=self._cfg(),
                               tau_max=self._tau_max())
        self._store(out)
        self.debug("step 6. checking exercise boundary ...")
        self.debug("exercise boundary = {0}".format(self.shared_B))
        if self.DEBUG:
            self.debug("match condition err = {0}".format(self.check_value_match_condition2()))
        return out["price"]

    def _store(self, out):
        """results of batch.price_one (one packed device-to-host copy) into the reference's attributes"""
        self.shared_B0 = list(out["B0"])
        self.shared_B = list(out["B"])
        self.num_iters = out["iters"]
        self.error = out["err"]
        errs = out["iter_errs"]
        for k in range(self.num_iters):
            self.debug("  iter = {0}, err = {1}".format(k + 1, errs[k]))
            self.iter_records.append((k + 1, float(errs[k])))
        self.european_price = out["european"]

    def compute_exercise_boundary(self):
        """reference :105-133 (seed + damped Jacobi loop); fills shared_B0, shared_B, error, num_iters, iter_records"""
        if self.tau_min != 0:
            self._nan_solve()
            return
        if len(self.shared_tau) != self.collocation_num + 1:
            self.set_collocation_points()
        self.debug("step 5. starting iteration ...")
        out = _batch.price_one(self.r, self.q, self.sigma, self.K, self.T, self.K, self._is_put(), cfg=self._cfg(),
                               tau_max=self._tau_max())
        eur = self.european_price
        self._store(out)
        self.european_price = eur

    def test_numerical_integration(self):
        """reference :78-86 (DEBUG self-check of the quadrature substitution on the seed boundary)"""
        if not self.DEBUG:
            return
        self.set_initial_guess()
        tau, s0 = 3.0, 2
        analy_res = s0 * 0.5 * (np.exp(tau * tau) - 1)
        num_res = self.quadrature_sum(self.test_integrand, tau, s0, self.quadrature_num)
        self.debug("analytical sol = {0}, numerical sol = {1}, err = {2}".format(analy_res, num_res, abs(analy_res - num_res)))

    @staticmethod
    def test_integrand(tau, S, u, Bu):
        return S * u * np.exp(u * u)

    def american_value_with_known_boundary(self, tau, s0, r, q, sigma, K):
        """reference :92-103: European value + the m-point price integral on self.shared_B"""
        if self.tau_min != 0:
            return np.float64(np.nan)
        p, e = _batch.price_known_boundary_batch(np.asarray(self.shared_B, dtype=np.float64), self.r, self.q, self.sigma,
                                                 self.K, self.T, s0, self._is_put(), t=self.T - tau, cfg=self._cfg(),
                                                 tau_max=self._tau_max())
        v12 = p.item() - e.item()
        if (r, q, sigma, K) == (self.r, self.q, self.sigma, self.K):
            v = e.item()
        elif self._is_put():
            v = float(europ.EuropeanOption.european_put_value(tau, s0, r, q, sigma, K))
        else:
            v = float(europ.EuropeanOption.european_call_value(tau, s0, r, q, sigma, K))
        self.european_price = np.float64(v)
        return np.float64(p.item()) if v == e.item() else np.float64(v + v12)

    def iterate_once(self, tau, B):
        """reference :135-142: one Jacobi sweep; every node is updated from the same previous iterate self.shared_B"""
        ev = self._eval(tau, B)
        return list(ev["Bnew"].cpu().numpy()[0])

    def iterate_once_foreach_tau(self, tau_i, B_i):
        """reference :144-165"""
        return np.float64(self._eval([tau_i], [B_i])["Bnew"].item())

    def _eval(self, tau, B, **over):
        return _batch.eval_points_batch(np.asarray(self.shared_B, dtype=np.float64), np.asarray(tau, dtype=np.float64)[None, :],
                                        np.asarray(B, dtype=np.float64)[None, :], *self._opt(), self._is_put(),
                                        cfg=self._cfg(**over))

    def compute_integration_terms(self, tau, num_points):
        """reference :167-195: Gauss-Legendre nodes, u_k and the interpolated boundary B(u_k) (always recomputed)"""
        self.y, self.w = legendre.leggauss(num_points)
        self.shared_u = tau - tau * np.square(1 + self.y) / 4.0
        if self.tau_min != 0:
            self.shared_Bu = np.full(len(self.y), np.nan)
            return
        Bu = _batch.interp_boundary_batch(np.asarray(self.shared_B, dtype=np.float64), self.shared_u[None, :], self.r, self.q,
                                          self.K, self.tau_max, self._is_put(), cfg=self._cfg())
        self.shared_Bu = Bu.cpu().numpy()[0]

    def set_collocation_points(self):
        """reference :221-223"""
        cheby_points = intrp.ChebyshevInterpolation.get_std_cheby_points(self.collocation_num)
        self.shared_tau = self.to_orig_point(cheby_points, self.tau_min, self.tau_max)

    def debug(self, message):
        if self.DEBUG == True:  # noqa: E712  (the reference's exact test)
            print(message)
            print("")

    def norm1_error(self, x1, x2):
        """reference :230-233 (the 2-norm, despite the name)"""
        return alg.norm(np.abs(np.array(x1) - np.array(x2)))

    def to_cheby_point(self, x, x_min, x_max):
        return np.sqrt(4 * (x - x_min) / (x_max - x_min)) - 1

    def to_orig_point(self, c, x_min, x_max):
        return np.square(c + 1) * (x_max - x_min) / 4 + x_min

    def jac(self, a, b, x):
        return 0.5 * (b - a) * (1 + x)

    def B_at_zero(self):
        """reference :246-256"""
        if self.option_type == qd.OptionType.Call:
            return self.K if self.r <= self.q else self.r / self.q * self.K
        return self.K if self.r >= self.q else self.r / self.q * self.K

    def set_initial_guess(self):
        """reference :258-265: QD seed at every collocation node (one kernel launch for all nodes)"""
        tau = np.asarray(self.shared_tau, dtype=np.float64)
        res = _batch.qd_boundary_batch(tau, self.r, self.q, self.sigma, self.K, self._is_put()).cpu().numpy()
        self.shared_B = list(res)
        self.shared_B0 = list(res)

    def compute_f_and_fprime(self, tau_i, B_i):
        """reference :267-279: returns [f, f'] (f' = 1 unless use_derivative)"""
        if tau_i == 0:
            return [self.B_at_zero(), 1]
        ev = self._eval([tau_i], [B_i])
        return [np.float64(ev["f"].item()), np.float64(ev["fprime"].item()) if self.use_derivative else 1]

    def compute_fprime_numerical(self, tau_i, B_i, B_i_old):
        """reference :281-288"""
        if B_i == B_i_old:
            return 0
        return (self.compute_f_and_fprime(tau_i, B_i)[0] - self.compute_f_and_fprime(tau_i, B_i_old)[0]) / (B_i - B_i_old)

    # the four integral forms are evaluated by the kernel of the subclass' solver id
    def N_func(self, tau, B):
        return np.float64(self._eval([tau], [B])["N"].item())

    def D_func(self, tau, B):
        return np.float64(self._eval([tau], [B])["D"].item())

    def Nprime_func(self, tau, Q):
        return np.float64(self._eval([max(1e-10, tau)], [Q], use_derivative=True)["Nprime"].item())

    def Dprime_func(self, tau, Q):
        return np.float64(self._eval([max(1e-10, tau)], [Q], use_derivative=True)["Dprime"].item())

    # ------------------------------------------------------------------ scalar helpers (interactive inspection only;
    # the solve path never calls them — the kernels carry their own fused versions)
    def dminus(self, tau, z):
        return (np.log(z) + (self.r - self.q) * tau - 0.5 * self.sigma * self.sigma * tau) / (self.sigma * np.sqrt(tau))

    def dplus(self, tau, z):
        return (np.log(z) + (self.r - self.q) * tau + 0.5 * self.sigma * self.sigma * tau) / (self.sigma * np.sqrt(tau))

    def CDF_neg_dminus(self, tau, z):
        if tau == 0:
            return 0 if z > 1 else 1
        return _Phi(-self.dminus(tau, z))

    def CDF_pos_dminus(self, tau, z):
        if tau == 0:
            return 1 if z > 1 else 0
        return _Phi(self.dminus(tau, z))

    def CDF_neg_dplus(self, tau, z):
        if tau == 0:
            return 0 if z > 1 else 1
        return _Phi(-self.dplus(tau, z))

    def CDF_pos_dplus(self, tau, z):
        if tau == 0:
            return 1 if z > 1 else 0
        return _Phi(self.dplus(tau, z))

    def PDF_dminus(self, tau, z):
        return 0 if tau == 0 else _phi(self.dminus(tau, z))

    def PDF_dplus(self, tau, z):
        return 0 if tau == 0 else _phi(self.dplus(tau, z))

    PDF_neg_dminus = PDF_dminus
    PDF_neg_dplus = PDF_dplus

    def v_integrand_1(self, tau, S, u, Bu):
        if self._is_put():
            return self.r * self.K * np.exp(-self.r * (tau - u)) * self.CDF_neg_dminus(tau - u, S / Bu)
        return self.q * S * np.exp(-self.q * (tau - u)) * self.CDF_pos_dplus(tau - u, S / Bu)

    def v_integrand_2(self, tau, S, u, Bu):
        if self._is_put():
            return self.q * S * np.exp(-self.q * (tau - u)) * self.CDF_neg_dplus(tau - u, S / Bu)
        return self.r * self.K * np.exp(-self.r * (tau - u)) * self.CDF_pos_dminus(tau - u, S / Bu)

    def v_integrand_12(self, tau, S, u, Bu):
        if self._is_put():
            return self.v_integrand_1(tau, S, u, Bu) - self.v_integrand_2(tau, S, u, Bu)
        return self.v_integrand_1(tau, S, u, Bu) - self.v_integrand_2(tau, S, u, Bu)

    def quadrature_sum(self, integrand, tau, S, num_points):
        """reference :372-388: sum of a HOST callable over the rule; u_k and B(u_k) come from the device"""
        self.compute_integration_terms(tau, num_points)
        u, Bu = self.shared_u, self.shared_Bu
        assert len(u) == len(Bu) and len(u) == len(self.w)
        if tau == 0:
            return 0
        ans = 0
        for i in range(len(u)):
            ans += integrand(tau, S, u[i], Bu[i]) * self.w[i] * self.jac(0, tau, self.y[i])
        return ans

    # ------------------------------------------------------------------ diagnostics (reference :390-447)
    def check_value_match_condition1(self):
        left, right = [], []
        for tau_i, B_i in zip(self.shared_tau, self.shared_B):
            left.append(self.K - B_i if self._is_put() else B_i - self.K)
            right.append(self.american_value_with_known_boundary(tau_i, B_i, self.r, self.q, self.sigma, self.K))
        return self.norm1_error(left, right)

    def check_value_match_condition2(self):
        tau = np.asarray(self.shared_tau, dtype=np.float64)
        B = np.asarray(self.shared_B, dtype=np.float64)
        act = tau != 0
        ev = self._eval(tau[act], B[act])
        Nv, Dv = ev["N"].cpu().numpy()[0], ev["D"].cpu().numpy()[0]
        left = Nv * self.K * np.exp(-self.r * tau[act])
        right = Dv * B[act] * np.exp(-self.q * tau[act])
        return self.norm1_error(left, right)

    def check_value_match_condition3(self):
        tau = np.asarray(self.shared_tau, dtype=np.float64)
        B = np.asarray(self.shared_B, dtype=np.float64)
        f = self._eval(tau, B)["f"].cpu().numpy()[0]
        return self.norm1_error(B, f)

    def check_f_with_B(self, B=np.linspace(50, 150, 30)):
        """reference :427-447 without the plot/exit(): returns (B, f, fprime) at tau = 0.2"""
        if self.option_type == qd.OptionType.Call and self.q == 0:
            return
        ev = self._eval(np.full(len(B), 0.2), B, use_derivative=True)
        return B, ev["f"].cpu().numpy()[0], ev["fprime"].cpu().numpy()[0]
