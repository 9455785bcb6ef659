"""Drop-in for the reference's QDplusAmericanOptionSolver.py: OptionType and QDplus with the same constructor and
methods; roots, residuals and the QD+ price come from the CUDA leaf kernels (faop_qd_boundary_batch,
faop_qd_residual_batch, faop_qd_price_batch).  The boundary root is found on the device by the reference's own root finder —
MINPACK hybrd as scipy.optimize.root(hybr, x0=K) drives it (reference :73), restated for one unknown in csrc/faop_seed.cu."""
from enum import Enum

import numpy as np

try:
    from . import batch as _batch
    from . import EuropeanOptionSolver as europ
except ImportError:
    from fastamericanoptionpricing_b200 import batch as _batch
    from fastamericanoptionpricing_b200 import EuropeanOptionSolver as europ


class OptionType(Enum):
    Call = 1
    Put = 2


class QDplus:
    """QD+ alogrithm for computing approximated american option price (reference QDplusAmericanOptionSolver.py:12-218)"""

    def __init__(self, riskfree, dividend, volatility, strike, option_type):
        self.r = riskfree
        self.q = dividend
        self.sigma = volatility
        self.K = strike
        self.option_type = option_type
        self.option_indicator = 1 if option_type == OptionType.Call else -1
        self.exercise_boundary = 0
        self.tolerance = 1e-10

    def _put(self):
        return self.option_type == OptionType.Put

    def price(self, tau, S):
        """reference :44-67 (prints the boundary residual `err`, as the reference does)"""
        if tau == 0:
            self.exercise_boundary = self.K
            return max(S - self.K, 0.0)
        P, Bd = _batch.qd_price_batch(tau, S, self.r, self.q, self.sigma, self.K, self._put())
        self.exercise_boundary = float(Bd.item())
        err = self.exercise_boundary_func(self.exercise_boundary, tau)
        print("err = ", err)
        return np.float64(P.item())

    def compute_exercise_boundary(self, tau):
        """reference :69-76"""
        if tau == 0:
            return self.B_at_zero()
        return np.float64(_batch.qd_boundary_batch(tau, self.r, self.q, self.sigma, self.K, self._put()).item())

    def B_at_zero(self):
        """reference :78-88"""
        if self.option_type == OptionType.Call:
            return self.K if self.r <= self.q else self.r / self.q * self.K
        return self.K if self.r >= self.q else self.r / self.q * self.K

    def exercise_boundary_func(self, S, tau):
        """reference :110-125; accepts a float or a numpy array of spots (test_QDplus.py:20-21)"""
        if tau == 0:
            return 0 if isinstance(S, float) else np.ones(np.size(S)) * 0
        Sa = np.asarray(S, dtype=np.float64)
        F = _batch.qd_residual_batch(Sa.reshape(-1), tau, self.r, self.q, self.sigma, self.K, self._put()).cpu().numpy()
        return np.float64(F[0]) if Sa.ndim == 0 else F.reshape(Sa.shape)

    # ------------------------------------------------------------------ internals of price() (reference :90-218), host scalars.
    # price() itself runs in faop_qd_price_batch; these are kept, with the reference's signatures, for callers that poke at the
    # pieces.  compute_miscellaneous fills the same v_* attributes in the same order (European legs through the device leaf kernels).
    def compute_miscellaneous(self, tau, S):
        """reference :90-107"""
        E = europ.EuropeanOption
        self.v_N = self.N()
        self.v_M = self.M()
        self.v_h = self.h(tau)
        self.v_qQD = self.q_QD(tau)
        self.v_qQDdot = self.q_QD_dot()
        self.v_d1 = E.d1(tau, S, self.r, self.q, self.sigma, self.K)
        self.v_d2 = E.d2(tau, S, self.r, self.q, self.sigma, self.K)
        self.v_p = E.european_put_value(tau, S, self.r, self.q, self.sigma, self.K) if self._put() \
            else E.european_call_value(tau, S, self.r, self.q, self.sigma, self.K)
        self.v_theta = E.european_option_theta(tau, S, self.r, self.q, self.sigma, self.K)
        self.v_dlogSdh = self.dlogSdh(tau, S)
        self.v_c = self.c(tau, S)
        self.v_c0 = self.c0(tau, S)
        self.v_b = self.b(tau, S)

    def q_QD(self, tau):
        """reference :127-134"""
        root = np.sqrt((self.v_N - 1) * (self.v_N - 1) + 4 * self.v_M / self.v_h)
        return -0.5 * (self.v_N - 1) + 0.5 * self.option_indicator * root

    def q_QD_dot(self):
        """reference :136-140"""
        return self.v_M / (self.v_h * self.v_h * np.sqrt((self.v_N - 1) * (self.v_N - 1) + 4 * self.v_M / self.v_h))

    def c0(self, tau, S):
        """reference :142-159"""
        den = 2 * self.v_qQD + self.v_N - 1
        return -(1 - self.v_h) * self.v_M / den * (1 / self.v_h - self.v_theta * np.exp(self.r * tau) / (self.r * (self.K - S - self.v_p))
                                                   + self.v_qQDdot / den)

    def c(self, tau, S):
        """reference :161-174 (Phi(-d1) for both option types, as upstream)"""
        from scipy.special import ndtr
        den = 2 * self.v_qQD + self.v_N - 1
        return self.c0(tau, S) - ((1 - self.v_h) * self.v_M) / den * (
            (1 - np.exp(-self.q * tau) * ndtr(-self.v_d1)) / (self.K - S - self.v_p) + self.v_qQD / S) * self.v_dlogSdh

    def b(self, tau, S):
        """reference :176-183"""
        return ((1 - self.v_h) * self.v_M * self.v_qQDdot) / (2 * (2 * self.v_qQD + self.v_N - 1))

    def dlogSdh(self, tau, S):
        """reference :185-209"""
        from scipy.special import ndtr
        r, q, sg, K = self.r, self.q, self.sigma, self.K
        pdf = np.exp(-0.5 * self.v_d1 * self.v_d1) / np.sqrt(2 * np.pi)
        eq = np.exp(-q * tau)
        dFdh = self.v_qQD * self.v_theta * np.exp(r * tau) / r + self.v_qQDdot * (K - S - self.v_p) \
            + (S * q * eq * ndtr(-self.v_d1)) / (r * (1 - self.v_h)) \
            - (S * eq * pdf) / (2 * r * tau * (1 - self.v_h)) * (2 * np.log(S / K) / (sg * np.sqrt(tau)) - self.v_d1)
        dFdS = (1 - self.v_qQD) * (1 - eq * ndtr(-self.v_d1)) + (eq * pdf) / (sg * np.sqrt(tau))
        return -dFdh / dFdS

    def h(self, tau):
        """reference :211-212"""
        return 1 - np.exp(-self.r * tau)

    def M(self):
        """reference :214-215"""
        return 2 * self.r / (self.sigma * self.sigma)

    def N(self):
        """reference :217-218"""
        return 2 * (self.r - self.q) / (self.sigma * self.sigma)
