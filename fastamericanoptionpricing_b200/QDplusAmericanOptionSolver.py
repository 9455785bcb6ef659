"""Drop-in for the reference's QDplusAmericanOptionSolver.py: OptionType and QDplus with the same constructor and
methods; roots, residuals and the QD+ price come from the CUDA leaf kernels (faop_qd_boundary_batch,
faop_qd_residual_batch, faop_qd_price_batch).  The boundary root uses a safeguarded Newton iteration from x0 = K
on the device instead of MINPACK hybr (reference :73)."""
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
