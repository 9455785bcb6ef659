"""Drop-in for the reference's FastAmericanOptionSolverB.py.  N/D use the cdf-only form
(reference FastAmericanOptionSolverB.py:5-39, derivative terms :41-75, call-side quirk of :57 kept), evaluated by
the solver-1 CUDA kernels."""
try:
    from .FastAmericanOptionSolverBase import *  # noqa: F401,F403
    from .FastAmericanOptionSolverBase import FastAmericanOptionSolver
except ImportError:
    from FastAmericanOptionSolverBase import *  # noqa: F401,F403
    from FastAmericanOptionSolverBase import FastAmericanOptionSolver


class FastAmericanOptionSolverB(FastAmericanOptionSolver):
    _SOLVER = "B"

    def N_integrand(self, tau, B_tau, u, B_u):
        """reference :25-31 (host scalar helper)"""
        if self._is_put():
            return np.exp(self.r * u) * self.CDF_pos_dminus(tau - u, B_tau / B_u)
        return np.exp(self.r * u) * self.CDF_neg_dminus(tau - u, B_tau / B_u)

    def D_integrand(self, tau, B_tau, u, B_u):
        """reference :33-39 (host scalar helper)"""
        if self._is_put():
            return np.exp(self.q * u) * self.CDF_pos_dplus(tau - u, B_tau / B_u)
        return np.exp(self.q * u) * self.CDF_neg_dplus(tau - u, B_tau / B_u)
