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

    def Nprime_integrand(self, tau, B_tau, u, B_u):
        """reference :59-66 (host scalar helper; the pdf is even, so the put and call branches coincide)"""
        tau = max(1e-10, tau)
        return np.exp(self.r * u) * self.PDF_dminus(tau - u, B_tau / B_u) / (self.sigma * np.sqrt(tau - u) * B_tau)

    def Dprime_integrand(self, tau, B_tau, u, B_u):
        """reference :68-75; the call branch upstream uses the d- density (kept)"""
        tau = max(1e-10, tau)
        pdf = self.PDF_dplus if self._is_put() else self.PDF_dminus
        return np.exp(self.q * u) * pdf(tau - u, B_tau / B_u) / (self.sigma * np.sqrt(tau - u) * B_tau)
