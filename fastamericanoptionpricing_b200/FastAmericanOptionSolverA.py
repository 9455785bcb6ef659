"""Drop-in for the reference's FastAmericanOptionSolverA.py (`from FastAmericanOptionSolverA import *` re-exports
qd, np, plt, europ, intrp like the reference's star-import chain).  N/D use the pdf-heavy K12/K3 form
(reference FastAmericanOptionSolverA.py:5-44, derivative terms :46-83), evaluated by the solver-0 CUDA kernels."""
try:
    from .FastAmericanOptionSolverBase import *  # noqa: F401,F403
    from .FastAmericanOptionSolverBase import FastAmericanOptionSolver
except ImportError:
    from FastAmericanOptionSolverBase import *  # noqa: F401,F403
    from FastAmericanOptionSolverBase import FastAmericanOptionSolver


class FastAmericanOptionSolverA(FastAmericanOptionSolver):
    _SOLVER = "A"

    def K12(self, tau, B):
        """reference :24-25 — the quadrature runs in the CUDA evaluation kernel (faop_alo_eval_points_batch, I1)"""
        return np.float64(self._eval([tau], [B])["I1"].item())

    def K3(self, tau, B):
        """reference :34-35 (I2 of the evaluation kernel)"""
        return np.float64(self._eval([tau], [B])["I2"].item())

    def K12_integrand(self, tau, B_tau, u, B_u):
        """reference :27-32 (host scalar helper)"""
        if self._is_put():
            return np.exp(self.q * u) * self.CDF_pos_dplus(tau - u, B_tau / B_u) \
                + np.exp(self.q * u) / (self.sigma * np.sqrt(tau - u)) * self.PDF_dplus(tau - u, B_tau / B_u)
        return -np.exp(self.q * u) * self.CDF_neg_dplus(tau - u, B_tau / B_u) \
            + np.exp(self.q * u) / (self.sigma * np.sqrt(tau - u)) * self.PDF_neg_dplus(tau - u, B_tau / B_u)

    def K3_integrand(self, tau, B_tau, u, B_u):
        """reference :37-44 (host scalar helper)"""
        return np.exp(self.r * u) / (self.sigma * np.sqrt(tau - u)) * self.PDF_dminus(tau - u, B_tau / B_u)

    def Nprime_integrand(self, tau, B_tau, u, B_u):
        """reference :55-63 (host scalar helper; the pdf is even, so the put and call branches coincide)"""
        tau = max(1e-10, tau)
        s_ = tau - u
        z = B_tau / B_u
        return np.exp(self.r * u) * self.dminus(s_, z) / (B_tau * self.sigma * self.sigma * s_) * self.PDF_dminus(s_, z)

    def Dprime_integrand(self, tau, B_tau, u, B_u):
        """reference :76-83 (host scalar helper)"""
        tau = max(1e-10, tau)
        s_ = tau - u
        z = B_tau / B_u
        sv = self.sigma * np.sqrt(s_)
        return np.exp(self.q * u) * self.PDF_dplus(s_, z) / (sv * B_tau) * (1 - self.dplus(s_, z) / sv)
