"""Drop-in for the reference's CrankNicolsonOptionSolver.py: CNOptionSolver with the same constructor and knobs
(N, max_dt, USE_PSOR, tol, max_iter, omega, Smax); the time stepping runs in the CUDA kernel faop_cn_put_batch.
USE_PSOR=False reproduces the reference's default dense solve (no early-exercise projection, :81-82),
USE_PSOR=True its SOR-sweep-then-project loop (:84-107).  `USE_PROJECTED_THOMAS=True` (not in the reference)
selects the Brennan-Schwartz projected tridiagonal sweep used for throughput."""
import numpy as np

try:
    from . import batch as _batch
except ImportError:
    from fastamericanoptionpricing_b200 import batch as _batch


class CNOptionSolver:
    def __init__(self, riskfree, dividend, volatility, strike, maturity):
        """reference :6-25"""
        self.r = riskfree
        self.q = dividend
        self.sigma = volatility
        self.K = strike
        self.T = maturity
        self.Smax = 3 * strike
        self.S = []
        self.X = []
        self.N = 200
        self.max_dt = 1 / 12.0
        self.USE_PSOR = False
        self.USE_PROJECTED_THOMAS = False
        self.tol = 1e-5
        self.max_iter = 200
        self.omega = 1.2
        self.err = 0
        self.iter = 0

    def _run(self, S0):
        if self.Smax != 3 * self.K:
            raise NotImplementedError("the CUDA kernel fixes Smax = 3 K as the reference constructor does")
        mode = _batch.CN_MODE_PROJECTED if self.USE_PROJECTED_THOMAS else (_batch.CN_MODE_SOR if self.USE_PSOR else _batch.CN_MODE_DIRECT)
        price, X = _batch.cn_put_batch(self.r, self.q, self.sigma, self.K, self.T, S0, n_space=self.N, max_dt=self.max_dt,
                                       mode=mode, omega=self.omega, tol=self.tol, max_iter=self.max_iter, want_grid=True)
        self.S = np.linspace(0, self.Smax, self.N)
        self.X = X.cpu().numpy()[0]
        return np.float64(price.item())

    def solve(self, S0):
        """reference :27-34: returns np.interp(S0, S, X) of the final grid; self.S / self.X hold the grid.  (Upstream prints
        t / err / iters after every time step, :45; the whole countdown is one kernel here and nothing is printed.)"""
        return self._run(S0)

    def setInitialCondition(self):
        """reference :47-52: the grid and the payoff (host arrays; the kernel builds its own copy of both)"""
        self.S = np.linspace(0, self.Smax, self.N)
        self.X = np.maximum(self.K - self.S, 0)

    def solvePDE(self):
        """reference :36-45: the whole countdown from the payoff — the state the reference has right after setInitialCondition(),
        which is the only state its own solve() ever starts from; fills self.X"""
        self._run(self.K)

    def applyConstraint(self):
        """reference :106-107"""
        self.X = np.maximum(self.X, self.K - self.S)

    def _per_step(self, *a, **k):
        raise NotImplementedError("setCoeff / solveLinearSystem / solvePSOR are single time steps on the reference's dense N x N "
                                  "matrix; here the matrix never exists and all steps run inside one kernel (faop_cn_put_countdown_batch)")

    setCoeff = solveLinearSystem = solvePSOR = _per_step
