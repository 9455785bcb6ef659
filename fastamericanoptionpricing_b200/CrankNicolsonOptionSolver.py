"""Drop-in for the reference's CrankNicolsonOptionSolver.py: CNOptionSolver with the same constructor, knobs
(N, max_dt, USE_PSOR, tol, max_iter, omega, Smax), state (S, X, A, b, err, iter) and methods; all arithmetic runs in the
CUDA kernels behind faop_cn_put_batch / faop_cn_put_countdown_batch / faop_cn_rows_batch.
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
        self.A = []
        self.b = []
        self.N = 200
        self.max_dt = 1 / 12.0
        self.USE_PSOR = False
        self.USE_PROJECTED_THOMAS = False
        self.tol = 1e-5
        self.max_iter = 200
        self.omega = 1.2
        self.cached_dt = 0
        self.err = 0
        self.iter = 0

    def _mode(self):
        return _batch.CN_MODE_PROJECTED if self.USE_PROJECTED_THOMAS else (_batch.CN_MODE_SOR if self.USE_PSOR else _batch.CN_MODE_DIRECT)

    def _args(self):
        return (self.r, self.q, self.sigma, self.K)

    def _kernel(self, S0, mode, T=None, schedule=None, X_in=None):
        (price, X), err, it = _batch.cn_put_batch(*self._args(), self.T if T is None else T, S0, n_space=self.N,
                                                  max_dt=self.max_dt, mode=mode, omega=self.omega, tol=self.tol,
                                                  max_iter=self.max_iter, want_grid=True, Smax=self.Smax, X_in=X_in,
                                                  schedule=schedule, want_state=True)
        if mode == _batch.CN_MODE_SOR:
            self.err = np.float64(err.item())       # :103-104
            self.iter = int(it.item())
        return np.float64(price.item()), X.cpu().numpy()[0]

    def solve(self, S0):
        """reference :27-34: returns np.interp(S0, S, X) of the final grid; self.S / self.X hold the grid.  (Upstream prints
        t / err / iters after every time step, :45; the whole countdown is one kernel here and nothing is printed.)"""
        self.setInitialCondition()
        price, X = self._kernel(S0, self._mode())
        self.X = X if self.USE_PSOR or self.USE_PROJECTED_THOMAS else X.reshape(-1, 1)     # np.linalg.solve leaves (N, 1), :82
        return price

    def setInitialCondition(self):
        """reference :47-52: the grid, the payoff and the empty system (host arrays; the kernels build their own copies)"""
        self.S = np.linspace(0, self.Smax, self.N)
        self.A = np.zeros((self.N, self.N))
        self.b = np.zeros((self.N, 1))
        self.X = np.maximum(self.K - self.S, 0)

    def solvePDE(self):
        """reference :36-45: the countdown from T in steps of max_dt, starting from whatever self.X holds"""
        X0 = np.asarray(self.X, dtype=np.float64).reshape(-1)
        _, X = self._kernel(self.K, self._mode(), X_in=X0[None, :])
        self.X = X if self.USE_PSOR or self.USE_PROJECTED_THOMAS else X.reshape(-1, 1)

    def setCoeff(self, dt):
        """reference :55-79: fills self.A (dense, N x N) and self.b for one step of length dt from self.X.  The entries come
        from faop_cn_rows_batch; placing the three diagonals into the dense matrix is the only host work."""
        N = self.N
        X0 = np.asarray(self.X, dtype=np.float64).reshape(-1)
        lo, di, up, b = (t.cpu().numpy()[0] for t in _batch.cn_rows_batch(*self._args(), dt, X0[None, :], Smax=self.Smax))
        if not isinstance(self.A, np.ndarray) or self.A.shape != (N, N):
            self.A = np.zeros((N, N))
        i = np.arange(N - 1)
        self.A[i, i] = di[:N - 1]
        self.A[i[1:], i[1:] - 1] = lo[1:N - 1]
        self.A[i[1:], i[1:] + 1] = up[1:N - 1]
        self.A[-1, N - 4:N] = (-1, 4, -5, 2)
        self.b = b.reshape(N, 1).copy()
        self.b[-1] = 0
        self.cached_dt = dt

    def _one_step(self, mode):
        X0 = np.asarray(self.X, dtype=np.float64).reshape(-1)
        sched = (np.array([[self.cached_dt]], dtype=np.float64), np.array([1], dtype=np.int32))
        _, X = self._kernel(self.K, mode, schedule=sched, X_in=X0[None, :])
        return X

    def solveLinearSystem(self):
        """reference :81-82: one unprojected step of the length last given to setCoeff; X becomes (N, 1) as np.linalg.solve leaves it"""
        self.X = self._one_step(_batch.CN_MODE_DIRECT).reshape(-1, 1)

    def solvePSOR(self):
        """reference :84-104: one SOR-then-project solve of the step last given to setCoeff; sets self.err / self.iter"""
        shape = np.shape(self.X)
        self.X = self._one_step(_batch.CN_MODE_SOR).reshape(shape)

    def applyConstraint(self):
        """reference :106-107"""
        self.X = np.maximum(self.X, self.K - self.S)
