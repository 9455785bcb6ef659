"""Drop-in for the reference's ChebyshevInterpolation.py: coefficient fit and Clenshaw evaluation run in the
CUDA leaf kernels faop_cheb_fit_batch / faop_cheb_eval_batch."""
import numpy as np

try:
    from . import batch as _batch
except ImportError:
    from fastamericanoptionpricing_b200 import batch as _batch


class ChebyshevInterpolation:
    def __init__(self, ynodes, x_to_cheby, x_min, x_max):
        """reference ChebyshevInterpolation.py:4-12"""
        y = np.asarray(ynodes, dtype=np.float64).reshape(-1)
        self.n = len(y) - 1
        self.x_to_cheby = x_to_cheby
        self.x_min = x_min
        self.x_max = x_max
        self.a = list(_batch.cheb_fit_batch(y[None, :]).cpu().numpy()[0])

    @staticmethod
    def get_std_cheby_points(numpoints):
        """reference ChebyshevInterpolation.py:15-17 (an option-independent table, built on the host)"""
        i = np.arange(0, numpoints + 1)
        return np.cos(i * np.pi / numpoints)

    def value(self, zv):
        """reference ChebyshevInterpolation.py:19-29: maps zv with x_to_cheby (a host callable) and evaluates; returns a list"""
        to_cheby = zv
        if self.x_to_cheby is not None:
            to_cheby = self.x_to_cheby(zv, self.x_min, self.x_max)
        z = np.asarray(to_cheby, dtype=np.float64).reshape(-1)
        out = _batch.cheb_eval_batch(np.asarray(self.a, dtype=np.float64)[None, :], z[None, :]).cpu().numpy()[0]
        return list(out)

    def std_cheby_single_value(self, z):
        """reference ChebyshevInterpolation.py:31-41"""
        return self.value_std(np.array([z]))[0]

    def value_std(self, z):
        z = np.asarray(z, dtype=np.float64).reshape(-1)
        return _batch.cheb_eval_batch(np.asarray(self.a, dtype=np.float64)[None, :], z[None, :]).cpu().numpy()[0]

    def std_coeff(self, k, n_y):
        """reference ChebyshevInterpolation.py:43-51"""
        return _batch.cheb_fit_batch(np.asarray(n_y, dtype=np.float64)[None, :]).cpu().numpy()[0][k]
