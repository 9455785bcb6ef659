"""Drop-in for the reference's EuropeanOptionSolver.py: same class, static methods and argument order
(tau, s0, r, q, vol, strike); the arithmetic runs in the CUDA leaf kernels (faop_european_batch,
faop_european_greeks_batch).  Scalars in -> numpy float out; array arguments broadcast -> numpy array out."""
import numpy as np

try:
    from . import batch as _batch
except ImportError:  # flat import (package directory on sys.path, as the reference's scripts expect)
    from fastamericanoptionpricing_b200 import batch as _batch


def _ret(t, args):
    a = t.cpu().numpy()
    if all(np.ndim(x) == 0 for x in args):
        return np.float64(a[0])
    return a.reshape(np.broadcast(*args).shape)


def _flat(args):
    bs = np.broadcast_arrays(*[np.asarray(x, dtype=np.float64) for x in args])
    return [np.ascontiguousarray(b).reshape(-1) for b in bs]


class EuropeanOption:
    @staticmethod
    def european_put_value(tau, s0, r, q, vol, strike):
        """reference EuropeanOptionSolver.py:7-13"""
        a = (tau, s0, r, q, vol, strike)
        return _ret(_batch.european_batch(*_flat(a), True), a)

    @staticmethod
    def european_call_value(tau, s0, r, q, vol, strike):
        """reference EuropeanOptionSolver.py:16-22"""
        a = (tau, s0, r, q, vol, strike)
        return _ret(_batch.european_batch(*_flat(a), False), a)

    @staticmethod
    def european_option_theta(tau, s0, r, q, vol, strike):
        """reference EuropeanOptionSolver.py:25-32 (put theta)"""
        a = (tau, s0, r, q, vol, strike)
        return _ret(_batch.european_greeks_batch(*_flat(a))[0], a)

    @staticmethod
    def d1(tau, s0, r, q, vol, strike):
        """reference EuropeanOptionSolver.py:35-36"""
        a = (tau, s0, r, q, vol, strike)
        return _ret(_batch.european_greeks_batch(*_flat(a))[1], a)

    @staticmethod
    def d2(tau, s0, r, q, vol, strike):
        """reference EuropeanOptionSolver.py:39-40"""
        a = (tau, s0, r, q, vol, strike)
        return _ret(_batch.european_greeks_batch(*_flat(a))[2], a)
