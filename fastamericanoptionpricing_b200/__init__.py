"""B200-native batched American option pricer: the spectral-collocation (Andersen-Lake-Offengenden) hot path of
antdvid/FastAmericanOptionPricing as hand-written sm_100a CUDA behind a C ABI (include/faop.h).

Two surfaces:
  * batch API (fastamericanoptionpricing_b200.batch): torch tensors in/out, one warp per option on the GPU;
  * drop-in modules with the reference's own module and class names (FastAmericanOptionSolverA/B,
    QDplusAmericanOptionSolver, EuropeanOptionSolver, ChebyshevInterpolation, CrankNicolsonOptionSolver):
    scalar `solve()` calls are batches of one through the same kernels.  Put this directory on sys.path to keep
    the reference's flat imports (`from FastAmericanOptionSolverA import *`) working unchanged.

There is no CPU fallback: without a CUDA device or without csrc/libfaop.so every compute call raises.
"""
from .batch import (AloConfig, HostContext, MultiGpuHost, bjerksund_stensland_batch, bs_phi_batch, cheb_eval_batch, cheb_fit_batch, cn_put_batch, eval_nodes_batch,  # noqa: F401
                    eval_points_batch, european_batch, european_greeks_batch, greeks_batch, implied_vol_batch, interp_boundary_batch, ladder_batch, price_batch,
                    price_known_boundary_batch, qd_boundary_batch, qd_price_batch, qd_residual_batch, surface_batch)
from .sharding import ladder_batch_sharded, price_batch_sharded, shard_bounds  # noqa: F401

__version__ = "0.1.0"
