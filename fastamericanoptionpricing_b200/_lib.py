"""ctypes binding of libfaop.so (include/faop.h).  The CUDA library is the only compute path:
if it is missing or fails to load this module raises — there is no CPU or PyTorch fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfaop.so")

FLAG_WS_SCRATCH = 1
FLAG_SEED_NEWTON = 2
c_i32, c_i64, c_dbl, c_vp, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t


class FaopCfg(ctypes.Structure):
    """Mirror of `faop_cfg` (include/faop.h)."""
    _fields_ = [("solver", c_i32), ("use_derivative", c_i32), ("n_colloc", c_i32), ("n_quad", c_i32),
                ("n_price_quad", c_i32), ("max_iters", c_i32), ("iter_tol", c_dbl), ("eta", c_dbl),
                ("kernel", c_i32), ("flags", c_i32), ("tau_max", c_vp)]


class FaopSurface(ctypes.Structure):
    """Mirror of `faop_surface` (include/faop.h)."""
    _fields_ = [("K0", c_dbl), ("K1", c_dbl), ("T0", c_dbl), ("T1", c_dbl), ("V0", c_dbl), ("V1", c_dbl),
                ("S", c_dbl), ("r", c_dbl), ("q", c_dbl), ("nK", c_i32), ("nT", c_i32), ("nV", c_i32), ("is_put", c_i32)]


class FaopCnExtra(ctypes.Structure):
    """Mirror of `faop_cn_extra` (include/faop.h): CNOptionSolver.Smax per option, a start state, err / iter outputs."""
    _fields_ = [("Smax", c_vp), ("X_in", c_vp), ("err_out", c_vp), ("iter_out", c_vp)]


_P = ctypes.POINTER
_SIGNATURES = {
    "faop_version": (ctypes.c_int, []),
    "faop_launch_count": (c_i64, []),
    "faop_tables_bytes": (ctypes.c_int, [_P(FaopCfg), _P(c_sz)]),
    "faop_tables_build": (ctypes.c_int, [_P(FaopCfg), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "faop_gauss_legendre_host": (ctypes.c_int, [c_i32, c_vp, c_vp]),
    "faop_workspace_bytes": (ctypes.c_int, [_P(FaopCfg), c_i64, _P(c_sz)]),
    "faop_alo_price_batch": (ctypes.c_int, [_P(FaopCfg), c_i64] + [c_vp] * 19),
    "faop_alo_surface_batch": (ctypes.c_int, [_P(FaopCfg), _P(FaopSurface), c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "faop_alo_surface_dedup_workspace_bytes": (ctypes.c_int, [_P(FaopCfg), _P(FaopSurface), c_i64, c_i64, _P(c_sz)]),
    "faop_alo_surface_dedup_batch": (ctypes.c_int, [_P(FaopCfg), _P(FaopSurface), c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "faop_alo_ladder_workspace_bytes": (ctypes.c_int, [_P(FaopCfg), c_i64, _P(c_sz)]),
    "faop_alo_ladder_batch": (ctypes.c_int, [_P(FaopCfg), c_i64] + [c_vp] * 6 + [c_i32, c_vp, c_dbl, c_dbl] + [c_vp] * 4 + [c_sz, c_vp]),
    "faop_alo_eval_points_batch": (ctypes.c_int, [_P(FaopCfg), c_i64, c_i32] + [c_vp] * 20),
    "faop_alo_interp_boundary_batch": (ctypes.c_int, [_P(FaopCfg), c_i64, c_i32] + [c_vp] * 10),
    "faop_alo_price_known_boundary_batch": (ctypes.c_int, [_P(FaopCfg), c_i64] + [c_vp] * 13),
    "faop_alo_greeks_known_boundary_batch": (ctypes.c_int, [_P(FaopCfg), c_i64] + [c_vp] * 10 + [c_dbl, c_dbl] + [c_vp] * 5),
    "faop_iv_update_batch": (ctypes.c_int, [c_i64] + [c_vp] * 9),
    "faop_qd_boundary_batch": (ctypes.c_int, [c_i64] + [c_vp] * 8),
    "faop_qd_residual_batch": (ctypes.c_int, [c_i64] + [c_vp] * 9),
    "faop_qd_price_batch": (ctypes.c_int, [c_i64] + [c_vp] * 10),
    "faop_bjerksund_stensland_batch": (ctypes.c_int, [c_i64] + [c_vp] * 12),
    "faop_bs_phi_batch": (ctypes.c_int, [c_i64] + [c_vp] * 10),
    "faop_european_batch": (ctypes.c_int, [c_i64] + [c_vp] * 9),
    "faop_european_greeks_batch": (ctypes.c_int, [c_i64] + [c_vp] * 10),
    "faop_cheb_fit_batch": (ctypes.c_int, [c_i64, c_i32, c_vp, c_vp, c_vp]),
    "faop_cheb_eval_batch": (ctypes.c_int, [c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "faop_cn_workspace_bytes": (ctypes.c_int, [c_i64, c_i32, c_i32, _P(c_sz)]),
    "faop_cn_put_batch": (ctypes.c_int, [c_i64, c_i32, c_i32] + [c_vp] * 7 + [c_i32, c_dbl, c_dbl, c_i32, _P(FaopCnExtra)] + [c_vp] * 4),
    "faop_cn_put_countdown_batch": (ctypes.c_int, [c_i64, c_i32, c_i32] + [c_vp] * 7 + [c_dbl, c_dbl, c_i32, _P(FaopCnExtra)] + [c_vp] * 4),
    "faop_cn_rows_batch": (ctypes.c_int, [c_i64, c_i32] + [c_vp] * 12),
    "faop_ctx_create": (ctypes.c_int, [ctypes.c_int, _P(c_vp)]),
    "faop_ctx_destroy": (ctypes.c_int, [c_vp]),
    "faop_alo_price_batch_host": (ctypes.c_int, [c_vp, _P(FaopCfg), c_i64] + [c_vp] * 19),
    "faop_alo_ladder_batch_host": (ctypes.c_int, [c_vp, _P(FaopCfg), c_i64] + [c_vp] * 6 + [c_i32] + [c_vp] * 7),
    "faop_cn_put_countdown_batch_host": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32] + [c_vp] * 10),
    "faop_dev_math_batch": (ctypes.c_int, [c_i32, c_i64, c_vp, c_vp, c_vp]),
    "faop_measure_point_mix_peak": (ctypes.c_int, [ctypes.c_int, _P(c_dbl), _P(c_dbl), c_vp]),
    "faop_measure_fp64_peak": (ctypes.c_int, [ctypes.c_int, _P(c_dbl), _P(c_dbl), c_vp]),
}

_lib = None


class FaopError(RuntimeError):
    pass


def _try_build():
    """Compile csrc/ with nvcc when the shared library is absent and a CUDA toolkit is at hand (a fresh checkout);
    failure is not masked — load() still raises if the library does not appear."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)
    if nvcc and shutil.which("make"):
        subprocess.call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc"), "NVCC=" + nvcc])


def load():
    """Load libfaop.so (once).  Raises FaopError if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        _try_build()
    if not os.path.exists(LIB_PATH):
        raise FaopError(
            "libfaop.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C fastamericanoptionpricing_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise FaopError("failed to load %s: %s" % (LIB_PATH, e)) from e
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here means the .so is stale w.r.t. include/faop.h
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def kernel_source_hash():
    """sha256 (first 16 hex digits) of the CUDA sources.  profiles/kernel_profile.json records the hash of the build its ncu counters
    were taken from; bench.py marks the counters `stale` when the sources have changed since."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(_HERE, "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def declared_symbols():
    """Every entry point include/faop.h declares (kept in step by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        kind = "cudaError_t" if rc > 0 else "FAOP_E"
        raise FaopError("%s failed: %s %d" % (what, kind, rc))
