"""CPU: the C-ABI library loads, exports every symbol include/faop.h declares, and its host-only helpers work.
No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO
from fastamericanoptionpricing_b200 import _lib, batch


def _header_symbols():
    src = open(os.path.join(REPO, "include", "faop.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(faop_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libfaop.so does not export %s" % n
    assert sorted(names) == _lib.declared_symbols()       # the ctypes table covers the whole header
    assert lib.faop_version() == 100


def test_cfg_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.FaopCfg) == 6 * 4 + 2 * 8 + 2 * 4 + 8 and _lib.FaopCfg.tau_max.offset == 48
    assert ctypes.sizeof(_lib.FaopCnExtra) == 4 * 8
    assert ctypes.sizeof(_lib.FaopSurface) == 9 * 8 + 4 * 4


def test_tables_build_host_only():
    lib = _lib.load()
    cfg = batch.AloConfig(n=12, l=32, m=64)
    c = cfg.c()
    nb = ctypes.c_size_t()
    assert lib.faop_tables_bytes(ctypes.byref(c), ctypes.byref(nb)) == 0
    host = np.empty(nb.value // 8)
    yl, wl = batch.gauss_legendre(32)
    ym, wm = batch.gauss_legendre(64)
    assert lib.faop_tables_build(ctypes.byref(c), yl.ctypes.data, wl.ctypes.data, ym.ctypes.data, wm.ctypes.data,
                                 host.ctypes.data) == 0
    LP, MP, n = 32, 64, 12
    w, h, rh, sg = host[0:LP], host[LP:2 * LP], host[2 * LP:3 * LP], host[3 * LP:4 * LP]
    assert np.array_equal(w, wl) and np.allclose(h, (1 + yl) / 2, rtol=0, atol=0)
    assert np.allclose(rh * h, 1.0, rtol=1e-15) and np.allclose(sg, np.sqrt(1 - h * h), rtol=1e-13)
    off_c = 4 * LP + 4 * MP
    c_nodes = host[off_c:off_c + n + 1]
    assert np.array_equal(c_nodes, np.cos(np.arange(n + 1) * np.pi / n))     # ChebyshevInterpolation.py:15-17
    # interpolation matrix rows sum to 1 (interpolant of a constant) and reproduce node values at z = node
    off_M = (off_c + (n + 1) + 2 * n + 1) & ~1
    M = host[off_M:off_M + n * (n + 1) * LP].reshape(n, n + 1, LP)
    assert np.max(np.abs(M.sum(axis=1) - 1.0)) < 1e-13
    from oracle import alo_oracle as o                                    # checker only
    z = (1 + c_nodes[:n, None]) * sg[None, :] - 1
    Hj = np.random.default_rng(1).normal(size=n + 1)
    ref = o.cheb_eval(o.cheb_fit(Hj), z.reshape(-1)).reshape(n, LP)
    assert np.max(np.abs(np.einsum("ijk,j->ik", M, Hj) - ref)) < 1e-13
    # bad configurations are refused with FAOP_E codes, not crashes
    bad = batch.AloConfig(n=1).c()
    assert lib.faop_tables_bytes(ctypes.byref(bad), ctypes.byref(nb)) == -2
    bad = batch.AloConfig(l=129).c()
    assert lib.faop_tables_bytes(ctypes.byref(bad), ctypes.byref(nb)) == -2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.FaopError):
        batch.price_batch([0.04], [0.01], [0.2], [100.0], [3.0], [80.0], [1])
    with pytest.raises(_lib.FaopError):
        batch.european_batch([1.0], [80.0], [0.04], [0.01], [0.2], [100.0], [1])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import or load it."""
    pkg = os.path.join(REPO, "fastamericanoptionpricing_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|from\s+\.+oracle|liboracle|oracle/|oracle\.)", txt), (root, f)


def test_gauss_legendre_host_matches_numpy():
    """faop_gauss_legendre_host (for callers without numpy) against numpy.polynomial.legendre.leggauss."""
    from numpy.polynomial.legendre import leggauss
    lib = _lib.load()
    for n in (1, 2, 3, 24, 32, 48, 64, 101, 128):
        y, w = np.empty(n), np.empty(n)
        assert lib.faop_gauss_legendre_host(n, y.ctypes.data, w.ctypes.data) == 0
        yn, wn = leggauss(n)
        # numpy's weights carry ~1e-14 of rounding of their own (they do not sum to 2 exactly; these do)
        assert np.max(np.abs(y - yn)) <= 4e-16 and np.max(np.abs(w - wn)) <= 1e-14 and abs(w.sum() - 2) <= 1e-15
    assert lib.faop_gauss_legendre_host(0, y.ctypes.data, w.ctypes.data) == -1


def test_plain_c_caller_compiles_and_links():
    """tests/cabi/testA.c needs nothing but include/faop.h and libfaop.so (run on the GPU by test_gpu_parity)."""
    import subprocess
    import tempfile
    exe = os.path.join(tempfile.mkdtemp(prefix="faop_cabi_"), "testA")
    csrc = os.path.join(REPO, "fastamericanoptionpricing_b200", "csrc")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(REPO, "include"), os.path.join(REPO, "tests", "cabi", "testA.c"),
                           "-o", exe, "-L", csrc, "-lfaop", "-lm", "-Wl,-rpath," + csrc])
    assert os.path.exists(exe)


def test_tanh_sinh_rule_host():
    """batch.tanh_sinh (opt-in rule): symmetric, weights sum to 2, integrates an end-point singularity
    that defeats Gauss-Legendre at the same size."""
    y, w = batch.tanh_sinh(64)
    assert np.allclose(y, -y[::-1], atol=1e-15) and np.allclose(w, w[::-1], rtol=1e-12) and abs(w.sum() - 2) < 1e-10
    assert (np.diff(y) >= 0).all() and y[0] > -1 and y[-1] < 1
    f = lambda x: 1 / np.sqrt(1 - x * x)          # noqa: E731   integral over [-1, 1] = pi
    assert abs((w * f(y)).sum() - np.pi) < 1e-6
    yg, wg = batch.gauss_legendre(64)
    assert abs((wg * f(yg)).sum() - np.pi) > 1e-2
    assert abs((w * np.exp(y)).sum() - (np.e - 1 / np.e)) < 1e-12
    with pytest.raises(ValueError):
        batch.quadrature_rule("simpson", 8)


def test_bad_arguments_are_codes_not_crashes():
    """Every batch entry point refuses null pointers / negative sizes / unsupported modes with a FAOP_E* code before it
    touches CUDA (this runs without a GPU)."""
    lib = _lib.load()
    c = batch.AloConfig(n=12, l=32, m=64).c()
    cd = batch.AloConfig(n=12, l=32, m=64, use_derivative=True).c()
    nb = ctypes.c_size_t()
    N = ctypes.c_int64
    z = None
    assert lib.faop_alo_price_batch(ctypes.byref(c), N(-1), *([z] * 19)) == -1
    assert lib.faop_alo_price_batch(ctypes.byref(c), N(5), *([z] * 19)) == -1
    assert lib.faop_alo_price_batch(ctypes.byref(c), N(0), *([z] * 19)) == 0            # empty batch: nothing to do
    assert lib.faop_alo_price_batch(None, N(0), *([z] * 19)) == -1
    assert lib.faop_alo_ladder_workspace_bytes(ctypes.byref(c), N(-1), ctypes.byref(nb)) == -1
    assert lib.faop_alo_ladder_workspace_bytes(ctypes.byref(c), N(10), ctypes.byref(nb)) == 0 and nb.value > 10 * 13 * 8
    assert lib.faop_alo_ladder_batch(ctypes.byref(c), N(3), z, z, z, z, z, z, 4, z, 1.0, 2.0, z, z, z, z, 0, z) == -1
    assert lib.faop_alo_ladder_batch(ctypes.byref(cd), N(3), z, z, z, z, z, z, 4, z, 1.0, 2.0, z, z, z, z, 0, z) == -2   # no derivative mode
    assert lib.faop_alo_ladder_batch(ctypes.byref(c), N(0), z, z, z, z, z, z, 4, z, 1.0, 2.0, z, z, z, z, 0, z) == 0
    s = _lib.FaopSurface(50.0, 150.0, 0.1, 2.0, 0.1, 0.5, 100.0, 0.04, 0.01, 10, 10, 10, 1)
    assert lib.faop_alo_surface_dedup_workspace_bytes(ctypes.byref(c), ctypes.byref(s), N(0), N(1001), ctypes.byref(nb)) == -1   # past the end
    assert lib.faop_alo_surface_dedup_workspace_bytes(ctypes.byref(c), ctypes.byref(s), N(0), N(1000), ctypes.byref(nb)) == 0
    assert lib.faop_alo_surface_dedup_batch(ctypes.byref(c), ctypes.byref(s), N(0), N(10), z, z, z, z, 0, z) == -1
    assert lib.faop_alo_surface_dedup_batch(ctypes.byref(cd), ctypes.byref(s), N(0), N(10), z, z, z, z, 0, z) == -2
    assert lib.faop_alo_greeks_known_boundary_batch(ctypes.byref(c), N(4), *([z] * 10), 1e-3, 1e-4, z, z, z, z, z) == -1
    assert lib.faop_iv_update_batch(N(4), *([z] * 9)) == -1 and lib.faop_iv_update_batch(N(0), *([z] * 9)) == 0
    assert lib.faop_cn_put_countdown_batch(N(4), 200, 3, *([z] * 7), 1.2, 1e-5, 10, None, z, z, z, z) == -1       # mode out of range
    assert lib.faop_cn_put_countdown_batch(N(4), 5000, 2, *([z] * 7), 1.2, 1e-5, 10, None, z, z, z, z) == -2      # grid beyond the compiled limit
    assert lib.faop_cn_put_countdown_batch(N(4), 200, 2, *([z] * 7), 1.2, 1e-5, 10, None, z, z, z, z) == -1
    assert lib.faop_bjerksund_stensland_batch(N(2), *([z] * 12)) == -1 and lib.faop_bs_phi_batch(N(2), *([z] * 10)) == -1
    # the scratch flag only changes the size for configurations that use it
    plain, flagged = ctypes.c_size_t(), ctypes.c_size_t()
    for cfg, uses in ((batch.AloConfig(**{"solver": "B", "n": 24, "l": 64, "m": 128}), True), (batch.AloConfig(n=12, l=32, m=64), False)):
        assert lib.faop_workspace_bytes(ctypes.byref(cfg.c()), N(100), ctypes.byref(plain)) == 0
        assert lib.faop_workspace_bytes(ctypes.byref(cfg.c(_lib.FLAG_WS_SCRATCH)), N(100), ctypes.byref(flagged)) == 0
        assert (flagged.value > plain.value) == uses
