"""Synthetic option batches for BASELINE.json's configs.

The generators are the single definition of the C1..C5 workloads: the golden
fixtures, the parity tests and bench.py all draw from here, so "the same
seeded inputs" means literally the same arrays.  Pure numpy, no GPU needed.

Parameter ranges extend the reference's own random bulk test
(reference testBulk.py:36-44: r in [.001,.1], q in [0,.1], K in [50,150]).
"""
import numpy as np

C2_SEED = 20251019
C3_SEED = 20251020
C4_SEED = 20251021


def c1_testA():
    """Config 1: the single put of reference testA.py:7-19 (defaults n=12,l=24,m=48)."""
    return dict(r=0.04, q=0.01, sigma=0.2, K=100.0, T=3.0, S=80.0, t=0.0, is_put=True,
                n=12, l=24, m=48, iter_tol=1e-5, max_iters=20, use_derivative=False, solver="A")


def c2_batch(N=1_000_000, seed=C2_SEED):
    """Config 2: N random puts/calls for Solver A, n=12, l=32, m=64.

    Draw order is part of the definition (each draw is a length-N vector, so a
    prefix of the N=1M batch is NOT the N=64 batch).
    """
    g = np.random.default_rng(seed)
    is_put = g.random(N) < 0.5
    r = g.uniform(0.001, 0.10, N)
    q = g.uniform(0.0, 0.10, N)
    sigma = g.uniform(0.10, 0.60, N)
    T = g.uniform(0.05, 3.0, N)
    K = g.uniform(50.0, 150.0, N)
    S = K * g.uniform(0.5, 1.5, N)
    return dict(r=r, q=q, sigma=sigma, K=K, T=T, S=S, t=np.zeros(N), is_put=is_put)


C2_CFG = dict(solver="A", n=12, l=32, m=64, iter_tol=1e-5, max_iters=20, use_derivative=False)


def c3_batch(N=1_000_000, seed=C3_SEED):
    """Config 3: stability stress for Solver B, n=24, l=64, m=128.

    Drawn as C2, then option i is overwritten by category i % 3:
      0: low vol   sigma=U(.02,.10), T=U(1,5)
      1: long T    T=U(5,15), sigma=U(.1,.4)
      2: q > r     q = r + U(.005,.06)
    (category by index so that any prefix covers all three thirds).
    """
    b = c2_batch(N, seed)
    g = np.random.default_rng(seed + 1000)
    cat = np.arange(N) % 3
    lo_sig = g.uniform(0.02, 0.10, N)
    lo_T = g.uniform(1.0, 5.0, N)
    long_T = g.uniform(5.0, 15.0, N)
    long_sig = g.uniform(0.1, 0.4, N)
    dq = g.uniform(0.005, 0.06, N)
    b["sigma"] = np.where(cat == 0, lo_sig, np.where(cat == 1, long_sig, b["sigma"]))
    b["T"] = np.where(cat == 0, lo_T, np.where(cat == 1, long_T, b["T"]))
    b["q"] = np.where(cat == 2, b["r"] + dq, b["q"])
    b["category"] = cat
    return b


C3_CFG = dict(solver="B", n=24, l=64, m=128, iter_tol=1e-5, max_iters=200, use_derivative=False)


def c4_batch(N=100_000, seed=C4_SEED):
    """Config 4: N American puts for the Crank-Nicolson solver (K=100, 2000x2000 grid)."""
    g = np.random.default_rng(seed)
    r = g.uniform(0.01, 0.08, N)
    q = g.uniform(0.0, 0.06, N)
    sigma = g.uniform(0.15, 0.5, N)
    T = g.uniform(0.25, 3.0, N)
    K = np.full(N, 100.0)
    S = g.uniform(70.0, 130.0, N)
    return dict(r=r, q=q, sigma=sigma, K=K, T=T, S=S, t=np.zeros(N), is_put=np.ones(N, dtype=bool))


C4_CFG = dict(n_space=2000, n_steps=2000)


def c4_smax(b):
    """Upper end of the Crank-Nicolson grid for Config 4 (the CNOptionSolver.Smax attribute, reference
    CrankNicolsonOptionSolver.py:12,48): Smax = K max(3, exp(3 sigma sqrt(T))).  The reference's constructor default 3K sits only
    ln(3)/(sigma sqrt(T)) = 1.3 standard deviations above the strike for the widest options of the batch (sigma sqrt(T) = 0.87)
    and its zero-second-derivative closure row then costs up to 0.33 against the spectral price; three standard deviations keep
    the truncation below the 2000-node discretisation error (max |CN - spectral| 2.3e-3 on the batch, bar 5e-3)."""
    return b["K"] * np.maximum(3.0, np.exp(3.0 * b["sigma"] * np.sqrt(b["T"])))


def c5_surface(start=0, count=None, nK=1000, nT=1000, nV=100):
    """Config 5: implied-surface sweep, 1000 strikes x 1000 maturities x 100 vols.

    Flat index i -> (iK, iT, iV) with iV fastest: i = (iK*nT + iT)*nV + iV.
    Returns the slice [start, start+count) as arrays (for parity subsets); the
    CUDA path decodes the same index on the device.
    """
    total = nK * nT * nV
    if count is None:
        count = total - start
    i = np.arange(start, start + count, dtype=np.int64)
    iV = i % nV
    iT = (i // nV) % nT
    iK = i // (nV * nT)
    # np.linspace semantics: start + i*step (unfused), last point == stop exactly
    K = np.linspace(50.0, 150.0, nK)[iK]
    T = np.linspace(1.0 / 52, 3.0, nT)[iT]
    sigma = np.linspace(0.10, 0.60, nV)[iV]
    n = len(i)
    return dict(r=np.full(n, 0.04), q=np.full(n, 0.01), sigma=sigma, K=K, T=T, S=np.full(n, 100.0),
                t=np.zeros(n), is_put=np.ones(n, dtype=bool))


C5_CFG = dict(C2_CFG)


def stress_batch(N=6000, seed=123):
    """Hard corners on top of the C3 thirds (low vol x long T, long T, q > r): deep ITM/OTM spots, maturities down to 1e-4,
    r == q, q == 0 (calls take the European shortcut), r ~ 0.  Used by the stress parity test and its live-reference fixture."""
    b = {k: v[:N].copy() for k, v in c3_batch(N).items()}
    g = np.random.default_rng(seed)
    b["S"][0:500] = b["K"][0:500] * g.uniform(0.02, 0.3, 500)
    b["S"][500:1000] = b["K"][500:1000] * g.uniform(3.0, 20.0, 500)
    b["T"][1000:1500] = 10 ** g.uniform(-4, -2, 500)
    b["q"][1500:1800] = b["r"][1500:1800]
    b["q"][1800:2000] = 0.0
    b["r"][2000:2100] = 1e-6
    return b
