/* faop.h — C ABI of libfaop.so, the B200 (sm_100a) batched American option pricer.
 *
 * This is the drop-in boundary for the one data-parallel hot path of
 * antdvid/FastAmericanOptionPricing: the Andersen-Lake-Offengenden spectral-collocation
 * solve of the early-exercise boundary and the price integral, plus the leaf closed forms
 * and the Crank-Nicolson cross-check solver.  The reference is pure Python with no FFI of
 * its own; every entry point below names the reference interface (file:line, relative to
 * the reference root) whose work it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - plain C, no torch types: pointers + sizes.  All arrays are contiguous structure-of-arrays,
 *    FP64 unless stated; `is_put` is one byte per option (1 = put, 0 = call; the reference's
 *    OptionType enum Call=1/Put=2, QDplusAmericanOptionSolver.py:8-10, maps to this flag).
 *  - functions ending in `_host` take HOST pointers and do their own H2D/D2H staging through a
 *    faop_ctx; all others take DEVICE pointers owned by the caller, never allocate or free
 *    user-visible memory, are asynchronous on `cuda_stream` (a cudaStream_t, NULL = default
 *    stream) and keep no global mutable state.
 *  - return value: 0 on success, a positive cudaError_t value for CUDA failures, or a negative
 *    FAOP_E* code for bad arguments.  Never throws, never aborts.  Numerical failure is not an
 *    error: NaN propagates and non-convergence is visible only through iters/err, exactly as in
 *    the reference (FastAmericanOptionSolverBase.py:119-133).
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef FAOP_H
#define FAOP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAOP_VERSION 100

#define FAOP_EINVAL  (-1) /* bad argument (null pointer, size out of range) */
#define FAOP_ERANGE  (-2) /* configuration outside the compiled limits       */
#define FAOP_ENOMEM  (-3) /* host allocation failed                          */

#define FAOP_MAX_COLLOC 64  /* n  <= 64  (nodes n+1)                           */
#define FAOP_MAX_QUAD   128 /* l, m <= 128                                     */

/* Solver knobs = the reference's post-construction attributes
 * (FastAmericanOptionSolverBase.py:19-23, 46-47; eta is the local at :145). */
typedef struct faop_cfg {
    int32_t solver;         /* 0 = FastAmericanOptionSolverA, 1 = FastAmericanOptionSolverB  */
    int32_t use_derivative; /* self.use_derivative                                            */
    int32_t n_colloc;       /* self.collocation_num  (n; nodes = n+1)                         */
    int32_t n_quad;         /* self.quadrature_num   (l)                                      */
    int32_t n_price_quad;   /* self.integration_num  (m)                                      */
    int32_t max_iters;      /* self.max_iters                                                 */
    double  iter_tol;       /* self.iter_tol                                                  */
    double  eta;            /* 0.5 in the reference                                           */
    int32_t kernel;         /* 0 = auto (specialised kernel when one exists for (solver,n,l)),
                               1 = force the generic kernel; results agree to rounding          */
    int32_t flags;          /* FAOP_FLAG_* bits, 0 in the reference-equivalent default                        */
    const double* tau_max;  /* self.tau_max per option (FastAmericanOptionSolverBase.py:28, consumed at :185, :223): the end
                               of the collocation domain [0, tau_max].  NULL = the constructor's value, the maturity T.
                               A DEVICE array of N doubles for the device-pointer entry points, a HOST array for the
                               *_host ones.  (self.tau_min stays 0: upstream returns NaN after one sweep for any other
                               value, because z = sqrt(4 (u - tau_min)/(tau_max - tau_min)) - 1 meets u < tau_min.)   */
} faop_cfg;

/* `workspace` was sized by faop_workspace_bytes() with this same flag set: besides the QD seeds it then holds the
 * per-warp scratch in which the n > 12 kernels keep the iteration-invariant tables e^{q u_ik}, e^{r u_ik} (they stay in L2
 * and save two of Solver B's four exponentials per quadrature point).  Without the flag (or with workspace == NULL) those
 * values are recomputed; results are identical to rounding.                                                            */
#define FAOP_FLAG_WS_SCRATCH 1
/* The QD seed (set_initial_guess, FastAmericanOptionSolverBase.py:258-265) is found by default with the reference's own
 * root finder — MINPACK hybrd as scipy.optimize.root(hybr, x0=K) drives it (QDplusAmericanOptionSolver.py:73), one unknown,
 * restated in csrc/faop_seed.cu — because the reference keeps whatever iterate hybrd stops at.  This flag selects a
 * safeguarded Newton iteration with the analytic derivative instead (warm-started node to node, fast exp/erfcx): ~3x
 * cheaper, same root to ~1e-15 where the residual is well conditioned, but not the reference's iterate where hybrd exits
 * early or the residual is flat.                                                                                       */
#define FAOP_FLAG_SEED_NEWTON 2

/* ---- tables ---------------------------------------------------------------------------------
 * Gauss-Legendre nodes/weights are produced by the CALLER with numpy.polynomial.legendre.leggauss
 * (the reference's own call, FastAmericanOptionSolverBase.py:175) so the device sees the same bits.
 * faop_tables_build packs them, the Chebyshev nodes cos(i pi/n) (ChebyshevInterpolation.py:15-17),
 * the cosine table of the DCT-I fit (ChebyshevInterpolation.py:43-51) and, when it fits in shared
 * memory, the option-independent interpolation matrix M[(i,k), j] into one blob
 * of faop_tables_bytes() bytes in HOST memory; the caller copies the blob to the device.            */
int faop_tables_bytes(const faop_cfg* cfg, size_t* bytes);
int faop_tables_build(const faop_cfg* cfg, const double* y_l, const double* w_l, /* leggauss(l) */
                      const double* y_m, const double* w_m,                     /* leggauss(m) */
                      void* host_out);

/* Convenience for callers without numpy: the n-point Gauss-Legendre rule on [-1, 1] (nodes ascending), host code.  Agrees
 * with numpy.polynomial.legendre.leggauss to 1e-16 in the nodes and 1e-14 in the weights (numpy's own rounding), not bit for
 * bit; prices move by ~1e-13.                                                                                            */
int faop_gauss_legendre_host(int32_t n, double* y, double* w);

/* Device scratch needed by faop_alo_price_batch for N options: the QD seeds (used when B0_out is NULL) and, with
 * FAOP_FLAG_WS_SCRATCH, the invariant-table scratch (a fixed ~58 MB at n = 24, l = 64, independent of N). */
int faop_workspace_bytes(const faop_cfg* cfg, int64_t N, size_t* bytes);

/* ---- the hot path -----------------------------------------------------------------------------
 * Replaces FastAmericanOptionSolver.solve(t, s0) (FastAmericanOptionSolverBase.py:49-76) for a
 * batch: collocation grid (:221-223), QD seed (:258-265 -> QDplusAmericanOptionSolver.py:69-76),
 * the damped Jacobi fixed-point / diagonal-Newton loop (:105-165) over the Solver A or B integral
 * forms (FastAmericanOptionSolverA.py:5-83, FastAmericanOptionSolverB.py:5-75) and the price
 * integral (:92-103).  q == 0 calls take the European shortcut (:50-53): iters = 0, err = 1e6,
 * boundary rows NaN.
 *   t        nullable (t = 0)
 *   B0_in    nullable: N x (n+1) borrowed seed boundary (stage-wise parity); skips the seed kernel.  Its last column
 *            (tau = 0) is B_at_zero in every reference seed; the specialised kernels impose that value
 *   B, B0_out nullable: N x (n+1) row-major, node i at tau_i = T (1+cos(i pi/n))^2/4, last node tau=0
 *   iters, err nullable: realised iteration count and final ||B_old-B_new||_2 per option
 *   iter_errs nullable: N x max_iters, the per-iteration errors of self.iter_records (:130), NaN padded
 *   workspace  faop_workspace_bytes(cfg, N) bytes of device memory, laid out [QD seeds][scratch].  When B0_in or B0_out is
 *            given the seeds live there instead and the scratch starts at offset 0: faop_workspace_bytes(cfg, 0) bytes are
 *            enough then, and NULL is allowed at the price of the FAOP_FLAG_WS_SCRATCH speed-up                   */
int faop_alo_price_batch(const faop_cfg* cfg, int64_t N,
                         const double* r, const double* q, const double* sigma, const double* K,
                         const double* T, const double* t, const double* S, const uint8_t* is_put,
                         const void* tables_dev, const double* B0_in,
                         double* price, double* european, double* B, double* B0_out,
                         int32_t* iters, double* err, double* iter_errs, void* workspace, void* cuda_stream);

/* Config-5 sweep: option i of [first, first+count) is decoded on the device from the flat index
 * i = (iK*nT + iT)*nV + iV into K = linspace(K0,K1,nK)[iK], T = linspace(T0,T1,nT)[iT],
 * sigma = linspace(V0,V1,nV)[iV] (numpy.linspace arithmetic), all puts or all calls with common
 * S, r, q, t = 0.  Only prices (and optionally iters) are written back.  `workspace` (device, any size
 * >= 256*(8+n+1)*8 bytes; ~0.5 GB is plenty) bounds the chunk the sweep is processed in.               */
typedef struct faop_surface {
    double K0, K1, T0, T1, V0, V1, S, r, q;
    int32_t nK, nT, nV, is_put;
} faop_surface;
int faop_alo_surface_batch(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count,
                           const void* tables_dev, double* price, int32_t* iters /* nullable */,
                           void* workspace, size_t workspace_bytes, void* cuda_stream);

/* The same sweep with the boundary solves deduplicated: the early-exercise boundary is homogeneous of
 * degree one in the strike, so all strikes sharing (T, sigma) share one solve at K = 1; the reference's ABSOLUTE
 * stopping rule (FastAmericanOptionSolverBase.py:119) is honoured per strike (a strike takes the first iterate whose
 * error, scaled by K, is <= iter_tol), then every option gets its own price integral (:92-103).  Same index order,
 * outputs and argument meaning as faop_alo_surface_batch; prices agree with it to rounding (<= 1e-9 absolute).
 * Not available with use_derivative (FAOP_ERANGE).  Exact while a group records < 48 iterates between the loosest and
 * the tightest threshold (always, unless the sweep contracts slower than ~0.98 per iteration).
 * Workspace: faop_alo_surface_dedup_workspace_bytes for the same (first, count).                                      */
int faop_alo_surface_dedup_workspace_bytes(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count,
                                           size_t* bytes);
int faop_alo_surface_dedup_batch(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count,
                                 const void* tables_dev, double* price, int32_t* iters /* nullable */,
                                 void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Strike ladder (option chain / scenario grid): G scenarios (r, q, sigma, T, S, type), each priced at the same nK strikes
 * K[0..nK) -> price[g * nK + k] (iters likewise, nullable).  One boundary solve per scenario shared by its strikes, exactly
 * as in faop_alo_surface_dedup_batch; every (scenario, strike) pair gets the result FastAmericanOptionSolver.solve(0, S)
 * would give for that strike (FastAmericanOptionSolverBase.py:49-76) to rounding.  K_min / K_max: the extremes of the ladder
 * (host values; they bound the stopping thresholds the shared solve has to cover).  All array pointers are device pointers. */
int faop_alo_ladder_workspace_bytes(const faop_cfg* cfg, int64_t G, size_t* bytes);
int faop_alo_ladder_batch(const faop_cfg* cfg, int64_t G, const double* r, const double* q, const double* sigma,
                          const double* T, const double* S, const uint8_t* is_put, int32_t nK, const double* K,
                          double K_min, double K_max, const void* tables_dev, double* price, int32_t* iters /* nullable */,
                          void* workspace, size_t workspace_bytes, void* cuda_stream);

/* N, D, N', D', f, f' (and the damped update B_new) at `npts` arbitrary (tau, Q) points per option on a GIVEN
 * boundary B (N x (n+1)) — one Jacobi sweep's internals:
 * FastAmericanOptionSolver.compute_f_and_fprime / N_func / D_func / Nprime_func / Dprime_func /
 * iterate_once_foreach_tau (FastAmericanOptionSolverBase.py:144-165, 267-279 and the A/B files).
 * pts_tau, pts_B and all outputs are N x npts (outputs nullable).  f' is the analytic derivative when
 * cfg->use_derivative, else 1 (as the reference returns); tau == 0 gives f = B_at_zero, f' = 1.
 * I1, I2 (nullable) are the two quadrature integrals themselves: Solver A K12 and K3 (FastAmericanOptionSolverA.py:24-44),
 * Solver B the D- and N-integrals (FastAmericanOptionSolverB.py:25-39).                                              */
int faop_alo_eval_points_batch(const faop_cfg* cfg, int64_t N, int32_t npts,
                               const double* r, const double* q, const double* sigma, const double* K,
                               const double* T, const uint8_t* is_put, const void* tables_dev,
                               const double* B, const double* pts_tau, const double* pts_B,
                               double* Nv, double* Dv, double* Np, double* Dp, double* f, double* fp,
                               double* Bnew, double* I1, double* I2, void* cuda_stream);

/* B(u) at `npts` arbitrary u per option from the Chebyshev interpolant of ln(B/X)^2:
 * compute_integration_terms (FastAmericanOptionSolverBase.py:167-195).  u, Bu are N x npts.         */
int faop_alo_interp_boundary_batch(const faop_cfg* cfg, int64_t N, int32_t npts,
                                   const double* r, const double* q, const double* K, const double* T,
                                   const uint8_t* is_put, const void* tables_dev,
                                   const double* B, const double* u, double* Bu, void* cuda_stream);

/* Price integral on a GIVEN boundary: american_value_with_known_boundary
 * (FastAmericanOptionSolverBase.py:92-103, 211-219); tau = T - t may differ from T.              */
int faop_alo_price_known_boundary_batch(const faop_cfg* cfg, int64_t N,
                              const double* r, const double* q, const double* sigma, const double* K,
                              const double* T, const double* t, const double* S, const uint8_t* is_put,
                              const void* tables_dev, const double* B /* N x (n+1) */,
                              double* price, double* european, void* cuda_stream);

/* Delta, gamma, theta by bump-and-reprice on a GIVEN boundary (the boundary depends neither on the spot nor on the valuation
 * date, so the three are differences of american_value_with_known_boundary, FastAmericanOptionSolverBase.py:92-103):
 * central differences in S with h = bump_S_rel * S, one-sided dV/dt with step min(bump_t, T - t).  Outputs nullable.
 * SURVEY 8f row 2; vega / rho need re-solved boundaries and are composed by the caller (batch.greeks_batch).            */
int faop_alo_greeks_known_boundary_batch(const faop_cfg* cfg, int64_t N,
                              const double* r, const double* q, const double* sigma, const double* K,
                              const double* T, const double* t, const double* S, const uint8_t* is_put,
                              const void* tables_dev, const double* B /* N x (n+1) */, double bump_S_rel, double bump_t,
                              double* price, double* delta, double* gamma, double* theta, void* cuda_stream);
/* One step of a bracketed Illinois root search for the implied volatility (price is increasing in sigma): consumes
 * `price` = the American price at the current `sigma`, updates the bracket state and writes the next trial sigma.
 * Initial state: sigma = lo, flo = fhi = NaN, side = -2; a target outside [price(lo), price(hi)] ends as sigma = NaN.     */
int faop_iv_update_batch(int64_t N, const double* target, const double* price, double* sigma, double* lo, double* hi,
                         double* flo, double* fhi, int32_t* side, void* cuda_stream);

/* ---- leaf kernels -----------------------------------------------------------------------------*/
/* QDplus.compute_exercise_boundary(tau) (QDplusAmericanOptionSolver.py:69-76; residual :110-125):
 * safeguarded Newton from x0 = K instead of MINPACK hybr; tau == 0 -> B_at_zero (:78-88).         */
int faop_qd_boundary_batch(int64_t N, const double* r, const double* q, const double* sigma,
                           const double* K, const double* tau, const uint8_t* is_put,
                           double* B, void* cuda_stream);
/* QDplus.exercise_boundary_func(S, tau) residual itself (QDplusAmericanOptionSolver.py:110-125). */
int faop_qd_residual_batch(int64_t N, const double* r, const double* q, const double* sigma,
                           const double* K, const double* tau, const double* S, const uint8_t* is_put,
                           double* F, void* cuda_stream);
/* QDplus.price(tau, S) (QDplusAmericanOptionSolver.py:44-67), quirks included; boundary nullable. */
int faop_qd_price_batch(int64_t N, const double* r, const double* q, const double* sigma,
                        const double* K, const double* tau, const double* S, const uint8_t* is_put,
                        double* price, double* boundary, void* cuda_stream);
/* BksdStld.price(S) (BjerksundStenslandSolver.py:15-21): Bjerksund-Stensland 1993 flat-boundary American CALL with
 * computeExerciseBoundary (:38-43) and computeAlpha (:23-24).  The reference reads the module globals `K`, `T` at :40/:42
 * (it only runs as __main__); Kb, Tb carry them (nullable: the option's own K, T).  is_put nullable (= all calls); a put
 * is priced by the symmetry of the reference's __main__ (:53): call with S<->K, r<->q.  boundary (the trigger X) nullable. */
int faop_bjerksund_stensland_batch(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                                   const double* T, const double* S, const uint8_t* is_put, const double* Kb,
                                   const double* Tb, double* price, double* boundary, void* cuda_stream);
/* BksdStld.phi(S, T, gamma, H, X) (BjerksundStenslandSolver.py:26-33). */
int faop_bs_phi_batch(int64_t N, const double* S, const double* T, const double* gamma, const double* H, const double* X,
                      const double* r, const double* q, const double* sigma, double* out, void* cuda_stream);
/* EuropeanOption.european_put_value / european_call_value (EuropeanOptionSolver.py:7-22). */
int faop_european_batch(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                        const double* sigma, const double* K, const uint8_t* is_put,
                        double* price, void* cuda_stream);
/* EuropeanOption.european_option_theta / d1 / d2 (EuropeanOptionSolver.py:25-40); outputs nullable. */
int faop_european_greeks_batch(int64_t N, const double* tau, const double* S, const double* r,
                               const double* q, const double* sigma, const double* K,
                               double* theta, double* d1, double* d2, void* cuda_stream);
/* ChebyshevInterpolation.__init__/std_coeff (ChebyshevInterpolation.py:4-12, 43-51): y is N x (n+1). */
int faop_cheb_fit_batch(int64_t N, int32_t n, const double* y, double* a, void* cuda_stream);
/* ChebyshevInterpolation.value/std_cheby_single_value (:19-41): z is N x npts, already in [-1,1]. */
int faop_cheb_eval_batch(int64_t N, int32_t n, int32_t npts, const double* a, const double* z,
                         double* out, void* cuda_stream);

/* ---- Crank-Nicolson (CNOptionSolver, CrankNicolsonOptionSolver.py:5-107), American puts --------
 * One CTA per option; grid S_i = linspace(0, Smax, n_space) (:48), Smax = 3K (:12) unless `extra->Smax` says otherwise;
 * the dt schedule of solvePDE's float countdown (:36-45) is built by the caller: dts is N x max_steps, n_steps[i] entries valid.
 *   mode 0: unprojected direct solve  == the reference default USE_PSOR=False (np.linalg.solve, :81-82)
 *   mode 1: the reference's SOR-sweep-then-project loop (:84-107), operation for operation
 *   mode 2: projected (Brennan-Schwartz) tridiagonal sweep — the throughput mode (not in the reference)
 * X_out nullable: N x n_space final grid values (the reference's self.X).
 * `extra` (nullable, every member nullable) carries the rest of the CNOptionSolver object state: the `Smax` attribute
 * (:12, consumed at :48) per option; `X_in`, an N x n_space state to start from instead of setInitialCondition's payoff
 * (what solveLinearSystem / solvePSOR see in self.X when called on their own, :81-104); `err_out` / `iter_out`, the
 * attributes self.err / self.iter after the last solvePSOR (:103-104; zeros, as the constructor leaves them, in modes 0/2). */
typedef struct {
    const double* Smax;
    const double* X_in;
    double* err_out;
    int32_t* iter_out;
} faop_cn_extra;
int faop_cn_workspace_bytes(int64_t N, int32_t n_space, int32_t mode, size_t* bytes);
int faop_cn_put_batch(int64_t N, int32_t n_space, int32_t mode,
                      const double* r, const double* q, const double* sigma, const double* K,
                      const double* S0, const double* dts, const int32_t* n_steps, int32_t max_steps,
                      double omega, double tol, int32_t max_iter, const faop_cn_extra* extra,
                      double* price, double* X_out, void* workspace, void* cuda_stream);

/* The same solver with the time stepping of CNOptionSolver.solvePDE itself (:36-45): `while t > 0: dt = min(t, max_dt); ...;
 * t -= dt`, run on the device in the same IEEE arithmetic from per-option T and max_dt — no N x steps schedule to build or
 * copy (1.6 GB at C4).  Identical results to faop_cn_put_batch fed with that countdown's schedule.                        */
int faop_cn_put_countdown_batch(int64_t N, int32_t n_space, int32_t mode,
                                const double* r, const double* q, const double* sigma, const double* K,
                                const double* T, const double* S0, const double* max_dt,
                                double omega, double tol, int32_t max_iter, const faop_cn_extra* extra,
                                double* price, double* X_out, void* workspace, void* cuda_stream);
/* CNOptionSolver.setCoeff(dt) (:55-79) as data: the three diagonals and the right-hand side of ONE time step from the state X
 * (N x n_space each; row n_space-1 is the constant closure -1, 4, -5, 2 | 0, reported as lo = -5, di = 2, up = 0, b = 0).
 * Smax nullable (3K). */
int faop_cn_rows_batch(int64_t N, int32_t n_space, const double* r, const double* q, const double* sigma, const double* K,
                       const double* Smax, const double* dt, const double* X,
                       double* lo, double* di, double* up, double* b, void* cuda_stream);

/* ---- host-buffer convenience (what a non-torch caller binds) ----------------------------------
 * A context owns a device, a stream, device scratch, pinned staging buffers and the table blobs it
 * has built; it grows on demand and is freed by faop_ctx_destroy.  One context per host thread.    */
typedef struct faop_ctx faop_ctx;
int faop_ctx_create(int device, faop_ctx** out);
int faop_ctx_destroy(faop_ctx* ctx);
/* Same semantics as faop_alo_price_batch with HOST arrays (y/w tables as in faop_tables_build);
 * H2D, kernels and D2H run inside the call, which returns after the results are in host memory.  */
int faop_alo_price_batch_host(faop_ctx* ctx, const faop_cfg* cfg, int64_t N,
                              const double* r, const double* q, const double* sigma, const double* K,
                              const double* T, const double* t, const double* S, const uint8_t* is_put,
                              const double* y_l, const double* w_l, const double* y_m, const double* w_m,
                              const double* B0_in, double* price, double* european, double* B,
                              double* B0_out, int32_t* iters, double* err);

/* faop_alo_ladder_batch and faop_cn_put_countdown_batch (modes 0 and 2) with HOST arrays through the same context. */
int faop_alo_ladder_batch_host(faop_ctx* ctx, const faop_cfg* cfg, int64_t G,
                               const double* r, const double* q, const double* sigma, const double* T, const double* S,
                               const uint8_t* is_put, int32_t nK, const double* K,
                               const double* y_l, const double* w_l, const double* y_m, const double* w_m,
                               double* price /* G x nK */, int32_t* iters /* nullable */);
int faop_cn_put_countdown_batch_host(faop_ctx* ctx, int64_t N, int32_t n_space, int32_t mode,
                                     const double* r, const double* q, const double* sigma, const double* K,
                                     const double* T, const double* S0, const double* max_dt,
                                     const double* Smax /* nullable: 3K */,
                                     double* price, double* X_out /* nullable, N x n_space */);

/* Diagnostics */
int faop_version(void);
/* Kernels launched by this library in this process since load (for bench.py's gpu_launches). */
int64_t faop_launch_count(void);
/* Unit-test hook for the device math the specialised kernel is built from (csrc/faop_fastmath.cuh):
 * kind 0 = fast_exp, 1 = mills_half (Phi(-y)/exp(-y^2/2)), 2 = fast_sqrt_pos, 3 = fast_rcp, 4 = Phi via fast path,
 * 5 = erfcx(x)/2 (x >= 0), 6 = sqrt(max(0, x)), 7 = e^x + e^(x-1) (two interleaved chains),
 * 8 = erfcx(x)/2 degree-14 variant, 9 = e^x degree-8 variant (the quadrature-point versions), 10 = fast_log,
 * 11 = Phi(x) as the leaf kernels compute it (erfc based), 12 = erfcx(x)/2 degree-16 variant. */
int faop_dev_math_batch(int32_t kind, int64_t N, const double* x, double* out, void* cuda_stream);
/* FP64 FMA-pipe peak microbenchmark (the roofline denominator; MEASURED_PEAKS.json has no FP64
 * figure): launches `launches` saturating DFMA kernels, returns measured TFLOP/s (2 flops/DFMA). */
int faop_measure_fp64_peak(int launches, double* tflops, double* ms_per_launch, void* cuda_stream);
/* Transcendental-mix microbenchmark: the flagship kernel's per-quadrature-point math (12 DFMA interpolation, sqrt, two
 * degree-8 exp, one degree-16 erfcx, combine) on register operands, no memory traffic: quadrature points per second.   */
int faop_measure_point_mix_peak(int launches, double* points_per_s, double* ms_per_launch, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* FAOP_H */
