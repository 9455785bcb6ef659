// faop_alo.cu — the ALO spectral-collocation hot path: QD seed kernel, boundary fixed-point kernel
// (generic: any n <= 64, l, m <= 128, Solver A/B, with/without derivative), price integral.
// Replaces FastAmericanOptionSolver.solve (FastAmericanOptionSolverBase.py:49-165) for a batch.
#include "faop_alo.cuh"
#include "faop_fastmath.cuh"
#include "faop_kernels.h"

namespace faop {

// ---------------------------------------------------------------- QD seed: one thread per option
// set_initial_guess (FastAmericanOptionSolverBase.py:258-265) -> QDplus.compute_exercise_boundary per node
// (QDplusAmericanOptionSolver.py:69-76, residual :110-125).  The reference starts MINPACK hybr from x0 = K at every node;
// here a thread walks its option's nodes from the smallest tau upwards and starts Newton at the previous node's root
// (the QD residual has a single root, so the limit is the same; ~5 instead of ~8 evaluations per node), with the
// fast exp / erfcx of faop_fastmath.cuh (their 1e-14 relative error moves the root by the same order).
FAOP_D double seed_cdf(double x, double E) {  // Phi(x) from E = exp(-x^2/2)
    double xa[1] = {fabs(x) * kInvSqrt2}, G[1];
    half_erfcx_n<1, false>(xa, G);
    double ec = E * G[0];
    return (x <= 0.0) ? ec : 1.0 - ec;
}

__global__ void __launch_bounds__(128) qd_seed_kernel(AloArgs a) {
    const int n = a.tl.n, nn = n + 1;
    int64_t opt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (opt >= a.N) return;
    const double r = a.r[opt], q = a.q[opt], sg = a.sigma[opt], K = a.K[opt], T = a.tau_max ? a.tau_max[opt] : a.T[opt];
    const bool put = a.is_put[opt] != 0;
    double* out = a.seeds_out + opt * nn;
    if (q == 0 && !put) {  // European shortcut: no boundary is solved
        for (int j = 0; j <= n; ++j) out[j] = nan("");
        return;
    }
    const double s2 = sg * sg;
    const double Nn1 = 2 * (r - q) / s2 - 1.0, M4 = 4.0 * (2 * r / s2), invK = 1.0 / K;
    double S = K;   // x0 = K for the first (smallest-tau) node, QDplusAmericanOptionSolver.py:73
    for (int j = n; j >= 0; --j) {
        const double c = a.tables[a.tl.off_c + j];
        const double tau = (c + 1) * (c + 1) * T / 4;  // to_orig_point, FastAmericanOptionSolverBase.py:239-240
        if (tau == 0) {
            out[j] = b_at_zero(r, q, K, put);
            continue;
        }
        double ee[2] = {-r * tau, -q * tau}, ed[2];
        fast_exp_n<2, false>(ee, ed);
        const double dr = ed[0], dq = ed[1];
        const double h = 1.0 - dr;
        const double disc = sqrt(Nn1 * Nn1 + M4 / h);
        const double qQD = put ? -0.5 * Nn1 - 0.5 * disc : -0.5 * Nn1 + 0.5 * disc;   // :127-134
        const double svt = sg * sqrt(tau), rsvt = 1.0 / svt, drift = (r - q) * tau;
        double prev = 1.79e308;
        for (int it = 0; it < 100; ++it) {
            const double d1 = (log(S * invK) + drift) * rsvt + 0.5 * svt;
            const double d2 = d1 - svt;
            double ex[2] = {-0.5 * d1 * d1, -0.5 * d2 * d2}, E[2];
            fast_exp_n<2, false>(ex, E);
            double F, dF;
            if (put) {
                const double P1 = seed_cdf(-d1, E[0]), P2 = seed_cdf(-d2, E[1]);
                const double p = K * dr * P2 - S * dq * P1;
                const double g = 1 - dq * P1;
                F = g * S + qQD * (K - S - p);                                   // :124
                dF = (1 - qQD) * g + dq * E[0] * kInvSqrt2Pi * rsvt;             // :207-208
            } else {
                const double P1 = seed_cdf(d1, E[0]), P2 = seed_cdf(d2, E[1]);
                const double cv = S * dq * P1 - K * dr * P2;
                const double g = 1 - dq * P1;
                F = g * S - qQD * (S - K - cv);                                  // :122
                dF = (1 - qQD) * g - dq * E[0] * kInvSqrt2Pi * rsvt;
            }
            double Sn = S - F / dF;
            if (!(Sn > 0) || !(Sn < 1.79e308)) Sn = 0.5 * S;
            if (Sn > 2.0 * S) Sn = 2.0 * S;
            const double d = fabs(Sn - S);
            // converged (the next Newton step would be ~d^2), or stalled at the rounding floor of a flat residual
            const bool done = d <= 1e-10 * fabs(Sn) || (d >= 0.5 * prev && prev <= 1e-7 * fabs(Sn));
            prev = d;
            S = Sn;
            if (done) break;
        }
        out[j] = S;
    }
}

// One Jacobi sweep's quadrature: for every active node i the reduced sums s1..s4 land in shared memory.
template <int SOLVER, bool DERIV>
FAOP_D void sweep_sums(const WarpSmem& w, const CtaTables& tb, const TableLayout& tl, const Opt& o, int lane) {
    const int n = tl.n;
    for (int i = 0; i < n; ++i) {
        double tau = w.tau[i];
        if (tau == 0) continue;  // uniform across the warp
        const NodeC c = node_from_smem(w, o, i);
        double L = w.L[i];
        double opc = 1.0 + tb.c[i];
        double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        for (int k = lane; k < tl.LP; k += 32) {
            double h = tb.hl[k], sg_k = tb.sgl[k];
            double z = fma(opc, sg_k, -1.0);  // to_cheby_point(u_ik): sqrt(4 u/T) - 1 = (1+c_i) sqrt(g_k) - 1
            double Hu = clenshaw(n, w.A, z);
            double root = sqrt(Hu < 0.0 ? 0.0 : Hu);  // np.maximum(0, H): NaN propagates (fmax would drop it)
            double lnz = fma(o.om, root, L);
            alo_point<SOLVER, DERIV>(o, c, lnz, tb.wl[k], h, tb.rhl[k], sg_k * sg_k, a1, a2, a3, a4);
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (DERIV) { a3 = warp_sum(a3); a4 = warp_sum(a4); }
        if (lane == 0) { w.s1[i] = a1; w.s2[i] = a2; w.s3[i] = a3; w.s4[i] = a4; }
    }
    __syncwarp();
}

constexpr int kGenWarps = 8;

// ---------------------------------------------------------------- generic boundary + price kernel
template <int SOLVER, bool DERIV>
__global__ void __launch_bounds__(kGenWarps * 32) alo_generic_kernel(AloArgs a) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        if (o.q == 0 && !o.put) {  // European shortcut, FastAmericanOptionSolverBase.py:50-53
            if (a.snap) {
                if (lane == 0) a.nsnap[opt] = -1;
                continue;
            }
            double e = euro_value(o.T - o.t, o.S, o.r, o.q, o.sg, o.K, false);
            if (lane == 0) {
                a.price[opt] = e;
                if (a.european) a.european[opt] = e;
                if (a.iters) a.iters[opt] = 0;
                if (a.err) a.err[opt] = 1e6;
            }
            if (a.B) for (int j = lane; j <= n; j += 32) a.B[opt * nn + j] = nan("");
            if (a.iter_errs) for (int j = lane; j < a.cfg.max_iters; j += 32) a.iter_errs[opt * a.cfg.max_iters + j] = nan("");
            continue;
        }
        setup_nodes(w, tb, n, o, lane);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        const double lnXK = log(o.X / o.K);
        int count = 0;
        double err = 1.0;
        int ns = 0;                  // snapshot mode only
        double best = 1.79e308;
        while (err > a.cfg.iter_tol && count < a.cfg.max_iters) {
            ++count;
            fit_boundary(w, tb, n, o.X, lane);
            sweep_sums<SOLVER, DERIV>(w, tb, tl, o, lane);
            double ss = 0;
            for (int j = lane; j <= n; j += 32) {
                double Bj = w.B[j], Bn;
                if (w.tau[j] == 0) {
                    Bn = o.X;  // iterate_once_foreach_tau: tau_i == 0 -> B_at_zero
                } else {
                    NodeC c = node_from_smem(w, o, j);
                    NodeOut no = alo_node<SOLVER, DERIV>(o, c, Bj, w.L[j], lnXK, w.disc[j], w.s1[j], w.s2[j], w.s3[j], w.s4[j]);
                    double fp = DERIV ? no.fp : 0.0;
                    Bn = Bj + a.cfg.eta * (Bj - no.f) / (fp - 1.0);
                }
                double d = fabs(Bj - Bn);
                ss = fma(d, d, ss);
                w.B[j] = Bn;
            }
            err = sqrt(warp_sum(ss));  // norm1_error is the 2-norm over all n+1 nodes (:230-233)
            if (a.iter_errs && lane == 0) a.iter_errs[opt * a.cfg.max_iters + (count - 1)] = err;  // iter_records (:130)
            __syncwarp();
            if (a.snap && err <= a.snap_tol_hi && err < best) {
                best = err;
                snap_append(a, opt, ns, err, count, w, tb, n, o.X, lane);
            }
        }
        if (a.snap) {
            if (!(ns > 0 && best == err)) snap_append(a, opt, ns, 0.0, count, w, tb, n, o.X, lane);
            if (lane == 0) a.nsnap[opt] = ns;
            continue;
        }
        if (a.iter_errs) for (int j = count + lane; j < a.cfg.max_iters; j += 32) a.iter_errs[opt * a.cfg.max_iters + j] = nan("");
        if (a.B) for (int j = lane; j <= n; j += 32) a.B[opt * nn + j] = w.B[j];
        double eur;
        double v = price_integral(w, tb, tl, o, lane, &eur);
        if (lane == 0) {
            a.price[opt] = v;
            if (a.european) a.european[opt] = eur;
            if (a.iters) a.iters[opt] = count;
            if (a.err) a.err[opt] = err;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- N, D, N', D', f, f' on a given boundary
// compute_f_and_fprime / N_func / D_func / Nprime_func / Dprime_func at npts arbitrary (tau, Q) pairs per option
// (FastAmericanOptionSolverBase.py:267-279; the reference evaluates them at collocation nodes during the solve and
// at free points in its diagnostics, e.g. check_f_with_B :427-447).  The boundary interpolant comes from a.seeds_in.
template <int SOLVER, bool DERIV>
__global__ void __launch_bounds__(kGenWarps * 32) alo_eval_points_kernel(AloArgs a, EvalOut eo) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        const double lnXK = log(o.X / o.K);
        fit_boundary(w, tb, n, o.X, lane);
        for (int p = 0; p < a.npts; ++p) {
            const int64_t at = opt * a.npts + p;
            const double tau = a.pts_tau[at], Q = a.pts_B[at];
            NodeOut no;
            double i1 = 0.0, i2 = 0.0;   // quadrature_sum returns 0 at tau == 0 (:382-383)
            if (tau == 0) {  // compute_f_and_fprime: tau == 0 -> [B_at_zero, 1]
                no.Nv = no.Dv = no.Np = no.Dp = nan("");
                no.f = o.X;
                no.fp = 1.0;
            } else {
                const NodeC c = node_consts(o, tau);
                const double L = log(Q / o.X);
                const double zf = sqrt(4 * tau / o.Tc);
                double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
                for (int k = lane; k < tl.LP; k += 32) {
                    double h = tb.hl[k], sg_k = tb.sgl[k];
                    double z = fma(zf, sg_k, -1.0);
                    double Hu = clenshaw(n, w.A, z);
                    double root = sqrt(Hu < 0.0 ? 0.0 : Hu);
                    double lnz = fma(o.om, root, L);
                    alo_point<SOLVER, DERIV>(o, c, lnz, tb.wl[k], h, tb.rhl[k], sg_k * sg_k, a1, a2, a3, a4);
                }
                a1 = warp_sum(a1);
                a2 = warp_sum(a2);
                if (DERIV) { a3 = warp_sum(a3); a4 = warp_sum(a4); }
                no = alo_node<SOLVER, DERIV>(o, c, Q, L, lnXK, o.K * exp(-tau * (o.r - o.q)), a1, a2, a3, a4);
                i1 = c.tau * a1;
                i2 = (SOLVER == 0) ? c.tau * c.cinv * kInvSqrt2Pi * a2 : c.tau * a2;
            }
            if (lane == 0) {
                if (eo.Nv) eo.Nv[at] = no.Nv;
                if (eo.Dv) eo.Dv[at] = no.Dv;
                if (eo.Np) eo.Np[at] = no.Np;
                if (eo.Dp) eo.Dp[at] = no.Dp;
                if (eo.f) eo.f[at] = no.f;
                if (eo.fp) eo.fp[at] = no.fp;
                if (eo.I1) eo.I1[at] = i1;
                if (eo.I2) eo.I2[at] = i2;
                if (eo.Bnew) {  // iterate_once_foreach_tau (:144-165)
                    double fp = DERIV ? no.fp : 0.0;
                    eo.Bnew[at] = (tau == 0) ? o.X : Q + a.cfg.eta * (Q - no.f) / (fp - 1.0);
                }
            }
        }
        __syncwarp();
    }
}

// B(u) at npts arbitrary u per option: compute_integration_terms (FastAmericanOptionSolverBase.py:167-195).
__global__ void __launch_bounds__(kGenWarps * 32) alo_interp_kernel(AloArgs a, double* Bu) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        fit_boundary(w, tb, n, o.X, lane);
        for (int p = lane; p < a.npts; p += 32) {
            double u = a.pts_tau[opt * a.npts + p];
            double z = sqrt(4 * (u - 0.0) / (o.Tc - 0.0)) - 1;  // to_cheby_point (:235-237), tau_min = 0
            double Hu = clenshaw(n, w.A, z);
            double root = sqrt(Hu < 0.0 ? 0.0 : Hu);
            Bu[opt * a.npts + p] = exp(-o.om * root) * o.X;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- price integral on a given boundary
__global__ void __launch_bounds__(kGenWarps * 32) alo_price_known_kernel(AloArgs a) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        double eur;
        double v = price_integral(w, tb, tl, o, lane, &eur);
        if (lane == 0) {
            a.price[opt] = v;
            if (a.european) a.european[opt] = eur;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- Greeks by bump-and-reprice on a given boundary
// The boundary does not depend on the spot or on the valuation date, so delta, gamma and theta are differences of the price
// integral (american_value_with_known_boundary, FastAmericanOptionSolverBase.py:92-103) evaluated on the SAME boundary:
//   delta = (V(S+h) - V(S-h)) / 2h,  gamma = (V(S+h) - 2 V(S) + V(S-h)) / h^2,  h = bump_S_rel * S
//   theta = (V(t + dt) - V(t)) / dt  (calendar-time derivative dV/dt, one-sided; dt is clipped to the remaining life)
// SURVEY 8f row 2; the reference itself only offers european_option_theta.
__global__ void __launch_bounds__(kGenWarps * 32) alo_greeks_known_kernel(AloArgs a, double bump_S_rel, double bump_t,
                                                                          double* delta, double* gamma, double* theta) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        Opt o = load_option(a, opt);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        double eur;
        const double S0 = o.S, t0 = o.t;
        const double v0 = price_integral(w, tb, tl, o, lane, &eur);
        const double h = bump_S_rel * S0;
        o.S = S0 + h;
        const double vp = price_integral(w, tb, tl, o, lane, &eur);
        o.S = S0 - h;
        const double vm = price_integral(w, tb, tl, o, lane, &eur);
        o.S = S0;
        double dt = fmin(bump_t, o.T - t0);
        o.t = t0 + dt;
        const double vt = price_integral(w, tb, tl, o, lane, &eur);
        if (lane == 0) {
            if (a.price) a.price[opt] = v0;
            if (delta) delta[opt] = (vp - vm) / (2.0 * h);
            if (gamma) gamma[opt] = (vp - 2.0 * v0 + vm) / (h * h);
            if (theta) theta[opt] = (vt - v0) / dt;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- host-side launchers
static int generic_grid(int64_t N) {
    int64_t ctas = (N + kGenWarps - 1) / kGenWarps;
    int64_t cap = (int64_t)kSMs * 16;
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

static size_t generic_smem(const TableLayout& tl) {
    return sizeof(double) * (size_t)(tl.small + kGenWarps * WarpSmem::doubles(tl.n + 1));
}

int launch_qd_seed(const AloArgs& a, cudaStream_t st) {
    if (a.N == 0) return 0;
    if (!(a.cfg.flags & FAOP_FLAG_SEED_NEWTON)) return launch_qd_seed_hybr(a, st);
    int64_t blocks = (a.N + 127) / 128;
    qd_seed_kernel<<<(unsigned)blocks, 128, 0, st>>>(a);
    count_launch();
    return launch_status();
}

template <int SOLVER, bool DERIV>
static int launch_generic_t(const AloArgs& a, cudaStream_t st) {
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_generic_kernel<SOLVER, DERIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_generic_kernel<SOLVER, DERIV><<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a);
    count_launch();
    return launch_status();
}

int launch_alo_generic(const AloArgs& a, cudaStream_t st) {
    if (a.N == 0) return 0;
    if (a.cfg.solver == 0) return a.cfg.use_derivative ? launch_generic_t<0, true>(a, st) : launch_generic_t<0, false>(a, st);
    return a.cfg.use_derivative ? launch_generic_t<1, true>(a, st) : launch_generic_t<1, false>(a, st);
}

template <int SOLVER, bool DERIV>
static int launch_eval_t(const AloArgs& a, const EvalOut& eo, cudaStream_t st) {
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_eval_points_kernel<SOLVER, DERIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_eval_points_kernel<SOLVER, DERIV><<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a, eo);
    count_launch();
    return launch_status();
}

int launch_alo_eval_nodes(const AloArgs& a, const EvalOut& eo, cudaStream_t st) {
    if (a.N == 0 || a.npts == 0) return 0;
    if (a.cfg.solver == 0) return a.cfg.use_derivative ? launch_eval_t<0, true>(a, eo, st) : launch_eval_t<0, false>(a, eo, st);
    return a.cfg.use_derivative ? launch_eval_t<1, true>(a, eo, st) : launch_eval_t<1, false>(a, eo, st);
}

int launch_alo_interp(const AloArgs& a, double* Bu, cudaStream_t st) {
    if (a.N == 0 || a.npts == 0) return 0;
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_interp_kernel<<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a, Bu);
    count_launch();
    return launch_status();
}

int launch_alo_greeks_known(const AloArgs& a, double bump_S_rel, double bump_t, double* delta, double* gamma, double* theta,
                            cudaStream_t st) {
    if (a.N == 0) return 0;
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_greeks_known_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_greeks_known_kernel<<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a, bump_S_rel, bump_t, delta, gamma, theta);
    count_launch();
    return launch_status();
}

int launch_alo_price_known(const AloArgs& a, cudaStream_t st) {
    if (a.N == 0) return 0;
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_price_known_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_price_known_kernel<<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a);
    count_launch();
    return launch_status();
}

}  // namespace faop
