// faop_alo_fast.cu — specialised boundary + price kernel for the flagship configuration
// (FastAmericanOptionSolverA, no derivative, n = 12 collocation nodes, l <= 32 quadrature points):
// BASELINE.json configs[1] (C2), configs[4] (C5) and the reference's own testA case (C1) — and, as the SOLVER = 1
// instantiation, FastAmericanOptionSolverB at the same sizes (the reference's testB case: per point two Phi instead of
// one Phi and two phi, FastAmericanOptionSolverB.py:25-39, with both invariant tables w e^{qu}, w e^{ru} in shared memory).
//
// One option per warp, lane = quadrature index k.  What makes it fast (all FP64-issue savings; the kernel sits at 95 % of the
// issue bound `2 x FP64 + other` warp instructions per scheduler):
//   * H(u_ik) comes from the option-independent interpolation matrix staged once per CTA in shared memory as
//     M2[i][j/2][k][2] (one 16-byte load feeds two DFMAs, 12 DFMA per point instead of a Clenshaw recurrence; the j = n column
//     multiplies H_n = 0 and is dropped); the iterate H_j lives in registers;
//   * e^{q u_ik} (iteration invariant) is computed once per option into shared memory, with the quadrature weight folded in;
//   * per point ~78 FP64 instructions: sqrt by rsqrt.approx + one third-order step, d+- as x = d/sqrt2 from packed per-node
//     constants (four 16-byte broadcast loads), two exponentials with merged arguments exp(ln w + qu - d+^2/2) (degree-8 fits),
//     Phi through erfcx(x)/2 = Q(t)/(x + c) (degree 16, one reciprocal), integer-ALU clamps and sign tests; two nodes advance in
//     lock step (6 independent polynomial chains per lane), coefficients are constant-bank operands;
//   * the 24 per-node sums of one sweep are reduced 4 nodes at a time through a shared-memory transpose (8 STS + 8 LDS +
//     9 DADD + 2 shuffles, row pitch 36) and closed by lanes 0..12 in one pass; the stopping test is warp-uniform and uses the
//     squared norm;
//   * NaN outcomes (a node outside (0, inf)) are decided once per sweep on the full-accuracy log instead of being carried
//     through the clamped fast paths.
// 16 warps per CTA, one CTA per SM, grid = min(#SM, ceil(N/16)) persistent CTAs striding over the batch.  Static striding: a
// warp prices ~420 options of a 1M batch, its total sweep count has a relative spread of 1.5 sqrt(420) / (19.3 x 420) = 0.4 %, so
// a work counter would buy nothing at these sizes (the snapshot mode's tail is equally averaged).
#include "faop_alo.cuh"
#include "faop_fastmath.cuh"
#include "faop_alo_fast.cuh"
#include "faop_kernels.h"

namespace faop {

constexpr int NC = 12;        // collocation n
constexpr int NN = NC + 1;    // nodes
constexpr int kGroup = 4;     // nodes per transposing reduction
constexpr int kPointErfcxTier = 2;   // erfcx fit at the quadrature points: 1 = degree 14 (3e-11), 2 = degree 16 (1.4e-12)

template <int SOLVER>
struct FastWarpSmem {
    double eq[NC * 32];   // w_k e^{q u_ik}, [i][k]
    double er[SOLVER == 1 ? NC * 32 : 1];   // w_k e^{r u_ik} (Solver B only)
    NodePack node[16];
    double B[16], L[16], H[16], A[16];
    double tau[16], disc[16];
    double sN[16], sD[16];
    double s3[16], s4[16];   // derivative sums (use_derivative, Solver A): N'- and D'-integrals
    double par[8];   // X, 1/X, ln(X/K)
    double red[8 * 36];   // transpose buffer of the per-group reduction, row pitch 36 = 4 (mod 16): conflict-free both ways
};

template <int kFastWarps, bool SNAP = false, int SOLVER = 0, bool DERIV = false, bool ERRS = false>
__global__ void __launch_bounds__(kFastWarps * 32, 1) alo_fast_A12_kernel(AloArgs a) {
    static_assert(!DERIV || (SOLVER == 0 && !SNAP), "the derivative instantiation exists for Solver A only");
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    double* s_tab = smem;
    double* s_M = smem + tl.small;          // re-laid out as M2[i][j/2][k][j&1], j < 12: one 16-byte load feeds two DFMAs
    FastWarpSmem<SOLVER>* s_w = reinterpret_cast<FastWarpSmem<SOLVER>*>(s_M + NC * NC * 32);
    for (int i = threadIdx.x; i < tl.small; i += blockDim.x) smem[i] = a.tables[i];
    for (int e = threadIdx.x; e < NC * NC * 32; e += blockDim.x) {
        const int k = e & 31, j = (e >> 5) % NC, i = e / (32 * NC);
        // the j = n column multiplies H_n = ln(X/X)^2 = 0 (the tau = 0 node is pinned to X) and is dropped
        s_M[((i * (NC / 2) + (j >> 1)) * 32 + k) * 2 + (j & 1)] = a.tables[tl.off_M + (i * NN + j) * 32 + k];
    }
    __syncthreads();
    const CtaTables tb(s_tab, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FastWarpSmem<SOLVER>& ws = s_w[warp];

    // lane constants of the l-point rule
    const double h = tb.hl[lane], rh = tb.rhl[lane], sgk = tb.sgl[lane], wk = tb.wl[lane];
    const double g = sgk * sgk;            // 1 - h^2
    const double lnw = log(wk);            // the quadrature weight rides in the exponents: w e^{a} = e^{a + ln w}

    // generic-kernel view of the same per-warp arrays, for the price integral
    WarpSmem gw(ws.B, 16);
    gw.B = ws.B; gw.L = ws.L; gw.H = ws.H; gw.A = ws.A;

    const int64_t stride = (int64_t)gridDim.x * kFastWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kFastWarps + warp; opt < a.N; opt += stride) {
        double om, oq, orr;     // the only option scalars the sweep needs in registers
        int sgn;
        bool degenerate;
        {
            const Opt o = load_option(a, opt);
            if (o.q == 0 && !o.put) {  // European shortcut, FastAmericanOptionSolverBase.py:50-53
                if (SNAP) {
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
                if (a.B && lane < NN) a.B[opt * NN + lane] = nan("");
                if (ERRS) for (int j = lane; j < a.cfg.max_iters; j += 32) a.iter_errs[opt * a.cfg.max_iters + j] = nan("");
                continue;
            }
            // ---- per-option setup: node constants (lanes 0..12) and the invariant e^{q u_ik} table
            if (lane < NN) {
                double c = tb.c[lane];
                double tau = (c + 1) * (c + 1) * o.Tc / 4;      // to_orig_point, FastAmericanOptionSolverBase.py:239-240
                ws.tau[lane] = tau;
                const NodePack np = make_node_pack(o, tau);
                ws.node[lane] = np;
                ws.disc[lane] = o.K * exp(-tau * (o.r - o.q));
                // the tau = 0 node is pinned to X: the reference's seed has it (QDplus.compute_exercise_boundary(0) ->
                // B_at_zero) and every sweep re-imposes it; a borrowed B0 row is normalised the same way
                ws.B[lane] = (lane == NC) ? o.X : a.seeds_in[opt * NN + lane];
            }
            if (lane == 0) {
                ws.par[0] = o.X;
                ws.par[1] = 1.0 / o.X;
                ws.par[2] = log(o.X / o.K);
            }
            om = o.om; oq = o.q; orr = o.r;
            sgn = o.put ? 0 : (int)0x80000000;   // sign flip of the call side, applied with integer XOR
            degenerate = !(o.Tc > 0);     // T == 0: every node has tau == 0 and is pinned to X
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NC; i += 4) {       // invariant tables, four interleaved chains with constant-bank coefficients
            double eqa[4], eqv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) eqa[c] = oq * ws.tau[i + c] * g;
            fast_exp_n<4, false>(eqa, eqv);
#pragma unroll
            for (int c = 0; c < 4; ++c) ws.eq[(i + c) * 32 + lane] = wk * eqv[c];
            if (SOLVER == 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c) eqa[c] = orr * ws.tau[i + c] * g;
                fast_exp_n<4, false>(eqa, eqv);
#pragma unroll
                for (int c = 0; c < 4; ++c) ws.er[(i + c) * 32 + lane] = wk * eqv[c];
            }
        }

        int count = 0;
        // the stopping rule ||B_old - B_new||_2 > tol (Base.py:119) is tested on the squares (err holds the squared norm
        // inside the loop), which keeps the square root out of the sweep; snapshot mode records the norm itself
        double err = 1.0;
        const double tol_cmp = SNAP ? a.cfg.iter_tol : (a.cfg.iter_tol < 0 ? -1.0 : a.cfg.iter_tol * a.cfg.iter_tol);
        int ns = 0;                  // snapshot mode only
        double best = 1.79e308;
        while (err > tol_cmp && count < a.cfg.max_iters) {
            ++count;
            // ---- L_j = ln(B_j/X), H_j = L_j^2 (FastAmericanOptionSolverBase.py:184), broadcast through shared memory
            bool bad = false;
            if (lane < NN) {
                double L = (lane == NC) ? 0.0 : fast_log(ws.B[lane] * ws.par[1]);
                ws.L[lane] = L;
                ws.node[lane].L = L;
                ws.H[lane] = L * L;
                bad = !(fabs(L) <= 1.79e308);
            }
            // A node outside (0, inf) (a diverging iterate, a NaN seed, X = rK/q = 0): upstream's H = ln(B/X)^2 (Base.py:184) is
            // NaN / inf there, the Chebyshev fit spreads it to every coefficient, so this sweep returns NaN at every active node,
            // its error is NaN and `while iter_err > tol` stops on it.  That outcome is written down directly — decided on the
            // full-accuracy log, once per sweep — rather than pushed through the point loop, whose exp / erfcx / sqrt clamp on
            // the integer high word and do not carry a NaN of either sign through.
            if (__any_sync(kFull, bad)) {
                if (lane < NC) ws.B[lane] = nan("");
                err = nan("");
                if (ERRS && lane == 0) a.iter_errs[opt * a.cfg.max_iters + (count - 1)] = err;
                __syncwarp();
                break;
            }
            __syncwarp();
            double Hj[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) Hj[j] = ws.H[j];

            // ---- quadrature of K12 / K3 at the 12 active nodes: 4 nodes per transposing reduction, the nodes of a
            //      group advanced two at a time in lock step (2 sqrt, 4 exp and 2 erfcx chains in flight per lane)
            if (!degenerate) {
#pragma unroll 1
                for (int grp = 0; grp < NC / kGroup; ++grp) {
                    double acc[8];
#pragma unroll
                    for (int pr = 0; pr < kGroup / 2; ++pr) {
                        const int i0 = grp * kGroup + pr * 2;
                        double Hu[2], root[2];
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            const double2* Mi = reinterpret_cast<const double2*>(s_M) + ((i0 + p) * (NC / 2)) * 32 + lane;
                            // H(u_ik) = sum_{j<n} M[i][j][k] H_j
                            double2 m = Mi[0];
                            double acc0 = m.x * Hj[0];
                            acc0 = fma(m.y, Hj[1], acc0);
#pragma unroll
                            for (int jp = 1; jp < NC / 2; ++jp) {
                                m = Mi[jp * 32];
                                acc0 = fma(m.x, Hj[2 * jp], acc0);
                                acc0 = fma(m.y, Hj[2 * jp + 1], acc0);
                            }
                            Hu[p] = acc0;
                        }
                        sqrt_clamped_n<2>(Hu, root);        // sqrt(max(0, H)), NaN propagates
                        double ex[4], xa[4], xp[2], xm[2], cis[2];
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            const NodePack& nd = ws.node[i0 + p];
                            const double lnz = fma(om, root[p], nd.L);
                            xp[p] = fma(lnz, nd.cinv2 * rh, nd.apc2 * h);        // d+ / sqrt2
                            xm[p] = fma(-nd.ssu2, h, xp[p]);                       // d- / sqrt2
                            ex[2 * p] = fma(-xp[p], xp[p], fma(nd.qtau, g, lnw));        // ln w + q u - d+^2/2
                            ex[2 * p + 1] = fma(-xm[p], xm[p], fma(nd.rtau, g, lnw));    // ln w + r u - d-^2/2
                            xa[2 * p] = fabs(xp[p]);
                            xa[2 * p + 1] = fabs(xm[p]);
                            cis[p] = nd.cis;
                        }
                        double E[4];
                        fast_exp_n<4, true>(ex, E);         // E[2p] = w e^{qu} phi(d+) sqrt(2 pi), E[2p+1] = w e^{ru} phi(d-) sqrt(2 pi)
                        if (SOLVER == 0) {
                            double xa2[2] = {xa[0], xa[2]}, G[2];
                            half_erfcx_n<2, kPointErfcxTier>(xa2, G);
#pragma unroll
                            for (int p = 0; p < 2; ++p) {
                                // put: +w e^{qu} Phi(d+); call: -w e^{qu} Phi(-d+)   (FastAmericanOptionSolverA.py:27-32)
                                const double ec = E[2 * p] * G[p];
                                const double eqw = ws.eq[(i0 + p) * 32 + lane];
                                const bool neg = (__double2hiint(xp[p]) ^ sgn) < 0;   // om * d+ < 0 (both branches agree at 0)
                                double cdf = neg ? ec : eqw - ec;
                                cdf = __hiloint2double(__double2hiint(cdf) ^ sgn, __double2loint(cdf));  // times om
                                if (!DERIV) {
                                    acc[4 * pr + 2 * p] = E[2 * p + 1];                          // -> K3  = tau cinv/sqrt(2 pi) * sum
                                    acc[4 * pr + 2 * p + 1] = fma(cis[p], E[2 * p], h * cdf);    // -> K12 = tau * sum
                                } else {
                                    // Nprime_integrand / Dprime_integrand (FastAmericanOptionSolverA.py:55-83) in the same scaling:
                                    //   s3 += (w/h) d- e^{ru}phi(d-) sqrt(2pi),  s4 += w e^{qu}phi(d+) sqrt(2pi) (1 - d+ cinv/h)
                                    acc[4 * p] = E[2 * p + 1];
                                    acc[4 * p + 1] = fma(cis[p], E[2 * p], h * cdf);
                                    acc[4 * p + 2] = (1.41421356237309504880 * rh) * xm[p] * E[2 * p + 1];
                                    acc[4 * p + 3] = E[2 * p] * fma(-2.0 * rh * ws.node[i0 + p].cinv2, xp[p], 1.0);
                                }
                            }
                            if (DERIV) {    // two nodes x four sums per reduction
                                const double tot = transpose_reduce8(acc, ws.red, lane);
                                if ((lane & 3) == 0) {
                                    const int v = lane >> 2, node = i0 + (v >> 2);
                                    double* dst = (v & 2) ? ((v & 1) ? ws.s4 : ws.s3) : ((v & 1) ? ws.sD : ws.sN);
                                    dst[node] = tot;
                                }
                            }
                        } else {
                            double G[4];
                            half_erfcx_n<4, kPointErfcxTier>(xa, G);
#pragma unroll
                            for (int p = 0; p < 2; ++p) {
                                // w e^{qu} Phi(om d+) and w e^{ru} Phi(om d-)   (FastAmericanOptionSolverB.py:25-39)
                                const double ecp = E[2 * p] * G[2 * p], ecm = E[2 * p + 1] * G[2 * p + 1];
                                const bool negp = (__double2hiint(xp[p]) ^ sgn) < 0;
                                const bool negm = (__double2hiint(xm[p]) ^ sgn) < 0;
                                const double cdfp = negp ? ecp : ws.eq[(i0 + p) * 32 + lane] - ecp;
                                const double cdfm = negm ? ecm : ws.er[(i0 + p) * 32 + lane] - ecm;
                                acc[4 * pr + 2 * p] = h * cdfm;          // -> N integral = tau * sum
                                acc[4 * pr + 2 * p + 1] = h * cdfp;      // -> D integral = tau * sum
                            }
                        }
                    }
                    if (!DERIV) {
                        const double tot = transpose_reduce8(acc, ws.red, lane);
                        // lanes 8 ii .. 8 ii + 3 hold the K3 sum of node grp*4+ii, lanes 8 ii + 4 .. 8 ii + 7 its K12 sum
                        if ((lane & 3) == 0) {
                            const int node = grp * kGroup + (lane >> 3);
                            if (lane & 4) ws.sD[node] = tot; else ws.sN[node] = tot;
                        }
                    }
                }
            }
            __syncwarp();
            // ---- close N, D, f and the damped update at the nodes (lanes 0..12)
            double dd = 0.0;
            if (lane < NN) {
                const double Bj = ws.B[lane];
                double Bn;
                const double tau = ws.tau[lane];
                if (tau == 0) {
                    Bn = ws.par[0];                // iterate_once_foreach_tau: tau_i == 0 -> B_at_zero
                } else {
                    const NodePack nd = ws.node[lane];
                    const double LK = nd.L + ws.par[2];
                    const double xpK = fma(LK, nd.cinv2, nd.apc2);      // d+(tau_i, B_i/K) / sqrt2
                    const double xmK = xpK - nd.ssu2;
                    double exK[2] = {-xpK * xpK, -xmK * xmK}, EK[2];
                    fast_exp_n<2, false>(exK, EK);      // full-accuracy table, the two chains interleaved
                    const double Ep = EK[0], Em = EK[1];
                    double Nv, Dv;
                    if (SOLVER == 0) {
                        const double K3 = tau * nd.cis * ws.sN[lane];
                        const double K12 = tau * ws.sD[lane];
                        Nv = Em * nd.cis + orr * K3;                                               // A.py:5-10
                        Dv = Ep * nd.cis + om * fast_half_erfcx_cdf(om * xpK, Ep) + oq * K12;      // A.py:13-19
                    } else {
                        Nv = fast_half_erfcx_cdf(om * xmK, Em) + orr * (tau * ws.sN[lane]);        // B.py:5-13
                        Dv = fast_half_erfcx_cdf(om * xpK, Ep) + oq * (tau * ws.sD[lane]);         // B.py:15-23
                    }
                    const double rD = fast_rcp(Dv);
                    const double f = ws.disc[lane] * Nv * rD;                                               // Base.py:273
                    double fp = 0.0;
                    if (DERIV) {    // N', D' (FastAmericanOptionSolverA.py:46-74), f' (Base.py:276-278)
                        const double rB = fast_rcp(Bj);
                        const double Nsum = ws.s3[lane] * kInvSqrt2Pi * tau * rB / (2.0 * nd.ssu2 * nd.ssu2);
                        const double Dsum = tau * nd.cis * rB * ws.s4[lane];
                        const double Np = -(1.41421356237309504880 * xmK) * (Em * kInvSqrt2Pi) * rB / (2.0 * nd.ssu2 * nd.ssu2) - orr * Nsum;
                        const double Dp = Ep * nd.cis * rB * fma(-2.0 * nd.cinv2, xpK, 1.0) + oq * Dsum;
                        fp = ws.disc[lane] * (Np * rD - Dp * Nv * rD * rD);
                    }
                    Bn = Bj + a.cfg.eta * (Bj - f) / (fp - 1.0);                                            // Base.py:164
                }
                dd = fabs(Bj - Bn);
                ws.B[lane] = Bn;
            }
            err = warp_sum(dd * dd);               // norm1_error is the 2-norm over all n+1 nodes (:230-233)
            if (SNAP) err = sqrt(err);
            if (ERRS && lane == 0) a.iter_errs[opt * a.cfg.max_iters + (count - 1)] = sqrt(err);  // iter_records (:130)
            __syncwarp();
            if (SNAP && err <= a.snap_tol_hi && err < best) {
                best = err;
                snap_append(a, opt, ns, err, count, gw, tb, NC, ws.par[0], lane);
            }
        }
        if (SNAP) {
            if (!(ns > 0 && best == err)) snap_append(a, opt, ns, 0.0, count, gw, tb, NC, ws.par[0], lane);
            if (lane == 0) a.nsnap[opt] = ns;
            __syncwarp();
            continue;
        }
        if (count > 0) err = sqrt(err);            // not SNAP here: back to the norm the reference reports (self.error); NaN stays NaN
        if (ERRS)
            for (int j = count + lane; j < a.cfg.max_iters; j += 32) a.iter_errs[opt * a.cfg.max_iters + j] = nan("");
        if (a.B && lane < NN) a.B[opt * NN + lane] = ws.B[lane];
        double eur;
        const Opt o = load_option(a, opt);
        const double v = price_integral_fast(gw, tb, tl, o, lane, &eur);
        if (lane == 0) {
            a.price[opt] = v;
            if (a.european) a.european[opt] = eur;
            if (a.iters) a.iters[opt] = count;
            if (a.err) a.err[opt] = err;
        }
        __syncwarp();
    }
}

template <int kFastWarps, bool SNAP = false, int SOLVER = 0, bool DERIV = false, bool ERRS = false>
static int launch_fast_t(const AloArgs& a, cudaStream_t st) {
    const size_t sm = sizeof(double) * (size_t)(a.tl.small + NC * NC * 32) + sizeof(FastWarpSmem<SOLVER>) * kFastWarps;
    FAOP_CUDA_OK(cudaFuncSetAttribute(alo_fast_A12_kernel<kFastWarps, SNAP, SOLVER, DERIV, ERRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int64_t ctas = (a.N + kFastWarps - 1) / kFastWarps;
    int grid = (int)(ctas < kSMs ? ctas : kSMs);
    alo_fast_A12_kernel<kFastWarps, SNAP, SOLVER, DERIV, ERRS><<<grid, kFastWarps * 32, sm, st>>>(a);
    count_launch();
    return launch_status();
}

// cfg.kernel: 0 = default (16 warps/CTA, <= 128 registers/thread); tuning variants 2 / 3 = 12 / 20 warps per CTA
int launch_alo_fast(const AloArgs& a, cudaStream_t st) {
    // Solver B with the derivative is NaN by construction upstream and stays on the generic kernel
    const bool match = (!a.cfg.use_derivative || (a.cfg.solver == 0 && !a.snap)) && a.tl.n == NC && a.tl.LP == 32 &&
                       a.tl.has_matrix && (a.iter_errs == nullptr || !a.snap);
    if (!match) return -1000;
    if (a.N == 0) return 0;
    if (a.iter_errs) {      // per-iteration error history (the drop-in classes' iter_records): its own instantiations
        if (a.cfg.use_derivative) return launch_fast_t<16, false, 0, true, true>(a, st);
        return a.cfg.solver == 1 ? launch_fast_t<16, false, 1, false, true>(a, st) : launch_fast_t<16, false, 0, false, true>(a, st);
    }
    if (a.cfg.use_derivative) return launch_fast_t<16, false, 0, true>(a, st);
    if (a.cfg.solver == 1) return a.snap ? launch_fast_t<16, true, 1>(a, st) : launch_fast_t<16, false, 1>(a, st);
    if (a.snap) return launch_fast_t<16, true>(a, st);
    if (a.cfg.kernel == 2) return launch_fast_t<12>(a, st);
    if (a.cfg.kernel == 3) return launch_fast_t<20>(a, st);
    return launch_fast_t<16>(a, st);
}

}  // namespace faop
