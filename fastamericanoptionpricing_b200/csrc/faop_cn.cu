// faop_cn.cu — batched Crank-Nicolson American put solver (CNOptionSolver, CrankNicolsonOptionSolver.py:5-107).
//
// Grid S_i = linspace(0, 3K, N), payoff max(K - S, 0), time stepping over the host-built dt schedule of solvePDE's
// float countdown (:36-45).  Rows (setCoeff, :55-79), with a = sigma S_i/dS and m = (r-q) S_i/dS:
//   alpha = dt/4 (a^2 - m), beta = dt/2 (r + a^2), gamma = dt/4 (a^2 + m)
//   row 0      : (1+beta_0) X_0 = (1-beta_0) X_0^old
//   rows 1..N-2: -alpha X_{i-1} + (1+beta) X_i - gamma X_{i+1} = alpha X^old_{i-1} + (1-beta) X^old_i + gamma X^old_{i+1}
//   row N-1    : -X_{N-4} + 4 X_{N-3} - 5 X_{N-2} + 2 X_{N-1} = 0
// mode 0 (the reference default, np.linalg.solve, no early-exercise projection) and mode 2 (Brennan-Schwartz
// projected back-substitution max(., K - S_i), the throughput mode) run ONE CTA PER OPTION with the grid values in
// registers (NPT consecutive nodes per thread): per time step the right-hand side is a 3-point stencil, the forward
// elimination d'_i = P_i + Q_i d'_{i-1} is an affine recurrence and the (projected) back-substitution
// X_i = max(d'_i + R_i X_{i+1}, g_i) a max-affine one (R_i >= 0), both evaluated as parallel scans of composed maps:
// thread-local composition, shuffle scan inside the warp, a short pass over the warp totals in shared memory.
// The 4-term last row is closed exactly by writing X_{N-2}, X_{N-3}, X_{N-4} as affine functions of X_{N-1}.
// mode 1 is the reference's SOR-sweep-then-project loop (:84-107) operation for operation (Gauss-Seidel is
// sequential: one thread per option, small grids only) and reproduces its quirks (leaked index, j = N-1 self term).
// Compiled with -fmad=false: the reference (numpy) never fuses a*b+c, and mode 1 is bit-comparable to it.
#include "faop_kernels.h"

namespace faop {

constexpr int kCnThreads = 256;
constexpr int kCnWarps = kCnThreads / 32;
// longest countdown the kernels accept: 2^24 steps of max_dt (finite); a max_dt of ~0 must not hang the GPU
__device__ __forceinline__ double mdt_guard(double max_dt) { return (max_dt > 0) ? fmin(max_dt * 16777216.0, 1.79e308) : 0.0; }
constexpr unsigned kFullMask = 0xffffffffu;

struct CnArgs {
    int64_t N;
    int n_space, max_steps, mode, max_iter;
    const double *r, *q, *sigma, *K, *S0, *dts;
    const int32_t* n_steps;
    const double *T, *max_dt;   // countdown mode (dts == nullptr): the kernel runs solvePDE's own float countdown
    const double *Smax;         // nullable: CNOptionSolver.Smax per option (:12, used at :48); nullptr = 3K
    const double *X_in;         // nullable: N x n_space state to start from instead of the payoff (single-step calls)
    double omega, tol;
    double *price, *X_out, *ws;
    double *err_out;            // nullable: CNOptionSolver.err / .iter after the last solvePSOR (:103-104), mode 1
    int32_t *iter_out;
};

__device__ __forceinline__ double cn_smax(const CnArgs& a, int64_t opt) { return a.Smax ? a.Smax[opt] : 3.0 * a.K[opt]; }

__device__ __forceinline__ double grid_point(int i, int n, double Smax) {
    // np.linspace(0, Smax, N) (:48): i * step + 0, last point == Smax exactly
    if (i == n - 1) return Smax;
    double step = (Smax - 0.0) / (double)(n - 1);
    return (double)i * step + 0.0;
}

// np.interp(S0, S, X) on the uniform grid; X read through a functor
template <typename XF>
__device__ double interp_uniform(double S0, int n, double K, XF X) {   // K here = Smax, the upper end of the grid
    double S_first = 0.0, S_last = K;
    if (!(S0 > S_first)) return X(0);
    if (!(S0 < S_last)) return X(n - 1);
    double dS = grid_point(1, n, K) - grid_point(0, n, K);
    if (!(dS > 0) || !(dS < 1.79e308)) return nan("");     // K = inf / NaN / <= 0: no grid to interpolate on (and no index to search)
    const double jf = S0 / dS;
    int j = (jf >= (double)(n - 2)) ? n - 2 : (jf > 0 ? (int)jf : 0);
    // the uniform-grid guess is off by at most one cell (rounding of i * step): bounded corrections, never a scan
    for (int c = 0; c < 2 && j > 0 && grid_point(j, n, K) > S0; ++c) --j;
    for (int c = 0; c < 2 && j < n - 2 && grid_point(j + 1, n, K) <= S0; ++c) ++j;
    double xj = X(j), xj1 = X(j + 1);
    double sj = grid_point(j, n, K), sj1 = grid_point(j + 1, n, K);
    double slope = (xj1 - xj) / (sj1 - sj);
    return slope * (S0 - sj) + xj;
}

// ---------------------------------------------------------------- modes 0 / 2: one CTA per option, scans
// affine map x -> A x + B
struct Aff { double A, B; };
// max-affine map x -> max(A x + B, C), A >= 0
struct MaxAff { double A, B, C; };

__device__ __forceinline__ Aff compose(const Aff& outer, const Aff& inner) {  // outer(inner(x))
    return {outer.A * inner.A, outer.A * inner.B + outer.B};
}
__device__ __forceinline__ MaxAff compose(const MaxAff& outer, const MaxAff& inner) {
    // max(a1 max(a2 x + b2, c2) + b1, c1) = max(a1 a2 x + a1 b2 + b1, max(a1 c2 + b1, c1)),  a1 >= 0
    return {outer.A * inner.A, outer.A * inner.B + outer.B, fmax(outer.A * inner.C + outer.B, outer.C)};
}
__device__ __forceinline__ double apply(const Aff& f, double x) { return f.A * x + f.B; }
__device__ __forceinline__ double apply(const MaxAff& f, double x) { return fmax(f.A * x + f.B, f.C); }

__device__ __forceinline__ Aff shfl_up_map(const Aff& m, int d) {
    return {__shfl_up_sync(kFullMask, m.A, d), __shfl_up_sync(kFullMask, m.B, d)};
}
__device__ __forceinline__ MaxAff shfl_down_map(const MaxAff& m, int d) {
    return {__shfl_down_sync(kFullMask, m.A, d), __shfl_down_sync(kFullMask, m.B, d), __shfl_down_sync(kFullMask, m.C, d)};
}

// max without fmax's NaN canonicalisation (one DSETP + two selects): the operands here are never NaN unless the inputs are
__device__ __forceinline__ double dmax(double x, double y) { return x > y ? x : y; }

// Everything that depends only on (option, dt) is computed once per distinct dt (normally once per option): the scaled
// rows Q_i = alpha_i/den_i, O_i = (1-beta_i)/den_i, R_i = gamma_i/den_i of the elimination, and the A-components (products
// of Q resp. R) of every composed map the two scans form — thread-local, per shuffle level, per warp.  A time step then
// only moves the B (and, projected, C) components: one FMA per composition.  Q, R, X and D live in registers, O and the
// payoff in shared memory; three barriers per step (halo exchange, forward warp totals, backward warp totals + closure data).
// Slot layout: the N-1 scanned nodes 0..N-2 occupy the LAST N-1 of the kCnThreads*NPT slots (node i = slot - off), the
// unused slots in front carry all-zero rows (identity-free padding that nobody reads), and X_{N-1} — fixed by the closure
// row, not by the scans — is a scalar every thread holds.  Every slot therefore runs the same unpredicated code.
template <int NPT, bool PROJ>
__global__ void __launch_bounds__(kCnThreads, (NPT <= 8 ? 2 : 1)) cn_scan_kernel(CnArgs a) {
    constexpr int NS = kCnThreads * NPT;  // slots
    constexpr double kNegBig = -1.79e308;
    extern __shared__ double sm[];
    double* s_Q = sm;                    // [j * T + t] for slot t*NPT + j
    double* s_R = s_Q + NS;
    double* s_O = s_R + NS;
    double* s_seq = s_O + NS;            // natural (node) order scratch (c' recurrence, fallback back-substitution, output)
    double* s_G = s_seq + NS + 2;        // payoff K - S_i, [j * T + t]   (s_seq holds up to NS + 1 nodes: n - 1 <= NS)
    double* s_eL = s_G + NS;             // [warp] X of the warp's first slot / last slot (halo of the neighbours)
    double* s_eR = s_eL + kCnWarps;
    double* s_fwA = s_eR + kCnWarps;     // forward warp totals: A (per dt) and B (per step)
    double* s_fwB = s_fwA + kCnWarps;
    double* s_bwA = s_fwB + kCnWarps;    // backward warp totals: A (per dt), B and C (per step)
    double* s_bwB = s_bwA + kCnWarps;
    double* s_bwC = s_bwB + kCnWarps;
    double* s_clo = s_bwC + kCnWarps;    // D_{N-4}, D_{N-3}, D_{N-2}
    double* s_fwC = s_clo + 4;           // [warp][w]: d' entering `warp` = sum_w s_fwC[warp][w] * (forward total B of warp w); per dt
    double* s_bwK = s_fwC + kCnWarps * kCnWarps;   // unprojected: [warp][w] coefficients of the backward totals, [warp][kCnWarps] of X_{N-1}
    double* s_misc = s_bwK + kCnWarps * (kCnWarps + 1);          // [0] flag: some R_i < 0 outside the thread of node 0, [1] payoff at the last node
    double* s_dl = s_misc + 8;           // [warp] D of the warp's last slot (the neighbour above rebuilds that slot's X from it: no halo barrier)

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n = a.n_space;
    const int off = NS - (n - 1);        // node i lives in slot i + off
    auto at_node = [&](int i) { const int sl = i + off; const int tt = sl / NPT; return (sl - tt * NPT) * kCnThreads + tt; };
    for (int64_t opt = blockIdx.x; opt < a.N; opt += gridDim.x) {
        double X[NPT], Q[NPT], R[NPT];
        double xlast;                    // X_{N-1}
        {
            const double K = a.K[opt], Sm = cn_smax(a, opt);
            const double* xin = a.X_in ? a.X_in + opt * n : nullptr;
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                const int i = t * NPT + j - off;
                const double S = grid_point(i >= 0 ? i : 0, n, Sm);
                s_G[j * kCnThreads + t] = (i >= 0) ? K - S : kNegBig;     // payoff (applyConstraint, :106-107); own slots only
                X[j] = (i >= 0) ? (xin ? xin[i] : fmax(K - S, 0.0)) : 0.0;   // setInitialCondition, :48-52
                Q[j] = R[j] = 0.0;
            }
            xlast = xin ? xin[n - 1] : fmax(K - grid_point(n - 1, n, Sm), 0.0);
        }
        double Af[5], Ab[5], excA = 0.0, nxtA = 0.0, cR3 = 0.0, cR4 = 0.0, cinv = 0.0;
        // halo without a barrier of its own: hr (used by lane 31) = X of the first slot of the warp above = the value `v` that
        // entered this warp in the last back-substitution; hl (used by lane 0) = X of the last slot of the warp below, rebuilt
        // from that slot's D (published before the back-substitution barrier) and its per-dt constants Rl, Gl
        double hl = 0.0, hr = 0.0, Rl = 0.0, Gl = 0.0;
        // time stepping: either the caller's schedule, or solvePDE's countdown `while t > 0: dt = min(t, max_dt); ...; t -= dt`
        // (CrankNicolsonOptionSolver.py:36-45) in the same IEEE arithmetic, including its last ~1e-16-long step
        const bool sched = a.dts != nullptr;
        const int nsteps = sched ? a.n_steps[opt] : 0;
        double trem = sched ? 0.0 : a.T[opt];
        // T = inf / NaN, or more than 2^24 steps of max_dt, would keep the device busy (almost) forever: no steps at all then
        // (upstream would not return either); checked once, outside the time loop
        if (!sched && !(trem <= mdt_guard(a.max_dt[opt]))) trem = 0.0;
        const double mdt = sched ? 0.0 : a.max_dt[opt];
        double dt_prev = -1.0;
        __syncthreads();                          // the previous option's output pass is done with s_seq
        if (lane == 0) s_eL[warp] = X[0];
        if (lane == 31) s_eR[warp] = X[NPT - 1];
        __syncthreads();
        hl = (warp > 0) ? s_eR[warp - 1] : 0.0;
        hr = (warp < kCnWarps - 1) ? s_eL[warp + 1] : xlast;   // node N-2 sees X_{N-1}
        for (int st = 0; sched ? st < nsteps : trem > 0; ++st) {
            const double dt = sched ? a.dts[opt * a.max_steps + st] : fmin(trem, mdt);
            if (!sched) {
                if (!(dt > 0)) break;     // max_dt <= 0 never terminates upstream
                trem -= dt;
            }
            if (dt != dt_prev) {
                // ---- coefficients for this dt (setCoeff, :55-79) and the c' recurrence of the elimination
                __syncthreads();   // nobody still reads the previous dt's rows or s_seq
                const double r = a.r[opt], q = a.q[opt], sg = a.sigma[opt], K = cn_smax(a, opt);   // cold path: not kept in registers; K = grid end here
                const double dS = grid_point(1, n, K) - grid_point(0, n, K);
                if (t == 0) s_misc[1] = a.K[opt] - K;           // payoff at the last node (floor of X_{N-1})
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    const int i = t * NPT + j - off;
                    double al = 0.0, be = 0.0, ga = 0.0;
                    if (i >= 0) {
                        double S = grid_point(i, n, K);
                        double aa = sg * S / dS;
                        double m = (r - q) * S / dS;
                        al = 0.25 * dt * (aa * aa - m);
                        be = 0.5 * dt * (r + aa * aa);
                        ga = 0.25 * dt * (aa * aa + m);
                        s_seq[i] = 1.0 + be;      // diagonal, node order, for the sequential recurrence below
                    }
                    s_Q[j * kCnThreads + t] = al;
                    s_R[j * kCnThreads + t] = ga;
                    s_O[j * kCnThreads + t] = (i >= 0) ? 1.0 - be : 0.0;
                }
                __syncthreads();
                if (t == 0) {
                    // den_i = (1+beta_i) - (-alpha_i)(c'_{i-1}),  c'_i = -gamma_i / den_i   (rows 0..N-2)
                    double cp = 0.0;
                    int neg = 0;
                    const int t0 = off / NPT;     // the thread holding node 0: its own maps are applied one by one, never composed
                    for (int i = 0; i < n - 1; ++i) {
                        const int at = at_node(i);
                        double al = s_Q[at], ga = s_R[at];
                        double den = s_seq[i] + al * cp;
                        double rd = 1.0 / den;
                        double Ri = ga * rd;
                        s_Q[at] = al * rd;
                        s_R[at] = Ri;
                        s_O[at] = s_O[at] * rd;
                        cp = -Ri;
                        if (Ri < 0.0 && (i + off) / NPT > t0) neg = 1;
                    }
                    s_misc[0] = (double)neg;
                }
                __syncthreads();
                // ---- step-invariant A-components of the composed maps
                double aF = 1.0, aB = 1.0;
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    Q[j] = s_Q[j * kCnThreads + t];
                    R[j] = s_R[j * kCnThreads + t];
                    aF *= Q[j];
                    aB *= R[j];
                }
#pragma unroll
                for (int lv = 0; lv < 5; ++lv) {
                    const int d = 1 << lv;
                    // multiplier of the incoming B at this level; 0 where the lane has no partner, so that the
                    // per-step update is an unconditional FMA
                    Af[lv] = (lane >= d) ? aF : 0.0;
                    double o = __shfl_up_sync(kFullMask, aF, d);
                    if (lane >= d) aF *= o;
                    Ab[lv] = (lane + d < 32) ? aB : 0.0;
                    o = __shfl_down_sync(kFullMask, aB, d);
                    if (lane + d < 32) aB *= o;
                }
                excA = __shfl_up_sync(kFullMask, aF, 1);           // inclusive A of the lane below
                nxtA = __shfl_down_sync(kFullMask, aB, 1);         // inclusive A of the lane above
                if (lane == 31) s_fwA[warp] = aF;
                if (lane == 0) s_bwA[warp] = aB;
                {   // closure of the 4-term last row: X_{N-2..N-4} are affine in x = X_{N-1} with slopes m2, m3, m4
                    const double m2 = s_R[at_node(n - 2)];
                    cR3 = s_R[at_node(n - 3)];
                    cR4 = s_R[at_node(n - 4)];
                    const double m3 = cR3 * m2, m4 = cR4 * m3;
                    cinv = 1.0 / (2.0 - m4 + 4.0 * m3 - 5.0 * m2);
                }
                if (warp > 0) {    // constants of the last slot of the warp below
                    Rl = s_R[(NPT - 1) * kCnThreads + warp * 32 - 1];
                    Gl = s_G[(NPT - 1) * kCnThreads + warp * 32 - 1];
                }
                __syncthreads();   // warp totals A visible
                if (t < kCnWarps * kCnWarps) {
                    // d' entering warp `wr` = sum over w < wr of (prod_{w < v < wr} A_v) B_w: the products are per dt
                    const int wr = t / kCnWarps, w = t % kCnWarps;
                    double cw = 0.0;
                    if (w < wr) {
                        cw = 1.0;
                        for (int v = w + 1; v < wr; ++v) cw *= s_fwA[v];
                    }
                    s_fwC[t] = cw;
                    if (!PROJ) {
                        // value entering warp `wr` from above = sum over w > wr of (prod_{wr < v < w} A_v) B_w + (prod_{v > wr} A_v) X_{N-1}
                        double ck = 0.0;
                        if (w > wr) {
                            ck = 1.0;
                            for (int v = wr + 1; v < w; ++v) ck *= s_bwA[v];
                        }
                        s_bwK[wr * (kCnWarps + 1) + w] = ck;
                        if (w == 0) {
                            double cx = 1.0;
                            for (int v = wr + 1; v < kCnWarps; ++v) cx *= s_bwA[v];
                            s_bwK[wr * (kCnWarps + 1) + kCnWarps] = cx;
                        }
                    }
                }
                dt_prev = dt;
                __syncthreads();   // the warp-level coefficients are visible
            }
            // ---- right-hand side, already scaled by 1/den: P_i = Q_i X_{i-1} + O_i X_i + R_i X_{i+1}
            double xl = __shfl_up_sync(kFullMask, X[NPT - 1], 1);
            double xr = __shfl_down_sync(kFullMask, X[0], 1);
            if (lane == 0) xl = hl;
            if (lane == 31) xr = hr;
            double D[NPT];
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                const double xm = (j == 0) ? xl : X[j - 1];
                const double xp = (j == NPT - 1) ? xr : X[j + 1];
                D[j] = fma(R[j], xp, fma(s_O[j * kCnThreads + t], X[j], Q[j] * xm));   // P_j for now
            }
            // ---- forward elimination d'_i = P_i + Q_i d'_{i-1}: prefix scan of the B-components
            double b = D[0];
#pragma unroll
            for (int j = 1; j < NPT; ++j) b = fma(Q[j], b, D[j]);
#pragma unroll
            for (int lv = 0; lv < 5; ++lv) {
                const int d = 1 << lv;
                b = fma(Af[lv], __shfl_up_sync(kFullMask, b, d), b);
            }
            if (lane == 31) s_fwB[warp] = b;
            __syncthreads();
            double din = 0.0;  // d' entering this warp: a dot product with per-dt coefficients (zero for w >= warp), no selects
#pragma unroll
            for (int w = 0; w < kCnWarps - 1; ++w) din = fma(s_fwC[warp * kCnWarps + w], s_fwB[w], din);
            const double eb = __shfl_up_sync(kFullMask, b, 1);
            double dprev = (lane == 0) ? din : fma(excA, din, eb);
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                dprev = fma(Q[j], dprev, D[j]);
                D[j] = dprev;
            }
            if (t >= kCnThreads - 3) {      // D of the last three scanned nodes N-4, N-3, N-2 (slots NS-3 .. NS-1)
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    const int sl = t * NPT + j;
                    if (sl >= NS - 3) s_clo[sl - (NS - 3)] = D[j];
                }
            }
            const bool fallback = PROJ && s_misc[0] != 0.0;
            if (!fallback) {
                // ---- (projected) back-substitution X_i = max(D_i + R_i X_{i+1}, g_i): suffix scan of (B, C)
                double B = 0.0, C = kNegBig;
#pragma unroll
                for (int j = NPT - 1; j >= 0; --j) {
                    if (PROJ) C = dmax(fma(R[j], C, D[j]), s_G[j * kCnThreads + t]);
                    B = fma(R[j], B, D[j]);
                }
#pragma unroll
                for (int lv = 0; lv < 5; ++lv) {
                    const int d = 1 << lv;
                    const double oB = __shfl_down_sync(kFullMask, B, d);
                    double oC = 0.0;
                    if (PROJ) oC = __shfl_down_sync(kFullMask, C, d);
                    if (PROJ && lane + d < 32) C = dmax(fma(Ab[lv], oC, B), C);
                    B = fma(Ab[lv], oB, B);
                }
                if (lane == 0) {
                    s_bwB[warp] = B;
                    if (PROJ) s_bwC[warp] = C;
                }
                if (lane == 31) s_dl[warp] = D[NPT - 1];
                __syncthreads();
                // last row: x = X_{N-1} from D_{N-4..N-2}
                const double p2 = s_clo[2];
                const double p3 = fma(cR3, p2, s_clo[1]);
                const double p4 = fma(cR4, p3, s_clo[0]);
                xlast = (p4 - 4.0 * p3 + 5.0 * p2) * cinv;
                if (PROJ) xlast = dmax(xlast, s_misc[1]);
                double v = xlast;
                if (PROJ) {
                    for (int w = kCnWarps - 1; w > warp; --w)       // warp-uniform trip count: only the warps above this one
                        v = dmax(fma(s_bwA[w], v, s_bwB[w]), s_bwC[w]);
                } else {    // affine maps superpose: a dot product with per-dt coefficients
                    const double* ck = s_bwK + warp * (kCnWarps + 1);
                    v = ck[kCnWarps] * xlast;
#pragma unroll
                    for (int w = 1; w < kCnWarps; ++w) v = fma(ck[w], s_bwB[w], v);
                }
                const double nB = __shfl_down_sync(kFullMask, B, 1);
                double nC = 0.0;
                if (PROJ) nC = __shfl_down_sync(kFullMask, C, 1);
                double xin = v;
                if (lane != 31) {
                    xin = fma(nxtA, v, nB);
                    if (PROJ) xin = dmax(xin, nC);
                }
#pragma unroll
                for (int j = NPT - 1; j >= 0; --j) {
                    double xn = fma(R[j], xin, D[j]);
                    if (PROJ) xn = dmax(xn, s_G[j * kCnThreads + t]);
                    X[j] = xn;
                    xin = xn;
                }
                hr = (warp < kCnWarps - 1) ? v : xlast;
                if (warp > 0) {     // X of the last slot of the warp below: the same arithmetic its owner performs (lane 0's X[0] enters it)
                    hl = fma(Rl, X[0], s_dl[warp - 1]);
                    if (PROJ) hl = dmax(hl, Gl);
                }
            } else {
                // some R_i < 0 away from S = 0 (q >> r with tiny sigma): the max-affine maps are not monotone there,
                // so this step falls back to a sequential projected back-substitution by one thread
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    const int i = t * NPT + j - off;
                    if (i >= 0) s_seq[i] = D[j];
                }
                __syncthreads();
                if (t == 0) {
                    const double K = a.K[opt], Sm = cn_smax(a, opt);
                    const double p2 = s_clo[2];
                    const double p3 = fma(cR3, p2, s_clo[1]);
                    const double p4 = fma(cR4, p3, s_clo[0]);
                    double x = dmax((p4 - 4.0 * p3 + 5.0 * p2) * cinv, s_misc[1]);
                    s_seq[n - 1] = x;
                    for (int i = n - 2; i >= 0; --i) {
                        x = dmax(fma(s_R[at_node(i)], x, s_seq[i]), K - grid_point(i, n, Sm));
                        s_seq[i] = x;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < NPT; ++j) {
                    const int i = t * NPT + j - off;
                    X[j] = (i >= 0) ? s_seq[i] : 0.0;
                }
                xlast = s_seq[n - 1];
                {   // halos from the sequential result
                    const int il = warp * 32 * NPT - 1 - off, ir = (warp + 1) * 32 * NPT - off;
                    hl = (warp > 0 && il >= 0) ? s_seq[il] : 0.0;
                    hr = (warp < kCnWarps - 1) ? s_seq[ir] : xlast;
                }
                __syncthreads();   // before the next step's forward totals / the next fallback overwrite s_seq
            }
        }
        // ---- output: X (optional) and np.interp(S0, S, X) (:32-34)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            const int i = t * NPT + j - off;
            if (i >= 0) {
                s_seq[i] = X[j];
                if (a.X_out) a.X_out[opt * n + i] = X[j];
            }
        }
        if (t == 0) {
            s_seq[n - 1] = xlast;
            if (a.X_out) a.X_out[opt * n + n - 1] = xlast;
        }
        __syncthreads();
        if (t == 0) a.price[opt] = interp_uniform(a.S0[opt], n, cn_smax(a, opt), [&](int i) { return s_seq[i]; });
    }
}

// ---------------------------------------------------------------- mode 1: the reference's SOR-then-project loop
// One thread per option; arrays in global scratch laid out [node][option] so that neighbouring threads coalesce.
__global__ void __launch_bounds__(64) cn_sor_kernel(CnArgs a) {
    int64_t opt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (opt >= a.N) return;
    const int n = a.n_space;
    const int64_t ld = a.N;
    double* X = a.ws + opt;              // X[i] at X[i*ld]
    double* b = X + (int64_t)n * ld;
    double* lo = b + (int64_t)n * ld;
    double* di = lo + (int64_t)n * ld;
    double* up = di + (int64_t)n * ld;
    double* xo = up + (int64_t)n * ld;
    const double r = a.r[opt], q = a.q[opt], sg = a.sigma[opt], K = a.K[opt], Sm = cn_smax(a, opt);
    const double omega = a.omega;
    for (int i = 0; i < n; ++i) X[i * ld] = a.X_in ? a.X_in[opt * n + i] : fmax(K - grid_point(i, n, Sm), 0.0);
    const double dS = grid_point(1, n, Sm) - grid_point(0, n, Sm);
    double last_err = 0.0;
    int last_it = 0;
    const bool sched = a.dts != nullptr;
    const int nsteps = sched ? a.n_steps[opt] : 0;
    double trem = sched ? 0.0 : a.T[opt];
    if (!sched && !(trem <= mdt_guard(a.max_dt[opt]))) trem = 0.0;
    const double mdt = sched ? 0.0 : a.max_dt[opt];
    for (int st = 0; sched ? st < nsteps : trem > 0; ++st) {
        const double dt = sched ? a.dts[opt * a.max_steps + st] : fmin(trem, mdt);
        if (!sched) {
            if (!(dt > 0)) break;
            trem -= dt;
        }
        for (int i = 0; i < n - 1; ++i) {
            double S = grid_point(i, n, Sm);
            double aa = sg * S / dS;
            double alpha = 0.25 * dt * (aa * aa - (r - q) * S / dS);
            double beta = 0.5 * dt * (r + aa * aa);
            double gamma = 0.25 * dt * (aa * aa + (r - q) * S / dS);
            if (i == 0) {
                b[0] = X[0] * (1 - beta); lo[0] = 0; di[0] = 1 + beta; up[0] = 0;
            } else {
                b[i * ld] = alpha * X[(i - 1) * ld] + (1 - beta) * X[i * ld] + gamma * X[(i + 1) * ld];
                lo[i * ld] = -alpha; di[i * ld] = 1 + beta; up[i * ld] = -gamma;
            }
        }
        int it = 0;
        double err = 1000.0;
        while (err > a.tol && it < a.max_iter) {   // solvePSOR, :84-104
            ++it;
            for (int i = 0; i < n; ++i) xo[i * ld] = X[i * ld];
            for (int k = 0; k < n - 1; ++k) {
                double d = di[k * ld];
                double xk = (1 - omega) * X[k * ld] + omega * b[k * ld] / d;
                xk -= up[k * ld] * X[(k + 1) * ld] * omega / d;
                xk -= lo[k * ld] * X[(k == 0 ? n - 1 : k - 1) * ld] * omega / d;   // k == 0 reads X[-1] with a zero coefficient
                X[k * ld] = xk;
            }
            const int k = n - 2;   // the loop index leaks into the last-row update (:98)
            X[(n - 1) * ld] = (1 - omega) * X[k * ld] + omega * b[k * ld] / di[k * ld];
            const double lastc[4] = {-1.0, 4.0, -5.0, 2.0};
            for (int jj = 0; jj < 4; ++jj) X[(n - 1) * ld] -= lastc[jj] * X[(n - 4 + jj) * ld] * omega / 2.0;
            double ss = 0.0;
            for (int i = 0; i < n; ++i) {
                double v = fmax(X[i * ld], K - grid_point(i, n, Sm));   // applyConstraint on the whole vector
                X[i * ld] = v;
                double dd = xo[i * ld] - v;
                ss += dd * dd;
            }
            err = sqrt(ss);
        }
        last_err = err;        // self.err / self.iter, :103-104
        last_it = it;
    }
    if (a.err_out) a.err_out[opt] = last_err;
    if (a.iter_out) a.iter_out[opt] = last_it;
    if (a.X_out) for (int i = 0; i < n; ++i) a.X_out[opt * n + i] = X[i * ld];
    a.price[opt] = interp_uniform(a.S0[opt], n, Sm, [&](int i) { return X[i * ld]; });
}

// ---------------------------------------------------------------- setCoeff (:55-79) as data: the rows and the right-hand side
// of ONE time step from a given state, one thread per (option, node).  lo / di / up are the three diagonals (row 0 has only
// the diagonal; row N-1 is the constant 4-term closure -1, 4, -5, 2 and is reported as lo = -5, di = 2, up = 0, b = 0).
__global__ void cn_rows_kernel(int64_t N, int n, const double* r, const double* q, const double* sigma, const double* K,
                               const double* Smax, const double* dt, const double* X, double* lo, double* di, double* up,
                               double* b) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N * n) return;
    const int64_t opt = g / n;
    const int i = (int)(g - opt * n);
    const double Sm = Smax ? Smax[opt] : 3.0 * K[opt];
    const double dS = grid_point(1, n, Sm) - grid_point(0, n, Sm);
    const double* x = X + opt * n;
    double vlo = 0.0, vdi = 0.0, vup = 0.0, vb = 0.0;
    if (i == n - 1) {
        vlo = -5.0; vdi = 2.0;
    } else {
        const double S = grid_point(i, n, Sm);
        const double aa = sigma[opt] * S / dS;
        const double m = (r[opt] - q[opt]) * S / dS;
        const double alpha = 0.25 * dt[opt] * (aa * aa - m);
        const double beta = 0.5 * dt[opt] * (r[opt] + aa * aa);
        const double gamma = 0.25 * dt[opt] * (aa * aa + m);
        vdi = 1 + beta;
        if (i == 0) {
            vb = x[0] * (1 - beta);
        } else {
            vb = alpha * x[i - 1] + (1 - beta) * x[i] + gamma * x[i + 1];
            vlo = -alpha;
            vup = -gamma;
        }
    }
    lo[g] = vlo; di[g] = vdi; up[g] = vup; b[g] = vb;
}

template <int NPT, bool PROJ>
static int launch_scan(const CnArgs& a, cudaStream_t st) {
    constexpr int NS = kCnThreads * NPT;
    size_t sm = sizeof(double) * (size_t)(5 * NS + 2 + 7 * kCnWarps + kCnWarps * (2 * kCnWarps + 1) + 8 + kCnWarps);
    FAOP_CUDA_OK(cudaFuncSetAttribute(cn_scan_kernel<NPT, PROJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int64_t grid = a.N < (int64_t)kSMs * 8 ? a.N : (int64_t)kSMs * 8;
    cn_scan_kernel<NPT, PROJ><<<(unsigned)grid, kCnThreads, sm, st>>>(a);
    count_launch();
    return launch_status();
}

template <bool PROJ>
static int launch_scan_npt(const CnArgs& a, cudaStream_t st) {
    int npt = (a.n_space - 1 + kCnThreads - 1) / kCnThreads;   // the N-1 scanned nodes; X_{N-1} is a scalar
    if (npt <= 1) return launch_scan<1, PROJ>(a, st);
    if (npt <= 2) return launch_scan<2, PROJ>(a, st);
    if (npt <= 4) return launch_scan<4, PROJ>(a, st);
    if (npt <= 8) return launch_scan<8, PROJ>(a, st);
    if (npt <= 16) return launch_scan<16, PROJ>(a, st);
    return FAOP_ERANGE;
}

}  // namespace faop

using namespace faop;

extern "C" {

int faop_cn_workspace_bytes(int64_t N, int32_t n_space, int32_t mode, size_t* bytes) {
    if (!bytes || N < 0 || n_space < 5) return FAOP_EINVAL;
    *bytes = (mode == 1) ? sizeof(double) * 6 * (size_t)N * (size_t)n_space : 0;
    return 0;
}

static int cn_launch(CnArgs& a, const faop_cn_extra* ex, void* workspace, cudaStream_t st) {
    a.Smax = ex ? ex->Smax : nullptr;
    a.X_in = ex ? ex->X_in : nullptr;
    a.err_out = ex ? ex->err_out : nullptr;
    a.iter_out = ex ? ex->iter_out : nullptr;
    a.ws = (double*)workspace;
    if (a.mode == 1) {
        cn_sor_kernel<<<(unsigned)((a.N + 63) / 64), 64, 0, st>>>(a);
        count_launch();
        return launch_status();
    }
    // the direct modes run no SOR loop: err / iter keep the constructor's zeros (:24-25)
    if (a.err_out) FAOP_CUDA_OK(cudaMemsetAsync(a.err_out, 0, sizeof(double) * (size_t)a.N, st));
    if (a.iter_out) FAOP_CUDA_OK(cudaMemsetAsync(a.iter_out, 0, sizeof(int32_t) * (size_t)a.N, st));
    return a.mode == 2 ? launch_scan_npt<true>(a, st) : launch_scan_npt<false>(a, st);
}

int faop_cn_put_batch(int64_t N, int32_t n_space, int32_t mode, const double* r, const double* q, const double* sigma,
                      const double* K, const double* S0, const double* dts, const int32_t* n_steps, int32_t max_steps,
                      double omega, double tol, int32_t max_iter, const faop_cn_extra* extra, double* price, double* X_out,
                      void* workspace, void* cuda_stream) {
    if (N < 0 || n_space < 5 || mode < 0 || mode > 2 || max_steps < 0) return FAOP_EINVAL;
    if (n_space > kCnThreads * 16) return FAOP_ERANGE;
    if (N == 0) return 0;
    if (!r || !q || !sigma || !K || !S0 || !dts || !n_steps || !price) return FAOP_EINVAL;
    if (mode == 1 && !workspace) return FAOP_EINVAL;
    CnArgs a;
    a.N = N; a.n_space = n_space; a.max_steps = max_steps; a.mode = mode; a.max_iter = max_iter;
    a.r = r; a.q = q; a.sigma = sigma; a.K = K; a.S0 = S0; a.dts = dts; a.n_steps = n_steps;
    a.T = nullptr; a.max_dt = nullptr;
    a.omega = omega; a.tol = tol; a.price = price; a.X_out = X_out;
    return cn_launch(a, extra, workspace, (cudaStream_t)cuda_stream);
}

int faop_cn_put_countdown_batch(int64_t N, int32_t n_space, int32_t mode, const double* r, const double* q,
                                const double* sigma, const double* K, const double* T, const double* S0, const double* max_dt,
                                double omega, double tol, int32_t max_iter, const faop_cn_extra* extra, double* price,
                                double* X_out, void* workspace, void* cuda_stream) {
    if (N < 0 || n_space < 5 || mode < 0 || mode > 2) return FAOP_EINVAL;
    if (n_space > kCnThreads * 16) return FAOP_ERANGE;
    if (N == 0) return 0;
    if (!r || !q || !sigma || !K || !T || !S0 || !max_dt || !price) return FAOP_EINVAL;
    if (mode == 1 && !workspace) return FAOP_EINVAL;
    CnArgs a;
    a.N = N; a.n_space = n_space; a.max_steps = 0; a.mode = mode; a.max_iter = max_iter;
    a.r = r; a.q = q; a.sigma = sigma; a.K = K; a.S0 = S0; a.dts = nullptr; a.n_steps = nullptr;
    a.T = T; a.max_dt = max_dt;
    a.omega = omega; a.tol = tol; a.price = price; a.X_out = X_out;
    return cn_launch(a, extra, workspace, (cudaStream_t)cuda_stream);
}

int faop_cn_rows_batch(int64_t N, int32_t n_space, const double* r, const double* q, const double* sigma, const double* K,
                       const double* Smax, const double* dt, const double* X, double* lo, double* di, double* up, double* b,
                       void* cuda_stream) {
    if (N < 0 || n_space < 5) return FAOP_EINVAL;
    if (N == 0) return 0;
    if (!r || !q || !sigma || !K || !dt || !X || !lo || !di || !up || !b) return FAOP_EINVAL;
    const int64_t tot = N * n_space;
    cn_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(N, n_space, r, q, sigma, K, Smax, dt, X,
                                                                                      lo, di, up, b);
    count_launch();
    return launch_status();
}

}  // extern "C"
