// faop_alo.cuh — device building blocks of the ALO boundary solve, shared by the generic and the
// specialised kernels.  One option per warp; lane = quadrature index (k, k+32, ...).
//
// Formulation (validated against the oracle to <1e-13 relative on the boundary):
//   tau_i - u_ik = tau_i h_k^2,  u_ik = tau_i g_k,  h_k = (1+y_k)/2,  g_k = 1 - h_k^2
//   sigma sqrt(tau_i-u) = sigma sqrt(tau_i) h_k            (no cancellation, no per-point sqrt)
//   ln(B_i/B(u)) = L_i + om sqrt(H(u)),  L_i = ln(B_i/X),  om = +1 put / -1 call   (log space)
//   d+ = lnz/(sigma sqrt(tau_i) h_k) + ((r-q)+sigma^2/2) sqrt(tau_i) h_k / sigma,  d- = d+ - sigma sqrt(tau_i) h_k
//   e^{qu} phi(d+) = exp(q u - d+^2/2)/sqrt(2 pi)   (one exp),   likewise e^{ru} phi(d-)
//   e^{qu} Phi(x)  = 1/2 exp(q u - x^2/2) erfcx(-x/sqrt2)            for x <= 0
//                  = e^{qu} - 1/2 exp(q u - x^2/2) erfcx(x/sqrt2)    for x  > 0
//   quadrature weight W_ik = w_k tau_i h_k, and W_ik/(sigma sqrt(tau_i-u)) = w_k sqrt(tau_i)/sigma
#pragma once
#include "faop_common.cuh"
#include "faop_kernels.h"

namespace faop {

constexpr unsigned kFull = 0xffffffffu;

FAOP_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Per-option scalars every lane keeps in registers.
struct Opt {
    double r, q, sg, K, T, t, S, X, om;
    double Tc;   // end of the collocation domain: the tau_max attribute (Base.py:28), = T unless the caller changed it
    bool put;
};

// Per-node constants (depend on option and node only, not on the iterate).
struct NodeC {
    double tau;    // tau_i
    double cinv;   // 1/(sigma sqrt(tau_i))
    double apc;    // ((r-q) + sigma^2/2) sqrt(tau_i)/sigma
    double ssu;    // sigma sqrt(tau_i)
    double qtau;   // q tau_i
    double rtau;   // r tau_i
};

FAOP_D NodeC node_consts(const Opt& o, double tau) {
    NodeC c;
    double sqt = sqrt(tau);
    c.tau = tau;
    c.ssu = o.sg * sqt;
    c.cinv = 1.0 / c.ssu;
    c.apc = ((o.r - o.q) + 0.5 * o.sg * o.sg) * sqt / o.sg;
    c.qtau = o.q * tau;
    c.rtau = o.r * tau;
    return c;
}

// e^{a} Phi(x) given E = exp(a - x^2/2): see header.  ea = e^{a} is only used when x > 0.
FAOP_D double scaled_cdf(double x, double E, double ea) {
    double ec = 0.5 * E * erfcx(fabs(x) * kInvSqrt2);
    return (x <= 0) ? ec : ea - ec;
}

// One quadrature point of the boundary equations.  Accumulates the lane-local partial sums
//   Solver A (FastAmericanOptionSolverA.py:24-44, 55-83):
//     a1 += w h [+-e^{qu} Phi(+-d+)] + (cinv/sqrt(2pi)) w e^{qu}phi(d+)      -> K12 = tau a1
//     a2 += w e^{ru}phi(d-) sqrt(2pi)                                         -> K3  = tau cinv/sqrt(2pi) a2
//     a3 += (w/h) d- e^{ru}phi(d-) sqrt(2pi)   a4 += w e^{qu}phi(d+) sqrt(2pi) (1 - d+ cinv/h)     (derivative sums)
//   Solver B (FastAmericanOptionSolverB.py:25-39, 59-75):
//     a1 += w h e^{qu} Phi(om d+)   a2 += w h e^{ru} Phi(om d-)   a3 += w e^{ru}phi(d-) sqrt(2pi)   a4 += w e^{qu}phi(d+) sqrt(2pi)
template <int SOLVER, bool DERIV>
FAOP_D void alo_point(const Opt& o, const NodeC& c, double lnz, double w, double h, double rh, double g,
                      double& a1, double& a2, double& a3, double& a4) {
    double inv = c.cinv * rh;
    double dp = fma(lnz, inv, c.apc * h);
    double dm = fma(-c.ssu, h, dp);
    double qu = c.qtau * g, ru = c.rtau * g;
    double Eq = exp(fma(-0.5 * dp, dp, qu));
    double Er = exp(fma(-0.5 * dm, dm, ru));
    if (SOLVER == 0) {
        double eq = exp(qu);
        double cdf = o.om * scaled_cdf(o.om * dp, Eq, eq);  // put: +e^{qu}Phi(d+); call: -e^{qu}Phi(-d+)
        a1 = fma(w * h, cdf, a1);
        a1 = fma(c.cinv * kInvSqrt2Pi * w, Eq, a1);
        a2 = fma(w, Er, a2);
        if (DERIV) {
            a3 = fma(w * rh * dm, Er, a3);
            a4 = fma(w * Eq, 1.0 - dp * inv, a4);
        }
    } else {
        double eq = exp(qu), er = exp(ru);
        a1 = fma(w * h, scaled_cdf(o.om * dp, Eq, eq), a1);
        a2 = fma(w * h, scaled_cdf(o.om * dm, Er, er), a2);
        if (DERIV) {
            a3 = fma(w, Er, a3);
            a4 = fma(w, Eq, a4);
        }
    }
}

struct NodeOut {
    double Nv, Dv, Np, Dp, f, fp;
};

// Node-level closure of N, D (, N', D') and f, f' from the reduced sums
// (FastAmericanOptionSolverA.py:5-22, 46-74; FastAmericanOptionSolverB.py:5-23, 41-57;
//  compute_f_and_fprime, FastAmericanOptionSolverBase.py:267-279).
// L = ln(B_i/X), lnXK = ln(X/K), disc = K e^{-(r-q) tau_i}.
template <int SOLVER, bool DERIV>
FAOP_D NodeOut alo_node(const Opt& o, const NodeC& c, double Bi, double L, double lnXK, double disc,
                        double s1, double s2, double s3, double s4) {
    NodeOut out;
    double LK = L + lnXK;
    double dpK = fma(LK, c.cinv, c.apc);
    double dmK = dpK - c.ssu;
    double pm = norm_pdf(dmK), pp = norm_pdf(dpK);
    out.Np = 0.0;
    out.Dp = 0.0;
    if (SOLVER == 0) {
        double K12 = c.tau * s1;
        double K3 = c.tau * c.cinv * kInvSqrt2Pi * s2;
        out.Nv = pm * c.cinv + o.r * K3;
        out.Dv = pp * c.cinv + o.om * norm_cdf(o.om * dpK) + o.q * K12;
        if (DERIV) {
            double s2g = o.sg * o.sg;
            double Nsum = s3 * kInvSqrt2Pi / (Bi * s2g);
            double Dsum = c.tau * c.cinv * kInvSqrt2Pi / Bi * s4;
            out.Np = -dmK * pm / (Bi * s2g * c.tau) - o.r * Nsum;
            out.Dp = pp * c.cinv / Bi * (1.0 - dpK * c.cinv) + o.q * Dsum;
        }
    } else {
        out.Nv = norm_cdf(o.om * dmK) + o.r * (c.tau * s2);
        out.Dv = norm_cdf(o.om * dpK) + o.q * (c.tau * s1);
        if (DERIV) {
            double sc = c.tau * c.cinv * kInvSqrt2Pi / Bi;
            double Nsum = sc * s3, Dsum = sc * s4;
            if (o.put) {
                out.Np = pm * c.cinv / Bi + o.r * Nsum;
                out.Dp = pp * c.cinv / Bi + o.q * Dsum;
            } else {  // quirk kept: the call's D' integrates Nprime_integrand (FastAmericanOptionSolverB.py:57)
                out.Np = -pm * c.cinv / Bi - o.r * Nsum;
                out.Dp = -pp * c.cinv / Bi - o.q * Nsum;
            }
        }
    }
    out.f = disc * out.Nv / out.Dv;
    out.fp = 1.0;
    if (DERIV) out.fp = disc * (out.Np / out.Dv - out.Dp * out.Nv / (out.Dv * out.Dv));
    return out;
}

// One point of the price integral, v_integrand_12 (FastAmericanOptionSolverBase.py:211-219),
// without the weight: put  r K e^{-r s} Phi(-d-) - q S e^{-q s} Phi(-d+),  s = tau h^2
//                     call q S e^{-q s} Phi(+d+) - r K e^{-r s} Phi(+d-)
FAOP_D double price_point(const Opt& o, const NodeC& c, double lnz, double h, double rh) {
    double inv = c.cinv * rh;
    double dp = fma(lnz, inv, c.apc * h);
    double dm = fma(-c.ssu, h, dp);
    double h2 = h * h;
    double er = exp(-c.rtau * h2), eq = exp(-c.qtau * h2);
    double a = o.r * o.K * er * norm_cdf(-o.om * dm);
    double b = o.q * o.S * eq * norm_cdf(-o.om * dp);
    return o.om * (a - b);
}

// ---------------------------------------------------------------- per-warp shared-memory view (generic kernels)
struct WarpSmem {
    double *B, *L, *H, *A, *tau, *cinv, *disc, *s1, *s2, *s3, *s4;
    static __host__ __device__ int doubles(int nn) { return 11 * nn; }
    __device__ WarpSmem(double* base, int nn) {
        B = base; L = B + nn; H = L + nn; A = H + nn; tau = A + nn; cinv = tau + nn; disc = cinv + nn;
        s1 = disc + nn; s2 = s1 + nn; s3 = s2 + nn; s4 = s3 + nn;
    }
};

struct CtaTables {
    const double *wl, *hl, *rhl, *sgl, *wm, *hm, *rhm, *sgm, *c, *cs;
    __device__ CtaTables(const double* s, const TableLayout& t) {
        wl = s + t.off_wl; hl = s + t.off_hl; rhl = s + t.off_rhl; sgl = s + t.off_sgl;
        wm = s + t.off_wm; hm = s + t.off_hm; rhm = s + t.off_rhm; sgm = s + t.off_sgm;
        c = s + t.off_c; cs = s + t.off_cos;
    }
};

FAOP_D void stage_tables(double* s_tab, const double* g_tab, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
}

FAOP_D Opt load_option(const AloArgs& a, int64_t opt) {
    Opt o;
    o.r = a.r[opt]; o.q = a.q[opt]; o.sg = a.sigma[opt]; o.K = a.K[opt]; o.T = a.T[opt];
    o.Tc = a.tau_max ? a.tau_max[opt] : o.T;
    o.t = a.t ? a.t[opt] : 0.0;
    o.S = a.S ? a.S[opt] : 0.0;
    o.put = a.is_put[opt] != 0;
    o.X = b_at_zero(o.r, o.q, o.K, o.put);
    o.om = o.put ? 1.0 : -1.0;
    return o;
}

// L_j = ln(B_j/X), H_j = L_j^2 (FastAmericanOptionSolverBase.py:184) and the DCT-I fit
// a_k = (2/n) sum'' H_j cos(j k pi/n) (ChebyshevInterpolation.py:43-51), lanes over j / k.
FAOP_D void fit_boundary(const WarpSmem& w, const CtaTables& tb, int n, double X, int lane) {
    for (int j = lane; j <= n; j += 32) {
        double L = log(w.B[j] / X);
        w.L[j] = L;
        w.H[j] = L * L;
    }
    __syncwarp();
    const int n2 = 2 * n;
    for (int k = lane; k <= n; k += 32) {
        double ans = 0.0;
        int idx = 0;  // (j*k) mod 2n
        for (int j = 0; j <= n; ++j) {
            double term = w.H[j] * tb.cs[idx];
            if (j == 0 || j == n) term *= 0.5;
            ans += term;
            idx += k;
            if (idx >= n2) idx -= n2;
        }
        w.A[k] = ans * (2.0 / n);
    }
    __syncwarp();
}

FAOP_D NodeC node_from_smem(const WarpSmem& w, const Opt& o, int j) {
    NodeC c;
    c.tau = w.tau[j];
    c.cinv = w.cinv[j];
    c.ssu = o.sg * sqrt(c.tau);
    c.apc = ((o.r - o.q) + 0.5 * o.sg * o.sg) * (c.ssu / (o.sg * o.sg));
    c.qtau = o.q * c.tau;
    c.rtau = o.r * c.tau;
    return c;
}

FAOP_D void setup_nodes(const WarpSmem& w, const CtaTables& tb, int n, const Opt& o, int lane) {
    for (int j = lane; j <= n; j += 32) {
        double c = tb.c[j];
        double tau = (c + 1) * (c + 1) * o.Tc / 4;
        w.tau[j] = tau;
        w.cinv[j] = 1.0 / (o.sg * sqrt(tau));
        w.disc[j] = o.K * exp(-tau * (o.r - o.q));
    }
}

// Snapshot mode: append [err, count, a_0..a_n] of the boundary currently in w.B (see AloArgs::snap).  Overwrites w.L/H/A.
// The catch-all row (err == 0) always lands: if the table is full it replaces the last row.
FAOP_D void snap_append(const AloArgs& a, int64_t opt, int& ns, double err, int count, const WarpSmem& w,
                        const CtaTables& tb, int n, double X, int lane) {
    if (ns >= a.snap_cap) {
        if (err != 0.0) return;
        ns = a.snap_cap - 1;
    }
    fit_boundary(w, tb, n, X, lane);
    double* dst = a.snap + ((int64_t)opt * a.snap_cap + ns) * (n + 3);
    if (lane == 0) {
        dst[0] = err;
        dst[1] = (double)count;
    }
    for (int k = lane; k <= n; k += 32) dst[2 + k] = w.A[k];
    ++ns;
    __syncwarp();
}

// Price integral on the boundary currently in w.B: american_value_with_known_boundary
// (FastAmericanOptionSolverBase.py:92-103).  Returns price; *eur gets the European value.
FAOP_D double price_integral(const WarpSmem& w, const CtaTables& tb, const TableLayout& tl, const Opt& o, int lane,
                             double* eur) {
    double tau = o.T - o.t;
    double e = euro_value(tau, o.S, o.r, o.q, o.sg, o.K, o.put);
    *eur = e;
    if (tau == 0) return e + 0.0;  // quadrature_sum returns 0 at tau == 0 (:382-383)
    fit_boundary(w, tb, tl.n, o.X, lane);
    NodeC c = node_consts(o, tau);
    double LS = log(o.S / o.X);
    double zf = sqrt(4 * tau / o.Tc);  // z = sqrt(4 u/tau_max) - 1 with u = tau g_k
    double acc = 0;
    for (int k = lane; k < tl.MP; k += 32) {
        double h = tb.hm[k];
        double z = fma(zf, tb.sgm[k], -1.0);
        double Hu = clenshaw(tl.n, w.A, z);
        double root = sqrt(Hu < 0.0 ? 0.0 : Hu);  // np.maximum(0, H): NaN propagates (fmax would drop it)
        double lnz = fma(o.om, root, LS);
        acc = fma(tb.wm[k] * h, price_point(o, c, lnz, h, tb.rhm[k]), acc);
    }
    acc = warp_sum(acc);
    return e + acc * tau;
}

}  // namespace faop
