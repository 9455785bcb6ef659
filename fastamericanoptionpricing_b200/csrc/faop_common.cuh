// faop_common.cuh — shared definitions: table-blob layout, launch bookkeeping, FP64 device math.
// All arithmetic is IEEE FP64; the only integers are indices, counters and the put/call flag.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/faop.h"

#define FAOP_HD __host__ __device__ __forceinline__
#define FAOP_D __device__ __forceinline__

namespace faop {

constexpr double kInvSqrt2Pi = 0.3989422804014326779399460599343819;
constexpr double kInvSqrt2 = 0.7071067811865475244008443621048490;
constexpr int kSMs = 148;  // B200

// ---------------------------------------------------------------- table blob layout (in doubles)
// [wl hl rhl sgl : 4 x LP][wm hm rhm sgm : 4 x MP][c : n+1][costab : 2n][M : n*(n+1)*LP, optional]
//   (yl, wl) = leggauss(l) padded to LP = 32*ceil(l/32) with (y = 0, w = 0)
//   hl = (1+y)/2, rhl = 1/hl, sgl = sqrt(1 - hl^2)       (tau - u = tau hl^2, u = tau (1 - hl^2))
//   likewise for the m-point price rule
//   c_i = cos(i pi/n); costab[j] = cos(j pi/n), j < 2n   (cos(i k pi/n) = costab[(i k) mod 2n])
//   M[i][j][k]: H(u_ik) = sum_j M[i][j][k] H_j, the option-independent interpolation matrix
struct TableLayout {
    int n, l, m, LP, MP;
    int off_wl, off_hl, off_rhl, off_sgl, off_wm, off_hm, off_rhm, off_sgm, off_c, off_cos, off_M;
    int has_matrix;
    int small;  // doubles before the matrix (what every kernel stages in shared memory)
    int total;  // doubles
};

// matrix kept only when the whole per-CTA table set stays well inside shared memory
constexpr int kMaxMatrixDoubles = 12 * 13 * 32;  // (n = 12, l <= 32): 39 KB

FAOP_HD TableLayout table_layout(int n, int l, int m) {
    TableLayout t;
    t.n = n; t.l = l; t.m = m;
    t.LP = 32 * ((l + 31) / 32);
    t.MP = 32 * ((m + 31) / 32);
    t.off_wl = 0;
    t.off_hl = t.off_wl + t.LP;
    t.off_rhl = t.off_hl + t.LP;
    t.off_sgl = t.off_rhl + t.LP;
    t.off_wm = t.off_sgl + t.LP;
    t.off_hm = t.off_wm + t.MP;
    t.off_rhm = t.off_hm + t.MP;
    t.off_sgm = t.off_rhm + t.MP;
    t.off_c = t.off_sgm + t.MP;
    t.off_cos = t.off_c + (n + 1);
    int end = t.off_cos + 2 * n;
    end = (end + 1) & ~1;
    t.small = end;
    t.off_M = end;
    int msz = n * (n + 1) * t.LP;
    t.has_matrix = (msz <= kMaxMatrixDoubles) ? 1 : 0;
    t.total = end + (t.has_matrix ? msz : 0);
    return t;
}

// ---------------------------------------------------------------- normal distribution
FAOP_HD double norm_pdf(double x) { return exp(-0.5 * x * x) * kInvSqrt2Pi; }
FAOP_HD double norm_cdf(double x) { return 0.5 * erfc(-x * kInvSqrt2); }

// FastAmericanOptionSolverBase.py:246-256 / QDplusAmericanOptionSolver.py:78-88
FAOP_HD double b_at_zero(double r, double q, double K, bool is_put) {
    if (is_put) return (r >= q) ? K : r / q * K;
    return (r <= q) ? K : r / q * K;
}

// ---------------------------------------------------------------- European closed forms
// EuropeanOptionSolver.py:35-40
FAOP_HD double euro_d1(double tau, double s0, double r, double q, double vol, double K) {
    double svt = vol * sqrt(tau);
    return log(s0 * exp((r - q) * tau) / K) / svt + 0.5 * svt;
}
// EuropeanOptionSolver.py:7-22 (tau == 0 -> intrinsic)
FAOP_HD double euro_value(double tau, double s0, double r, double q, double vol, double K, bool is_put) {
    if (tau == 0) return is_put ? fmax(0.0, K - s0) : fmax(0.0, s0 - K);
    double d1 = euro_d1(tau, s0, r, q, vol, K);
    double d2 = d1 - vol * sqrt(tau);
    double dr = exp(-r * tau), dq = exp(-q * tau);
    if (is_put) return K * dr * norm_cdf(-d2) - s0 * dq * norm_cdf(-d1);
    return s0 * dq * norm_cdf(d1) - K * dr * norm_cdf(d2);
}
// EuropeanOptionSolver.py:25-32 (put theta; r and tau floored at 1e-10)
FAOP_HD double euro_theta(double tau, double s0, double r, double q, double vol, double K) {
    r = fmax(r, 1e-10);
    tau = fmax(tau, 1e-10);
    double d1 = euro_d1(tau, s0, r, q, vol, K);
    double d2 = d1 - vol * sqrt(tau);
    return r * K * exp(-r * tau) * norm_cdf(-d2) - q * s0 * exp(-q * tau) * norm_cdf(-d1) -
           vol * s0 * exp(-q * tau) * norm_pdf(d1) / (2 * sqrt(tau));
}

// ---------------------------------------------------------------- QD seed
// Per-(option, tau) invariants of the QD residual, QDplusAmericanOptionSolver.py:127-134, 211-218
struct QdConst {
    double qQD, svt, drift, dr, dq;  // drift = (r-q) tau, dr = e^{-r tau}, dq = e^{-q tau}
};
FAOP_HD QdConst qd_const(double tau, double r, double q, double sg, bool is_put) {
    QdConst c;
    double s2 = sg * sg;
    double Nn = 2 * (r - q) / s2, M = 2 * r / s2, h = 1 - exp(-r * tau);
    double disc = sqrt((Nn - 1) * (Nn - 1) + 4 * M / h);
    c.qQD = is_put ? -0.5 * (Nn - 1) - 0.5 * disc : -0.5 * (Nn - 1) + 0.5 * disc;
    c.svt = sg * sqrt(tau);
    c.drift = (r - q) * tau;
    c.dr = exp(-r * tau);
    c.dq = exp(-q * tau);
    return c;
}
// Residual F(S) of QDplusAmericanOptionSolver.py:110-125 and dF/dS (:207-208; call derived likewise)
FAOP_HD void qd_resid(const QdConst& c, double S, double K, bool is_put, double* F, double* dF) {
    double d1 = (log(S / K) + c.drift) / c.svt + 0.5 * c.svt;
    double d2 = d1 - c.svt;
    if (is_put) {
        double P1 = norm_cdf(-d1), P2 = norm_cdf(-d2);
        double p = K * c.dr * P2 - S * c.dq * P1;
        double g = 1 - c.dq * P1;
        *F = g * S + c.qQD * (K - S - p);
        *dF = (1 - c.qQD) * g + c.dq * norm_pdf(d1) / c.svt;
    } else {
        double P1 = norm_cdf(d1), P2 = norm_cdf(d2);
        double cv = S * c.dq * P1 - K * c.dr * P2;
        double g = 1 - c.dq * P1;
        *F = g * S - c.qQD * (S - K - cv);
        *dF = (1 - c.qQD) * g - c.dq * norm_pdf(d1) / c.svt;
    }
}
// QDplus.compute_exercise_boundary, QDplusAmericanOptionSolver.py:69-76: root from x0 = K.
// Safeguarded Newton (positivity, at most doubling per step) instead of MINPACK hybr (xtol 1.49e-8).
FAOP_HD double qd_boundary(double tau, double r, double q, double sg, double K, bool is_put) {
    if (tau == 0) return b_at_zero(r, q, K, is_put);
    QdConst c = qd_const(tau, r, q, sg, is_put);
    double S = K, prev = 1.79e308;
    for (int it = 0; it < 100; ++it) {
        double F, dF;
        qd_resid(c, S, K, is_put, &F, &dF);
        double Sn = S - F / dF;
        if (!(Sn > 0) || !(Sn < 1.79e308)) Sn = 0.5 * S;
        if (Sn > 2.0 * S) Sn = 2.0 * S;
        double d = fabs(Sn - S);
        // converged (the next Newton step would be ~d^2), or stalled at the rounding floor of a flat residual
        bool done = d <= 1e-10 * fabs(Sn) || (d >= 0.5 * prev && prev <= 1e-7 * fabs(Sn));
        prev = d;
        S = Sn;
        if (done) break;
    }
    return S;
}

// ---------------------------------------------------------------- Chebyshev (ChebyshevInterpolation.py)
// Clenshaw, :31-41
template <typename Ptr>
FAOP_HD double clenshaw(int n, Ptr a, double z) {
    double b0 = a[n] * 0.5, b1 = 0, b2 = 0;
    double z2 = 2 * z;
    for (int k = n - 1; k >= 0; --k) {
        b2 = b1;
        b1 = b0;
        b0 = a[k] + z2 * b1 - b2;
    }
    return 0.5 * (b0 - b2);
}

// ---------------------------------------------------------------- launch bookkeeping
extern unsigned long long g_launches;  // defined in faop_api.cu
inline void count_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }

#define FAOP_CUDA_OK(expr)                        \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return (int)_e;    \
    } while (0)

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

}  // namespace faop
