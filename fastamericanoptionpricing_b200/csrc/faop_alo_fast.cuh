// faop_alo_fast.cuh — pieces shared by the specialised boundary kernels (faop_alo_fast.cu, faop_alo_clen.cu).
#pragma once
#include "faop_alo.cuh"
#include "faop_fastmath.cuh"

namespace faop {

// Node constants the sweep reads per point, packed so that a lane fetches them with four 16-byte broadcast loads.
// The d+- arguments are carried as x = d/sqrt2 (exp(-d^2/2) = exp(-x^2), Phi(-d) = erfcx(x)/2 exp(-x^2)).
struct __align__(16) NodePack {
    double L;      // ln(B_i/X), refreshed every iteration
    double cinv2;  // 1/(sigma sqrt(2 tau_i))
    double apc2;   // ((r-q) + sigma^2/2) sqrt(tau_i/2)/sigma
    double ssu2;   // sigma sqrt(tau_i/2)
    double qtau;   // q tau_i
    double rtau;   // r tau_i
    double cis;    // 1/(sigma sqrt(2 pi tau_i))
    double pad;
};

FAOP_D NodePack make_node_pack(const Opt& o, double tau) {
    NodePack np;
    const double sqt = sqrt(tau);
    np.L = 0.0;
    np.cinv2 = kInvSqrt2 / (o.sg * sqt);
    np.apc2 = kInvSqrt2 * ((o.r - o.q) + 0.5 * o.sg * o.sg) * sqt / o.sg;
    np.ssu2 = kInvSqrt2 * o.sg * sqt;
    np.qtau = o.q * tau;
    np.rtau = o.r * tau;
    np.cis = kInvSqrt2Pi / (o.sg * sqt);
    np.pad = 0.0;
    return np;
}

// Reduce 8 per-lane values over the 32 lanes through a shared-memory transpose: every lane stores its 8 partial sums
// (row v, column lane), then lane l sums the 8 entries {l&3, (l&3)+4, ...} of row v = l>>2 and two xor-shuffles finish
// the row.  On return every lane of the 4-lane group l>>2 holds the total of value index l>>2.  Costs 8 STS + 8 LDS +
// 9 DADD + 2 shuffles per 4 nodes, about half the issue slots of the select-and-shuffle butterfly it replaces.
FAOP_D double transpose_reduce8(const double (&v)[8], double* red, int lane) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k * 36 + lane] = v[k];
    __syncwarp();
    const double* row = red + (lane >> 2) * 36 + (lane & 3);
    double s0 = row[0] + row[4], s1 = row[8] + row[12], s2 = row[16] + row[20], s3 = row[24] + row[28];
    double t = (s0 + s1) + (s2 + s3);
    t += __shfl_xor_sync(kFull, t, 1);
    t += __shfl_xor_sync(kFull, t, 2);
    __syncwarp();   // the buffer is rewritten by the next group
    return t;
}

// Price integral on the boundary in w.B (american_value_with_known_boundary, FastAmericanOptionSolverBase.py:92-103) with the
// full-accuracy fits of faop_fastmath.cuh (exp 4e-16, erfcx 2e-14 relative: <= 1e-12 on the price) instead of libdevice,
// two quadrature points per lane advanced in lock step.  Same arithmetic layout as price_integral (faop_alo.cuh); the
// European leg keeps libdevice because it is also returned on its own.
FAOP_D double price_integral_fast(const WarpSmem& w, const CtaTables& tb, const TableLayout& tl, const Opt& o, int lane,
                                  double* eur) {
    const double tau = o.T - o.t;
    const double e = euro_value(tau, o.S, o.r, o.q, o.sg, o.K, o.put);
    *eur = e;
    if (tau == 0) return e + 0.0;  // quadrature_sum returns 0 at tau == 0 (:382-383)
    const int n = tl.n;
    bool bad = false;
    for (int j = lane; j <= n; j += 32) {   // L_j = ln(B_j/X), H_j = L_j^2 and the DCT-I fit (fit_boundary)
        const double L = fast_log(w.B[j] / o.X);
        w.L[j] = L;
        w.H[j] = L * L;
        bad |= !(fabs(L) <= 1.79e308);
    }
    // A boundary node outside (0, inf) (or X = 0), a spot outside (0, inf), tau < 0: upstream takes the log / square root of it
    // and returns NaN (Base.py:184, :237, :307).  Decided here, on full-accuracy values: the reduced-cost exp / erfcx / sqrt
    // below clamp on the integer high word and are not required to carry a NaN of either sign through.
    const double LS = fast_log(o.S / o.X);
    bad |= !(fabs(LS) <= 1.79e308) || !(tau > 0);
    if (__any_sync(kFull, bad)) return nan("");
    __syncwarp();
    const int n2 = 2 * n;
    for (int k = lane; k <= n; k += 32) {
        double ans = 0.0;
        int idx = 0;
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
    const NodeC c = node_consts(o, tau);
    const double zf = sqrt(4 * tau / o.Tc);
    const double cinv2 = c.cinv * kInvSqrt2, apc2 = c.apc * kInvSqrt2, ssu2 = c.ssu * kInvSqrt2;
    const double rK = o.r * o.K, qS = o.q * o.S;
    double acc = 0.0;
    for (int k0 = lane; k0 < tl.MP; k0 += 64) {
        const int k1 = (k0 + 32 < tl.MP) ? k0 + 32 : k0;      // second slot (weight forced to 0 when it has no point of its own)
        const double h[2] = {tb.hm[k0], tb.hm[k1]};
        double Hu[2], root[2];
        {
            double b0[2], b1[2] = {0.0, 0.0}, b2[2] = {0.0, 0.0};
            const double z2[2] = {2.0 * fma(zf, tb.sgm[k0], -1.0), 2.0 * fma(zf, tb.sgm[k1], -1.0)};
            b0[0] = b0[1] = w.A[n] * 0.5;
            for (int k = n - 1; k >= 0; --k) {
                const double ak = w.A[k];
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    b2[s2] = b1[s2];
                    b1[s2] = b0[s2];
                    b0[s2] = fma(z2[s2], b1[s2], ak - b2[s2]);
                }
            }
            Hu[0] = 0.5 * (b0[0] - b2[0]);
            Hu[1] = 0.5 * (b0[1] - b2[1]);
        }
        sqrt_clamped_n<2>(Hu, root);
        double ex[4], xa[4], xs[4];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            const double lnz = fma(o.om, root[s2], LS);
            const double xp = fma(lnz, cinv2 * (s2 ? tb.rhm[k1] : tb.rhm[k0]), apc2 * h[s2]);   // d+ / sqrt2
            const double xm = fma(-ssu2, h[s2], xp);                                             // d- / sqrt2
            const double h2 = h[s2] * h[s2];
            ex[2 * s2] = fma(-xm, xm, -c.rtau * h2);      // -r (tau - u) - d-^2/2
            ex[2 * s2 + 1] = fma(-xp, xp, -c.qtau * h2);  // -q (tau - u) - d+^2/2
            xs[2 * s2] = -o.om * xm; xs[2 * s2 + 1] = -o.om * xp;        // Phi(-om d-), Phi(-om d+)
            xa[2 * s2] = fabs(xm); xa[2 * s2 + 1] = fabs(xp);
        }
        double E[4], G[4], ed[4];
        fast_exp_n<4, false>(ex, E);
        half_erfcx_n<4, false>(xa, G);
        {   // the plain discount factors, needed where Phi's argument is positive: e^{-r s} Phi(y) = e^{-r s} - E G
            double exd[4] = {-c.rtau * h[0] * h[0], -c.qtau * h[0] * h[0], -c.rtau * h[1] * h[1], -c.qtau * h[1] * h[1]};
            fast_exp_n<4, false>(exd, ed);
        }
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            const double cm = E[2 * s2] * G[2 * s2], cp = E[2 * s2 + 1] * G[2 * s2 + 1];
            const double Pm = (xs[2 * s2] <= 0.0) ? cm : ed[2 * s2] - cm;           // e^{-r s} Phi(-om d-)
            const double Pp = (xs[2 * s2 + 1] <= 0.0) ? cp : ed[2 * s2 + 1] - cp;   // e^{-q s} Phi(-om d+)
            const double wgt = (s2 == 1 && k1 == k0) ? 0.0 : (s2 ? tb.wm[k1] : tb.wm[k0]) * h[s2];
            acc = fma(wgt, o.om * (rK * Pm - qS * Pp), acc);                          // v_integrand_12 (:211-219)
        }
    }
    acc = warp_sum(acc);
    return e + acc * tau;
}

}  // namespace faop
