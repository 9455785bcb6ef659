// faop_fastmath.cuh — FP64 building blocks tuned for the DFMA pipe of sm_100a.
// The hot loop is bound by FP64 instruction issue, so every function here is written for the smallest DFMA count
// at ~1e-14..1e-15 relative accuracy (the parity bars are 1e-10 on the boundary and 1e-9 on the price):
//   fast_exp      14 FP64 ops (libdevice exp: ~25)       x <= 709; < -708 flushes to 0; NaN propagates
//   fast_rcp       3 FP64 ops + MUFU.RCP64H              normal positive arguments
//   mills_half    25 FP64 ops (libdevice erfcx: ~50)     G(y) = Phi(-y) / exp(-y^2/2), y >= 0
//   fast_sqrt_pos  7 FP64 ops + MUFU.RSQ64H              x > 0 (callers clamp)
// Coefficients: Chebyshev fits computed with mpmath at 50 digits (degree 10 for exp on |r| <= ln2/2, max rel 4.4e-16;
// degree 18 for (y+5) G(y) in t = 1 - 10/(y+5), max rel 1.9e-14 on y in [0, 45]).
#pragma once
#include "faop_common.cuh"

namespace faop {

// Polynomial coefficients live in constant memory so that every DFMA takes its coefficient as a constant-bank operand
// (c[bank][offset]) instead of two 32-bit immediate moves into registers per coefficient.
// kExpC[k]: exp(r) = sum_k kExpC[k] r^k on |r| <= ln2/2.
// kMillsC[k]: (x + kMillsShift) * erfcx(x)/2 = sum_k kMillsC[k] t^k with t = 1 - 2 kMillsShift/(x + kMillsShift)
//             (the degree-18 fit of (y+5) Phi(-y)/exp(-y^2/2) in y = sqrt2 x, divided by sqrt2).
static __constant__ double kExpC[11] = {1.0, 1.00000000000000666e+00, 5.00000000000001887e-01, 1.66666666665542140e-01,
                                 4.16666666664877935e-02, 8.33333338573419465e-03, 1.38888889523942471e-03,
                                 1.98411701867620645e-04, 2.48014853976660804e-05, 2.76402154411630899e-06,
                                 2.76326708447496886e-07};
// Reduced-degree fits for the quadrature points, where the terms enter N and D with weight ~ r tau, q tau <= 0.3
// (max rel 1.1e-12 for exp, 3.0e-11 for erfcx/2; the node-level closure keeps the full-accuracy tables):
static __constant__ double kExp8[9] = {1.00000000000001354e+00, 9.99999999979748533e-01, 4.99999999994376054e-01,
                                       1.66666668913440119e-01, 4.16666670410351936e-02, 8.33326603932144604e-03,
                                       1.38888016753182446e-03, 1.99158990839349160e-04, 2.48844948414224168e-05};
// (x + 3) erfcx(x)/2 = sum_k kErfcx14[k] t^k, t = 1 - 6/(x + 3)
static __constant__ double kErfcx14[15] = {
    5.37003453547791554e-01, -4.41697226390743936e-01, 2.95114178084050105e-01, -1.55233638330990936e-01,
    5.97638905924903277e-02, -1.34135356524479627e-02, -6.06289213047839443e-04, 1.51662855789247821e-03,
    -2.73829690347347061e-04, -1.37352260196071919e-04, 5.30630226822841820e-05, 1.37163947469176610e-05,
    -7.83813889147139852e-06, -1.19644291282268299e-06, 7.67685587370318563e-07};
// the same form at degree 16 (max rel 1.4e-12): the long-maturity configurations (C3: r tau up to 1.5) run on this one
static __constant__ double kErfcx16[17] = {
    5.37003453544169895e-01, -4.41697226572994928e-01, 2.95114178547531192e-01, -1.55233631526699944e-01,
    5.97638808610307173e-02, -1.34136091408635482e-02, -6.06211371911752317e-04, 1.51697851049094291e-03,
    -2.74135460896897365e-04, -1.38207715201892501e-04, 5.37152728700476733e-05, 1.48362796563143242e-05,
    -8.60891969407260209e-06, -1.94304210166222180e-06, 1.24198041290133434e-06, 1.99095216049436270e-07,
    -1.18566812020442530e-07};
#define FAOP_S2 0.70710678118654752440
static __constant__ double kMillsC[19] = {
    7.69193049750053093e-01 * FAOP_S2, -6.65382502890239702e-01 * FAOP_S2, 4.95305615971961821e-01 * FAOP_S2,
    -3.13533156660161738e-01 * FAOP_S2, 1.65020366104179528e-01 * FAOP_S2, -6.91186389062564549e-02 * FAOP_S2,
    2.07950676583355080e-02 * FAOP_S2, -2.99337512829167757e-03 * FAOP_S2, -7.97875070561768315e-04 * FAOP_S2,
    5.44891678680881932e-04 * FAOP_S2, -5.93541001006240258e-05 * FAOP_S2, -4.97429386785415854e-05 * FAOP_S2,
    1.66306732961883153e-05 * FAOP_S2, 3.94749285187075274e-06 * FAOP_S2, -2.69546304181993080e-06 * FAOP_S2,
    -3.01053062131189312e-07 * FAOP_S2, 3.67057836016596013e-07 * FAOP_S2, 1.86058824387743594e-08 * FAOP_S2,
    -3.23812507239800613e-08 * FAOP_S2};
constexpr double kMillsShift = 5.0 * FAOP_S2;

template <int NCH, int TIER> FAOP_D void half_erfcx_n(const double (&x)[NCH], double (&out)[NCH]);

FAOP_D double fast_rcp(double a) {
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a));  // ~20 good bits
    double e = fma(-a, x0, 1.0);
    double e2 = fma(e, e, e);          // e + e^2: third-order update, error ~ e^3
    return fma(x0, e2, x0);
}

FAOP_D double fast_exp(double x) {
    const double kL2E = 1.4426950408889634074, kMagic = 6755399441055744.0;  // 1.5 * 2^52
    double t = fma(x, kL2E, kMagic);
    int n = __double2loint(t);
    double kf = t - kMagic;
    double r = fma(kf, -6.93147180559945286e-01, x);
    r = fma(kf, -2.31904681384629956e-17, r);
    double p = 2.76326708447496886e-07;
    p = fma(p, r, 2.76402154411630899e-06);
    p = fma(p, r, 2.48014853976660804e-05);
    p = fma(p, r, 1.98411701867620645e-04);
    p = fma(p, r, 1.38888889523942471e-03);
    p = fma(p, r, 8.33333338573419465e-03);
    p = fma(p, r, 4.16666666664877935e-02);
    p = fma(p, r, 1.66666666665542140e-01);
    p = fma(p, r, 5.00000000000001887e-01);
    p = fma(p, r, 1.00000000000000666e+00);
    p = fma(p, r, 1.0);
    double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return (x < -708.0) ? 0.0 : res;
}

// G(y) = Phi(-y) / exp(-y^2/2) = erfcx(y / sqrt2) / 2 for y >= 0 (NaN propagates; y = +inf gives 0).
FAOP_D double mills_half(double y) {
    const double c = 5.0;
    double w = fast_rcp(y + c);
    if (__double2hiint(y) == 0x7ff00000) w = 0.0;   // y = +inf (integer test: no FP64-pipe slot); NaN falls through
    double t = fma(-2.0 * c, w, 1.0);
    double q = -3.23812507239800613e-08;
    q = fma(q, t, 1.86058824387743594e-08);
    q = fma(q, t, 3.67057836016596013e-07);
    q = fma(q, t, -3.01053062131189312e-07);
    q = fma(q, t, -2.69546304181993080e-06);
    q = fma(q, t, 3.94749285187075274e-06);
    q = fma(q, t, 1.66306732961883153e-05);
    q = fma(q, t, -4.97429386785415854e-05);
    q = fma(q, t, -5.93541001006240258e-05);
    q = fma(q, t, 5.44891678680881932e-04);
    q = fma(q, t, -7.97875070561768315e-04);
    q = fma(q, t, -2.99337512829167757e-03);
    q = fma(q, t, 2.07950676583355080e-02);
    q = fma(q, t, -6.91186389062564549e-02);
    q = fma(q, t, 1.65020366104179528e-01);
    q = fma(q, t, -3.13533156660161738e-01);
    q = fma(q, t, 4.95305615971961821e-01);
    q = fma(q, t, -6.65382502890239702e-01);
    q = fma(q, t, 7.69193049750053093e-01);
    return q * w;
}

// ln(a) for positive normal a: a = m 2^e with m in [sqrt(1/2), sqrt 2), ln m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716,
// odd series through s^21 (truncation 2.4e-17 relative).  Anything else (zero, negative, denormal, inf, NaN) takes the
// libdevice path so that the reference's NaN / -inf behaviour (numpy log of a non-positive boundary) is kept.
static __constant__ double kLogC[10] = {2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0, 2.0 / 17.0,
                                        2.0 / 19.0, 2.0 / 21.0};
FAOP_D double fast_log(double a) {
    const int hi = __double2hiint(a);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(a);
    int e = (hi >> 20) - 1023;
    int mhi = (hi & 0x000fffff) | 0x3ff00000;
    if (mhi >= 0x3ff6a09f) { mhi -= 0x00100000; e += 1; }     // m > ~sqrt2 -> m/2
    const double m = __hiloint2double(mhi, __double2loint(a));
    const double s = (m - 1.0) * fast_rcp(m + 1.0);
    const double z = s * s;
    double p = kLogC[9];
#pragma unroll
    for (int k = 8; k >= 0; --k) p = fma(p, z, kLogC[k]);
    const double lm = fma(s * z, p, 2.0 * s);
    const double ef = (double)e;
    return fma(ef, 6.93147180559945286e-01, fma(ef, 2.31904681384629956e-17, lm));
}

// e^{a} Phi(x) from E = exp(a - x^2/2) (and ea = e^{a}, used only for x > 0)
FAOP_D double fast_scaled_cdf(double x, double E, double ea) {
    double ec = E * mills_half(fabs(x));
    return (x <= 0.0) ? ec : ea - ec;
}

// Phi(sqrt2 x) from E = exp(-x^2): the node-level closure (one value per lane, not on the hot path).
FAOP_D double fast_half_erfcx_cdf(double x, double E) {
    double xa[1] = {fabs(x)}, G[1];
    half_erfcx_n<1, false>(xa, G);
    double ec = E * G[0];
    return (x <= 0.0) ? ec : 1.0 - ec;
}

// sqrt(x) for x > 0 and normal; callers clamp non-positive inputs.  NaN propagates.
FAOP_D double fast_sqrt_pos(double x) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));  // ~20 good bits of 1/sqrt(x)
    double g = x * y0, h = 0.5 * y0;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    return fma(g, r, g);
}

// ---- chain-interleaved versions: NCH independent evaluations advance in lock step so that the FP64 pipe always has
// NCH independent DFMAs in flight per warp (a lone Horner chain is latency-bound).

// exp(x[c]) for NCH arguments.  Arguments below -700 are clamped to ~-700 on the high word with one integer min (the
// result, < 1e-304, is zero for every sum it enters); NaN propagates.  LOW selects the degree-8 table.
template <int NCH, bool LOW = false>
FAOP_D void fast_exp_n(const double (&xin)[NCH], double (&out)[NCH]) {
    const double kL2E = 1.4426950408889634074, kMagic = 6755399441055744.0;
    double x[NCH], t[NCH], r[NCH], p[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        unsigned hi = min((unsigned)__double2hiint(xin[c]), 0xC085E000u);   // 0xC085E000 = high word of -700.0
        x[c] = __hiloint2double((int)hi, __double2loint(xin[c]));
        t[c] = fma(x[c], kL2E, kMagic);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        double kf = t[c] - kMagic;
        // one-step reduction: the rounding of ln2 costs |kf| * 2.3e-17 relative, < 1.4e-15 for every argument above -40
        r[c] = fma(kf, -6.93147180559945309e-01, x[c]);
        p[c] = LOW ? kExp8[8] : kExpC[10];
    }
#pragma unroll
    for (int k = (LOW ? 7 : 9); k >= 0; --k) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) p[c] = fma(p[c], r[c], LOW ? kExp8[k] : kExpC[k]);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        out[c] = __hiloint2double(__double2hiint(p[c]) + (__double2loint(t[c]) << 20), __double2loint(p[c]));
}

// erfcx(x[c]) / 2 for NCH arguments x >= 0.  The argument is clamped to ~1e300 on the high word (integer min), so
// x = +inf gives ~0; a NaN argument always arrives with a NaN exp factor, which keeps the product NaN.
// TIER 0 (false): degree 18, 2e-14;  TIER 1 (true): degree 14, 3e-11;  TIER 2: degree 16, 1.4e-12.
template <int NCH, int TIER = 0>
FAOP_D void half_erfcx_n(const double (&xin)[NCH], double (&out)[NCH]) {
    constexpr int DEG = TIER == 0 ? 18 : (TIER == 1 ? 14 : 16);
    const double sh = TIER == 0 ? kMillsShift : 3.0;
    double w[NCH], t[NCH], q[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        int hi = min(__double2hiint(xin[c]), 0x7E37E43C);                    // high word of 1e300
        double x = __hiloint2double(hi, __double2loint(xin[c]));
        w[c] = fast_rcp(x + sh);
        t[c] = fma(-2.0 * sh, w[c], 1.0);
        q[c] = TIER == 0 ? kMillsC[DEG] : (TIER == 1 ? kErfcx14[DEG] : kErfcx16[DEG]);
    }
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) q[c] = fma(q[c], t[c], TIER == 0 ? kMillsC[k] : (TIER == 1 ? kErfcx14[k] : kErfcx16[k]));
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) out[c] = q[c] * w[c];
}

// sqrt(max(0, x[c])) for NCH arguments, branch-free; values below ~1e-289 (or negative) give 0, NaN propagates.
// One rsqrt seed (>= 20 bits) and one third-order step: g (1 + r + 3/2 r^2) with r = 1/2 - g h, error ~ r^3.
template <int NCH>
FAOP_D void sqrt_clamped_n(const double (&x)[NCH], double (&out)[NCH]) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x[c]));
        double g = x[c] * y0, h = 0.5 * y0;
        double r = fma(-g, h, 0.5);
        double r2 = fma(1.5 * r, r, r);
        g = fma(g, r2, g);
        // x < ~1e-289 or negative (signed test on the high word, integer ALU); NaN is not tiny and stays NaN
        out[c] = (__double2hiint(x[c]) < 0x03C00000) ? 0.0 : g;
    }
}

}  // namespace faop
