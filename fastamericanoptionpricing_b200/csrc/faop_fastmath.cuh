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

// e^{a} Phi(x) from E = exp(a - x^2/2) (and ea = e^{a}, used only for x > 0)
FAOP_D double fast_scaled_cdf(double x, double E, double ea) {
    double ec = E * mills_half(fabs(x));
    return (x <= 0.0) ? ec : ea - ec;
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

}  // namespace faop
