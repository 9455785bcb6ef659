/* CPU oracle (plain C): scalar restatement of the reference's spectral-collocation American pricer.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_build/liboracle.so by oracle/Makefile; loaded only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg.  The product
 * (fastamericanoptionpricing_b200/) never links or calls it.
 *
 * One option at a time, the way the reference's Python objects work, in the reference's operation
 * order; options are spread over pthreads only to use the host cores.  Every function cites the
 * reference file:line it follows (paths relative to antdvid/FastAmericanOptionPricing).
 * Pinned through tests/golden/ref_*.npz (outputs of the unmodified reference, oracle/gen_golden.py)
 * and against oracle/alo_oracle.py (the numpy restatement) in tests/test_oracle_golden.py.
 * The QD seed root follows the reference's root finder: hybr1() restates MINPACK hybrd for one unknown as
 * scipy.optimize.root(method="hybr", x0=K) drives it (QDplusAmericanOptionSolver.py:73), like alo_oracle.py's
 * hybr1 (which is bit-identical to scipy); safeguarded Newton is kept as seed mode 0.  Phi(x) = erfc(-x/sqrt2)/2
 * from libm stands in for scipy.special.ndtr (<= 1 ulp apart).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define INV_SQRT_2PI 0.3989422804014326779399460599343819
#define INV_SQRT_2   0.7071067811865475244008443621048490

static double Phi(double x) { return 0.5 * erfc(-x * INV_SQRT_2); }
/* Conditioning probe of the QD seed (oracle_set_seed_probe): Phi inside the seed residual is scaled by (1 + k ulp).  The
 * spread of the roots over k = 0, +-1, +-2 is how far the reference's own seed moves when its normal cdf changes in the
 * last bit — which is what any two correct implementations (scipy's ndtr, libm's erfc, libdevice's erfc) differ by.  For flat
 * residuals (tiny tau with X = rK/q << K: catastrophic cancellation in K - S - p, QDplusAmericanOptionSolver.py:124) that
 * spread reaches 1e-6; everywhere else it is a few ulp.  0 outside such probes. */
static int g_seed_probe_ulps = 0;
void oracle_set_seed_probe(int ulps) { g_seed_probe_ulps = ulps; }
static double PhiS(double x) { return Phi(x) * (1.0 + g_seed_probe_ulps * 2.220446049250313e-16); }
static double phi(double x) { return exp(-0.5 * x * x) * INV_SQRT_2PI; }

/* ---- EuropeanOptionSolver.py:7-40 ---- */
static double eu_d1(double tau, double s0, double r, double q, double vol, double K) {
    return log(s0 * exp((r - q) * tau) / K) / (vol * sqrt(tau)) + 0.5 * vol * sqrt(tau);
}
static double eu_put(double tau, double s0, double r, double q, double vol, double K) {
    if (tau == 0) return fmax(0.0, K - s0);
    double d1 = eu_d1(tau, s0, r, q, vol, K), d2 = d1 - vol * sqrt(tau);
    return K * exp(-r * tau) * Phi(-d2) - s0 * exp(-q * tau) * Phi(-d1);
}
static double eu_call(double tau, double s0, double r, double q, double vol, double K) {
    if (tau == 0) return fmax(0.0, s0 - K);
    double d1 = eu_d1(tau, s0, r, q, vol, K), d2 = d1 - vol * sqrt(tau);
    return s0 * exp(-q * tau) * Phi(d1) - K * exp(-r * tau) * Phi(d2);
}

/* ---- FastAmericanOptionSolverBase.py:246-256 ---- */
static double b_at_zero(double r, double q, double K, int is_put) {
    if (is_put) return (r >= q) ? K : r / q * K;
    return (r <= q) ? K : r / q * K;
}

/* ---- QDplusAmericanOptionSolver.py:110-134 residual, :207-208 dF/dS (call derived likewise) ---- */
static void qd_resid(double S, double tau, double r, double q, double sg, double K, int is_put, double* F, double* dF) {
    double Nn = 2 * (r - q) / (sg * sg), M = 2 * r / (sg * sg), h = 1 - exp(-r * tau);
    double disc = sqrt((Nn - 1) * (Nn - 1) + 4 * M / h);
    double qQD = is_put ? -0.5 * (Nn - 1) - 0.5 * disc : -0.5 * (Nn - 1) + 0.5 * disc;
    double svt = sg * sqrt(tau);
    double d1 = eu_d1(tau, S, r, q, sg, K), d2 = d1 - svt, eq = exp(-q * tau);
    if (is_put) {
        double p = K * exp(-r * tau) * PhiS(-d2) - S * eq * PhiS(-d1);
        *F = (1 - eq * PhiS(-d1)) * S + qQD * (K - S - p);
        *dF = (1 - qQD) * (1 - eq * PhiS(-d1)) + eq * phi(d1) / svt;
    } else {
        double c = S * eq * PhiS(d1) - K * exp(-r * tau) * PhiS(d2);
        *F = (1 - eq * PhiS(d1)) * S - qQD * (S - K - c);
        *dF = (1 - qQD) * (1 - eq * PhiS(d1)) - eq * phi(d1) / svt;
    }
}

/* QDplusAmericanOptionSolver.py:69-76, Newton from x0 = K (see header for the hybr deviation) */
static double qd_boundary(double tau, double r, double q, double sg, double K, int is_put) {
    if (tau == 0) return b_at_zero(r, q, K, is_put);
    double S = K, prev = 1.79e308;
    for (int it = 0; it < 100; ++it) {
        double F, dF;
        qd_resid(S, tau, r, q, sg, K, is_put, &F, &dF);
        double Sn = S - F / dF;
        if (!isfinite(Sn) || Sn <= 0) Sn = 0.5 * S;
        if (Sn > 2.0 * S) Sn = 2.0 * S;
        double d = fabs(Sn - S);
        /* converged (the next Newton step would be ~d^2), or stalled at the rounding floor of a flat residual */
        int done = d <= 1e-10 * fabs(Sn) || (d >= 0.5 * prev && prev <= 1e-7 * fabs(Sn));
        prev = d;
        S = Sn;
        if (done) break;
    }
    return S;
}

/* MINPACK hybrd for ONE unknown with scipy.optimize.root's defaults (xtol 1.49012e-8, maxfev 400, epsfcn = eps, mode 1,
 * factor 100): see hybr1 in alo_oracle.py for the derivation (r = -J, q = -1, enorm = fabs, no rotations).  Returns the
 * iterate hybrd leaves in x, whatever its info code (the reference ignores `success`). */
static double qd_F(double S, double tau, double r, double q, double sg, double K, int is_put) {
    double F, dF;
    qd_resid(S, tau, r, q, sg, K, is_put, &F, &dF);
    return F;
}
static double hybr1_qd(double tau, double r, double q, double sg, double K, int is_put, int* nfev_out) {
    const double epsmch = 2.220446049250313e-16, xtol = 1.49012e-08, factor = 100.0;
    const int maxfev = 400;
    double x = K;
    double fvec = qd_F(x, tau, r, q, sg, K, is_put);
    int nfev = 1, it = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    double fnorm = fabs(fvec), diag = 0, delta = 0, xnorm = 0;
    const double eps = sqrt(epsmch);
    for (;;) {
        int jeval = 1;
        double h = eps * fabs(x);
        if (h == 0.0) h = eps;
        double J = (qd_F(x + h, tau, r, q, sg, K, is_put) - fvec) / h;
        ++nfev;
        double acn = fabs(J);
        if (it == 1) {
            diag = acn != 0.0 ? acn : 1.0;
            xnorm = fabs(diag * x);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        double qtf, rr, Q;
        if (J != 0.0) { qtf = -fvec; rr = -J; Q = -1.0; } else { qtf = fvec; rr = -0.0; Q = 1.0; }
        diag = fmax(diag, acn);
        for (;;) {
            double temp = rr;
            if (temp == 0.0) { temp = epsmch * fabs(rr); if (temp == 0.0) temp = epsmch; }
            double xg = qtf / temp, p;
            double qnorm = fabs(diag * xg);
            if (qnorm <= delta) p = xg;
            else {
                double wa1 = (rr * qtf) / diag;
                double gnorm = fabs(wa1), sgnorm = 0.0, alpha = delta / qnorm;
                if (gnorm != 0.0) {
                    wa1 = (wa1 / gnorm) / diag;
                    double t2 = fabs(rr * wa1);
                    sgnorm = (gnorm / t2) / t2;
                    alpha = 0.0;
                    if (sgnorm < delta) {
                        double bnorm = fabs(qtf);
                        double tt = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
                        double dq = delta / qnorm, sd = sgnorm / delta;
                        tt = tt - dq * (sd * sd) + sqrt((tt - dq) * (tt - dq) + (1.0 - dq * dq) * (1.0 - sd * sd));
                        alpha = (dq * (1.0 - sd * sd)) / tt;
                    }
                }
                double tt = (1.0 - alpha) * fmin(sgnorm, delta);
                p = tt * wa1 + alpha * xg;
            }
            double w1 = -p, w2 = x + w1, w3 = diag * w1;
            double pnorm = fabs(w3);
            if (it == 1) delta = fmin(delta, pnorm);
            double w4 = qd_F(w2, tau, r, q, sg, K, is_put);
            ++nfev;
            double fnorm1 = fabs(w4), actred = -1.0;
            if (fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            w3 = qtf + rr * w1;
            temp = fabs(w3);
            double prered = 0.0;
            if (temp < fnorm) prered = 1.0 - (temp / fnorm) * (temp / fnorm);
            double ratio = 0.0;
            if (prered > 0.0) ratio = actred / prered;
            if (ratio < 0.1) { ncsuc = 0; ++ncfail; delta = 0.5 * delta; }
            else {
                ncfail = 0; ++ncsuc;
                if (ratio >= 0.5 || ncsuc > 1) delta = fmax(delta, pnorm / 0.5);
                if (fabs(ratio - 1.0) <= 0.1) delta = pnorm / 0.5;
            }
            if (ratio >= 1e-4) { x = w2; w2 = diag * x; fvec = w4; xnorm = fabs(w2); fnorm = fnorm1; ++it; }
            ++nslow1;
            if (actred >= 0.001) nslow1 = 0;
            if (jeval) ++nslow2;
            if (actred >= 0.1) nslow2 = 0;
            int info = 0;
            if (delta <= xtol * xnorm || fnorm == 0.0) info = 1;
            if (!info) {
                if (nfev >= maxfev) info = 2;
                if (0.1 * fmax(0.1 * delta, pnorm) <= epsmch * xnorm) info = 3;
                if (nslow2 == 5) info = 4;
                if (nslow1 == 10) info = 5;
            }
            if (info) { if (nfev_out) *nfev_out = nfev; return x; }
            if (ncfail == 2) break;
            double sm = Q * w4;
            w2 = (sm - w3) / pnorm;
            w1 = diag * ((diag * w1) / pnorm);
            if (ratio >= 1e-4) qtf = sm;
            rr = rr + w2 * w1;
            jeval = 0;
        }
    }
}
static double qd_seed(int mode, double tau, double r, double q, double sg, double K, int is_put) {
    if (mode == 0) return qd_boundary(tau, r, q, sg, K, is_put);
    if (tau == 0) return b_at_zero(r, q, K, is_put);
    return hybr1_qd(tau, r, q, sg, K, is_put, NULL);
}

/* ---- FastAmericanOptionSolverBase.py:306-310 ---- */
static void dpm(double s, double z, double r, double q, double sg, double* dm, double* dp) {
    double lz = log(z);
    *dm = (lz + (r - q) * s - 0.5 * sg * sg * s) / (sg * sqrt(s));
    *dp = (lz + (r - q) * s + 0.5 * sg * sg * s) / (sg * sqrt(s));
}
/* CDF_* with the s == 0 limits (:312-346), PDF_* (:348-370) */
static double cdf_s(int sign, double d, double s, double z) {
    if (s == 0) return (sign > 0) ? (z > 1 ? 1.0 : 0.0) : (z > 1 ? 0.0 : 1.0);
    return Phi(sign * d);
}
static double pdf_s(double d, double s) { return s == 0 ? 0.0 : phi(d); }

typedef struct {
    int solver, deriv, n, l, m, max_iters;
    double tol, eta;
    const double *yl, *wl, *ym, *wm;
    const double* costab; /* (n+1)x(n+1): cos(i k pi / n) */
    int seed;             /* 1 = hybr as upstream (QDplusAmericanOptionSolver.py:73), 0 = safeguarded Newton */
} cfg_t;

/* ChebyshevInterpolation.py:43-51 */
static void cheb_fit(const cfg_t* c, const double* y, double* a) {
    int n = c->n;
    for (int k = 0; k <= n; ++k) {
        double ans = 0;
        for (int i = 0; i <= n; ++i) {
            double term = y[i] * c->costab[i * (n + 1) + k];
            if (i == 0 || i == n) term *= 0.5;
            ans += term;
        }
        a[k] = ans * (2.0 / n);
    }
}
/* ChebyshevInterpolation.py:31-41 */
static double clenshaw(int n, const double* a, double z) {
    double b0 = a[n] * 0.5, b1 = 0, b2 = 0;
    for (int k = n - 1; k >= 0; --k) {
        b2 = b1; b1 = b0;
        b0 = a[k] + 2 * z * b1 - b2;
    }
    return 0.5 * (b0 - b2);
}
/* compute_integration_terms, FastAmericanOptionSolverBase.py:167-195, one quadrature point */
static double boundary_at(const cfg_t* c, const double* a, double u, double T, double X, int is_put) {
    double z = sqrt(4 * (u - 0.0) / (T - 0.0)) - 1;
    double Hu = clenshaw(c->n, a, z);
    double root = sqrt(Hu < 0.0 ? 0.0 : Hu); /* np.maximum propagates NaN, fmax would not */
    return (is_put ? exp(-root) : exp(root)) * X;
}

typedef struct { double r, q, sg, K, T, X; int is_put; } opt_t;

/* N, D, N', D' at one active node: FastAmericanOptionSolverA.py:5-83 / FastAmericanOptionSolverB.py:5-75 */
static void nd_node(const cfg_t* c, const opt_t* o, const double* a, double tau, double Bi,
                    double* Nv, double* Dv, double* Np, double* Dp) {
    double r = o->r, q = o->q, sg = o->sg, K = o->K;
    int put = o->is_put, l = c->l;
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    for (int k = 0; k < l; ++k) {
        double y = c->yl[k], w = c->wl[k];
        double u = tau - tau * ((1 + y) * (1 + y)) / 4.0;
        double Bu = boundary_at(c, a, u, o->T, o->X, put);
        double s = tau - u, z = Bi / Bu, dm, dp;
        dpm(s, z, r, q, sg, &dm, &dp);
        double jac = 0.5 * (tau - 0) * (1 + y);
        double ssu = sg * sqrt(s), eq = exp(q * u), er = exp(r * u);
        if (c->solver == 0) {
            double cdfp = put ? cdf_s(+1, dp, s, z) : -cdf_s(-1, dp, s, z);
            s1 += (eq * cdfp + eq / ssu * pdf_s(dp, s)) * w * jac;       /* K12, A.py:24-30 */
            s2 += (er / ssu * pdf_s(dm, s)) * w * jac;                    /* K3,  A.py:37-44 */
            if (c->deriv) {
                s3 += (er * dm / (Bi * sg * sg * s) * pdf_s(dm, s)) * w * jac;            /* A.py:55-63 */
                s4 += (eq * pdf_s(dp, s) / (ssu * Bi) * (1 - dp / ssu)) * w * jac;        /* A.py:76-83 */
            }
        } else {
            int sgn = put ? +1 : -1;
            s1 += (eq * cdf_s(sgn, dp, s, z)) * w * jac;                  /* D_integrand, B.py:33-39 */
            s2 += (er * cdf_s(sgn, dm, s, z)) * w * jac;                  /* N_integrand, B.py:25-31 */
            if (c->deriv) {
                s3 += (er * pdf_s(dm, s) / (ssu * Bi)) * w * jac;         /* Nprime_integrand, B.py:59-66 */
                s4 += (eq * pdf_s(dp, s) / (ssu * Bi)) * w * jac;         /* Dprime_integrand (put), B.py:68-72 */
            }
        }
    }
    double dmK, dpK, svt = sg * sqrt(tau), zK = Bi / K;
    dpm(tau, zK, r, q, sg, &dmK, &dpK);
    *Np = *Dp = 0;
    if (c->solver == 0) {
        *Nv = pdf_s(dmK, tau) / svt + r * s2;
        *Dv = pdf_s(dpK, tau) / svt + (put ? cdf_s(+1, dpK, tau, zK) : -cdf_s(-1, dpK, tau, zK)) + q * s1;
        if (c->deriv) {
            *Np = -dmK * pdf_s(dmK, tau) / (Bi * sg * sg * tau) - r * s3;
            *Dp = pdf_s(dpK, tau) / (svt * Bi) * (1 - dpK / svt) + q * s4;
        }
    } else {
        int sgn = put ? +1 : -1;
        *Nv = cdf_s(sgn, dmK, tau, zK) + r * s2;
        *Dv = cdf_s(sgn, dpK, tau, zK) + q * s1;
        if (c->deriv) {
            if (put) {
                *Np = pdf_s(dmK, tau) / (svt * Bi) + r * s3;
                *Dp = pdf_s(dpK, tau) / (svt * Bi) + q * s4;
            } else { /* quirk kept: the call's D' integrates Nprime_integrand, B.py:57 */
                *Np = -pdf_s(dmK, tau) / (svt * Bi) - r * s3;
                *Dp = -pdf_s(dpK, tau) / (svt * Bi) - q * s3;
            }
        }
    }
}

/* solve(t, s0), FastAmericanOptionSolverBase.py:49-165 and :92-103 */
/* Tc = the `tau_max` attribute, the end of the collocation domain (Base.py:28, consumed at :185 and :223); = T by default */
static void solve_one(const cfg_t* c, double r, double q, double sg, double K, double T, double Tc, double t, double S, int is_put,
                      const double* B0_in, double* price, double* european, double* Bout, double* B0out,
                      int32_t* iters, double* err) {
    int n = c->n;
    if (q == 0 && !is_put) { /* European shortcut, :50-53 */
        double e = eu_call(T - t, S, r, q, sg, K);
        *price = e; *european = e; *iters = 0; *err = 1e6;
        for (int i = 0; i <= n; ++i) { if (Bout) Bout[i] = NAN; if (B0out) B0out[i] = NAN; }
        return;
    }
    opt_t o = {r, q, sg, K, Tc, b_at_zero(r, q, K, is_put), is_put};
    double tau[n + 1], B[n + 1], Bn[n + 1], H[n + 1], a[n + 1];
    for (int i = 0; i <= n; ++i) {
        double ci = cos(i * M_PI / n);                    /* Cheb.py:15-17 */
        tau[i] = (ci + 1) * (ci + 1) * (Tc - 0.0) / 4 + 0.0; /* Base.py:239-240 */
    }
    for (int i = 0; i <= n; ++i) B[i] = B0_in ? B0_in[i] : qd_seed(c->seed, tau[i], r, q, sg, K, is_put);
    if (B0out) memcpy(B0out, B, sizeof(double) * (n + 1));
    int count = 0;
    double e = 1;
    while (e > c->tol && count < c->max_iters) {
        ++count;
        for (int i = 0; i <= n; ++i) { double lg = log(B[i] / o.X); H[i] = lg * lg; }
        cheb_fit(c, H, a);
        double ss = 0;
        for (int i = 0; i <= n; ++i) {
            if (tau[i] == 0) { Bn[i] = o.X; }
            else {
                double Nv, Dv, Np, Dp;
                nd_node(c, &o, a, tau[i], B[i], &Nv, &Dv, &Np, &Dp);
                double disc = K * exp(-tau[i] * (r - q));
                double f = disc * Nv / Dv, fp = 0.0;
                if (c->deriv) fp = disc * (Np / Dv - Dp * Nv / (Dv * Dv));
                Bn[i] = B[i] + c->eta * (B[i] - f) / (fp - 1);
            }
            double d = fabs(B[i] - Bn[i]);
            ss += d * d;
        }
        e = sqrt(ss);
        memcpy(B, Bn, sizeof(double) * (n + 1));
    }
    *iters = count; *err = e;
    if (Bout) memcpy(Bout, B, sizeof(double) * (n + 1));
    /* price integral */
    double tt = T - t;
    double v = is_put ? eu_put(tt, S, r, q, sg, K) : eu_call(tt, S, r, q, sg, K);
    *european = v;
    double v12 = 0;
    if (tt != 0) {
        for (int i = 0; i <= n; ++i) { double lg = log(B[i] / o.X); H[i] = lg * lg; }
        cheb_fit(c, H, a);
        for (int k = 0; k < c->m; ++k) {
            double y = c->ym[k], w = c->wm[k];
            double u = tt - tt * ((1 + y) * (1 + y)) / 4.0;
            double Bu = boundary_at(c, a, u, Tc, o.X, is_put);
            double s = tt - u, z = S / Bu, dm, dp;
            dpm(s, z, r, q, sg, &dm, &dp);
            double g;
            if (is_put) g = r * K * exp(-r * s) * cdf_s(-1, dm, s, z) - q * S * exp(-q * s) * cdf_s(-1, dp, s, z);
            else        g = q * S * exp(-q * s) * cdf_s(+1, dp, s, z) - r * K * exp(-r * s) * cdf_s(+1, dm, s, z);
            v12 += g * w * (0.5 * (tt - 0) * (1 + y));
        }
    }
    *price = v + v12;
}

/* ------------------------------------------------------------------ batch drivers (pthreads) */
typedef struct {
    cfg_t c; int64_t N;
    const double *r, *q, *sg, *K, *T, *t, *S, *tau_max; const uint8_t* is_put; const double* B0_in;
    double *price, *european, *B, *B0out; int32_t* iters; double* err;
    int64_t next; pthread_mutex_t mu;
} job_t;

static void* worker(void* p) {
    job_t* j = (job_t*)p;
    int nn = j->c.n + 1;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t b = j->next; j->next += 64;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->N) break;
        int64_t e = b + 64 < j->N ? b + 64 : j->N;
        for (int64_t i = b; i < e; ++i)
            solve_one(&j->c, j->r[i], j->q[i], j->sg[i], j->K[i], j->T[i], j->tau_max ? j->tau_max[i] : j->T[i],
                      j->t ? j->t[i] : 0.0, j->S[i], j->is_put[i],
                      j->B0_in ? j->B0_in + i * nn : NULL, &j->price[i], &j->european[i],
                      j->B ? j->B + i * nn : NULL, j->B0out ? j->B0out + i * nn : NULL, &j->iters[i], &j->err[i]);
    }
    return NULL;
}

int oracle_alo_price_batch(int solver, int use_derivative, int n, int l, int m, int max_iters, double tol, double eta,
                           int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                           const double* T, const double* t, const double* S, const uint8_t* is_put,
                           const double* y_l, const double* w_l, const double* y_m, const double* w_m,
                           const double* B0_in, double* price, double* european, double* B, double* B0_out,
                           int32_t* iters, double* err, int nthreads, int seed_mode, const double* tau_max /* nullable: T */) {
    job_t j;
    memset(&j, 0, sizeof j);
    double* ct = (double*)malloc(sizeof(double) * (n + 1) * (n + 1));
    for (int i = 0; i <= n; ++i) for (int k = 0; k <= n; ++k) ct[i * (n + 1) + k] = cos(i * k * M_PI / n);
    cfg_t c = {solver, use_derivative, n, l, m, max_iters, tol, eta, y_l, w_l, y_m, w_m, ct, seed_mode};
    j.tau_max = tau_max;
    j.c = c; j.N = N; j.r = r; j.q = q; j.sg = sigma; j.K = K; j.T = T; j.t = t; j.S = S; j.is_put = is_put;
    j.B0_in = B0_in; j.price = price; j.european = european; j.B = B; j.B0out = B0_out; j.iters = iters; j.err = err;
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    pthread_t th[nthreads];
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker, &j);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&j.mu);
    free(ct);
    return 0;
}

int oracle_qd_boundary_batch(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                             const double* tau, const uint8_t* is_put, double* B, int seed_mode, int32_t* nfev /* nullable */) {
    for (int64_t i = 0; i < N; ++i) {
        int nf = 0;
        if (seed_mode == 1 && tau[i] != 0) B[i] = hybr1_qd(tau[i], r[i], q[i], sigma[i], K[i], is_put[i], &nf);
        else B[i] = qd_seed(seed_mode, tau[i], r[i], q[i], sigma[i], K[i], is_put[i]);
        if (nfev) nfev[i] = nf;
    }
    return 0;
}

int oracle_european_batch(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                          const double* sigma, const double* K, const uint8_t* is_put, double* price) {
    for (int64_t i = 0; i < N; ++i)
        price[i] = is_put[i] ? eu_put(tau[i], S[i], r[i], q[i], sigma[i], K[i]) : eu_call(tau[i], S[i], r[i], q[i], sigma[i], K[i]);
    return 0;
}

/* ------------------------------------------------------------------ Crank-Nicolson, CrankNicolsonOptionSolver.py:27-107
 * mode 0: unprojected direct solve (the reference default, np.linalg.solve, :81-82) as an exact tridiagonal solve
 * mode 1: the reference's SOR-sweep-then-project loop (:84-107) operation for operation
 * mode 2: Brennan-Schwartz projected back-substitution (not in the reference; throughput mode of the CUDA path)
 * dt[] is the host-built schedule of `while t > 0: dt = min(t, max_dt); t -= dt` (:36-45).                          */
static double cn_one(double r, double q, double sg, double K, double Smax, double S0, int N, const double* dts, int nsteps,
                     int mode, double omega, double tol, int max_iter, double* Xout, double* err_out, int* iter_out) {
    double *S = malloc(sizeof(double) * N * 8), *X = S + N, *b = X + N, *lo = b + N, *di = lo + N, *up = di + N, *cp = up + N, *dp = cp + N;
    /* np.linspace(0, Smax, N) (:48; Smax = 3K unless the caller changed the attribute, :12): i*step, last point exact */
    double step = (Smax - 0.0) / (N - 1);
    for (int i = 0; i < N; ++i) S[i] = i * step + 0.0;
    S[N - 1] = Smax;
    double last_err = 0.0; int last_it = 0;   /* CNOptionSolver.err / .iter after the last solvePSOR (:103-104) */
    for (int i = 0; i < N; ++i) X[i] = fmax(K - S[i], 0.0);
    double dS = S[1] - S[0];
    for (int st = 0; st < nsteps; ++st) {
        double dt = dts[st];
        for (int i = 0; i < N - 1; ++i) {
            double a = sg * S[i] / dS;
            double alpha = 0.25 * dt * (a * a - (r - q) * S[i] / dS);
            double beta = 0.5 * dt * (r + a * a);
            double gamma = 0.25 * dt * (a * a + (r - q) * S[i] / dS);
            if (i == 0) { b[0] = X[0] * (1 - beta); lo[0] = 0; di[0] = 1 + beta; up[0] = 0; }
            else { b[i] = alpha * X[i - 1] + (1 - beta) * X[i] + gamma * X[i + 1]; lo[i] = -alpha; di[i] = 1 + beta; up[i] = -gamma; }
        }
        if (mode == 1) {
            int it = 0; double err = 1000;
            while (err > tol && it < max_iter) {
                ++it;
                memcpy(cp, X, sizeof(double) * N); /* x_old */
                for (int k = 0; k < N - 1; ++k) {
                    double xk = (1 - omega) * X[k] + omega * b[k] / di[k];
                    xk -= up[k] * X[k + 1] * omega / di[k];
                    xk -= lo[k] * X[k == 0 ? N - 1 : k - 1] * omega / di[k];
                    X[k] = xk;
                }
                int k = N - 2;
                X[N - 1] = (1 - omega) * X[k] + omega * b[k] / di[k];
                const double lastc[4] = {-1.0, 4.0, -5.0, 2.0};
                for (int jj = 0; jj < 4; ++jj) X[N - 1] -= lastc[jj] * X[N - 4 + jj] * omega / 2.0;
                double ss = 0;
                for (int i = 0; i < N; ++i) { X[i] = fmax(X[i], K - S[i]); double d = cp[i] - X[i]; ss += d * d; }
                err = sqrt(ss);
            }
            last_err = err; last_it = it;
        } else {
            /* rows 0..N-2 are tridiagonal; forward elimination gives X_i = dp_i - cp_i X_{i+1}.  The 4-term last row
             * (-1, 4, -5, 2 | 0) is closed by writing X_{N-2}, X_{N-3}, X_{N-4} as affine functions of x = X_{N-1}. */
            cp[0] = up[0] / di[0]; dp[0] = b[0] / di[0];
            for (int k = 1; k < N - 1; ++k) {
                double den = di[k] - lo[k] * cp[k - 1];
                cp[k] = up[k] / den; dp[k] = (b[k] - lo[k] * dp[k - 1]) / den;
            }
            double p2 = dp[N - 2], m2 = -cp[N - 2];
            double p3 = dp[N - 3] - cp[N - 3] * p2, m3 = -cp[N - 3] * m2;
            double p4 = dp[N - 4] - cp[N - 4] * p3, m4 = -cp[N - 4] * m3;
            X[N - 1] = (p4 - 4.0 * p3 + 5.0 * p2) / (2.0 - m4 + 4.0 * m3 - 5.0 * m2);
            if (mode == 2) X[N - 1] = fmax(X[N - 1], K - S[N - 1]);
            for (int k = N - 2; k >= 0; --k) {
                X[k] = dp[k] - cp[k] * X[k + 1];
                if (mode == 2) X[k] = fmax(X[k], K - S[k]);
            }
        }
    }
    /* np.interp(S0, S, X) */
    double v;
    if (S0 <= S[0]) v = X[0];
    else if (S0 >= S[N - 1]) v = X[N - 1];
    else {
        int j = (int)(S0 / dS); if (j > N - 2) j = N - 2;
        while (j > 0 && S[j] > S0) --j;
        while (j < N - 2 && S[j + 1] <= S0) ++j;
        double slope = (X[j + 1] - X[j]) / (S[j + 1] - S[j]);
        v = slope * (S0 - S[j]) + X[j];
    }
    if (Xout) memcpy(Xout, X, sizeof(double) * N);
    if (err_out) *err_out = last_err;
    if (iter_out) *iter_out = last_it;
    free(S);
    return v;
}

int oracle_cn_put_batch(int64_t Nopt, int n_space, int mode, const double* r, const double* q, const double* sigma,
                        const double* K, const double* Smax /* nullable: 3K */, const double* S0,
                        const double* dts /* Nopt x max_steps */, const int32_t* nsteps,
                        int max_steps, double omega, double tol, int max_iter, double* price, double* Xout /* nullable */,
                        double* err_out /* nullable */, int32_t* iter_out /* nullable */) {
    for (int64_t i = 0; i < Nopt; ++i)
        price[i] = cn_one(r[i], q[i], sigma[i], K[i], Smax ? Smax[i] : 3 * K[i], S0[i], n_space, dts + i * max_steps, nsteps[i],
                          mode, omega, tol, max_iter, Xout ? Xout + i * n_space : NULL, err_out ? err_out + i : NULL,
                          iter_out ? iter_out + i : NULL);
    return 0;
}
