// faop_seed.cu — the QD seed with the reference's own root finder (the default seed of every boundary solve).
//
// set_initial_guess (FastAmericanOptionSolverBase.py:258-265) calls QDplus.compute_exercise_boundary per collocation
// node, which is `scipy.optimize.root(self.exercise_boundary_func, x0=self.K, args=(tau,))` (QDplusAmericanOptionSolver.py:73):
// MINPACK hybrd (Powell's hybrid method) with scipy's defaults.  The reference returns whatever iterate hybrd stops at
// (xtol 1.49e-8 on the trust region, or one of its slow-progress exits — `success` is never read), and because the damped
// fixed point only contracts by ~1/2 per sweep and stops at an absolute 1e-5, that iterate — not the exact root — is what
// the final boundary carries.  So the seed kernel runs the same algorithm: hybrd specialised to one unknown
// (forward-difference Jacobian with h = sqrt(eps)|x|, dogleg step, trust-region update, Broyden rank-one update, the five
// exit tests), one thread per (option, node), every node started from x0 = K.  With one unknown the QR factorisation is
// r = -J, q = -1, every norm is |.| and r1updt / r1mpyq apply no rotation; the same restatement on the host (the test suite holds it)
// returns scipy's iterate bit for bit on the host.
// The residual (QDplusAmericanOptionSolver.py:110-125 through EuropeanOptionSolver.py:7-40) is evaluated in the reference's
// operation order with libdevice's exp / log / erfc, and the file is compiled with -fmad=false: numpy never fuses a*b+c, and
// for flat residuals (tiny tau, X = rK/q << K) the iterate hybrd stops at depends on the last bits of F.
#include "faop_kernels.h"

namespace faop {

struct SeedC {      // per (option, tau) invariants, formed exactly as the reference forms them
    double K, qQD, Ed, vs, hv, Ker, eq;
    bool put;
};

__device__ __forceinline__ double seed_Phi(double x) { return 0.5 * erfc(-x * kInvSqrt2); }

// One expression for puts and calls, so that a warp holding both does not split: with om = +1 (put) / -1 (call),
// P1 = Phi(-om d1), P2 = Phi(-om d2) and v = K e^{-r tau} P2 - S e^{-q tau} P1, the put's residual (:124) is
// (1 - e^{-q tau} P1) S + qQD ((K - S) - v) with p = v, and the call's (:122), (1 - e^{-q tau} P1) S - qQD ((S - K) - c) with
// c = -v, is the same number bit for bit: negation is exact, so (S - K) - c = -((K - S) - v) and -qQD (-X) = qQD X.
__device__ __forceinline__ double seed_F(const SeedC& c, double S) {
    const double d1 = log(S * c.Ed / c.K) / c.vs + c.hv;          // EuropeanOptionSolver.py:35-36
    const double d2 = d1 - c.vs;                                   // :39-40
    const double om = c.put ? 1.0 : -1.0;
    const double P1 = 0.5 * erfc((om * d1) * kInvSqrt2);           // Phi(-om d1) = erfc(om d1 / sqrt2) / 2
    const double P2 = 0.5 * erfc((om * d2) * kInvSqrt2);
    const double v = c.Ker * P2 - S * c.eq * P1;                   // :13 (put value; the call value :22 is -v)
    return (1 - c.eq * P1) * S + c.qQD * (c.K - S - v);
}

// hybrd for one unknown as a state machine around ONE residual evaluation per trip: `phase` says what the value just computed
// is (0 the initial F(x0), 1 the forward-difference point F(x + h), 2 the trial point F(x + p)).  Same arithmetic, same order as
// the textbook nesting (outer loop = Jacobian by differences, inner loop = dogleg steps with Broyden updates); written this way
// because the lanes of a warp sit in different places of that nesting, and with the residual inlined at three sites they ran
// one after the other (8 of 32 lanes active on average) instead of side by side.
struct Hybr {
    double x, xe, fvec, fnorm, diag, delta, xnorm, qtf, rr, Q, h, w1, w2, w3, pnorm;
    int nfev, it, ncsuc, ncfail, nslow1, nslow2, phase;
    bool jeval;
    __device__ __forceinline__ void start(double x0) {
        x = xe = x0;
        fvec = fnorm = diag = delta = xnorm = qtf = rr = Q = h = w1 = w2 = w3 = pnorm = 0.0;
        nfev = 0; it = 1; ncsuc = ncfail = nslow1 = nslow2 = phase = 0;
        jeval = true;
    }
    // consumes fe = F(xe); returns true when hybrd has stopped (x is the answer), otherwise xe holds the next point to evaluate
    __device__ __forceinline__ bool advance(double fe) {
        const double epsmch = 2.220446049250313e-16, xtol = 1.49012e-08, factor = 100.0;
        const int maxfev = 400;
        const double eps = 1.4901161193847656e-08;     // sqrt(max(epsfcn, epsmch)), epsfcn = eps
        ++nfev;
        bool step = false;                         // next: a dogleg trial point (true) or a forward-difference point (false)
        if (phase == 0) {
            fvec = fe;
            fnorm = fabs(fe);
        } else if (phase == 1) {                   // fdjac1 + qrfac: rdiag = -J, acnorm = |J|
            const double J = (fe - fvec) / h;
            const double acn = fabs(J);
            if (it == 1) {
                diag = acn != 0.0 ? acn : 1.0;
                xnorm = fabs(diag * x);
                delta = factor * xnorm;
                if (delta == 0.0) delta = factor;
            }
            if (J != 0.0) { qtf = -fvec; rr = -J; Q = -1.0; } else { qtf = fvec; rr = -0.0; Q = 1.0; }
            diag = fmax(diag, acn);
            jeval = true;
            step = true;
        } else {                                   // the trial point has been evaluated
            const double w4 = fe;
            const double fnorm1 = fabs(w4);
            double actred = -1.0;
            if (fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            w3 = qtf + rr * w1;
            const double temp = fabs(w3);
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
            const bool done = (delta <= xtol * xnorm || fnorm == 0.0)                      // info 1
                              || nfev >= maxfev                                            // info 2
                              || 0.1 * fmax(0.1 * delta, pnorm) <= epsmch * xnorm          // info 3
                              || nslow2 == 5 || nslow1 == 10                               // info 4 / 5
                              // NaN anywhere: every comparison above is false, upstream spins until maxfev; the iterate no longer changes
                              || !(fnorm1 == fnorm1) || !(delta == delta);
            if (done) return true;
            if (ncfail != 2) {
                const double sm = Q * w4;          // Broyden update of r (= -J) and of (q transpose) fvec
                w2 = (sm - w3) / pnorm;
                w1 = diag * ((diag * w1) / pnorm);
                if (ratio >= 1e-4) qtf = sm;
                rr = rr + w2 * w1;
                jeval = false;
                step = true;
            }                                      // ncfail == 2: recompute the Jacobian by forward differences (step stays false)
        }
        if (step) {                                // dogleg
            double temp = rr;
            if (temp == 0.0) { temp = epsmch * fabs(rr); if (temp == 0.0) temp = epsmch; }
            const double xg = qtf / temp;
            const double qnorm = fabs(diag * xg);
            double p;
            if (qnorm <= delta) p = xg;
            else {
                double wa1 = (rr * qtf) / diag;
                const double gnorm = fabs(wa1);
                double sgnorm = 0.0, alpha = delta / qnorm;
                if (gnorm != 0.0) {
                    wa1 = (wa1 / gnorm) / diag;
                    const double t2 = fabs(rr * wa1);
                    sgnorm = (gnorm / t2) / t2;
                    alpha = 0.0;
                    if (sgnorm < delta) {
                        const double bnorm = fabs(qtf);
                        double tt = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
                        const double dq = delta / qnorm, sd = sgnorm / delta;
                        tt = tt - dq * (sd * sd) + sqrt((tt - dq) * (tt - dq) + (1.0 - dq * dq) * (1.0 - sd * sd));
                        alpha = (dq * (1.0 - sd * sd)) / tt;
                    }
                }
                const double tt = (1.0 - alpha) * fmin(sgnorm, delta);
                p = tt * wa1 + alpha * xg;
            }
            w1 = -p;
            w2 = x + w1;
            w3 = diag * w1;
            pnorm = fabs(w3);
            if (it == 1) delta = fmin(delta, pnorm);
            xe = w2;
            phase = 2;
        } else {
            h = eps * fabs(x);                     // fdjac1's step
            if (h == 0.0) h = eps;
            xe = x + h;
            phase = 1;
        }
        return false;
    }
};

// one root, the calling lane alone (leaf kernels)
__device__ double hybr1_seed(const SeedC& c) {
    Hybr s;
    s.start(c.K);
    while (!s.advance(seed_F(c, s.xe))) {}
    return s.x;
}

__device__ __forceinline__ SeedC seed_consts(double tau, double r, double q, double sg, double K, bool put) {
    SeedC c;
    c.K = K;
    c.put = put;
    const double Nn = 2 * (r - q) / (sg * sg), M = 2 * r / (sg * sg), h = 1 - exp(-r * tau);      // :211-218
    const double disc = sqrt((Nn - 1) * (Nn - 1) + 4 * M / h);
    c.qQD = put ? -0.5 * (Nn - 1) - 0.5 * disc : -0.5 * (Nn - 1) + 0.5 * disc;                    // :127-134
    c.Ed = exp((r - q) * tau);
    c.vs = sg * sqrt(tau);
    c.hv = 0.5 * sg * sqrt(tau);
    c.Ker = K * exp(-r * tau);
    c.eq = exp(-q * tau);
    return c;
}

// Persistent warps with LANE-LEVEL REFILL.  hybrd needs 9-12 evaluations at the long nodes, 18-30 at the shortest and now and then
// 50; with one thread per root and the lanes of a warp left to themselves only a third of the lanes is active on average (ncu:
// 4.4 k FP64 warp instructions per option against 1.4 k if all 32 lanes worked; 11.3 ms per 1M C2 options).  Measured steps to
// this form: a vote per trip so that finished lanes wait, 13.2 ms (the slowest lane of 32 sets the pace); dealing items to lanes as
// they finish from NODE-major ranges, 13.9 ms (the warps that own the short nodes finish last); OPTION-major ranges, 10.2 ms; one
// residual expression for puts and calls (seed_F), so that a warp holding both does not split, and refills batched by the
// quarter-warp, 6.7 ms with 23 of 32 lanes active on average and the residual itself at 31-32.  So every warp owns a contiguous range of the
// N n root problems — OPTION-major (item g = node g % n of option g / n), so that every range holds the same mix of cheap and
// expensive nodes — and deals them to its lanes one at a time: a lane that has finished takes the next item of the range at the
// top of the next trip.  One residual evaluation per trip for all 32 lanes side by side; the lane-private parts (taking an item,
// hybrd's bookkeeping) are what diverges.
constexpr int kSeedThreads = 128;
constexpr int kSeedRefill = 8;
__global__ void __launch_bounds__(kSeedThreads) qd_seed_hybr_kernel(AloArgs a, int refill) {
    const int n = a.tl.n, nn = n + 1;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kSeedThreads + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kSeedThreads) >> 5;
    const int64_t total = a.N * n;
    // the pinned node tau_n = 0 (B_at_zero) and the European shortcut rows, grid-stride
    for (int64_t o = (int64_t)blockIdx.x * kSeedThreads + threadIdx.x; o < a.N; o += (int64_t)gridDim.x * kSeedThreads) {
        const bool put = a.is_put[o] != 0;
        a.seeds_out[o * nn + n] = (a.q[o] == 0 && !put) ? nan("") : b_at_zero(a.r[o], a.q[o], a.K[o], put);
    }
    const int64_t per = (total + nwarps - 1) / nwarps;
    int64_t next = warp * per;                    // warp-uniform cursor into this warp's range
    const int64_t end = (next + per < total) ? next + per : total;
    Hybr s;
    SeedC c;
    c.K = 1.0; c.qQD = -1.0; c.Ed = 1.0; c.vs = 1.0; c.hv = 0.5; c.Ker = 1.0; c.eq = 1.0; c.put = true;    // harmless idle problem
    s.start(1.0);
    bool have = false;
    double* dst = nullptr;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        // deal new items only once a quarter of the lanes is idle (or nothing is running): the per-item setup (four exponentials,
        // two square roots, a handful of divisions) then runs with >= 8 lanes instead of one or two at a time
        if (next < end && (__popc(need) >= refill || need == 0xffffffffu)) {                 // warp-uniform
            if (!have) {
                const int64_t g = next + __popc(need & ((1u << lane) - 1));
                if (g < end) {
                    const int64_t opt = g / n;
                    const int j = (int)(g - opt * n);
                    const double r = a.r[opt], q = a.q[opt], sg = a.sigma[opt], K = a.K[opt];
                    const double T = a.tau_max ? a.tau_max[opt] : a.T[opt];      // collocation domain [0, tau_max], Base.py:28,223
                    const bool put = a.is_put[opt] != 0;
                    const double cj = a.tables[a.tl.off_c + j];
                    const double tau = (cj + 1) * (cj + 1) * (T - 0.0) / 4 + 0.0;      // to_orig_point, Base.py:239-240
                    dst = a.seeds_out + opt * nn + j;
                    if (q == 0 && !put) *dst = nan("");            // European shortcut (Base.py:50-53): no boundary is solved
                    else if (tau == 0) *dst = b_at_zero(r, q, K, put);
                    else {
                        c = seed_consts(tau, r, q, sg, K, put);
                        s.start(K);
                        have = true;
                    }
                }
            }
            next += __popc(need);
        }
        if (!__any_sync(0xffffffffu, have)) {
            if (next >= end) break;
            continue;                             // a trip that only dealt trivial items
        }
        const double fe = seed_F(c, s.xe);        // all 32 lanes side by side (idle lanes evaluate their last problem again)
        if (have && s.advance(fe)) {
            *dst = s.x;
            have = false;
        }
    }
}

// QDplus.compute_exercise_boundary per element (QDplusAmericanOptionSolver.py:69-76), the reference's root finder
__global__ void __launch_bounds__(128) qd_boundary_kernel(int64_t N, const double* r, const double* q, const double* sigma,
                                                               const double* K, const double* tau, const uint8_t* is_put,
                                                               double* B) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const bool put = is_put[i] != 0;
    B[i] = (tau[i] == 0) ? b_at_zero(r[i], q[i], K[i], put) : hybr1_seed(seed_consts(tau[i], r[i], q[i], sigma[i], K[i], put));
}

// QDplus.exercise_boundary_func, QDplusAmericanOptionSolver.py:110-125 (tau == 0 -> 0)
__global__ void __launch_bounds__(256) qd_residual_kernel(int64_t N, const double* r, const double* q, const double* sigma,
                                                          const double* K, const double* tau, const double* S,
                                                          const uint8_t* is_put, double* F) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (tau[i] == 0) { F[i] = 0.0; return; }
    F[i] = seed_F(seed_consts(tau[i], r[i], q[i], sigma[i], K[i], is_put[i] != 0), S[i]);
}

// QDplus.price, QDplusAmericanOptionSolver.py:44-67, with compute_miscellaneous' c, b (:142-218).
// Quirks kept: tau == 0 returns max(S-K, 0) and boundary K for puts too (:45-47); the put-form
// correction (K - Sb - p(Sb)) is used for calls as well (:67); c() uses Phi(-d1) for both types (:172);
// theta is the put theta (EuropeanOptionSolver.py:25-32).
__global__ void __launch_bounds__(128) qd_price_kernel(int64_t N, const double* r_, const double* q_, const double* sigma_,
                                                       const double* K_, const double* tau_, const double* S_,
                                                       const uint8_t* is_put, double* price, double* boundary) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double r = r_[i], q = q_[i], sg = sigma_[i], K = K_[i], tau = tau_[i], S = S_[i];
    bool put = is_put[i] != 0;
    if (tau == 0) {
        price[i] = fmax(S - K, 0.0);
        if (boundary) boundary[i] = K;
        return;
    }
    double Sb = hybr1_seed(seed_consts(tau, r, q, sg, K, put));      // compute_exercise_boundary, :69-76
    if (boundary) boundary[i] = Sb;
    double ind = put ? -1.0 : 1.0;
    double s2 = sg * sg;
    double Nn = 2 * (r - q) / s2, M = 2 * r / s2, h = 1 - exp(-r * tau);
    double disc = sqrt((Nn - 1) * (Nn - 1) + 4 * M / h);
    double qQD = -0.5 * (Nn - 1) + ind * 0.5 * disc;
    double qQDdot = M / (h * h * disc);
    double svt = sg * sqrt(tau);
    double d1 = euro_d1(tau, Sb, r, q, sg, K);
    double p = euro_value(tau, Sb, r, q, sg, K, put);
    double theta = euro_theta(tau, Sb, r, q, sg, K);
    double eq = exp(-q * tau), ert = exp(r * tau);
    double Pm1 = norm_cdf(-d1), pd1 = norm_pdf(d1);
    double KSp = K - Sb - p;
    double dFdh = qQD * theta * ert / r + qQDdot * KSp + (Sb * q * eq * Pm1) / (r * (1 - h)) -
                  (Sb * eq * pd1) / (2 * r * tau * (1 - h)) * (2 * log(Sb / K) / svt - d1);
    double dFdS = (1 - qQD) * (1 - eq * Pm1) + (eq * pd1) / svt;
    double dlogSdh = -dFdh / dFdS;
    double den = 2 * qQD + Nn - 1;
    double c0 = -(1 - h) * M / den * (1 / h - (theta * ert) / (r * KSp) + qQDdot / den);
    double c = c0 - ((1 - h) * M) / den * ((1 - eq * Pm1) / KSp + qQD / Sb) * dlogSdh;
    double b = ((1 - h) * M * qQDdot) / (2 * den);
    if (ind * (Sb - S) <= 0) {
        price[i] = ind * (S - K);
        return;
    }
    double pS = euro_value(tau, S, r, q, sg, K, put);
    double lg = log(S / Sb);
    price[i] = pS + KSp / (1 - b * (lg * lg) - c * lg) * pow(S / Sb, qQD);
}

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int launch_qd_seed_hybr(const AloArgs& a, cudaStream_t st) {
    if (a.N == 0) return 0;
    // persistent: enough warps to fill the machine (5 CTAs of 4 warps per SM at ~96 registers), never more warps than 64-item ranges
    const int64_t want = (a.N * a.tl.n + 64 * 4 - 1) / (64 * 4);
    const int64_t cap = (int64_t)kSMs * 5;      // 4 or 5 CTAs per SM and a refill threshold of 4..12 lanes measure the same (60.5-61.5 ms per C2 step)
    qd_seed_hybr_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), kSeedThreads, 0, st>>>(a, kSeedRefill);
    count_launch();
    return launch_status();
}

int launch_qd_boundary(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                       const double* tau, const uint8_t* is_put, double* B, cudaStream_t st) {
    if (N == 0) return 0;
    qd_boundary_kernel<<<blocks_for(N, 128), 128, 0, st>>>(N, r, q, sigma, K, tau, is_put, B);
    count_launch();
    return launch_status();
}
int launch_qd_residual(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                       const double* tau, const double* S, const uint8_t* is_put, double* F, cudaStream_t st) {
    if (N == 0) return 0;
    qd_residual_kernel<<<blocks_for(N, 256), 256, 0, st>>>(N, r, q, sigma, K, tau, S, is_put, F);
    count_launch();
    return launch_status();
}
int launch_qd_price(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                    const double* tau, const double* S, const uint8_t* is_put, double* price, double* boundary,
                    cudaStream_t st) {
    if (N == 0) return 0;
    qd_price_kernel<<<blocks_for(N, 128), 128, 0, st>>>(N, r, q, sigma, K, tau, S, is_put, price, boundary);
    count_launch();
    return launch_status();
}

}  // namespace faop
