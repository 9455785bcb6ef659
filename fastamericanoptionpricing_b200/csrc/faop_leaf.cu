// faop_leaf.cu — leaf kernels: European closed forms (the QD+ leaf kernels live in faop_seed.cu), Chebyshev fit and
// Clenshaw evaluation, and the FP64 FMA-pipe peak microbenchmark used as the roofline denominator.
// One thread per element, coalesced structure-of-arrays loads, grid sized to the batch.
#include "faop_kernels.h"
#include "faop_fastmath.cuh"

namespace faop {

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// EuropeanOption.european_put_value / european_call_value, EuropeanOptionSolver.py:7-22
__global__ void __launch_bounds__(256) european_kernel(int64_t N, const double* tau, const double* S, const double* r,
                                                       const double* q, const double* sigma, const double* K,
                                                       const uint8_t* is_put, double* price) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    price[i] = euro_value(tau[i], S[i], r[i], q[i], sigma[i], K[i], is_put[i] != 0);
}

// european_option_theta / d1 / d2, EuropeanOptionSolver.py:25-40
__global__ void __launch_bounds__(256) european_greeks_kernel(int64_t N, const double* tau, const double* S, const double* r,
                                                              const double* q, const double* sigma, const double* K,
                                                              double* theta, double* d1, double* d2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (theta) theta[i] = euro_theta(tau[i], S[i], r[i], q[i], sigma[i], K[i]);
    if (d1 || d2) {
        double v = euro_d1(tau[i], S[i], r[i], q[i], sigma[i], K[i]);
        if (d1) d1[i] = v;
        if (d2) d2[i] = v - sigma[i] * sqrt(tau[i]);
    }
}

// ChebyshevInterpolation.std_coeff, ChebyshevInterpolation.py:43-51: one thread per (row, k).
__global__ void __launch_bounds__(256) cheb_fit_kernel(int64_t N, int n, const double* y, double* a) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nn = n + 1;
    if (idx >= N * nn) return;
    int64_t row = idx / nn;
    int k = (int)(idx - row * nn);
    const double* yr = y + row * nn;
    double ans = 0.0;
    for (int i = 0; i <= n; ++i) {
        double term = yr[i] * cospi((double)((int64_t)i * k % (2 * n)) / (double)n);
        if (i == 0 || i == n) term *= 0.5;
        ans += term;
    }
    a[idx] = ans * (2.0 / n);
}

// ChebyshevInterpolation.std_cheby_single_value, ChebyshevInterpolation.py:31-41: one thread per (row, point).
__global__ void __launch_bounds__(256) cheb_eval_kernel(int64_t N, int n, int npts, const double* a, const double* z,
                                                        double* out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * npts) return;
    int64_t row = idx / npts;
    out[idx] = clenshaw(n, a + row * (n + 1), z[idx]);
}

// Config-5 surface decode: flat index -> (K, T, sigma) with numpy.linspace arithmetic (i*step + start, unfused;
// last point == stop), index order i = (iK*nT + iT)*nV + iV as in workloads.c5_surface.
__global__ void __launch_bounds__(256) surface_decode_kernel(faop_surface s, double stepK, double stepT, double stepV,
                                                             int64_t first, int64_t count, double* r, double* q,
                                                             double* sigma, double* K, double* T, double* S, uint8_t* is_put) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    int64_t i = first + j;
    int iV = (int)(i % s.nV);
    int64_t rest = i / s.nV;
    int iT = (int)(rest % s.nT);
    int iK = (int)(rest / s.nT);
    K[j] = (iK == s.nK - 1 && s.nK > 1) ? s.K1 : __dadd_rn(__dmul_rn((double)iK, stepK), s.K0);
    T[j] = (iT == s.nT - 1 && s.nT > 1) ? s.T1 : __dadd_rn(__dmul_rn((double)iT, stepT), s.T0);
    sigma[j] = (iV == s.nV - 1 && s.nV > 1) ? s.V1 : __dadd_rn(__dmul_rn((double)iV, stepV), s.V0);
    r[j] = s.r;
    q[j] = s.q;
    S[j] = s.S;
    is_put[j] = (uint8_t)(s.is_put != 0);
}

int launch_surface_decode(const faop_surface& s, int64_t first, int64_t count, double* r, double* q, double* sigma,
                          double* K, double* T, double* S, uint8_t* is_put, cudaStream_t st) {
    if (count == 0) return 0;
    double stepK = s.nK > 1 ? (s.K1 - s.K0) / (s.nK - 1) : 0.0;
    double stepT = s.nT > 1 ? (s.T1 - s.T0) / (s.nT - 1) : 0.0;
    double stepV = s.nV > 1 ? (s.V1 - s.V0) / (s.nV - 1) : 0.0;
    surface_decode_kernel<<<blocks_for(count, 256), 256, 0, st>>>(s, stepK, stepT, stepV, first, count, r, q, sigma, K, T, S, is_put);
    count_launch();
    return launch_status();
}

// Unit-test hook for the device math of faop_fastmath.cuh: out[i] = f_kind(x[i]).
__global__ void __launch_bounds__(256) dev_math_kernel(int kind, int64_t N, const double* x, double* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double v = x[i], r;
    switch (kind) {
        case 0: r = fast_exp(v); break;
        case 1: r = mills_half(v); break;
        case 2: r = fast_sqrt_pos(v); break;
        case 3: r = fast_rcp(v); break;
        case 4: r = fast_scaled_cdf(v, fast_exp(-0.5 * v * v), 1.0); break;   // Phi(v)
        case 5: { double a[1] = {v}, o[1]; half_erfcx_n<1, false>(a, o); r = o[0]; break; }      // erfcx(v)/2
        case 6: { double a[1] = {v}, o[1]; sqrt_clamped_n<1>(a, o); r = o[0]; break; }    // sqrt(max(0, v))
        case 8: { double a[1] = {v}, o[1]; half_erfcx_n<1, true>(a, o); r = o[0]; break; }       // erfcx(v)/2, degree 14
        case 9: { double a[1] = {v}, o[1]; fast_exp_n<1, true>(a, o); r = o[0]; break; }         // e^v, degree 8
        case 12: { double a[1] = {v}, o[1]; half_erfcx_n<1, 2>(a, o); r = o[0]; break; }          // erfcx(v)/2, degree 16
        case 10: r = fast_log(v); break;
        case 11: r = norm_cdf(v); break;                                        // Phi(v), the library-accuracy form
        case 7: { double a[2] = {v, v - 1.0}, o[2]; fast_exp_n<2>(a, o); r = o[0] + o[1]; break; }  // e^v + e^(v-1)
        default: r = nan("");
    }
    out[i] = r;
}
int launch_dev_math(int kind, int64_t N, const double* x, double* out, cudaStream_t st) {
    if (N == 0) return 0;
    dev_math_kernel<<<blocks_for(N, 256), 256, 0, st>>>(kind, N, x, out);
    count_launch();
    return launch_status();
}

// FP64 FMA-pipe peak: 8 independent DFMA chains per thread, 4096 rounds, 8 resident CTAs/SM.  Multiplier and addend are
// kernel parameters (uniform operands): on sm_100a a DFMA with at most two distinct 64-bit REGISTER sources retires at
// 63.6 lanes/clk/SM (36.98 TFLOP/s at 1965 MHz), one with three distinct register sources at only 53-54 lanes/clk/SM
// (register-file read bandwidth) - measured with tools/dfma_forms.cu, profiles/r1_dfma_operand_forms.txt.
constexpr int kPeakIters = 4096;
constexpr int kPeakChains = 8;
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, double m, double b) {
    double x[kPeakChains];
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) x[c] = 1.0 + c + threadIdx.x * 1e-9;
#pragma unroll 1
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int c = 0; c < kPeakChains; ++c) x[c] = fma(x[c], m, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) s += x[c];
    if (s == 123456.789) sink[0] = s;  // never true; keeps the chains alive
}

// ---------------------------------------------------------------- implied volatility: one Illinois (regula falsi) step
// The American price is increasing in sigma, so price(sigma) - target has one root in a bracket [lo, hi].  State per
// option: the bracket, the residuals at its ends (NaN = not evaluated yet) and which end moved last.  Each call consumes
// the price just computed at `sigma` and writes the next trial point into `sigma`:
//   first two calls evaluate the ends of the bracket; afterwards the secant point of the bracket, with the retained end's
//   residual halved (in the secant formula only) whenever the same end moves twice in a row (Illinois), clipped into the
//   bracket's interior.
// A target outside [price(lo), price(hi)] pins sigma to NaN.  SURVEY 8f row 2 (callers of solve(); not in the reference).
__global__ void __launch_bounds__(256) iv_update_kernel(int64_t N, const double* target, const double* price, double* sigma,
                                                        double* lo, double* hi, double* flo, double* fhi, int32_t* side) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x = sigma[i], f = price[i] - target[i];
    double a = lo[i], b = hi[i], fa = flo[i], fb = fhi[i];
    const int state = side[i];
    if (state == -3) return;                    // target outside the bracket: stays NaN
    if (state == -2) {                          // first call: x == lo
        flo[i] = f; side[i] = -1; sigma[i] = b;
        return;
    }
    // state word: bits 0-1 which end moved last, bits 2-3 consecutive steps that shrank the bracket < 2x, bits 4-9 how often the
    // RETAINED end's residual has been halved (Illinois).  flo / fhi always hold the TRUE residuals (the caller's termination
    // test and its choice of the answer read them); the Illinois factor 2^-k is applied only when the secant point is formed.
    int sd = state & 3, slow = (state >> 2) & 3, k = (state >> 4) & 63;
    const double w_old = b - a;
    if (state == -1) {                          // second call: x == hi
        fb = f;
        if (!(fa <= 0.0) || !(fb >= 0.0)) { side[i] = -3; fhi[i] = fb; sigma[i] = nan(""); return; }
        sd = 0; slow = 0; k = 0;
    } else if (f == 0.0) {
        lo[i] = hi[i] = x; flo[i] = fhi[i] = 0.0;
        return;                                 // sigma[i] stays at the root
    } else if (f < 0.0) {
        a = x; fa = f;
        k = (sd == 1) ? min(k + 1, 60) : 0;
        sd = 1;
    } else {
        b = x; fb = f;
        k = (sd == 2) ? min(k + 1, 60) : 0;
        sd = 2;
    }
    const double w = b - a;
    slow = (state >= 0 && w > 0.5 * w_old) ? min(slow + 1, 3) : 0;
    const double sc = ldexp(1.0, -k);
    const double fas = (sd == 2) ? fa * sc : fa, fbs = (sd == 1) ? fb * sc : fb;     // the end that did not move is the stale one
    double xn = (fbs == fas) ? 0.5 * (a + b) : (a * fbs - b * fas) / (fbs - fas);
    if (slow >= 2) { xn = 0.5 * (a + b); slow = 0; }   // flat-then-steep residuals stall the secant: bisect
    if (!(xn > a + 1e-3 * w)) xn = a + 1e-3 * w;      // also catches NaN
    if (!(xn < b - 1e-3 * w)) xn = b - 1e-3 * w;
    if (w <= 4e-16 * b) xn = x;                        // bracket at rounding level: stay
    lo[i] = a; hi[i] = b; flo[i] = fa; fhi[i] = fb; side[i] = sd | (slow << 2) | (k << 4); sigma[i] = xn;
}
int launch_iv_update(int64_t N, const double* target, const double* price, double* sigma, double* lo, double* hi, double* flo,
                     double* fhi, int32_t* side, cudaStream_t st) {
    if (N == 0) return 0;
    iv_update_kernel<<<blocks_for(N, 256), 256, 0, st>>>(N, target, price, sigma, lo, hi, flo, fhi, side);
    count_launch();
    return launch_status();
}

// Transcendental-mix microbenchmark: the per-quadrature-point math of the flagship kernel — 12 DFMA of
// interpolation, sqrt, the d+- arguments, two degree-8 exp, one degree-16 erfcx, the combine — on register operands only,
// two points in lock step, 16 warps per SM: no shared-memory traffic, no reductions, no node closure.  Its rate in
// points/s is the speed of light of the point math on this GPU; options/s * sweeps * n * l of the real kernel over that rate
// says how much the surrounding machinery costs.
constexpr int kMixIters = 2048;
__global__ void __launch_bounds__(512, 1) point_mix_kernel(double* sink, double seed) {
    double Hj[12], acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int j = 0; j < 12; ++j) Hj[j] = 0.01 * (j + 1) + seed * threadIdx.x;
    double m0 = 0.3 + 1e-6 * threadIdx.x, m1 = 0.29 - 1e-6 * threadIdx.x;
    const double L = -0.2 + seed, cinv2rh = 1.7, apc2h = 0.05, ssu2h = 0.11, qtg = 0.02, rtg = 0.04, lnw = -3.0, cis = 0.8, h = 0.37, eqw = 0.9;
#pragma unroll 1
    for (int it = 0; it < kMixIters; ++it) {
        double Hu[2] = {0.0, 0.0}, root[2];
#pragma unroll
        for (int j = 0; j < 12; ++j) {      // the matrix entries are lane data in the real kernel; here two drifting registers
            Hu[0] = fma(m0, Hj[j], Hu[0]);
            Hu[1] = fma(m1, Hj[j], Hu[1]);
        }
        m0 = fma(m0, 0.999999, 1e-9);
        m1 = fma(m1, 0.999998, 2e-9);
        sqrt_clamped_n<2>(Hu, root);
        double ex[4], xa[2], xp[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const double lnz = L + root[p];
            xp[p] = fma(lnz, cinv2rh, apc2h);
            const double xm = xp[p] - ssu2h;
            ex[2 * p] = fma(-xp[p], xp[p], qtg + lnw);
            ex[2 * p + 1] = fma(-xm, xm, rtg + lnw);
            xa[p] = fabs(xp[p]);
        }
        double E[4], G[2];
        fast_exp_n<4, true>(ex, E);
        half_erfcx_n<2, 2>(xa, G);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const double ec = E[2 * p] * G[p];
            const double cdf = (__double2hiint(xp[p]) < 0) ? ec : eqw - ec;
            acc0 += E[2 * p + 1];
            acc1 += fma(cis, E[2 * p], h * cdf);
        }
    }
    const double s2 = acc0 + acc1;
    if (s2 == 123456.789) sink[0] = s2;
}
int launch_point_mix_peak(int launches, double* points_per_s, double* ms_per_launch, cudaStream_t st) {
    double* buf = nullptr;
    FAOP_CUDA_OK(cudaMalloc(&buf, sizeof(double)));
    cudaEvent_t e0, e1;
    FAOP_CUDA_OK(cudaEventCreate(&e0));
    FAOP_CUDA_OK(cudaEventCreate(&e1));
    const int blocks = kSMs * 4, threads = 512;
    for (int w = 0; w < 2; ++w) { point_mix_kernel<<<blocks, threads, 0, st>>>(buf, 1e-7); count_launch(); }
    FAOP_CUDA_OK(cudaEventRecord(e0, st));
    for (int i = 0; i < launches; ++i) { point_mix_kernel<<<blocks, threads, 0, st>>>(buf, 1e-7); count_launch(); }
    FAOP_CUDA_OK(cudaEventRecord(e1, st));
    FAOP_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    FAOP_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double points = 2.0 * (double)kMixIters * (double)blocks * threads * launches;
    if (points_per_s) *points_per_s = points / (ms * 1e-3);
    if (ms_per_launch) *ms_per_launch = ms / launches;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return launch_status();
}

// ---------------------------------------------------------------- Bjerksund-Stensland (1993) flat-boundary call
// BksdStld (BjerksundStenslandSolver.py:4-43).  phi (:26-33):
FAOP_D double bs_phi(double S, double T, double gamma, double H, double X, double r, double b, double sg) {
    const double s2 = sg * sg, svt = sg * sqrt(T);
    const double lamb = -r + gamma * b + 0.5 * gamma * (gamma - 1) * s2;
    const double kappa = 2 * b / s2 + (2 * gamma - 1.0);
    const double drift = (b + (gamma - 0.5) * s2) * T;
    const double d1 = -(log(S / H) + drift) / svt;
    const double d2 = -(log(X * X / (S * H)) + drift) / svt;
    return exp(lamb * T) * pow(S, gamma) * (norm_cdf(d1) - pow(X / S, kappa) * norm_cdf(d2));
}
// price (:15-21), computeAlpha (:23-24), computeExerciseBoundary (:38-43).  The reference reads the module globals
// `K` and `T` at :40 and :42 (it only runs as __main__); Kb / Tb carry those two values (nullable = the option's own
// strike and maturity).  is_put (nullable = all calls) prices a put by the symmetry the reference's __main__ uses (:53):
// put(S, K, r, q) = call(S' = K, K' = S, r' = q, q' = r).  No S >= X branch, as upstream.
__global__ void __launch_bounds__(128) bjerksund_stensland_kernel(int64_t N, const double* r_, const double* q_,
                                                                  const double* sigma_, const double* K_, const double* T_,
                                                                  const double* S_, const uint8_t* is_put, const double* Kb_,
                                                                  const double* Tb_, double* price, double* boundary) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double r = r_[i], q = q_[i], sg = sigma_[i], K = K_[i], T = T_[i], S = S_[i];
    if (is_put && is_put[i]) {
        double t = r; r = q; q = t;
        t = S; S = K; K = t;
    }
    const double Kb = Kb_ ? Kb_[i] : K, Tb = Tb_ ? Tb_[i] : T;
    const double b = r - q, s2 = sg * sg;
    const double hb = b / s2 - 0.5;
    const double beta = (0.5 - b / s2) + sqrt(hb * hb + 2.0 * r / s2);                  // :13
    const double B0 = fmax(K, r / (r - b) * Kb);                                        // :40
    const double Binf = beta / (beta - 1) * K;                                          // :41
    const double h = -(b * Tb + 2 * sg * sqrt(T)) * (K * K / ((Binf - B0) * B0));       // :42
    const double X = B0 + (Binf - B0) * (1 - exp(h));                                   // :43
    if (boundary) boundary[i] = X;
    const double alpha = (X - K) * pow(X, -beta);                                       // :24
    price[i] = alpha * pow(S, beta) - alpha * bs_phi(S, T, beta, X, X, r, b, sg) + bs_phi(S, T, 1.0, X, X, r, b, sg) -
               bs_phi(S, T, 1.0, K, X, r, b, sg) - K * bs_phi(S, T, 0.0, X, X, r, b, sg) + K * bs_phi(S, T, 0.0, K, X, r, b, sg);
}
// BksdStld.phi as a stand-alone component (:26-33)
__global__ void __launch_bounds__(256) bs_phi_kernel(int64_t N, const double* S, const double* T, const double* gamma,
                                                     const double* H, const double* X, const double* r, const double* q,
                                                     const double* sigma, double* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out[i] = bs_phi(S[i], T[i], gamma[i], H[i], X[i], r[i], r[i] - q[i], sigma[i]);
}

int launch_bjerksund_stensland(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                               const double* T, const double* S, const uint8_t* is_put, const double* Kb, const double* Tb,
                               double* price, double* boundary, cudaStream_t st) {
    if (N == 0) return 0;
    bjerksund_stensland_kernel<<<blocks_for(N, 128), 128, 0, st>>>(N, r, q, sigma, K, T, S, is_put, Kb, Tb, price, boundary);
    count_launch();
    return launch_status();
}
int launch_bs_phi(int64_t N, const double* S, const double* T, const double* gamma, const double* H, const double* X,
                  const double* r, const double* q, const double* sigma, double* out, cudaStream_t st) {
    if (N == 0) return 0;
    bs_phi_kernel<<<blocks_for(N, 256), 256, 0, st>>>(N, S, T, gamma, H, X, r, q, sigma, out);
    count_launch();
    return launch_status();
}
int launch_european(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                    const double* sigma, const double* K, const uint8_t* is_put, double* price, cudaStream_t st) {
    if (N == 0) return 0;
    european_kernel<<<blocks_for(N, 256), 256, 0, st>>>(N, tau, S, r, q, sigma, K, is_put, price);
    count_launch();
    return launch_status();
}
int launch_european_greeks(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                           const double* sigma, const double* K, double* theta, double* d1, double* d2, cudaStream_t st) {
    if (N == 0) return 0;
    european_greeks_kernel<<<blocks_for(N, 256), 256, 0, st>>>(N, tau, S, r, q, sigma, K, theta, d1, d2);
    count_launch();
    return launch_status();
}
int launch_cheb_fit(int64_t N, int n, const double* y, double* a, cudaStream_t st) {
    if (N == 0) return 0;
    cheb_fit_kernel<<<blocks_for(N * (n + 1), 256), 256, 0, st>>>(N, n, y, a);
    count_launch();
    return launch_status();
}
int launch_cheb_eval(int64_t N, int n, int npts, const double* a, const double* z, double* out, cudaStream_t st) {
    if (N == 0 || npts == 0) return 0;
    cheb_eval_kernel<<<blocks_for(N * npts, 256), 256, 0, st>>>(N, n, npts, a, z, out);
    count_launch();
    return launch_status();
}

int launch_fp64_peak(int launches, double* tflops, double* ms_per_launch, cudaStream_t st) {
    double* buf = nullptr;
    FAOP_CUDA_OK(cudaMalloc(&buf, sizeof(double)));
    cudaEvent_t e0, e1;
    FAOP_CUDA_OK(cudaEventCreate(&e0));
    FAOP_CUDA_OK(cudaEventCreate(&e1));
    const int blocks = kSMs * 8 * 4, threads = 256;  // 8 resident CTAs/SM x 4 waves
    for (int w = 0; w < 2; ++w) { fp64_peak_kernel<<<blocks, threads, 0, st>>>(buf, 0.999999, 1e-7); count_launch(); }
    FAOP_CUDA_OK(cudaEventRecord(e0, st));
    for (int i = 0; i < launches; ++i) { fp64_peak_kernel<<<blocks, threads, 0, st>>>(buf, 0.999999, 1e-7); count_launch(); }
    FAOP_CUDA_OK(cudaEventRecord(e1, st));
    FAOP_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    FAOP_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * kPeakChains * (double)kPeakIters * (double)blocks * threads * launches;
    if (tflops) *tflops = flops / (ms * 1e-3) / 1e12;
    if (ms_per_launch) *ms_per_launch = ms / launches;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return launch_status();
}

}  // namespace faop
