// faop_surface.cu — deduplicated implied-surface sweep (BASELINE.json configs[4]).
//
// The early-exercise boundary is homogeneous of degree one in the strike: B(tau; K) = K b(tau), with b depending on
// (r, q, sigma, T, type) only — the QD residual (QDplusAmericanOptionSolver.py:110-125) and f = K e^{-(r-q)tau} N/D
// (FastAmericanOptionSolverBase.py:267-279) both scale with K, so every Jacobi iterate does.  All strikes of a surface
// that share (T, sigma) therefore share ONE boundary solve at K = 1 and differ only in
//   * how many sweeps the reference would have run: its stopping rule ||B_old - B_new||_2 <= iter_tol (:119) is absolute,
//     i.e. K e_j <= tol for the normalised error e_j, so a strike stops at the first iterate with e_j <= tol / K.  The
//     solve (snapshot mode of the boundary kernels) records every iterate that is a new error minimum below tol / K_min
//     and runs on until tol / K_max; each strike picks the first recorded iterate that satisfies its own threshold;
//   * the m-point price integral (american_value_with_known_boundary, :92-103), evaluated here.
// Price kernel: one CTA per (T, sigma) group, one thread per strike.  Everything that does not depend on the strike is
// tabulated once per group in shared memory: sqrt(H(u_k)) for every recorded iterate (Clenshaw on the stored Chebyshev
// coefficients), the discounted weights and the d+- coefficients of the m quadrature points.  Per strike and point the
// work left is two FMAs and two Phi evaluations.
#include "faop_alo.cuh"
#include "faop_fastmath.cuh"
#include "faop_kernels.h"

namespace faop {

constexpr int kSurfThreads = 256;

FAOP_D double lin_at(int i, int n, double lo, double hi, double step) {   // numpy.linspace arithmetic
    return (i == n - 1 && n > 1) ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);
}

// Group parameters (K = 1) for the seed and boundary kernels: group g = iT * nV + iV.
__global__ void __launch_bounds__(256) surface_groups_kernel(SurfArgs a, double* r, double* q, double* sigma, double* K,
                                                             double* T, uint8_t* is_put) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.g_cnt) return;
    const int g = a.g_lo + j;
    const int iV = g % a.s.nV, iT = g / a.s.nV;
    r[j] = a.s.r;
    q[j] = a.s.q;
    K[j] = 1.0;
    T[j] = lin_at(iT, a.s.nT, a.s.T0, a.s.T1, a.stepT);
    sigma[j] = lin_at(iV, a.s.nV, a.s.V0, a.s.V1, a.stepV);
    is_put[j] = (uint8_t)(a.s.is_put != 0);
}

// Phi(y0), Phi(y1) from E0 = exp(-y0^2/2), E1 = exp(-y1^2/2): Phi(y) = E G(|y|/sqrt2) for y <= 0, 1 - E G otherwise,
// G = erfcx/2 (full-accuracy fit of faop_fastmath.cuh).
FAOP_D void phi_pair(double y0, double y1, double E0, double E1, double& P0, double& P1) {
    double xa[2] = {fabs(y0) * kInvSqrt2, fabs(y1) * kInvSqrt2}, G[2];
    half_erfcx_n<2, false>(xa, G);
    const double c0 = E0 * G[0], c1 = E1 * G[1];
    P0 = (y0 <= 0.0) ? c0 : 1.0 - c0;
    P1 = (y1 <= 0.0) ? c1 : 1.0 - c1;
}

// LADDER = false: surface mode, group g = iT * nV + iV decoded from the faop_surface, strikes linspace(K0, K1, nK), output
// in the surface's flat order.  LADDER = true: G explicit scenarios (r, q, sigma, T, S, type) x one common strike ladder,
// output price[g * nK + k].
template <bool LADDER>
__global__ void __launch_bounds__(kSurfThreads) surface_price_kernel(SurfArgs a) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, MP = tl.MP, cap = a.snap_cap;
    double* s_A = smem;            // w h tau r e^{-r tau h^2}        (times K per strike)
    double* s_Bq = s_A + MP;       // w h tau q S e^{-q tau h^2}
    double* s_inv = s_Bq + MP;     // 1 / (sigma sqrt(tau) h)
    double* s_apch = s_inv + MP;   // ((r-q) + sigma^2/2) sqrt(tau) h / sigma
    double* s_ssuh = s_apch + MP;  // sigma sqrt(tau) h
    double* s_e = s_ssuh + MP;     // [cap] recorded errors
    double* s_cnt = s_e + cap;     // [cap] sweep counts
    double* s_root = s_cnt + cap;  // [cap][MP] sqrt(max(0, H(u_k)))
    double* s_T = s_root + cap * MP;   // [cap][MP] exp(om sqrt(H(u_k)) + (r-q) tau h_k^2) = (X/B(u_k))^{-om} e^{(r-q)(tau-u_k)}
    const double* g_wm = a.tables + tl.off_wm;
    const double* g_hm = a.tables + tl.off_hm;
    const double* g_sgm = a.tables + tl.off_sgm;
    const int64_t rowlen = (int64_t)a.s.nT * a.s.nV;

    for (int j = blockIdx.x; j < a.g_cnt; j += gridDim.x) {
        double r, q, S, T, sg;
        bool put;
        int g = 0;
        if (LADDER) {
            r = a.g_r[j]; q = a.g_q[j]; S = a.g_S[j]; T = a.g_T[j]; sg = a.g_sigma[j];
            put = a.g_put[j] != 0;
        } else {
            g = a.g_lo + j;
            const int iV = g % a.s.nV, iT = g / a.s.nV;
            r = a.s.r; q = a.s.q; S = a.s.S;
            put = a.s.is_put != 0;
            T = lin_at(iT, a.s.nT, a.s.T0, a.s.T1, a.stepT);
            sg = lin_at(iV, a.s.nV, a.s.V0, a.s.V1, a.stepV);
        }
        const double om = put ? 1.0 : -1.0;
        const double x0 = b_at_zero(r, q, 1.0, put);
        const int ns = a.nsnap[j];
        const double* snap = a.snap + (int64_t)j * cap * (n + 3);
        const double sqt = sqrt(T), ssu = sg * sqt, cinv = 1.0 / ssu, apc = ((r - q) + 0.5 * sg * sg) * sqt / sg;
        __syncthreads();   // previous group's tables are no longer read
        for (int k = threadIdx.x; k < MP; k += kSurfThreads) {
            const double h = g_hm[k], wh = g_wm[k] * h * T, h2 = h * h;
            s_A[k] = wh * r * exp(-r * T * h2);
            s_Bq[k] = wh * q * S * exp(-q * T * h2);
            s_inv[k] = cinv / h;
            s_apch[k] = apc * h;
            s_ssuh[k] = ssu * h;
        }
        for (int e = threadIdx.x; e < ns; e += kSurfThreads) {
            s_e[e] = snap[e * (n + 3)];
            s_cnt[e] = snap[e * (n + 3) + 1];
        }
        for (int e = threadIdx.x; e < ns * MP; e += kSurfThreads) {
            const int sidx = e / MP, k = e - sidx * MP;
            // z = sqrt(4 u / T) - 1 with u = T g_k (t = 0): 2 sqrt(g_k) - 1 (to_cheby_point, Base.py:235-237)
            const double z = fma(2.0, g_sgm[k], -1.0);
            const double Hu = clenshaw(n, snap + sidx * (n + 3) + 2, z);
            const double root = sqrt(Hu < 0.0 ? 0.0 : Hu);    // NaN propagates
            s_root[e] = root;
            s_T[e] = exp(fma(om, root, (r - q) * T * g_hm[k] * g_hm[k]));
        }
        __syncthreads();
        const int k_lo = LADDER ? 0 : a.iK_lo, k_hi = LADDER ? a.nK - 1 : a.iK_hi;
        for (int iK = k_lo + threadIdx.x; iK <= k_hi; iK += kSurfThreads) {
            int64_t at;
            double K;
            if (LADDER) {
                at = (int64_t)j * a.nK + iK;
                K = a.ladder[iK];
            } else {
                const int64_t flat = (int64_t)iK * rowlen + g;
                if (flat < a.first || flat >= a.first + a.count) continue;
                at = flat - a.first;
                K = lin_at(iK, a.s.nK, a.s.K0, a.s.K1, a.stepK);
            }
            const double eur = euro_value(T, S, r, q, sg, K, put);
            double v = eur;
            int its = 0;
            if (ns >= 0 && T != 0) {       // ns < 0: European shortcut; tau == 0: quadrature_sum returns 0
                int sidx = ns - 1;
                for (int e = 0; e < ns - 1; ++e)
                    if (s_e[e] * K <= a.cfg.iter_tol) { sidx = e; break; }   // first recorded iterate the strike accepts
                its = (int)s_cnt[sidx];
                const double* root = s_root + sidx * MP;
                const double* Tk = s_T + sidx * MP;
                const double SX = S / (K * x0);
                const double LS = log(SX);
                double accA = 0.0, accB = 0.0;
                for (int k = 0; k < MP; ++k) {
                    const double lnz = fma(om, root[k], LS);
                    const double dp = fma(lnz, s_inv[k], s_apch[k]);
                    const double dm = dp - s_ssuh[k];
                    // one exponential per point: phi(d-)/phi(d+) = exp(d+ s - s^2/2) with s = sigma sqrt(tau) h, and
                    // d+ s - s^2/2 = ln(S/B(u)) + (r-q) tau h^2, whose strike-independent factor is tabulated (s_T)
                    double ex[1] = {-0.5 * dp * dp}, Ep[1];
                    fast_exp_n<1, false>(ex, Ep);
                    const double Em = Ep[0] * (SX * Tk[k]);
                    double Pm, Pp;
                    phi_pair(-om * dm, -om * dp, Em, Ep[0], Pm, Pp);
                    accA = fma(s_A[k], Pm, accA);
                    accB = fma(s_Bq[k], Pp, accB);
                }
                v = eur + om * (K * accA - accB);      // v_integrand_12, Base.py:211-219
            } else if (ns >= 0) {
                its = (int)s_cnt[ns - 1];
            }
            a.price[at] = v;
            if (a.iters) a.iters[at] = its;
        }
    }
}

__global__ void __launch_bounds__(256) fill_kernel(double* dst, double value, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = value;
}
int launch_fill(double* dst, double value, int64_t count, cudaStream_t st) {
    if (count == 0) return 0;
    fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(dst, value, count);
    count_launch();
    return launch_status();
}

int launch_surface_groups(const SurfArgs& a, double* r, double* q, double* sigma, double* K, double* T, uint8_t* is_put,
                          cudaStream_t st) {
    if (a.g_cnt == 0) return 0;
    surface_groups_kernel<<<(a.g_cnt + 255) / 256, 256, 0, st>>>(a, r, q, sigma, K, T, is_put);
    count_launch();
    return launch_status();
}

int launch_surface_price(const SurfArgs& a, cudaStream_t st) {
    if (a.g_cnt == 0 || (a.ladder ? a.nK == 0 : a.count == 0)) return 0;
    const size_t sm = sizeof(double) * (size_t)(5 * a.tl.MP + 2 * a.snap_cap + 2 * a.snap_cap * a.tl.MP);
    const int grid = a.g_cnt < kSMs * 8 ? a.g_cnt : kSMs * 8;
    if (a.ladder) {
        if (sm > 48 * 1024)
            FAOP_CUDA_OK(cudaFuncSetAttribute(surface_price_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        surface_price_kernel<true><<<grid, kSurfThreads, sm, st>>>(a);
    } else {
        if (sm > 48 * 1024)
            FAOP_CUDA_OK(cudaFuncSetAttribute(surface_price_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        surface_price_kernel<false><<<grid, kSurfThreads, sm, st>>>(a);
    }
    count_launch();
    return launch_status();
}

}  // namespace faop
