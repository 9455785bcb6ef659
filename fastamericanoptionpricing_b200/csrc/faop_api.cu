// faop_api.cu — the C ABI of libfaop.so (include/faop.h): argument checks, table construction,
// kernel dispatch, and the host-buffer convenience path.  No compute happens on the host except the
// one-time construction of the (option-independent) quadrature/collocation tables.
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "faop_kernels.h"

namespace faop {
unsigned long long g_launches = 0;

static int check_cfg(const faop_cfg* cfg) {
    if (!cfg) return FAOP_EINVAL;
    if (cfg->solver < 0 || cfg->solver > 1) return FAOP_EINVAL;
    if (cfg->n_colloc < 2 || cfg->n_colloc > FAOP_MAX_COLLOC) return FAOP_ERANGE;
    if (cfg->n_quad < 1 || cfg->n_quad > FAOP_MAX_QUAD) return FAOP_ERANGE;
    if (cfg->n_price_quad < 1 || cfg->n_price_quad > FAOP_MAX_QUAD) return FAOP_ERANGE;
    if (cfg->max_iters < 0) return FAOP_EINVAL;
    return 0;
}

static void fill_rule(double* w, double* h, double* rh, double* sg, const double* y, const double* wt, int cnt, int padded) {
    for (int k = 0; k < padded; ++k) {
        double yk = k < cnt ? y[k] : 0.0;
        double hk = (1.0 + yk) * 0.5;
        w[k] = k < cnt ? wt[k] : 0.0;
        h[k] = hk;
        rh[k] = 1.0 / hk;
        sg[k] = sqrt(0.5 * (1.0 - yk) * (1.0 + hk));  // sqrt(1 - h^2) without cancellation near y = 1
    }
}
}  // namespace faop

using namespace faop;

extern "C" {

int faop_version(void) { return FAOP_VERSION; }
int64_t faop_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int faop_tables_bytes(const faop_cfg* cfg, size_t* bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bytes) return FAOP_EINVAL;
    TableLayout tl = table_layout(cfg->n_colloc, cfg->n_quad, cfg->n_price_quad);
    *bytes = sizeof(double) * (size_t)tl.total;
    return 0;
}

int faop_tables_build(const faop_cfg* cfg, const double* y_l, const double* w_l, const double* y_m, const double* w_m,
                      void* host_out) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!y_l || !w_l || !y_m || !w_m || !host_out) return FAOP_EINVAL;
    const int n = cfg->n_colloc;
    TableLayout tl = table_layout(n, cfg->n_quad, cfg->n_price_quad);
    double* t = (double*)host_out;
    memset(t, 0, sizeof(double) * (size_t)tl.total);
    fill_rule(t + tl.off_wl, t + tl.off_hl, t + tl.off_rhl, t + tl.off_sgl, y_l, w_l, tl.l, tl.LP);
    fill_rule(t + tl.off_wm, t + tl.off_hm, t + tl.off_rhm, t + tl.off_sgm, y_m, w_m, tl.m, tl.MP);
    for (int i = 0; i <= n; ++i) t[tl.off_c + i] = cos(i * M_PI / n);  // ChebyshevInterpolation.py:15-17
    for (int j = 0; j < 2 * n; ++j) t[tl.off_cos + j] = cos(j * M_PI / n);
    t[tl.off_cos + n] = -1.0;
    if (tl.has_matrix) {
        // M[i][j][k] = value at z_ik = (1+c_i) sqrt(g_k) - 1 of the Chebyshev interpolant of the unit vector e_j
        // (DCT-I fit ChebyshevInterpolation.py:43-51 followed by Clenshaw :31-41), in long double.
        std::vector<long double> a(n + 1);
        for (int j = 0; j <= n; ++j) {
            for (int kk = 0; kk <= n; ++kk) {
                long double v = cosl((long double)((j * kk) % (2 * n)) * (long double)M_PIl / n);
                if (j == 0 || j == n) v *= 0.5L;
                a[kk] = v * (2.0L / n);
            }
            for (int i = 0; i < n; ++i) {
                for (int k = 0; k < tl.LP; ++k) {
                    long double z = (1.0L + (long double)t[tl.off_c + i]) * (long double)t[tl.off_sgl + k] - 1.0L;
                    long double b0 = a[n] * 0.5L, b1 = 0, b2 = 0;
                    for (int kk = n - 1; kk >= 0; --kk) {
                        b2 = b1;
                        b1 = b0;
                        b0 = a[kk] + 2 * z * b1 - b2;
                    }
                    t[tl.off_M + ((size_t)i * (n + 1) + j) * tl.LP + k] = (double)(0.5L * (b0 - b2));
                }
            }
        }
    }
    return 0;
}

// Gauss-Legendre rule for callers without numpy: Newton iteration on P_n in long double from the Chebyshev-like initial
// guess cos(pi (k + 3/4) / (n + 1/2)), nodes ascending as numpy.polynomial.legendre.leggauss returns them.  Agrees with
// numpy's eigenvalue-based rule to ~1e-16; callers that need the reference's exact bits pass numpy's arrays instead.
int faop_gauss_legendre_host(int32_t n, double* y, double* w) {
    if (n < 1 || n > FAOP_MAX_QUAD || !y || !w) return FAOP_EINVAL;
    for (int k = 0; k < (n + 1) / 2; ++k) {
        long double x = cosl(M_PIl * (k + 0.75L) / (n + 0.5L)), dp = 1.0L;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1.0L, p1 = x;
            for (int j = 2; j <= n; ++j) {
                long double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
                p0 = p1;
                p1 = p2;
            }
            if (n == 1) { p0 = 1.0L; p1 = x; }
            dp = n * (x * p1 - p0) / (x * x - 1.0L);
            long double dx = p1 / dp;
            x -= dx;
            if (fabsl(dx) < 1e-19L) break;
        }
        {   // derivative at the converged node for the weight
            long double p0 = 1.0L, p1 = x;
            for (int j = 2; j <= n; ++j) {
                long double p2 = ((2 * j - 1) * x * p1 - (j - 1) * p0) / j;
                p0 = p1;
                p1 = p2;
            }
            dp = n * (x * p1 - p0) / (x * x - 1.0L);
        }
        long double wt = 2.0L / ((1.0L - x * x) * dp * dp);
        y[k] = (double)(-x); y[n - 1 - k] = (double)x;
        w[k] = (double)wt; w[n - 1 - k] = (double)wt;
    }
    if (n & 1) y[n / 2] = 0.0;
    return 0;
}

int faop_workspace_bytes(const faop_cfg* cfg, int64_t N, size_t* bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bytes || N < 0) return FAOP_EINVAL;
    size_t b = sizeof(double) * (size_t)N * (size_t)(cfg->n_colloc + 1);
    b = (b + 255) & ~(size_t)255;
    if (cfg->flags & FAOP_FLAG_WS_SCRATCH) b += alo_clen_scratch_bytes(*cfg);
    *bytes = b;
    return 0;
}

static int fill_args(AloArgs& a, const faop_cfg* cfg, int64_t N, const double* r, const double* q, const double* sigma,
                     const double* K, const double* T, const double* t, const double* S, const uint8_t* is_put,
                     const void* tables_dev) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (N < 0) return FAOP_EINVAL;
    if (N > 0 && (!r || !q || !sigma || !K || !T || !is_put || !tables_dev)) return FAOP_EINVAL;
    memset(&a, 0, sizeof a);
    a.cfg = *cfg;
    a.tl = table_layout(cfg->n_colloc, cfg->n_quad, cfg->n_price_quad);
    a.N = N;
    a.r = r; a.q = q; a.sigma = sigma; a.K = K; a.T = T; a.t = t; a.S = S; a.is_put = is_put;
    a.tau_max = cfg->tau_max;
    a.tables = (const double*)tables_dev;
    return 0;
}

// seeds_ws: where the QD seeds go when B0_out is NULL; inv_scratch: nullable per-warp invariant-table scratch (faop_alo_clen.cu)
static int price_batch_impl(const faop_cfg* cfg, int64_t N, const double* r, const double* q, const double* sigma,
                            const double* K, const double* T, const double* t, const double* S, const uint8_t* is_put,
                            const void* tables_dev, const double* B0_in, double* price, double* european, double* B,
                            double* B0_out, int32_t* iters, double* err, double* iter_errs, double* seeds_ws,
                            double* inv_scratch, cudaStream_t st) {
    AloArgs a;
    int rc = fill_args(a, cfg, N, r, q, sigma, K, T, t, S, is_put, tables_dev);
    if (rc) return rc;
    if (N == 0) return 0;
    if (!S || !price) return FAOP_EINVAL;
    a.price = price; a.european = european; a.B = B; a.iters = iters; a.err = err; a.iter_errs = iter_errs;
    a.inv_scratch = inv_scratch;
    const size_t seed_bytes = sizeof(double) * (size_t)N * (size_t)(cfg->n_colloc + 1);
    if (B0_in) {
        a.seeds_in = B0_in;
        if (B0_out && B0_out != B0_in) FAOP_CUDA_OK(cudaMemcpyAsync(B0_out, B0_in, seed_bytes, cudaMemcpyDeviceToDevice, st));
    } else {
        a.seeds_out = B0_out ? B0_out : seeds_ws;
        if (!a.seeds_out) return FAOP_EINVAL;
        rc = launch_qd_seed(a, st);
        if (rc) return rc;
        a.seeds_in = a.seeds_out;
    }
    if (cfg->kernel != 1) {
        rc = launch_alo_fast(a, st);        // records the per-iteration error history too
        if (rc != -1000) return rc;
        if (!iter_errs) {                   // the Clenshaw kernels do not; the generic kernel serves those requests
            rc = launch_alo_clen(a, st);
            if (rc != -1000) return rc;
        }
    }
    return launch_alo_generic(a, st);
}

int faop_alo_price_batch(const faop_cfg* cfg, int64_t N, const double* r, const double* q, const double* sigma,
                         const double* K, const double* T, const double* t, const double* S, const uint8_t* is_put,
                         const void* tables_dev, const double* B0_in, double* price, double* european, double* B,
                         double* B0_out, int32_t* iters, double* err, double* iter_errs, void* workspace,
                         void* cuda_stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    double* scratch = nullptr;
    if (workspace && N > 0 && (cfg->flags & FAOP_FLAG_WS_SCRATCH) && alo_clen_scratch_bytes(*cfg)) {
        // layout: [seeds, only when they live in the workspace][scratch]
        const size_t seed_bytes = (B0_in || B0_out) ? 0 : sizeof(double) * (size_t)N * (size_t)(cfg->n_colloc + 1);
        scratch = (double*)((char*)workspace + ((seed_bytes + 255) & ~(size_t)255));
    }
    return price_batch_impl(cfg, N, r, q, sigma, K, T, t, S, is_put, tables_dev, B0_in, price, european, B, B0_out, iters, err,
                            iter_errs, (double*)workspace, scratch, (cudaStream_t)cuda_stream);
}

int faop_alo_surface_batch(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count,
                           const void* tables_dev, double* price, int32_t* iters, void* workspace,
                           size_t workspace_bytes, void* cuda_stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (cfg->tau_max) return FAOP_ERANGE;     // the surface is generated on the device: collocation domain = maturity
    if (!surf || first < 0 || count < 0 || surf->nK < 1 || surf->nT < 1 || surf->nV < 1) return FAOP_EINVAL;
    if (first + count > (int64_t)surf->nK * surf->nT * surf->nV) return FAOP_EINVAL;
    if (count == 0) return 0;
    if (!price || !workspace || !tables_dev) return FAOP_EINVAL;
    // per-option scratch: 6 parameter doubles + 8 B flag slot + the seed row
    const size_t per = sizeof(double) * (size_t)(7 + cfg->n_colloc + 1);
    int64_t chunk = (int64_t)(workspace_bytes / per) & ~(int64_t)255;
    if (chunk < 256) return FAOP_EINVAL;
    if (chunk > count) chunk = (count + 255) & ~(int64_t)255;
    double* w = (double*)workspace;
    double *d_r = w, *d_q = w + chunk, *d_s = w + 2 * chunk, *d_K = w + 3 * chunk, *d_T = w + 4 * chunk, *d_S = w + 5 * chunk;
    uint8_t* d_put = (uint8_t*)(w + 6 * chunk);
    double* d_seed = w + 7 * chunk;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    faop_cfg plain = *cfg;
    plain.flags &= ~FAOP_FLAG_WS_SCRATCH;      // the chunk workspace below holds the seeds only
    for (int64_t off = 0; off < count; off += chunk) {
        int64_t cnt = count - off < chunk ? count - off : chunk;
        rc = launch_surface_decode(*surf, first + off, cnt, d_r, d_q, d_s, d_K, d_T, d_S, d_put, st);
        if (rc) return rc;
        rc = faop_alo_price_batch(&plain, cnt, d_r, d_q, d_s, d_K, d_T, nullptr, d_S, d_put, tables_dev, nullptr, price + off,
                                  nullptr, nullptr, nullptr, iters ? iters + off : nullptr, nullptr, nullptr, d_seed, st);
        if (rc) return rc;
    }
    return 0;
}

// ---- deduplicated surface sweep (faop_surface.cu) ------------------------------------------------------------------
static int surf_snap_cap(const faop_cfg* cfg) { return cfg->max_iters + 1 < 48 ? cfg->max_iters + 1 : 48; }

static int surf_plan(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count, SurfArgs& a) {
    memset(&a, 0, sizeof a);
    a.cfg = *cfg;
    a.tl = table_layout(cfg->n_colloc, cfg->n_quad, cfg->n_price_quad);
    a.s = *surf;
    a.stepK = surf->nK > 1 ? (surf->K1 - surf->K0) / (surf->nK - 1) : 0.0;
    a.stepT = surf->nT > 1 ? (surf->T1 - surf->T0) / (surf->nT - 1) : 0.0;
    a.stepV = surf->nV > 1 ? (surf->V1 - surf->V0) / (surf->nV - 1) : 0.0;
    a.first = first;
    a.count = count;
    const int64_t G = (int64_t)surf->nT * surf->nV;
    a.iK_lo = (int)(first / G);
    a.iK_hi = (int)((first + (count > 0 ? count - 1 : 0)) / G);
    if (a.iK_hi == a.iK_lo && count > 0) {   // inside one strike row: only the groups the range touches
        a.g_lo = (int)(first % G);
        a.g_cnt = (int)count;
    } else {
        a.g_lo = 0;
        a.g_cnt = (int)G;
    }
    a.snap_cap = surf_snap_cap(cfg);
    return 0;
}

static size_t surf_ws_bytes(const SurfArgs& a) {
    const size_t G = (size_t)a.g_cnt, nn = (size_t)a.tl.n + 1;
    // r q sigma K T | is_put (one 8-byte slot each) | seeds | snapshot rows | row counts
    return sizeof(double) * (G * 6 + G * nn + G * (size_t)a.snap_cap * (nn + 2)) + sizeof(int32_t) * ((G + 1) & ~(size_t)1);
}

int faop_alo_surface_dedup_workspace_bytes(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count,
                                           size_t* bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!surf || !bytes || first < 0 || count < 0 || surf->nK < 1 || surf->nT < 1 || surf->nV < 1) return FAOP_EINVAL;
    if (first + count > (int64_t)surf->nK * surf->nT * surf->nV) return FAOP_EINVAL;
    SurfArgs a;
    surf_plan(cfg, surf, first, count, a);
    *bytes = (surf_ws_bytes(a) + 255) & ~(size_t)255;
    return 0;
}

int faop_alo_surface_dedup_batch(const faop_cfg* cfg, const faop_surface* surf, int64_t first, int64_t count,
                                 const void* tables_dev, double* price, int32_t* iters, void* workspace,
                                 size_t workspace_bytes, void* cuda_stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (cfg->tau_max) return FAOP_ERANGE;     // the surface is generated on the device: collocation domain = maturity
    if (!surf || first < 0 || count < 0 || surf->nK < 1 || surf->nT < 1 || surf->nV < 1) return FAOP_EINVAL;
    if (first + count > (int64_t)surf->nK * surf->nT * surf->nV) return FAOP_EINVAL;
    if (cfg->use_derivative) return FAOP_ERANGE;    // f' = N'/D - D'N/D^2 is not scale-free in the reference's form
    if (count == 0) return 0;
    if (!price || !workspace || !tables_dev) return FAOP_EINVAL;
    SurfArgs sa;
    surf_plan(cfg, surf, first, count, sa);
    if (workspace_bytes < surf_ws_bytes(sa)) return FAOP_EINVAL;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t G = (size_t)sa.g_cnt, nn = (size_t)sa.tl.n + 1;
    double* w = (double*)workspace;
    double *d_r = w, *d_q = w + G, *d_s = w + 2 * G, *d_K = w + 3 * G, *d_T = w + 4 * G;
    uint8_t* d_put = (uint8_t*)(w + 5 * G);
    double* d_seed = w + 6 * G;
    double* d_snap = d_seed + G * nn;
    int32_t* d_ns = (int32_t*)(d_snap + G * (size_t)sa.snap_cap * (nn + 2));
    sa.tables = (const double*)tables_dev;
    sa.snap = d_snap;
    sa.nsnap = d_ns;
    sa.price = price;
    sa.iters = iters;
    rc = launch_surface_groups(sa, d_r, d_q, d_s, d_K, d_T, d_put, st);
    if (rc) return rc;
    // one boundary solve per group at K = 1, stopping at the tightest threshold any strike of the range has
    double Klo = surf->K0 + sa.stepK * sa.iK_lo, Khi = surf->K0 + sa.stepK * sa.iK_hi;
    if (sa.iK_hi == surf->nK - 1 && surf->nK > 1) Khi = surf->K1;
    if (sa.iK_lo == surf->nK - 1 && surf->nK > 1) Klo = surf->K1;
    const double Kmin = Klo < Khi ? Klo : Khi, Kmax = Klo < Khi ? Khi : Klo;
    AloArgs a;
    rc = fill_args(a, cfg, (int64_t)G, d_r, d_q, d_s, d_K, d_T, nullptr, nullptr, d_put, tables_dev);
    if (rc) return rc;
    a.cfg.iter_tol = cfg->iter_tol / Kmax * (1.0 - 1e-12);     // never above any strike's own tol / K
    a.snap_tol_hi = cfg->iter_tol / Kmin * (1.0 + 1e-12);      // never below
    a.snap = d_snap;
    a.nsnap = d_ns;
    a.snap_cap = sa.snap_cap;
    a.seeds_out = d_seed;
    rc = launch_qd_seed(a, st);
    if (rc) return rc;
    a.seeds_in = d_seed;
    rc = cfg->kernel != 1 ? launch_alo_fast(a, st) : -1000;
    if (rc == -1000) rc = launch_alo_generic(a, st);
    if (rc) return rc;
    return launch_surface_price(sa, st);
}

// Strike ladder: G scenarios x nK strikes, one boundary solve per scenario (same machinery as the deduplicated surface).
static size_t ladder_ws_bytes(const faop_cfg* cfg, int64_t G) {
    const size_t nn = (size_t)cfg->n_colloc + 1, cap = (size_t)surf_snap_cap(cfg);
    return sizeof(double) * ((size_t)G * (1 + nn + cap * (nn + 2))) + sizeof(int32_t) * (((size_t)G + 1) & ~(size_t)1);
}
int faop_alo_ladder_workspace_bytes(const faop_cfg* cfg, int64_t G, size_t* bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!bytes || G < 0) return FAOP_EINVAL;
    *bytes = (ladder_ws_bytes(cfg, G) + 255) & ~(size_t)255;
    return 0;
}
int faop_alo_ladder_batch(const faop_cfg* cfg, int64_t G, const double* r, const double* q, const double* sigma,
                          const double* T, const double* S, const uint8_t* is_put, int32_t nK, const double* K,
                          double K_min, double K_max, const void* tables_dev, double* price, int32_t* iters,
                          void* workspace, size_t workspace_bytes, void* cuda_stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (G < 0 || nK < 0 || G > 0x7fffffff) return FAOP_EINVAL;
    if (cfg->use_derivative) return FAOP_ERANGE;
    if (cfg->tau_max) return FAOP_ERANGE;     // the ladder's price kernel maps u with the maturity
    if (G == 0 || nK == 0) return 0;
    if (!r || !q || !sigma || !T || !S || !is_put || !K || !price || !workspace || !tables_dev) return FAOP_EINVAL;
    if (!(K_min > 0) || !(K_max >= K_min)) return FAOP_EINVAL;
    if (workspace_bytes < ladder_ws_bytes(cfg, G)) return FAOP_EINVAL;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t nn = (size_t)cfg->n_colloc + 1;
    SurfArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.cfg = *cfg;
    sa.tl = table_layout(cfg->n_colloc, cfg->n_quad, cfg->n_price_quad);
    sa.g_cnt = (int)G;
    sa.snap_cap = surf_snap_cap(cfg);
    double* w = (double*)workspace;
    double* d_K1 = w;                       // the unit strike of the normalised solves
    double* d_seed = w + G;
    double* d_snap = d_seed + (size_t)G * nn;
    int32_t* d_ns = (int32_t*)(d_snap + (size_t)G * (size_t)sa.snap_cap * (nn + 2));
    sa.tables = (const double*)tables_dev;
    sa.snap = d_snap; sa.nsnap = d_ns; sa.price = price; sa.iters = iters;
    sa.g_r = r; sa.g_q = q; sa.g_sigma = sigma; sa.g_T = T; sa.g_S = S; sa.g_put = is_put; sa.ladder = K; sa.nK = nK;
    rc = launch_fill(d_K1, 1.0, G, st);
    if (rc) return rc;
    AloArgs a;
    rc = fill_args(a, cfg, G, r, q, sigma, d_K1, T, nullptr, nullptr, is_put, tables_dev);
    if (rc) return rc;
    a.cfg.iter_tol = cfg->iter_tol / K_max * (1.0 - 1e-12);
    a.snap_tol_hi = cfg->iter_tol / K_min * (1.0 + 1e-12);
    a.snap = d_snap; a.nsnap = d_ns; a.snap_cap = sa.snap_cap;
    a.seeds_out = d_seed;
    rc = launch_qd_seed(a, st);
    if (rc) return rc;
    a.seeds_in = d_seed;
    rc = cfg->kernel != 1 ? launch_alo_fast(a, st) : -1000;
    if (rc == -1000) rc = launch_alo_generic(a, st);
    if (rc) return rc;
    return launch_surface_price(sa, st);
}

int faop_alo_eval_points_batch(const faop_cfg* cfg, int64_t N, int32_t npts, const double* r, const double* q,
                               const double* sigma, const double* K, const double* T, const uint8_t* is_put,
                               const void* tables_dev, const double* B, const double* pts_tau, const double* pts_B,
                               double* Nv, double* Dv, double* Np, double* Dp, double* f, double* fp, double* Bnew,
                               double* I1, double* I2, void* cuda_stream) {
    AloArgs a;
    int rc = fill_args(a, cfg, N, r, q, sigma, K, T, nullptr, nullptr, is_put, tables_dev);
    if (rc) return rc;
    if (npts < 0) return FAOP_EINVAL;
    if (N == 0 || npts == 0) return 0;
    if (!B || !pts_tau || !pts_B) return FAOP_EINVAL;
    a.seeds_in = B;
    a.npts = npts; a.pts_tau = pts_tau; a.pts_B = pts_B;
    EvalOut eo = {Nv, Dv, Np, Dp, f, fp, Bnew, I1, I2};
    return launch_alo_eval_nodes(a, eo, (cudaStream_t)cuda_stream);
}

int faop_alo_interp_boundary_batch(const faop_cfg* cfg, int64_t N, int32_t npts, const double* r, const double* q,
                                   const double* K, const double* T, const uint8_t* is_put, const void* tables_dev,
                                   const double* B, const double* u, double* Bu, void* cuda_stream) {
    AloArgs a;
    int rc = fill_args(a, cfg, N, r, q, r /* sigma unused */, K, T, nullptr, nullptr, is_put, tables_dev);
    if (rc) return rc;
    if (npts < 0) return FAOP_EINVAL;
    if (N == 0 || npts == 0) return 0;
    if (!B || !u || !Bu) return FAOP_EINVAL;
    a.seeds_in = B;
    a.npts = npts; a.pts_tau = u;
    return launch_alo_interp(a, Bu, (cudaStream_t)cuda_stream);
}

int faop_alo_price_known_boundary_batch(const faop_cfg* cfg, int64_t N, const double* r, const double* q,
                                        const double* sigma, const double* K, const double* T, const double* t,
                                        const double* S, const uint8_t* is_put, const void* tables_dev, const double* B,
                                        double* price, double* european, void* cuda_stream) {
    AloArgs a;
    int rc = fill_args(a, cfg, N, r, q, sigma, K, T, t, S, is_put, tables_dev);
    if (rc) return rc;
    if (N == 0) return 0;
    if (!B || !S || !price) return FAOP_EINVAL;
    a.seeds_in = B;
    a.price = price;
    a.european = european;
    return launch_alo_price_known(a, (cudaStream_t)cuda_stream);
}

int faop_alo_greeks_known_boundary_batch(const faop_cfg* cfg, int64_t N, const double* r, const double* q,
                                         const double* sigma, const double* K, const double* T, const double* t,
                                         const double* S, const uint8_t* is_put, const void* tables_dev, const double* B,
                                         double bump_S_rel, double bump_t, double* price, double* delta, double* gamma,
                                         double* theta, void* cuda_stream) {
    AloArgs a;
    int rc = fill_args(a, cfg, N, r, q, sigma, K, T, t, S, is_put, tables_dev);
    if (rc) return rc;
    if (!(bump_S_rel > 0) || !(bump_t > 0)) return FAOP_EINVAL;
    if (N == 0) return 0;
    if (!B || !S) return FAOP_EINVAL;
    a.seeds_in = B;
    a.price = price;
    return launch_alo_greeks_known(a, bump_S_rel, bump_t, delta, gamma, theta, (cudaStream_t)cuda_stream);
}

int faop_iv_update_batch(int64_t N, const double* target, const double* price, double* sigma, double* lo, double* hi,
                         double* flo, double* fhi, int32_t* side, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!target || !price || !sigma || !lo || !hi || !flo || !fhi || !side))) return FAOP_EINVAL;
    return launch_iv_update(N, target, price, sigma, lo, hi, flo, fhi, side, (cudaStream_t)cuda_stream);
}

int faop_qd_boundary_batch(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                           const double* tau, const uint8_t* is_put, double* B, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!r || !q || !sigma || !K || !tau || !is_put || !B))) return FAOP_EINVAL;
    return launch_qd_boundary(N, r, q, sigma, K, tau, is_put, B, (cudaStream_t)cuda_stream);
}
int faop_qd_residual_batch(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                           const double* tau, const double* S, const uint8_t* is_put, double* F, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!r || !q || !sigma || !K || !tau || !S || !is_put || !F))) return FAOP_EINVAL;
    return launch_qd_residual(N, r, q, sigma, K, tau, S, is_put, F, (cudaStream_t)cuda_stream);
}
int faop_qd_price_batch(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                        const double* tau, const double* S, const uint8_t* is_put, double* price, double* boundary,
                        void* cuda_stream) {
    if (N < 0 || (N > 0 && (!r || !q || !sigma || !K || !tau || !S || !is_put || !price))) return FAOP_EINVAL;
    return launch_qd_price(N, r, q, sigma, K, tau, S, is_put, price, boundary, (cudaStream_t)cuda_stream);
}
int faop_bjerksund_stensland_batch(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                                   const double* T, const double* S, const uint8_t* is_put, const double* Kb,
                                   const double* Tb, double* price, double* boundary, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!r || !q || !sigma || !K || !T || !S || !price))) return FAOP_EINVAL;
    return launch_bjerksund_stensland(N, r, q, sigma, K, T, S, is_put, Kb, Tb, price, boundary, (cudaStream_t)cuda_stream);
}
int faop_bs_phi_batch(int64_t N, const double* S, const double* T, const double* gamma, const double* H, const double* X,
                      const double* r, const double* q, const double* sigma, double* out, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!S || !T || !gamma || !H || !X || !r || !q || !sigma || !out))) return FAOP_EINVAL;
    return launch_bs_phi(N, S, T, gamma, H, X, r, q, sigma, out, (cudaStream_t)cuda_stream);
}
int faop_european_batch(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                        const double* sigma, const double* K, const uint8_t* is_put, double* price, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!tau || !S || !r || !q || !sigma || !K || !is_put || !price))) return FAOP_EINVAL;
    return launch_european(N, tau, S, r, q, sigma, K, is_put, price, (cudaStream_t)cuda_stream);
}
int faop_european_greeks_batch(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                               const double* sigma, const double* K, double* theta, double* d1, double* d2,
                               void* cuda_stream) {
    if (N < 0 || (N > 0 && (!tau || !S || !r || !q || !sigma || !K))) return FAOP_EINVAL;
    return launch_european_greeks(N, tau, S, r, q, sigma, K, theta, d1, d2, (cudaStream_t)cuda_stream);
}
int faop_cheb_fit_batch(int64_t N, int32_t n, const double* y, double* a, void* cuda_stream) {
    if (N < 0 || n < 1 || (N > 0 && (!y || !a))) return FAOP_EINVAL;
    return launch_cheb_fit(N, n, y, a, (cudaStream_t)cuda_stream);
}
int faop_cheb_eval_batch(int64_t N, int32_t n, int32_t npts, const double* a, const double* z, double* out,
                         void* cuda_stream) {
    if (N < 0 || n < 1 || npts < 0 || (N > 0 && npts > 0 && (!a || !z || !out))) return FAOP_EINVAL;
    return launch_cheb_eval(N, n, npts, a, z, out, (cudaStream_t)cuda_stream);
}

int faop_dev_math_batch(int32_t kind, int64_t N, const double* x, double* out, void* cuda_stream) {
    if (N < 0 || (N > 0 && (!x || !out))) return FAOP_EINVAL;
    return launch_dev_math(kind, N, x, out, (cudaStream_t)cuda_stream);
}

int faop_measure_point_mix_peak(int launches, double* points_per_s, double* ms_per_launch, void* cuda_stream) {
    if (launches < 1) return FAOP_EINVAL;
    return launch_point_mix_peak(launches, points_per_s, ms_per_launch, (cudaStream_t)cuda_stream);
}
int faop_measure_fp64_peak(int launches, double* tflops, double* ms_per_launch, void* cuda_stream) {
    if (launches < 1) return FAOP_EINVAL;
    return launch_fp64_peak(launches, tflops, ms_per_launch, (cudaStream_t)cuda_stream);
}

// ------------------------------------------------------------------------------------------ host-buffer path
struct faop_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t stream2;   // second lane of the chunk pipeline: copies of one chunk overlap the kernels of the other
    void* dev;          // one growing device arena
    size_t dev_bytes;
    void* tables_dev;   // cached table blob
    faop_cfg tables_cfg;
    size_t tables_bytes;
    std::vector<double> tables_key;  // the y/w values the cached blob was built from
};

int faop_ctx_create(int device, faop_ctx** out) {
    if (!out) return FAOP_EINVAL;
    FAOP_CUDA_OK(cudaSetDevice(device));
    faop_ctx* c = new (std::nothrow) faop_ctx();
    if (!c) return FAOP_ENOMEM;
    c->device = device;
    c->dev = nullptr; c->dev_bytes = 0; c->tables_dev = nullptr; c->tables_bytes = 0;
    memset(&c->tables_cfg, 0, sizeof c->tables_cfg);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return (int)e; }
    e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return (int)e; }
    *out = c;
    return 0;
}

int faop_ctx_destroy(faop_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->dev) cudaFree(c->dev);
    if (c->tables_dev) cudaFree(c->tables_dev);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->stream2);
    delete c;
    return 0;
}

static int ctx_tables(faop_ctx* c, const faop_cfg* cfg, const double* y_l, const double* w_l, const double* y_m,
                      const double* w_m) {
    std::vector<double> key;
    key.insert(key.end(), y_l, y_l + cfg->n_quad);
    key.insert(key.end(), w_l, w_l + cfg->n_quad);
    key.insert(key.end(), y_m, y_m + cfg->n_price_quad);
    key.insert(key.end(), w_m, w_m + cfg->n_price_quad);
    bool same = c->tables_dev && c->tables_cfg.n_colloc == cfg->n_colloc && c->tables_cfg.n_quad == cfg->n_quad &&
                c->tables_cfg.n_price_quad == cfg->n_price_quad && key.size() == c->tables_key.size() &&
                memcmp(key.data(), c->tables_key.data(), key.size() * sizeof(double)) == 0;
    if (same) return 0;
    size_t bytes = 0;
    int rc = faop_tables_bytes(cfg, &bytes);
    if (rc) return rc;
    std::vector<char> host(bytes);
    rc = faop_tables_build(cfg, y_l, w_l, y_m, w_m, host.data());
    if (rc) return rc;
    if (c->tables_dev) { cudaFree(c->tables_dev); c->tables_dev = nullptr; }
    FAOP_CUDA_OK(cudaMalloc(&c->tables_dev, bytes));
    FAOP_CUDA_OK(cudaMemcpy(c->tables_dev, host.data(), bytes, cudaMemcpyHostToDevice));
    c->tables_cfg = *cfg;
    c->tables_bytes = bytes;
    c->tables_key.swap(key);
    return 0;
}

int faop_alo_price_batch_host(faop_ctx* c, const faop_cfg* cfg, int64_t N, const double* r, const double* q,
                              const double* sigma, const double* K, const double* T, const double* t, const double* S,
                              const uint8_t* is_put, const double* y_l, const double* w_l, const double* y_m,
                              const double* w_m, const double* B0_in, double* price, double* european, double* B,
                              double* B0_out, int32_t* iters, double* err) {
    if (!c) return FAOP_EINVAL;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (N < 0) return FAOP_EINVAL;
    if (N == 0) return 0;
    if (!r || !q || !sigma || !K || !T || !S || !is_put || !price || !y_l || !w_l || !y_m || !w_m) return FAOP_EINVAL;
    FAOP_CUDA_OK(cudaSetDevice(c->device));
    rc = ctx_tables(c, cfg, y_l, w_l, y_m, w_m);
    if (rc) return rc;
    const size_t nn = (size_t)cfg->n_colloc + 1;
    const size_t vec = ((size_t)N * sizeof(double) + 255) & ~(size_t)255;
    const size_t mat = ((size_t)N * nn * sizeof(double) + 255) & ~(size_t)255;
    // arena: r q sigma K T t S | price european err | is_put iters | B0 B | invariant-table scratch of the two stream lanes
    const size_t scr = (alo_clen_scratch_bytes(*cfg) + 255) & ~(size_t)255;
    size_t need = 10 * vec + 2 * vec + 2 * mat + 2 * scr + (cfg->tau_max ? vec : 0);
    if (need > c->dev_bytes) {
        if (c->dev) { cudaFree(c->dev); c->dev = nullptr; c->dev_bytes = 0; }
        FAOP_CUDA_OK(cudaMalloc(&c->dev, need));
        c->dev_bytes = need;
    }
    char* p = (char*)c->dev;
    auto take = [&](size_t b) { char* q_ = p; p += b; return q_; };
    double* d_r = (double*)take(vec); double* d_q = (double*)take(vec); double* d_s = (double*)take(vec);
    double* d_K = (double*)take(vec); double* d_T = (double*)take(vec); double* d_t = (double*)take(vec);
    double* d_S = (double*)take(vec); double* d_price = (double*)take(vec); double* d_eur = (double*)take(vec);
    double* d_err = (double*)take(vec); uint8_t* d_put = (uint8_t*)take(vec); int32_t* d_it = (int32_t*)take(vec);
    double* d_B0 = (double*)take(mat); double* d_B = (double*)take(mat);
    double* d_scr[2] = {scr ? (double*)take(scr) : nullptr, scr ? (double*)take(scr) : nullptr};
    double* d_tmax = cfg->tau_max ? (double*)take(vec) : nullptr;     // cfg->tau_max is a HOST array here
    faop_cfg dcfg = *cfg;
    // Chunk pipeline over two streams: while one chunk's kernels run, the other's inputs go up and results come back
    // (asynchronous when the caller's arrays are pinned).  Chunks are slices of the same arena.
    int nchunks = (int)(N / 262144);
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 4) nchunks = 4;
    const int64_t per = ((N + nchunks - 1) / nchunks + 255) & ~(int64_t)255;
    int idx = 0;
    for (int64_t off = 0; off < N; off += per, ++idx) {
        const int64_t cnt = (N - off < per) ? N - off : per;
        cudaStream_t st = (idx & 1) ? c->stream2 : c->stream;
        const size_t vb = (size_t)cnt * sizeof(double), mb = (size_t)cnt * nn * sizeof(double);
        FAOP_CUDA_OK(cudaMemcpyAsync(d_r + off, r + off, vb, cudaMemcpyHostToDevice, st));
        FAOP_CUDA_OK(cudaMemcpyAsync(d_q + off, q + off, vb, cudaMemcpyHostToDevice, st));
        FAOP_CUDA_OK(cudaMemcpyAsync(d_s + off, sigma + off, vb, cudaMemcpyHostToDevice, st));
        FAOP_CUDA_OK(cudaMemcpyAsync(d_K + off, K + off, vb, cudaMemcpyHostToDevice, st));
        FAOP_CUDA_OK(cudaMemcpyAsync(d_T + off, T + off, vb, cudaMemcpyHostToDevice, st));
        if (t) FAOP_CUDA_OK(cudaMemcpyAsync(d_t + off, t + off, vb, cudaMemcpyHostToDevice, st));
        FAOP_CUDA_OK(cudaMemcpyAsync(d_S + off, S + off, vb, cudaMemcpyHostToDevice, st));
        FAOP_CUDA_OK(cudaMemcpyAsync(d_put + off, is_put + off, (size_t)cnt, cudaMemcpyHostToDevice, st));
        if (B0_in) FAOP_CUDA_OK(cudaMemcpyAsync(d_B0 + off * nn, B0_in + off * nn, mb, cudaMemcpyHostToDevice, st));
        if (d_tmax) {
            FAOP_CUDA_OK(cudaMemcpyAsync(d_tmax + off, cfg->tau_max + off, vb, cudaMemcpyHostToDevice, st));
            dcfg.tau_max = d_tmax + off;
        }
        rc = price_batch_impl(&dcfg, cnt, d_r + off, d_q + off, d_s + off, d_K + off, d_T + off, t ? d_t + off : nullptr,
                              d_S + off, d_put + off, c->tables_dev, B0_in ? d_B0 + off * nn : nullptr, d_price + off,
                              european ? d_eur + off : nullptr, B ? d_B + off * nn : nullptr,
                              B0_in ? nullptr : d_B0 + off * nn, iters ? d_it + off : nullptr, err ? d_err + off : nullptr,
                              nullptr, nullptr, d_scr[idx & 1], st);
        if (rc) return rc;
        FAOP_CUDA_OK(cudaMemcpyAsync(price + off, d_price + off, vb, cudaMemcpyDeviceToHost, st));
        if (european) FAOP_CUDA_OK(cudaMemcpyAsync(european + off, d_eur + off, vb, cudaMemcpyDeviceToHost, st));
        if (err) FAOP_CUDA_OK(cudaMemcpyAsync(err + off, d_err + off, vb, cudaMemcpyDeviceToHost, st));
        if (iters) FAOP_CUDA_OK(cudaMemcpyAsync(iters + off, d_it + off, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (B) FAOP_CUDA_OK(cudaMemcpyAsync(B + off * nn, d_B + off * nn, mb, cudaMemcpyDeviceToHost, st));
        if (B0_out) FAOP_CUDA_OK(cudaMemcpyAsync(B0_out + off * nn, d_B0 + off * nn, mb, cudaMemcpyDeviceToHost, st));
    }
    FAOP_CUDA_OK(cudaStreamSynchronize(c->stream));
    FAOP_CUDA_OK(cudaStreamSynchronize(c->stream2));
    return 0;
}

static int ctx_arena(faop_ctx* c, size_t need) {
    if (need > c->dev_bytes) {
        if (c->dev) { cudaFree(c->dev); c->dev = nullptr; c->dev_bytes = 0; }
        FAOP_CUDA_OK(cudaMalloc(&c->dev, need));
        c->dev_bytes = need;
    }
    return 0;
}

// Host-buffer form of faop_alo_ladder_batch: G scenarios x nK strikes, host arrays in, price[G * nK] (and iters) out.
int faop_alo_ladder_batch_host(faop_ctx* c, const faop_cfg* cfg, int64_t G, const double* r, const double* q,
                               const double* sigma, const double* T, const double* S, const uint8_t* is_put, int32_t nK,
                               const double* K, const double* y_l, const double* w_l, const double* y_m, const double* w_m,
                               double* price, int32_t* iters) {
    if (!c) return FAOP_EINVAL;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (G < 0 || nK < 0) return FAOP_EINVAL;
    if (G == 0 || nK == 0) return 0;
    if (!r || !q || !sigma || !T || !S || !is_put || !K || !price || !y_l || !w_l || !y_m || !w_m) return FAOP_EINVAL;
    double Kmin = K[0], Kmax = K[0];
    for (int32_t k = 1; k < nK; ++k) { Kmin = K[k] < Kmin ? K[k] : Kmin; Kmax = K[k] > Kmax ? K[k] : Kmax; }
    FAOP_CUDA_OK(cudaSetDevice(c->device));
    rc = ctx_tables(c, cfg, y_l, w_l, y_m, w_m);
    if (rc) return rc;
    size_t wsb = 0;
    rc = faop_alo_ladder_workspace_bytes(cfg, G, &wsb);
    if (rc) return rc;
    const size_t vec = ((size_t)G * sizeof(double) + 255) & ~(size_t)255, kb = ((size_t)nK * sizeof(double) + 255) & ~(size_t)255;
    const size_t outb = ((size_t)G * nK * sizeof(double) + 255) & ~(size_t)255;
    rc = ctx_arena(c, 6 * vec + kb + 2 * outb + wsb);
    if (rc) return rc;
    char* p = (char*)c->dev;
    auto take = [&](size_t b) { char* q_ = p; p += b; return q_; };
    double* d_r = (double*)take(vec); double* d_q = (double*)take(vec); double* d_s = (double*)take(vec);
    double* d_T = (double*)take(vec); double* d_S = (double*)take(vec); uint8_t* d_put = (uint8_t*)take(vec);
    double* d_K = (double*)take(kb); double* d_price = (double*)take(outb); int32_t* d_it = (int32_t*)take(outb);
    void* d_ws = take(wsb);
    cudaStream_t st = c->stream;
    const size_t vb = (size_t)G * sizeof(double);
    FAOP_CUDA_OK(cudaMemcpyAsync(d_r, r, vb, cudaMemcpyHostToDevice, st));
    FAOP_CUDA_OK(cudaMemcpyAsync(d_q, q, vb, cudaMemcpyHostToDevice, st));
    FAOP_CUDA_OK(cudaMemcpyAsync(d_s, sigma, vb, cudaMemcpyHostToDevice, st));
    FAOP_CUDA_OK(cudaMemcpyAsync(d_T, T, vb, cudaMemcpyHostToDevice, st));
    FAOP_CUDA_OK(cudaMemcpyAsync(d_S, S, vb, cudaMemcpyHostToDevice, st));
    FAOP_CUDA_OK(cudaMemcpyAsync(d_put, is_put, (size_t)G, cudaMemcpyHostToDevice, st));
    FAOP_CUDA_OK(cudaMemcpyAsync(d_K, K, (size_t)nK * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = faop_alo_ladder_batch(cfg, G, d_r, d_q, d_s, d_T, d_S, d_put, nK, d_K, Kmin, Kmax, c->tables_dev, d_price,
                               iters ? d_it : nullptr, d_ws, wsb, st);
    if (rc) return rc;
    FAOP_CUDA_OK(cudaMemcpyAsync(price, d_price, (size_t)G * nK * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (iters) FAOP_CUDA_OK(cudaMemcpyAsync(iters, d_it, (size_t)G * nK * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FAOP_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

// Host-buffer form of faop_cn_put_countdown_batch (modes 0 and 2; the SOR mode's N x 6 n_space scratch is the caller's business).
int faop_cn_put_countdown_batch_host(faop_ctx* c, int64_t N, int32_t n_space, int32_t mode, const double* r, const double* q,
                                     const double* sigma, const double* K, const double* T, const double* S0,
                                     const double* max_dt, const double* Smax, double* price, double* X_out) {
    if (!c || N < 0 || (mode != 0 && mode != 2)) return FAOP_EINVAL;
    if (N == 0) return 0;
    if (!r || !q || !sigma || !K || !T || !S0 || !max_dt || !price) return FAOP_EINVAL;
    FAOP_CUDA_OK(cudaSetDevice(c->device));
    const size_t vec = ((size_t)N * sizeof(double) + 255) & ~(size_t)255;
    const size_t xb = X_out ? (((size_t)N * n_space * sizeof(double) + 255) & ~(size_t)255) : 0;
    int rc = ctx_arena(c, 9 * vec + xb);
    if (rc) return rc;
    double* d = (double*)c->dev;
    const size_t st_ = vec / sizeof(double);
    const double* src[8] = {r, q, sigma, K, T, S0, max_dt, Smax};
    cudaStream_t st = c->stream;
    for (int a = 0; a < 8; ++a)
        if (src[a]) FAOP_CUDA_OK(cudaMemcpyAsync(d + a * st_, src[a], (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st));
    double* d_price = d + 8 * st_;
    double* d_X = X_out ? d + 9 * st_ : nullptr;
    faop_cn_extra ex = {Smax ? d + 7 * st_ : nullptr, nullptr, nullptr, nullptr};
    rc = faop_cn_put_countdown_batch(N, n_space, mode, d, d + st_, d + 2 * st_, d + 3 * st_, d + 4 * st_, d + 5 * st_, d + 6 * st_, 1.2,
                                     1e-5, 0, &ex, d_price, d_X, nullptr, st);
    if (rc) return rc;
    FAOP_CUDA_OK(cudaMemcpyAsync(price, d_price, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (X_out) FAOP_CUDA_OK(cudaMemcpyAsync(X_out, d_X, (size_t)N * n_space * sizeof(double), cudaMemcpyDeviceToHost, st));
    FAOP_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
