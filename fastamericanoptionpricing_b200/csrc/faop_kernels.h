// faop_kernels.h — internal launch interface between the C-ABI layer (faop_api.cu) and the kernels.
#pragma once
#include "faop_common.cuh"

namespace faop {

struct AloArgs {
    faop_cfg cfg;
    TableLayout tl;
    int64_t N;
    const double *r, *q, *sigma, *K, *T, *t, *S;
    const double* tau_max;   // nullable: end of the collocation domain per option (the tau_max attribute, Base.py:28); T when null
    const uint8_t* is_put;
    const double* tables;    // device table blob (faop_tables_build)
    const double* seeds_in;  // N x (n+1) boundary the solve starts from (QD seed or borrowed B0)
    double* seeds_out;       // where qd_seed_kernel writes
    double *price, *european, *B;
    int32_t* iters;
    double* err;
    double* iter_errs;       // nullable: N x max_iters, per-iteration ||B_old - B_new||_2 (NaN padded)
    // eval / interpolation kernels: npts points per option at arbitrary (tau, B) / u
    int npts;
    const double *pts_tau, *pts_B;
    // snapshot mode (deduplicated surface sweep, faop_surface.cu): instead of pricing, the solve records every iterate
    // whose error is a new minimum <= snap_tol_hi as a row [err, count, a_0..a_n] (a = DCT-I fit of ln(B/X)^2) and the
    // final iterate as a catch-all row with err = 0; cfg.iter_tol is the LOWEST tolerance any sharing option needs.
    double* inv_scratch;     // nullable: per-warp scratch for e^{q u_ik}, e^{r u_ik} (faop_alo_clen.cu), kSMs*16 warps x 2*n*LP
    double* snap;            // N x snap_cap x (n+3)
    int32_t* nsnap;          // rows written per option; -1 = European shortcut (q == 0 call)
    int snap_cap;
    double snap_tol_hi;
};

struct EvalOut {
    double *Nv, *Dv, *Np, *Dp, *f, *fp, *Bnew;
    double *I1, *I2;   // the two quadrature integrals: A: K12, K3 (A.py:24-44); B: the D- and N-integrals (B.py:25-39)
};

int launch_qd_seed(const AloArgs& a, cudaStream_t st);        // dispatches on FAOP_FLAG_SEED_NEWTON
int launch_qd_seed_hybr(const AloArgs& a, cudaStream_t st);   // faop_seed.cu: the reference's root finder (default)
int launch_alo_generic(const AloArgs& a, cudaStream_t st);
int launch_alo_eval_nodes(const AloArgs& a, const EvalOut& eo, cudaStream_t st);
int launch_alo_price_known(const AloArgs& a, cudaStream_t st);
int launch_alo_greeks_known(const AloArgs& a, double bump_S_rel, double bump_t, double* delta, double* gamma, double* theta,
                            cudaStream_t st);
int launch_iv_update(int64_t N, const double* target, const double* price, double* sigma, double* lo, double* hi, double* flo,
                     double* fhi, int32_t* side, cudaStream_t st);
int launch_alo_interp(const AloArgs& a, double* Bu, cudaStream_t st);
// specialised kernels (faop_alo_fast.cu); returns -1000 when no specialisation matches the config
int launch_alo_fast(const AloArgs& a, cudaStream_t st);
// Clenshaw-based specialised kernels (faop_alo_clen.cu): Solver A/B, no derivative, 2 <= n <= 31, l <= 64
int launch_alo_clen(const AloArgs& a, cudaStream_t st);
size_t alo_clen_scratch_bytes(const faop_cfg& cfg);   // 0 when the config does not run there

int launch_qd_boundary(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                       const double* tau, const uint8_t* is_put, double* B, cudaStream_t st);
int launch_qd_residual(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                       const double* tau, const double* S, const uint8_t* is_put, double* F, cudaStream_t st);
int launch_qd_price(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                    const double* tau, const double* S, const uint8_t* is_put, double* price, double* boundary,
                    cudaStream_t st);
int launch_bjerksund_stensland(int64_t N, const double* r, const double* q, const double* sigma, const double* K,
                               const double* T, const double* S, const uint8_t* is_put, const double* Kb, const double* Tb,
                               double* price, double* boundary, cudaStream_t st);
int launch_bs_phi(int64_t N, const double* S, const double* T, const double* gamma, const double* H, const double* X,
                  const double* r, const double* q, const double* sigma, double* out, cudaStream_t st);
int launch_european(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                    const double* sigma, const double* K, const uint8_t* is_put, double* price, cudaStream_t st);
int launch_european_greeks(int64_t N, const double* tau, const double* S, const double* r, const double* q,
                           const double* sigma, const double* K, double* theta, double* d1, double* d2, cudaStream_t st);
int launch_cheb_fit(int64_t N, int n, const double* y, double* a, cudaStream_t st);
int launch_cheb_eval(int64_t N, int n, int npts, const double* a, const double* z, double* out, cudaStream_t st);

// deduplicated surface sweep (faop_surface.cu)
struct SurfArgs {
    faop_cfg cfg;
    TableLayout tl;
    faop_surface s;
    double stepK, stepT, stepV;
    int64_t first, count;    // flat option range to price
    int g_lo, g_cnt;         // groups (iT*nV + iV) solved: [g_lo, g_lo + g_cnt)
    int iK_lo, iK_hi;        // strike rows touched by the range
    const double* tables;
    const double* snap;
    const int32_t* nsnap;
    int snap_cap;
    double* price;
    int32_t* iters;
    // ladder mode (faop_alo_ladder_batch): g_cnt explicit scenarios x one strike ladder, output [g * nK + k]
    const double *g_r, *g_q, *g_sigma, *g_T, *g_S;
    const uint8_t* g_put;
    const double* ladder;
    int nK;
};
int launch_surface_groups(const SurfArgs& a, double* r, double* q, double* sigma, double* K, double* T, uint8_t* is_put,
                          cudaStream_t st);
int launch_surface_price(const SurfArgs& a, cudaStream_t st);
int launch_fill(double* dst, double value, int64_t count, cudaStream_t st);
int launch_surface_decode(const faop_surface& s, int64_t first, int64_t count, double* r, double* q, double* sigma,
                          double* K, double* T, double* S, uint8_t* is_put, cudaStream_t st);
int launch_dev_math(int kind, int64_t N, const double* x, double* out, cudaStream_t st);
int launch_fp64_peak(int launches, double* tflops, double* ms_per_launch, cudaStream_t st);
int launch_point_mix_peak(int launches, double* points_per_s, double* ms_per_launch, cudaStream_t st);

}  // namespace faop
