// faop_alo.cu — the ALO spectral-collocation hot path: QD seed kernel, boundary fixed-point kernel
// (generic: any n <= 64, l, m <= 128, Solver A/B, with/without derivative), price integral.
// Replaces FastAmericanOptionSolver.solve (FastAmericanOptionSolverBase.py:49-165) for a batch.
#include "faop_alo.cuh"
#include "faop_kernels.h"

namespace faop {

// ---------------------------------------------------------------- QD seed: one thread per (option, node)
// set_initial_guess (FastAmericanOptionSolverBase.py:258-265) -> QDplus.compute_exercise_boundary per node.
__global__ void __launch_bounds__(256) qd_seed_kernel(AloArgs a) {
    const int nn = a.tl.n + 1;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.N * nn) return;
    int64_t opt = idx / nn;
    int node = (int)(idx - opt * nn);
    double r = a.r[opt], q = a.q[opt], sg = a.sigma[opt], K = a.K[opt], T = a.T[opt];
    bool put = a.is_put[opt] != 0;
    double out;
    if (q == 0 && !put) {
        out = nan("");
    } else {
        double c = a.tables[a.tl.off_c + node];
        double tau = (c + 1) * (c + 1) * T / 4;  // to_orig_point, FastAmericanOptionSolverBase.py:239-240
        out = qd_boundary(tau, r, q, sg, K, put);
    }
    a.seeds_out[idx] = out;
}

// ---------------------------------------------------------------- per-warp shared-memory view (generic kernels)
struct WarpSmem {
    double *B, *L, *H, *A, *tau, *cinv, *disc, *s1, *s2, *s3, *s4;
    static __host__ __device__ int doubles(int nn) { return 11 * nn; }
    __device__ WarpSmem(double* base, int nn) {
        B = base; L = B + nn; H = L + nn; A = H + nn; tau = A + nn; cinv = tau + nn; disc = cinv + nn;
        s1 = disc + nn; s2 = s1 + nn; s3 = s2 + nn; s4 = s3 + nn;
    }
};

struct CtaTables {
    const double *wl, *hl, *rhl, *sgl, *wm, *hm, *rhm, *sgm, *c, *cs;
    __device__ CtaTables(const double* s, const TableLayout& t) {
        wl = s + t.off_wl; hl = s + t.off_hl; rhl = s + t.off_rhl; sgl = s + t.off_sgl;
        wm = s + t.off_wm; hm = s + t.off_hm; rhm = s + t.off_rhm; sgm = s + t.off_sgm;
        c = s + t.off_c; cs = s + t.off_cos;
    }
};

FAOP_D void stage_tables(double* s_tab, const double* g_tab, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
}

FAOP_D Opt load_option(const AloArgs& a, int64_t opt) {
    Opt o;
    o.r = a.r[opt]; o.q = a.q[opt]; o.sg = a.sigma[opt]; o.K = a.K[opt]; o.T = a.T[opt];
    o.t = a.t ? a.t[opt] : 0.0;
    o.S = a.S ? a.S[opt] : 0.0;
    o.put = a.is_put[opt] != 0;
    o.X = b_at_zero(o.r, o.q, o.K, o.put);
    o.om = o.put ? 1.0 : -1.0;
    return o;
}

// L_j = ln(B_j/X), H_j = L_j^2 (FastAmericanOptionSolverBase.py:184) and the DCT-I fit
// a_k = (2/n) sum'' H_j cos(j k pi/n) (ChebyshevInterpolation.py:43-51), lanes over j / k.
FAOP_D void fit_boundary(const WarpSmem& w, const CtaTables& tb, int n, double X, int lane) {
    for (int j = lane; j <= n; j += 32) {
        double L = log(w.B[j] / X);
        w.L[j] = L;
        w.H[j] = L * L;
    }
    __syncwarp();
    const int n2 = 2 * n;
    for (int k = lane; k <= n; k += 32) {
        double ans = 0.0;
        int idx = 0;  // (j*k) mod 2n
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
}

FAOP_D NodeC node_from_smem(const WarpSmem& w, const Opt& o, int j) {
    NodeC c;
    c.tau = w.tau[j];
    c.cinv = w.cinv[j];
    c.ssu = o.sg * sqrt(c.tau);
    c.apc = ((o.r - o.q) + 0.5 * o.sg * o.sg) * (c.ssu / (o.sg * o.sg));
    c.qtau = o.q * c.tau;
    c.rtau = o.r * c.tau;
    return c;
}

// One Jacobi sweep's quadrature: for every active node i the reduced sums s1..s4 land in shared memory.
template <int SOLVER, bool DERIV>
FAOP_D void sweep_sums(const WarpSmem& w, const CtaTables& tb, const TableLayout& tl, const Opt& o, int lane) {
    const int n = tl.n;
    for (int i = 0; i < n; ++i) {
        double tau = w.tau[i];
        if (tau == 0) continue;  // uniform across the warp
        const NodeC c = node_from_smem(w, o, i);
        double L = w.L[i];
        double opc = 1.0 + tb.c[i];
        double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        for (int k = lane; k < tl.LP; k += 32) {
            double h = tb.hl[k], sg_k = tb.sgl[k];
            double z = fma(opc, sg_k, -1.0);  // to_cheby_point(u_ik): sqrt(4 u/T) - 1 = (1+c_i) sqrt(g_k) - 1
            double Hu = clenshaw(n, w.A, z);
            double root = sqrt(Hu < 0.0 ? 0.0 : Hu);  // np.maximum(0, H): NaN propagates (fmax would drop it)
            double lnz = fma(o.om, root, L);
            alo_point<SOLVER, DERIV>(o, c, lnz, tb.wl[k], h, tb.rhl[k], sg_k * sg_k, a1, a2, a3, a4);
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (DERIV) { a3 = warp_sum(a3); a4 = warp_sum(a4); }
        if (lane == 0) { w.s1[i] = a1; w.s2[i] = a2; w.s3[i] = a3; w.s4[i] = a4; }
    }
    __syncwarp();
}

FAOP_D void setup_nodes(const WarpSmem& w, const CtaTables& tb, int n, const Opt& o, int lane) {
    for (int j = lane; j <= n; j += 32) {
        double c = tb.c[j];
        double tau = (c + 1) * (c + 1) * o.T / 4;
        w.tau[j] = tau;
        w.cinv[j] = 1.0 / (o.sg * sqrt(tau));
        w.disc[j] = o.K * exp(-tau * (o.r - o.q));
    }
}

// Price integral on the boundary currently in w.B: american_value_with_known_boundary
// (FastAmericanOptionSolverBase.py:92-103).  Returns price; *eur gets the European value.
FAOP_D double price_integral(const WarpSmem& w, const CtaTables& tb, const TableLayout& tl, const Opt& o, int lane,
                             double* eur) {
    double tau = o.T - o.t;
    double e = euro_value(tau, o.S, o.r, o.q, o.sg, o.K, o.put);
    *eur = e;
    if (tau == 0) return e + 0.0;  // quadrature_sum returns 0 at tau == 0 (:382-383)
    fit_boundary(w, tb, tl.n, o.X, lane);
    NodeC c = node_consts(o, tau);
    double LS = log(o.S / o.X);
    double zf = sqrt(4 * tau / o.T);  // z = sqrt(4 u/T) - 1 with u = tau g_k
    double acc = 0;
    for (int k = lane; k < tl.MP; k += 32) {
        double h = tb.hm[k];
        double z = fma(zf, tb.sgm[k], -1.0);
        double Hu = clenshaw(tl.n, w.A, z);
        double root = sqrt(Hu < 0.0 ? 0.0 : Hu);  // np.maximum(0, H): NaN propagates (fmax would drop it)
        double lnz = fma(o.om, root, LS);
        acc = fma(tb.wm[k] * h, price_point(o, c, lnz, h, tb.rhm[k]), acc);
    }
    acc = warp_sum(acc);
    return e + acc * tau;
}

constexpr int kGenWarps = 8;

// ---------------------------------------------------------------- generic boundary + price kernel
template <int SOLVER, bool DERIV>
__global__ void __launch_bounds__(kGenWarps * 32) alo_generic_kernel(AloArgs a) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        if (o.q == 0 && !o.put) {  // European shortcut, FastAmericanOptionSolverBase.py:50-53
            double e = euro_value(o.T - o.t, o.S, o.r, o.q, o.sg, o.K, false);
            if (lane == 0) {
                a.price[opt] = e;
                if (a.european) a.european[opt] = e;
                if (a.iters) a.iters[opt] = 0;
                if (a.err) a.err[opt] = 1e6;
            }
            if (a.B) for (int j = lane; j <= n; j += 32) a.B[opt * nn + j] = nan("");
            continue;
        }
        setup_nodes(w, tb, n, o, lane);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        const double lnXK = log(o.X / o.K);
        int count = 0;
        double err = 1.0;
        while (err > a.cfg.iter_tol && count < a.cfg.max_iters) {
            ++count;
            fit_boundary(w, tb, n, o.X, lane);
            sweep_sums<SOLVER, DERIV>(w, tb, tl, o, lane);
            double ss = 0;
            for (int j = lane; j <= n; j += 32) {
                double Bj = w.B[j], Bn;
                if (w.tau[j] == 0) {
                    Bn = o.X;  // iterate_once_foreach_tau: tau_i == 0 -> B_at_zero
                } else {
                    NodeC c = node_from_smem(w, o, j);
                    NodeOut no = alo_node<SOLVER, DERIV>(o, c, Bj, w.L[j], lnXK, w.disc[j], w.s1[j], w.s2[j], w.s3[j], w.s4[j]);
                    double fp = DERIV ? no.fp : 0.0;
                    Bn = Bj + a.cfg.eta * (Bj - no.f) / (fp - 1.0);
                }
                double d = fabs(Bj - Bn);
                ss = fma(d, d, ss);
                w.B[j] = Bn;
            }
            err = sqrt(warp_sum(ss));  // norm1_error is the 2-norm over all n+1 nodes (:230-233)
            __syncwarp();
        }
        if (a.B) for (int j = lane; j <= n; j += 32) a.B[opt * nn + j] = w.B[j];
        double eur;
        double v = price_integral(w, tb, tl, o, lane, &eur);
        if (lane == 0) {
            a.price[opt] = v;
            if (a.european) a.european[opt] = eur;
            if (a.iters) a.iters[opt] = count;
            if (a.err) a.err[opt] = err;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- N, D, N', D', f, f' on a given boundary
template <int SOLVER, bool DERIV>
__global__ void __launch_bounds__(kGenWarps * 32) alo_eval_nodes_kernel(AloArgs a, EvalOut eo) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        setup_nodes(w, tb, n, o, lane);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        const double lnXK = log(o.X / o.K);
        fit_boundary(w, tb, n, o.X, lane);
        sweep_sums<SOLVER, DERIV>(w, tb, tl, o, lane);
        for (int j = lane; j < n; j += 32) {
            NodeOut no;
            if (w.tau[j] == 0) {  // compute_f_and_fprime: tau == 0 -> [B_at_zero, 1]
                no.Nv = no.Dv = no.Np = no.Dp = nan("");
                no.f = o.X;
                no.fp = 1.0;
            } else {
                NodeC c = node_from_smem(w, o, j);
                no = alo_node<SOLVER, DERIV>(o, c, w.B[j], w.L[j], lnXK, w.disc[j], w.s1[j], w.s2[j], w.s3[j], w.s4[j]);
            }
            int64_t at = opt * n + j;
            if (eo.Nv) eo.Nv[at] = no.Nv;
            if (eo.Dv) eo.Dv[at] = no.Dv;
            if (eo.Np) eo.Np[at] = no.Np;
            if (eo.Dp) eo.Dp[at] = no.Dp;
            if (eo.f) eo.f[at] = no.f;
            if (eo.fp) eo.fp[at] = no.fp;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- price integral on a given boundary
__global__ void __launch_bounds__(kGenWarps * 32) alo_price_known_kernel(AloArgs a) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpSmem w(smem + tl.small + warp * WarpSmem::doubles(nn), nn);
    const int64_t stride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kGenWarps + warp; opt < a.N; opt += stride) {
        const Opt o = load_option(a, opt);
        for (int j = lane; j <= n; j += 32) w.B[j] = a.seeds_in[opt * nn + j];
        __syncwarp();
        double eur;
        double v = price_integral(w, tb, tl, o, lane, &eur);
        if (lane == 0) {
            a.price[opt] = v;
            if (a.european) a.european[opt] = eur;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- host-side launchers
static int generic_grid(int64_t N) {
    int64_t ctas = (N + kGenWarps - 1) / kGenWarps;
    int64_t cap = (int64_t)kSMs * 16;
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

static size_t generic_smem(const TableLayout& tl) {
    return sizeof(double) * (size_t)(tl.small + kGenWarps * WarpSmem::doubles(tl.n + 1));
}

int launch_qd_seed(const AloArgs& a, cudaStream_t st) {
    int64_t total = a.N * (a.tl.n + 1);
    if (total == 0) return 0;
    int64_t blocks = (total + 255) / 256;
    qd_seed_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    count_launch();
    return launch_status();
}

template <int SOLVER, bool DERIV>
static int launch_generic_t(const AloArgs& a, cudaStream_t st) {
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_generic_kernel<SOLVER, DERIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_generic_kernel<SOLVER, DERIV><<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a);
    count_launch();
    return launch_status();
}

int launch_alo_generic(const AloArgs& a, cudaStream_t st) {
    if (a.N == 0) return 0;
    if (a.cfg.solver == 0) return a.cfg.use_derivative ? launch_generic_t<0, true>(a, st) : launch_generic_t<0, false>(a, st);
    return a.cfg.use_derivative ? launch_generic_t<1, true>(a, st) : launch_generic_t<1, false>(a, st);
}

template <int SOLVER, bool DERIV>
static int launch_eval_t(const AloArgs& a, const EvalOut& eo, cudaStream_t st) {
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_eval_nodes_kernel<SOLVER, DERIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_eval_nodes_kernel<SOLVER, DERIV><<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a, eo);
    count_launch();
    return launch_status();
}

int launch_alo_eval_nodes(const AloArgs& a, const EvalOut& eo, cudaStream_t st) {
    if (a.N == 0) return 0;
    if (a.cfg.solver == 0) return a.cfg.use_derivative ? launch_eval_t<0, true>(a, eo, st) : launch_eval_t<0, false>(a, eo, st);
    return a.cfg.use_derivative ? launch_eval_t<1, true>(a, eo, st) : launch_eval_t<1, false>(a, eo, st);
}

int launch_alo_price_known(const AloArgs& a, cudaStream_t st) {
    if (a.N == 0) return 0;
    size_t sm = generic_smem(a.tl);
    if (sm > 48 * 1024)
        FAOP_CUDA_OK(cudaFuncSetAttribute(alo_price_known_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    alo_price_known_kernel<<<generic_grid(a.N), kGenWarps * 32, sm, st>>>(a);
    count_launch();
    return launch_status();
}

}  // namespace faop
