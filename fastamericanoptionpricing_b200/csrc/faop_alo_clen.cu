// faop_alo_clen.cu — specialised boundary + price kernel for the configurations whose interpolation matrix does not fit
// in shared memory: Solver A or B without derivative, any n <= 31 collocation nodes, l <= 64 quadrature
// points.  BASELINE.json configs[2] (C3: FastAmericanOptionSolverB, n = 24, l = 64, m = 128) runs here.
//
// Same structure as faop_alo_fast.cu (one option per warp, lane = quadrature index, 4 nodes per shared-memory transpose
// reduction, fast exp / erfcx / sqrt with constant-bank coefficients), except that
//   * H(u) is evaluated by Clenshaw from the DCT-I coefficients a_k of the current iterate (refitted once per iteration
//     by lanes 0..n, ChebyshevInterpolation.py:43-51), read from shared memory as broadcast loads;
//   * two point slots advance in lock step: with l > 32 (PPL = 2) the lane's two quadrature points k and k+32 of one
//     node, with l <= 32 (PPL = 1) the points of two consecutive nodes;
//   * e^{q u_ik}, e^{r u_ik} (iteration invariant, 2 n l doubles per option: 24.6 KB at C3) do not fit in shared memory next
//     to 16 warps' state; with SCR they live in a per-warp global scratch that stays in L2 (2368 resident warps x 24.6 KB
//     = 58 MB of the 126 MB), written once per option and re-read every sweep by the lane that wrote them (coalesced
//     256-byte rows); without a scratch they are recomputed per point.
// Solver B per point: E+ = exp(qu - d+^2/2), E- = exp(ru - d-^2/2), e^{qu}, e^{ru}, erfcx(|d+|/sqrt2)/2, erfcx(|d-|/sqrt2)/2
//   D-sum += w h e^{qu} Phi(om d+),  N-sum += w h e^{ru} Phi(om d-)           (FastAmericanOptionSolverB.py:25-39)
#include <string.h>
#include "faop_alo_fast.cuh"
#include "faop_kernels.h"

namespace faop {

constexpr int kClenWarps = 16;
constexpr int kClenMaxNodes = 32;   // n + 1 <= 32: one lane per node in the closing pass

struct ClenWarpSmem {
    NodePack node[kClenMaxNodes];
    double B[kClenMaxNodes], L[kClenMaxNodes], H[kClenMaxNodes], A[kClenMaxNodes];
    double tau[kClenMaxNodes], disc[kClenMaxNodes], opc[kClenMaxNodes];
    double sN[kClenMaxNodes], sD[kClenMaxNodes];
    double par[8];
    double red[8 * 36];
};

// Two Clenshaw recurrences in lock step (ChebyshevInterpolation.py:31-41), coefficients broadcast from shared memory.
FAOP_D void clenshaw2(int n, const double* A, const double (&z)[2], double (&out)[2]) {
    double b0[2], b1[2] = {0.0, 0.0}, b2[2] = {0.0, 0.0};
    const double an = A[n] * 0.5;
    const double z2[2] = {2.0 * z[0], 2.0 * z[1]};
    b0[0] = an;
    b0[1] = an;
#pragma unroll 4
    for (int k = n - 1; k >= 0; --k) {
        const double ak = A[k];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            b2[c] = b1[c];
            b1[c] = b0[c];
            b0[c] = fma(z2[c], b1[c], ak - b2[c]);
        }
    }
    out[0] = 0.5 * (b0[0] - b2[0]);
    out[1] = 0.5 * (b0[1] - b2[1]);
}

template <int SOLVER, int PPL, bool SCR>
__global__ void __launch_bounds__(kClenWarps * 32, 1) alo_clen_kernel(AloArgs a) {
    extern __shared__ double smem[];
    const TableLayout tl = a.tl;
    const int n = tl.n, nn = n + 1;
    ClenWarpSmem* s_w = reinterpret_cast<ClenWarpSmem*>(smem + tl.small);
    stage_tables(smem, a.tables, tl.small);
    const CtaTables tb(smem, tl);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ClenWarpSmem& ws = s_w[warp];

    // lane constants of the l-point rule: PPL quadrature points per lane
    double h[PPL], rh[PPL], sgk[PPL], g[PPL], wh[PPL];
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
        const int k = lane + 32 * p;
        h[p] = tb.hl[k]; rh[p] = tb.rhl[k]; sgk[p] = tb.sgl[k];
        g[p] = sgk[p] * sgk[p];
        wh[p] = tb.wl[k] * h[p];
    }
    WarpSmem gw(ws.B, kClenMaxNodes);
    gw.B = ws.B; gw.L = ws.L; gw.H = ws.H; gw.A = ws.A;
    // scratch rows of this warp: [(node * PPL + p) * 2 + {q, r}][lane]
    double* scr = SCR ? a.inv_scratch + ((size_t)blockIdx.x * kClenWarps + warp) * (size_t)(2 * n * PPL * 32) + lane : nullptr;

    const int64_t stride = (int64_t)gridDim.x * kClenWarps;
    for (int64_t opt = (int64_t)blockIdx.x * kClenWarps + warp; opt < a.N; opt += stride) {
        double om, oq, orr;
        int sgn;
        bool degenerate;
        {
            const Opt o = load_option(a, opt);
            if (o.q == 0 && !o.put) {  // European shortcut, FastAmericanOptionSolverBase.py:50-53
                double e = euro_value(o.T - o.t, o.S, o.r, o.q, o.sg, o.K, false);
                if (lane == 0) {
                    a.price[opt] = e;
                    if (a.european) a.european[opt] = e;
                    if (a.iters) a.iters[opt] = 0;
                    if (a.err) a.err[opt] = 1e6;
                }
                if (a.B && lane < nn) a.B[opt * nn + lane] = nan("");
                continue;
            }
            if (lane < nn) {
                const double c = tb.c[lane];
                const double tau = (c + 1) * (c + 1) * o.Tc / 4;   // to_orig_point, FastAmericanOptionSolverBase.py:239-240
                ws.tau[lane] = tau;
                ws.opc[lane] = 1.0 + c;
                ws.node[lane] = make_node_pack(o, tau);
                ws.disc[lane] = o.K * exp(-tau * (o.r - o.q));
                ws.B[lane] = (lane == n) ? o.X : a.seeds_in[opt * nn + lane];   // tau = 0 node pinned to X (see faop_alo_fast.cu)
            }
            if (lane == 0) {
                ws.par[0] = o.X;
                ws.par[1] = 1.0 / o.X;
                ws.par[2] = log(o.X / o.K);
            }
            om = o.om; oq = o.q; orr = o.r;
            sgn = o.put ? 0 : (int)0x80000000;
            degenerate = !(o.Tc > 0);
        }
        __syncwarp();
        if (SCR && !degenerate) {
            for (int i = 0; i < n; ++i) {
                const double qt = ws.node[i].qtau, rt = ws.node[i].rtau;
#pragma unroll
                for (int p = 0; p < PPL; ++p) {
                    double ex2[2] = {qt * g[p], rt * g[p]}, e2[2];
                    fast_exp_n<2, false>(ex2, e2);
                    __stcg(scr + ((i * PPL + p) * 2) * 32, e2[0]);
                    __stcg(scr + ((i * PPL + p) * 2 + 1) * 32, e2[1]);
                }
            }
        }

        int count = 0;
        double err = 1.0;
        while (err > a.cfg.iter_tol && count < a.cfg.max_iters) {
            ++count;
            // ---- L_j, H_j and the DCT-I fit a_k = (2/n) sum'' H_j cos(j k pi/n)
            bool bad = false;
            if (lane < nn) {
                const double L = (lane == n) ? 0.0 : fast_log(ws.B[lane] * ws.par[1]);
                ws.L[lane] = L;
                ws.node[lane].L = L;
                ws.H[lane] = L * L;
                bad = !(fabs(L) <= 1.79e308);
            }
            if (__any_sync(kFull, bad)) {    // see alo_fast_A12_kernel: the sweep upstream turns to NaN at every active node
                if (lane < n) ws.B[lane] = nan("");
                err = nan("");
                __syncwarp();
                break;
            }
            __syncwarp();
            if (lane < nn) {
                double ans = 0.0;
                int idx = 0;
                const int n2 = 2 * n;
                for (int j = 0; j <= n; ++j) {
                    double term = ws.H[j] * tb.cs[idx];
                    if (j == 0 || j == n) term *= 0.5;
                    ans += term;
                    idx += lane;
                    if (idx >= n2) idx -= n2;
                }
                ws.A[lane] = ans * (2.0 / n);
            }
            __syncwarp();

            if (!degenerate) {
#pragma unroll 1
                for (int grp = 0; grp < (n + 3) / 4; ++grp) {      // the last group of an n that is no multiple of 4 repeats node n-1
                    double acc[8];
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) acc[k8] = 0.0;
                    // slot pairs: PPL == 2 -> (node, k) and (node, k+32); PPL == 1 -> (node, k) and (node+1, k)
#pragma unroll
                    for (int pr = 0; pr < (PPL == 2 ? 4 : 2); ++pr) {
                        int nd_i[2];
                        double hh[2], rhh[2], gg[2], whh[2], z[2];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            nd_i[c] = min((PPL == 2) ? grp * 4 + pr : grp * 4 + 2 * pr + c, n - 1);
                            const int p = (PPL == 2) ? c : 0;
                            hh[c] = h[p]; rhh[c] = rh[p]; gg[c] = g[p]; whh[c] = wh[p];
                            z[c] = fma(ws.opc[nd_i[c]], sgk[p], -1.0);   // to_cheby_point(u_ik) = (1+c_i) sqrt(g_k) - 1
                        }
                        double Hu[2], root[2];
                        clenshaw2(n, ws.A, z, Hu);
                        sqrt_clamped_n<2>(Hu, root);
                        double ex[8], xa[4], xs[4];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const NodePack& nd = ws.node[nd_i[c]];
                            const double lnz = fma(om, root[c], nd.L);
                            const double xp = fma(lnz, nd.cinv2 * rhh[c], nd.apc2 * hh[c]);   // d+ / sqrt2
                            const double xm = fma(-nd.ssu2, hh[c], xp);                        // d- / sqrt2
                            const double aq = nd.qtau * gg[c], ar = nd.rtau * gg[c];           // q u, r u
                            ex[4 * c] = fma(-xp, xp, aq);
                            ex[4 * c + 1] = fma(-xm, xm, ar);
                            ex[4 * c + 2] = aq;
                            ex[4 * c + 3] = ar;
                            xs[2 * c] = xp; xs[2 * c + 1] = xm;
                            xa[2 * c] = fabs(xp); xa[2 * c + 1] = fabs(xm);
                        }
                        double inv[4];   // e^{qu}, e^{ru} of the two slots, fetched before the exp / erfcx chains that hide the L2 latency
                        if (SCR) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int p = (PPL == 2) ? c : 0;
                                const double* row = scr + ((nd_i[c] * PPL + p) * 2) * 32;
                                inv[2 * c] = __ldcg(row);
                                if (SOLVER == 1) inv[2 * c + 1] = __ldcg(row + 32);
                            }
                        }
                        double E[8], G[4];
                        if (SCR) {
                            double ex4[4] = {ex[0], ex[1], ex[4], ex[5]}, E4[4];
                            fast_exp_n<4, true>(ex4, E4);
                            E[0] = E4[0]; E[1] = E4[1]; E[4] = E4[2]; E[5] = E4[3];
                            E[2] = inv[0]; E[6] = inv[2];
                            if (SOLVER == 1) { E[3] = inv[1]; E[7] = inv[3]; }
                            if (SOLVER == 0) {
                                double xa2[2] = {xa[0], xa[2]}, G2[2];
                                half_erfcx_n<2, 2>(xa2, G2);
                                G[0] = G2[0]; G[2] = G2[1];
                            } else {
                                half_erfcx_n<4, 2>(xa, G);
                            }
                        } else if (SOLVER == 0) {
                            // Solver A needs e^{qu} but not e^{ru}, and one erfcx: reuse the B layout, skip the spare chains
                            double ex6[6] = {ex[0], ex[1], ex[2], ex[4], ex[5], ex[6]}, E6[6];
                            fast_exp_n<6, true>(ex6, E6);
                            E[0] = E6[0]; E[1] = E6[1]; E[2] = E6[2]; E[4] = E6[3]; E[5] = E6[4]; E[6] = E6[5];
                            double xa2[2] = {xa[0], xa[2]}, G2[2];
                            half_erfcx_n<2, 2>(xa2, G2);
                            G[0] = G2[0]; G[2] = G2[1];
                        } else {
                            fast_exp_n<8, true>(ex, E);
                            half_erfcx_n<4, 2>(xa, G);
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int slot = (PPL == 2) ? pr : 2 * pr + c;   // node index within the group
                            const double wk = whh[c] * rhh[c];               // w_k
                            if (SOLVER == 0) {
                                // K12 += w h [+-e^{qu}Phi(+-d+)] + cis w e^{qu}phi(d+) sqrt(2pi);  K3 += w e^{ru}phi(d-) sqrt(2pi)
                                const double ec = E[4 * c] * G[2 * c];
                                const bool neg = (__double2hiint(xs[2 * c]) ^ sgn) < 0;
                                double cdf = neg ? ec : E[4 * c + 2] - ec;
                                cdf = __hiloint2double(__double2hiint(cdf) ^ sgn, __double2loint(cdf));
                                acc[2 * slot] = fma(wk, E[4 * c + 1], acc[2 * slot]);
                                acc[2 * slot + 1] = fma(ws.node[nd_i[c]].cis * wk, E[4 * c], fma(whh[c], cdf, acc[2 * slot + 1]));
                            } else {
                                // put: Phi(+d), call: Phi(-d)  ->  Phi(om d)
                                const double ecp = E[4 * c] * G[2 * c], ecm = E[4 * c + 1] * G[2 * c + 1];
                                const bool negp = (__double2hiint(xs[2 * c]) ^ sgn) < 0;
                                const bool negm = (__double2hiint(xs[2 * c + 1]) ^ sgn) < 0;
                                const double cdfp = negp ? ecp : E[4 * c + 2] - ecp;   // e^{qu} Phi(om d+)
                                const double cdfm = negm ? ecm : E[4 * c + 3] - ecm;   // e^{ru} Phi(om d-)
                                acc[2 * slot] = fma(whh[c], cdfm, acc[2 * slot]);          // -> N integral
                                acc[2 * slot + 1] = fma(whh[c], cdfp, acc[2 * slot + 1]);  // -> D integral
                            }
                        }
                    }
                    const double tot = transpose_reduce8(acc, ws.red, lane);
                    if ((lane & 3) == 0) {
                        const int node = grp * 4 + (lane >> 3);
                        if (node < n) {
                            if (lane & 4) ws.sD[node] = tot; else ws.sN[node] = tot;
                        }
                    }
                }
            }
            __syncwarp();
            // ---- close N, D, f and the damped update at the nodes (lanes 0..n)
            double dd = 0.0;
            if (lane < nn) {
                const double Bj = ws.B[lane];
                double Bn;
                const double tau = ws.tau[lane];
                if (tau == 0) {
                    Bn = ws.par[0];
                } else {
                    const NodePack nd = ws.node[lane];
                    const double LK = nd.L + ws.par[2];
                    const double xpK = fma(LK, nd.cinv2, nd.apc2);
                    const double xmK = xpK - nd.ssu2;
                    const double Ep = fast_exp(-xpK * xpK), Em = fast_exp(-xmK * xmK);
                    double Nv, Dv;
                    if (SOLVER == 0) {
                        Nv = Em * nd.cis + orr * (tau * nd.cis * ws.sN[lane]);                                            // A.py:5-10
                        Dv = Ep * nd.cis + om * fast_half_erfcx_cdf(om * xpK, Ep) + oq * (tau * ws.sD[lane]);             // A.py:13-19
                    } else {
                        Nv = fast_half_erfcx_cdf(om * xmK, Em) + orr * (tau * ws.sN[lane]);                               // B.py:5-13
                        Dv = fast_half_erfcx_cdf(om * xpK, Ep) + oq * (tau * ws.sD[lane]);                                // B.py:15-23
                    }
                    const double f = ws.disc[lane] * Nv / Dv;
                    Bn = Bj + a.cfg.eta * (Bj - f) / (0.0 - 1.0);
                }
                dd = fabs(Bj - Bn);
                ws.B[lane] = Bn;
            }
            err = sqrt(warp_sum(dd * dd));
            __syncwarp();
        }
        if (a.B && lane < nn) a.B[opt * nn + lane] = ws.B[lane];
        double eur;
        const Opt o = load_option(a, opt);
        const double v = price_integral_fast(gw, tb, tl, o, lane, &eur);
        if (lane == 0) {
            a.price[opt] = v;
            if (a.european) a.european[opt] = eur;
            if (a.iters) a.iters[opt] = count;
            if (a.err) a.err[opt] = err;
        }
        __syncwarp();
    }
}

template <int SOLVER, int PPL, bool SCR>
static int launch_clen_t(const AloArgs& a, cudaStream_t st) {
    const size_t sm = sizeof(double) * (size_t)a.tl.small + sizeof(ClenWarpSmem) * kClenWarps;
    FAOP_CUDA_OK(cudaFuncSetAttribute(alo_clen_kernel<SOLVER, PPL, SCR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int64_t ctas = (a.N + kClenWarps - 1) / kClenWarps;
    int grid = (int)(ctas < kSMs ? ctas : kSMs);
    // The invariant-table scratch is written once per option and re-read every sweep: pin it in L2 for the launch (persisting
    // access-policy window on the launch stream, removed again afterwards) so that the other traffic does not evict it to DRAM.
    // (ncu, 32768 C3 options: DRAM writes 194 MB without the window, 9.9 MB = seeds + results with it; the time is the same.)
    bool window = false;
    if (SCR) {
        static int s_persist[64], s_window[64];      // per device: 0 = not asked yet, -1 = unsupported
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (s_persist[dev] == 0) {
            int mp = 0, mw = 0;
            cudaDeviceGetAttribute(&mp, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&mw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            if (mp > 0 && mw > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)mp) == cudaSuccess) {
                s_persist[dev] = mp;
                s_window[dev] = mw;
            } else {
                s_persist[dev] = -1;
            }
            cudaGetLastError();
        }
        if (s_persist[dev] > 0) {
            const size_t bytes = sizeof(double) * (size_t)grid * kClenWarps * 2 * (size_t)a.tl.n * PPL * 32;
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = (void*)a.inv_scratch;
            av.accessPolicyWindow.num_bytes = bytes < (size_t)s_window[dev] ? bytes : (size_t)s_window[dev];
            const double ratio = (double)s_persist[dev] / (double)av.accessPolicyWindow.num_bytes;
            av.accessPolicyWindow.hitRatio = ratio < 1.0 ? (float)ratio : 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
            cudaGetLastError();
        }
    }
    alo_clen_kernel<SOLVER, PPL, SCR><<<grid, kClenWarps * 32, sm, st>>>(a);
    count_launch();
    const int rc = launch_status();
    if (window) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);     // num_bytes = 0: no window
    }
    return rc;
}

static bool clen_match(const faop_cfg& cfg) {
    const int n = cfg.n_colloc, LP = 32 * ((cfg.n_quad + 31) / 32);
    // n = 12, l <= 32 Solver A runs in the matrix kernel (faop_alo_fast.cu)
    return !cfg.use_derivative && n >= 2 && n + 1 <= kClenMaxNodes && LP <= 64;
}

// Scratch for the invariant tables: one region per resident warp of the persistent grid.  Used for Solver B (two of four
// exponentials per point saved); Solver A would save one of three, which does not pay for the loads.
size_t alo_clen_scratch_bytes(const faop_cfg& cfg) {
    if (!clen_match(cfg) || cfg.solver != 1 || cfg.kernel == 1) return 0;
    const int LP = 32 * ((cfg.n_quad + 31) / 32);
    return sizeof(double) * (size_t)kSMs * kClenWarps * 2 * (size_t)cfg.n_colloc * LP;
}

int launch_alo_clen(const AloArgs& a, cudaStream_t st) {
    const int n = a.tl.n;
    const bool match = !a.cfg.use_derivative && n >= 2 && n + 1 <= kClenMaxNodes && a.tl.LP <= 64 && a.iter_errs == nullptr && a.snap == nullptr;
    if (!match) return -1000;
    if (a.N == 0) return 0;
    if (a.cfg.solver == 0) return a.tl.LP == 64 ? launch_clen_t<0, 2, false>(a, st) : launch_clen_t<0, 1, false>(a, st);
    if (a.inv_scratch) return a.tl.LP == 64 ? launch_clen_t<1, 2, true>(a, st) : launch_clen_t<1, 1, true>(a, st);
    return a.tl.LP == 64 ? launch_clen_t<1, 2, false>(a, st) : launch_clen_t<1, 1, false>(a, st);
}

}  // namespace faop
