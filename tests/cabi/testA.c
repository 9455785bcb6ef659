/* A plain-C caller of libfaop.so: the reference's testA.py case (testA.py:7-21: American put, r=.04, q=.01, sigma=.2, K=100,
 * T=3, S=80, FastAmericanOptionSolverA with n=12, l=24, m=48, iter_tol=1e-5, max_iters=20) through the host-buffer entry
 * point, plus a 1000-option batch for a ragged size.  No Python, no PyTorch: include/faop.h and -lfaop only.
 * Build: gcc -O2 -I include tests/cabi/testA.c -o testA -L fastamericanoptionpricing_b200/csrc -lfaop -lm
 * Exit code 0 and "ok" on success. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "faop.h"

int main(void) {
    faop_cfg cfg = {0 /* Solver A */, 0, 12, 24, 48, 20, 1e-5, 0.5, 0, 0};
    double yl[24], wl[24], ym[48], wm[48];
    if (faop_gauss_legendre_host(24, yl, wl) || faop_gauss_legendre_host(48, ym, wm)) return 2;
    faop_ctx* ctx = NULL;
    int rc = faop_ctx_create(0, &ctx);
    if (rc) { printf("faop_ctx_create failed: %d\n", rc); return 3; }
    enum { N = 1000 };
    static double r[N], q[N], sg[N], K[N], T[N], S[N], price[N], eur[N], err[N], B[N * 13];
    static unsigned char put[N];
    static int iters[N];
    for (int i = 0; i < N; ++i) {
        r[i] = 0.04; q[i] = 0.01; sg[i] = 0.2; K[i] = 100; T[i] = 3; put[i] = 1;
        S[i] = 80 + 0.04 * i;            /* option 0 is testA.py's */
    }
    rc = faop_alo_price_batch_host(ctx, &cfg, N, r, q, sg, K, T, NULL, S, put, yl, wl, ym, wm, NULL, price, eur, B, NULL,
                                   iters, err);
    if (rc) { printf("faop_alo_price_batch_host failed: %d\n", rc); return 4; }
    printf("testA put: price %.12f european %.12f iters %d err %.3e B(T)=%.6f B(0)=%.6f\n", price[0], eur[0], iters[0], err[0],
           B[0], B[12]);
    int bad = fabs(price[0] - 21.135486471314067) > 1e-9 || fabs(eur[0] - 17.78865238332743) > 1e-10 || iters[0] != 19 ||
              B[12] != 100.0;
    for (int i = 1; i < N; ++i)          /* American put: decreasing in the spot, above European and intrinsic-ish */
        if (!(price[i] < price[i - 1]) || !(price[i] >= eur[i])) bad = 1;
    /* strike ladder through the host entry point: scenario 0 = the testA option (its K = 100 rung must reproduce price[0]) */
    {
        double lr[2] = {0.04, 0.03}, lq[2] = {0.01, 0.05}, ls[2] = {0.2, 0.3}, lT[2] = {3.0, 1.0}, lS[2] = {80.0, 100.0};
        unsigned char lput[2] = {1, 0};
        double Kl[5] = {60, 80, 100, 120, 140}, lp[10];
        int lit[10];
        rc = faop_alo_ladder_batch_host(ctx, &cfg, 2, lr, lq, ls, lT, lS, lput, 5, Kl, yl, wl, ym, wm, lp, lit);
        if (rc) { printf("faop_alo_ladder_batch_host failed: %d\n", rc); return 5; }
        printf("ladder: put K=100 %.12f (%d sweeps), call rungs %.6f %.6f %.6f %.6f %.6f\n", lp[2], lit[2], lp[5], lp[6], lp[7], lp[8], lp[9]);
        if (fabs(lp[2] - price[0]) > 1e-9 || lit[2] != iters[0]) bad = 1;
        for (int k = 1; k < 5; ++k)
            if (!(lp[k] > lp[k - 1]) || !(lp[5 + k] < lp[5 + k - 1])) bad = 1;      /* puts rise, calls fall with the strike */
    }
    /* Crank-Nicolson through the host entry point: the anchor of the reference's __main__ (CrankNicolsonOptionSolver.py:110-128:
     * r = q = .04, sigma = .2, K = 100, T = 3, S0 = 80, N = 200, max_dt = 1/12, default mode without projection) */
    {
        double cr[1] = {0.04}, cq[1] = {0.04}, cs[1] = {0.2}, cK[1] = {100}, cT[1] = {3.0}, cS[1] = {80.0}, cdt[1] = {1.0 / 12.0}, cp[1];
        static double X[200];
        rc = faop_cn_put_countdown_batch_host(ctx, 1, 200, 0, cr, cq, cs, cK, cT, cS, cdt, NULL /* Smax = 3K */, cp, X);
        if (rc) { printf("faop_cn_put_countdown_batch_host failed: %d\n", rc); return 6; }
        printf("crank-nicolson: %.12f (X[0] = %.6f)\n", cp[0], X[0]);
        if (fabs(cp[0] - 22.01470571289995) > 1e-10) bad = 1;
    }
    faop_ctx_destroy(ctx);
    printf(bad ? "MISMATCH\n" : "ok\n");
    return bad;
}
