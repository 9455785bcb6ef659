// faop_alo_fast.cuh — pieces shared by the specialised boundary kernels (faop_alo_fast.cu, faop_alo_clen.cu).
#pragma once
#include "faop_alo.cuh"
#include "faop_fastmath.cuh"

namespace faop {

// Node constants the sweep reads per point, packed so that a lane fetches them with four 16-byte broadcast loads.
// The d+- arguments are carried as x = d/sqrt2 (exp(-d^2/2) = exp(-x^2), Phi(-d) = erfcx(x)/2 exp(-x^2)).
struct __align__(16) NodePack {
    double L;      // ln(B_i/X), refreshed every iteration
    double cinv2;  // 1/(sigma sqrt(2 tau_i))
    double apc2;   // ((r-q) + sigma^2/2) sqrt(tau_i/2)/sigma
    double ssu2;   // sigma sqrt(tau_i/2)
    double qtau;   // q tau_i
    double rtau;   // r tau_i
    double cis;    // 1/(sigma sqrt(2 pi tau_i))
    double pad;
};

FAOP_D NodePack make_node_pack(const Opt& o, double tau) {
    NodePack np;
    const double sqt = sqrt(tau);
    np.L = 0.0;
    np.cinv2 = kInvSqrt2 / (o.sg * sqt);
    np.apc2 = kInvSqrt2 * ((o.r - o.q) + 0.5 * o.sg * o.sg) * sqt / o.sg;
    np.ssu2 = kInvSqrt2 * o.sg * sqt;
    np.qtau = o.q * tau;
    np.rtau = o.r * tau;
    np.cis = kInvSqrt2Pi / (o.sg * sqt);
    np.pad = 0.0;
    return np;
}

// Reduce 8 per-lane values over the 32 lanes through a shared-memory transpose: every lane stores its 8 partial sums
// (row v, column lane), then lane l sums the 8 entries {l&3, (l&3)+4, ...} of row v = l>>2 and two xor-shuffles finish
// the row.  On return every lane of the 4-lane group l>>2 holds the total of value index l>>2.  Costs 8 STS + 8 LDS +
// 9 DADD + 2 shuffles per 4 nodes, about half the issue slots of the select-and-shuffle butterfly it replaces.
FAOP_D double transpose_reduce8(const double (&v)[8], double* red, int lane) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k * 36 + lane] = v[k];
    __syncwarp();
    const double* row = red + (lane >> 2) * 36 + (lane & 3);
    double s0 = row[0] + row[4], s1 = row[8] + row[12], s2 = row[16] + row[20], s3 = row[24] + row[28];
    double t = (s0 + s1) + (s2 + s3);
    t += __shfl_xor_sync(kFull, t, 1);
    t += __shfl_xor_sync(kFull, t, 2);
    __syncwarp();   // the buffer is rewritten by the next group
    return t;
}

}  // namespace faop
