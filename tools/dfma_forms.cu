// Which DFMA operand forms reach the FP64 pipe's peak on sm_100a?  (complements dfma_peak.cu)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* sink, double mp, double bp, int iters) {
    double x[8], bb[8];
    const double mr = sink[1 + (threadIdx.x & 1)], br = sink[3 + (threadIdx.x & 1)];
#pragma unroll
    for (int c = 0; c < 8; ++c) { x[c] = 1.0 + c + threadIdx.x * 1e-9; bb[c] = sink[8 + c]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (MODE == 0) x[c] = fma(x[c], mp, bp);      // A: multiplier and addend are kernel parameters
            if (MODE == 1) x[c] = fma(x[c], mr, br);      // B: both in regular registers, shared by the chains
            if (MODE == 2) x[c] = fma(x[c], mr, bb[c]);   // C: distinct addend register per chain
            if (MODE == 3) x[c] = fma(x[c], mp, bb[c]);   // D: parameter multiplier, distinct addend registers
            if (MODE == 4) x[c] = fma(x[c], mr, 1.0);     // E: register multiplier, immediate addend
            if (MODE == 5) x[c] = fma(x[c], x[c], br);    // F: x*x + b (two reads of the same register)
            if (MODE == 6) x[c] = fma(mr, br, x[c]);      // G: accumulate form, product of two shared registers
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    if (s == 123456.789) sink[0] = s;
}
template <int MODE> void run(const char* name) {
    double* sink; cudaMalloc(&sink, 512);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.999999 - 1e-9 * i; for (int i = 8; i < 64; ++i) h[i] = 1e-7 * i;
    cudaMemcpy(sink, h, 512, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096, blocks = 148 * 32, L = 20;
    k<MODE><<<blocks, 256>>>(sink, 0.999999, 1e-7, iters);
    cudaEventRecord(e0);
    for (int i = 0; i < L; ++i) k<MODE><<<blocks, 256>>>(sink, 0.999999, 1e-7, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 8.0 * iters * blocks * 256 * L;
    printf("%-62s %.1f lanes/clk/SM @1965MHz  %.2f TFLOP/s\n", name, inst / (ms * 1e-3) / 148 / 1.965e9, 2 * inst / (ms * 1e-3) / 1e12);
    cudaFree(sink);
}
int main() {
    run<0>("A fma(x, m_param, b_param)");
    run<1>("B fma(x, m_reg, b_reg) shared registers");
    run<2>("C fma(x, m_reg, bb[c]) distinct addend registers");
    run<3>("D fma(x, m_param, bb[c])");
    run<4>("E fma(x, m_reg, 1.0)");
    run<5>("F fma(x, x, b_reg)");
    run<6>("G fma(m_reg, b_reg, x)");
    return 0;
}
