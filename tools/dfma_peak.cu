// Standalone FP64-pipe microbenchmarks: what DFMA rate does a B200 SM sustain, with which operand forms?
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kC[16] = {0.999999, 1e-7, 0.99999, 2e-7, 0.9999, 3e-7, 0.999, 4e-7, 0.99, 5e-7, 0.9, 6e-7, 0.8, 7e-7, 0.7, 8e-7};
template <int CH, int MODE>
__global__ void __launch_bounds__(256) k(double* sink, double m0, double b0, int iters) {
    double x[CH];
    double m = m0, b = b0;
    double mr = sink[1 + (threadIdx.x & 1)];   // opaque per-thread value -> lives in a regular register
    double bb[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) bb[c] = sink[4 + c];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.0 + c + threadIdx.x * 1e-9;
    float f[8]; int iv[8]; float fs = 0; int is = 0;
    __shared__ double sh[256];
    sh[threadIdx.x] = mr;
#pragma unroll
    for (int c = 0; c < 8; ++c) { f[c] = (float)mr + c; iv[c] = threadIdx.x + c; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 14 || MODE == 17) {
#pragma unroll
            for (int c = 0; c < 8; ++c) f[c] = fmaf(f[c], 0.9999f, 0.5f);
        }
        if (MODE == 15 || MODE == 17) {
#pragma unroll
            for (int c = 0; c < 8; ++c) iv[c] = (iv[c] ^ (iv[c] >> 3)) + it;
        }
        if (MODE == 16) {
#pragma unroll
            for (int c = 0; c < 8; ++c) iv[c] += __double2loint(sh[(threadIdx.x + c * 32 + it) & 255]);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0 || MODE >= 14) x[c] = fma(x[c], mr, bb[c]);                 // reg, reg, reg
            if (MODE == 1) x[c] = fma(x[c], 0.999999, 1e-7);       // immediates / constant bank
            if (MODE == 2) x[c] = fma(x[c], x[(c + 1) % CH], b);   // three distinct registers
            if (MODE == 3) x[c] = x[c] * m;                        // DMUL
            if (MODE == 4) x[c] = x[c] + b;                        // DADD
            if (MODE == 5) x[c] = fma(x[c], m, kC[(2 * c + 1) % 16]);   // Horner-like: addend from __constant__
            if (MODE == 6) x[c] = fma(x[c], kC[(2 * c) % 16], b);       // multiplier from __constant__
            if (MODE == 7) x[c] = fma(x[c], m, 1.0);                    // 32-bit-encodable immediate addend
            if (MODE == 8) x[c] = (c & 1) ? fma(x[c], m0, bb[c]) : fma(x[c], mr, bb[c]);   // alternate UR / R multiplier
            if (MODE == 9) { double cc; asm volatile("ld.const.f64 %0, [%1];" : "=d"(cc) : "l"(kC + ((it + c) & 15))); x[c] = fma(x[c], mr, cc); }
            if (MODE == 11) { double cc = kC[c]; asm volatile("" : "+d"(cc)); x[c] = fma(x[c], mr, cc); }             // 1 opaque copy per DFMA
            if (MODE == 12) { if (c % 2 == 0) { double cc = kC[c]; asm volatile("" : "+d"(cc)); x[c] = fma(x[c], mr, cc); x[c + 1] = fma(x[c + 1], mr, cc); } }  // per 2
            if (MODE == 13) { if (c % 4 == 0) { double cc = kC[c]; asm volatile("" : "+d"(cc)); x[c] = fma(x[c], mr, cc); x[c + 1] = fma(x[c + 1], mr, cc); x[c + 2] = fma(x[c + 2], mr, cc); x[c + 3] = fma(x[c + 3], mr, cc); } }  // per 4
            if (MODE == 10) { if (c % 4 == 0) x[c] = fma(x[c], m0, bb[c]); else x[c] = fma(x[c], mr, bb[c]); }  // 1 in 4 with UR
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) { fs += f[c]; is += iv[c]; }
    double s = fs + (double)is;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    if (s == 123456.789) sink[0] = s;
}
template <int CH, int MODE>
void run(const char* name, int ctas_per_sm) {
    double* sink; cudaMalloc(&sink, 512); { double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.999999 - 1e-9 * i; for (int i = 4; i < 64; ++i) h[i] = 1e-7 * i; cudaMemcpy(sink, h, 512, cudaMemcpyHostToDevice); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096, blocks = 148 * ctas_per_sm * 4;
    k<CH, MODE><<<blocks, 256>>>(sink, 0.999999, 1e-7, iters);
    cudaEventRecord(e0);
    int L = 20;
    for (int i = 0; i < L; ++i) k<CH, MODE><<<blocks, 256>>>(sink, 0.999999, 1e-7, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)CH * iters * blocks * 256 * L;
    double per_clk_sm = inst / (ms * 1e-3) / 148 / 1.965e9;
    printf("%-28s ch=%2d ctas/sm=%d  %.2f T inst/s  = %.1f lanes/clk/SM @1965MHz  (%.2f TFLOP/s as FMA)\n", name, CH, ctas_per_sm,
           inst / (ms * 1e-3) / 1e12, per_clk_sm, 2 * inst / (ms * 1e-3) / 1e12);
    cudaFree(sink);
}
int main() {
    run<8, 0>("dfma reg,reg,reg", 8);
    run<16, 0>("dfma reg,reg,reg", 8);
    run<4, 0>("dfma reg,reg,reg", 8);
    run<8, 0>("dfma reg,reg,reg", 4);
    run<8, 0>("dfma reg,reg,reg", 2);
    run<8, 1>("dfma imm/const", 8);
    run<8, 2>("dfma 3 distinct regs", 8);
    run<8, 5>("dfma addend __constant__", 8);
    run<8, 6>("dfma multiplier __constant__", 8);
    run<8, 7>("dfma addend imm 1.0", 8);
    run<8, 8>("dfma alternate UR/R mult", 8);
    run<8, 10>("dfma 1-in-4 UR mult", 8);
    run<8, 9>("dfma + ld.const each", 8);
    run<8, 11>("dfma + opaque const copy 1:1", 8);
    run<8, 12>("dfma + opaque const copy 1:2", 8);
    run<8, 13>("dfma + opaque const copy 1:4", 8);
    run<8, 14>("8 dfma + 8 ffma", 8);
    run<8, 15>("8 dfma + 16 int alu", 8);
    run<8, 16>("8 dfma + 8 lds", 8);
    run<8, 17>("8 dfma + 8 ffma + 16 int", 8);
    run<8, 3>("dmul", 8);
    run<8, 4>("dadd", 8);
    // long run for clock sampling
    double* sink; cudaMalloc(&sink, 512);
    for (int i = 0; i < 60; ++i) k<8, 0><<<148 * 32, 256>>>(sink, 0.999999, 1e-7, 4096);
    cudaDeviceSynchronize();
    printf("long run done\n");
    return 0;
}
