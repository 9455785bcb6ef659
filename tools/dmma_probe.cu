// Does the FP64 tensor path (mma.sync.m8n8k4.f64, "DMMA") issue beside the DFMA pipe on B200, or through it?
// Three kernels on register operands, no memory traffic: DFMA chains alone, DMMA chains alone, both interleaved in the same warps.
// If the mixed kernel takes max(t_dfma, t_dmma) the pipes are separate (the interpolation H(u) = M H of the boundary kernels could be
// offloaded); if it takes the sum, they share the FP64 datapath.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) k(double* sink, int iters) {
    double x[NF > 0 ? NF : 1], acc[NM > 0 ? NM : 1][2];
    const double m = sink[1 + (threadIdx.x & 1)], b = sink[3];
#pragma unroll
    for (int c = 0; c < (NF > 0 ? NF : 1); ++c) x[c] = 1.0 + c + threadIdx.x * 1e-9;
#pragma unroll
    for (int c = 0; c < (NM > 0 ? NM : 1); ++c) acc[c][0] = acc[c][1] = 1e-3 * c;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int c = 0; c < NF; ++c) x[c] = fma(x[c], m, b);
#pragma unroll
            for (int c = 0; c < NM; ++c) dmma(acc[c], m, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < (NF > 0 ? NF : 1); ++c) s += x[c];
#pragma unroll
    for (int c = 0; c < (NM > 0 ? NM : 1); ++c) s += acc[c][0] + acc[c][1];
    if (s == 123456.789) sink[0] = s;
}

template <int NF, int NM>
double run(const char* name) {
    double* sink; cudaMalloc(&sink, 512);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.999999 - 1e-9 * i; h[3] = 1e-7;
    cudaMemcpy(sink, h, 512, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, blocks = 148 * 8, L = 10;
    k<NF, NM><<<blocks, 256>>>(sink, iters);
    cudaEventRecord(e0);
    for (int i = 0; i < L; ++i) k<NF, NM><<<blocks, 256>>>(sink, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= L;
    const double warps = (double)blocks * 8, trips = (double)iters * 4;
    const double dfma_flops = warps * trips * NF * 32 * 2, dmma_flops = warps * trips * NM * (8.0 * 8 * 4 * 2);
    printf("%-34s %8.3f ms   DFMA %6.2f TFLOP/s   DMMA %6.2f TFLOP/s  (%d DFMA + %d DMMA per trip)\n", name, ms, dfma_flops / ms / 1e9,
           dmma_flops / ms / 1e9, NF, NM);
    cudaFree(sink);
    return ms;
}

int main() {
    double a = run<8, 0>("DFMA only (8 chains)");
    double b = run<0, 8>("DMMA only (8 accumulators)");
    double c = run<8, 8>("both, interleaved");
    double d = run<8, 2>("8 DFMA + 2 DMMA");
    double e = run<16, 0>("DFMA only (16 chains)");
    printf("mixed / (dfma + dmma) = %.2f   mixed / max(dfma, dmma) = %.2f   -> %s\n", c / (a + b), c / (a > b ? a : b),
           c < 0.75 * (a + b) ? "the FP64 tensor path issues beside the DFMA pipe" : "DMMA and DFMA share the FP64 datapath");
    (void)d; (void)e;
    return 0;
}
