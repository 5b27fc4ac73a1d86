// Micro-benchmark: how many issue slots does an fp64 instruction take on this SM?
// K0: NF independent DFMA chains per thread.  K1: the same plus NI integer ops per DFMA.
// If an fp64 warp instruction blocks the scheduler's issue port for two cycles, time(K1 with 1 int op per
// DFMA) = 1.5 x time(K0); if integer work hides in the second cycle, time(K1) = time(K0).
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NFP>
__global__ void probe(double* out, int iters, double a, double b, unsigned m) {
    double x[8];
    unsigned v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; v[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (NFP) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int q = 0; q < NI; ++q) v[i] = (v[i] ^ m) + (v[i] >> 3);   // 2 ALU-ish ops (LOP3 + IADD/SHF mix)
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; t += v[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}
template <int NI, int NFP> float run(double* d, int iters, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NI, NFP><<<148, threads>>>(d, 10, 0.999, 1e-3, 0x5555u);
    cudaEventRecord(e0);
    probe<NI, NFP><<<148, threads>>>(d, iters, 0.999, 1e-3, 0x5555u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        const float t_f = run<0, 1>(d, iters, threads), t_fi1 = run<1, 1>(d, iters, threads), t_fi2 = run<2, 1>(d, iters, threads),
                    t_i1 = run<1, 0>(d, iters, threads), t_i2 = run<2, 0>(d, iters, threads);
        const double warps = threads / 32.0 / 4.0;       // per scheduler
        const double dfma = 8.0 * iters * warps;         // warp-DFMAs per scheduler
        printf("threads %4d: DFMA only %.3f ms (%.3f clk/DFMA at 1.965 GHz) | +1 int pair %.3f ms | +2 int pairs %.3f ms | int only x1 %.3f x2 %.3f ms\n",
               threads, t_f, t_f * 1e-3 * 1.965e9 / dfma, t_fi1, t_fi2, t_i1, t_i2);
    }
    return 0;
}
