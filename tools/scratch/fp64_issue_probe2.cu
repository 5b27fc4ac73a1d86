// Second issue probe: does an fp64 instruction hold the scheduler's issue port for two cycles?
// Mix DFMA with FFMA (full-rate pipe, one issue cycle, no pipe contention with fp64 or the ALU pipe).
//   hypothesis A (pipes overlap, issue 1/clk):  t(1 DFMA + n FFMA) = max(2.2, 1 + n) cycles
//   hypothesis B (fp64 blocks issue 2 cycles):  t = 2.2 + n cycles
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int ND>
__global__ void probe(double* out, int iters, double a, double b, float c, float d) {
    double x[8]; float v[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; for (int q = 0; q < 4; ++q) v[i][q] = threadIdx.x + i + q; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (ND) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int q = 0; q < NF; ++q) v[i][q] = fmaf(v[i][q], c, d);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; for (int q = 0; q < 4; ++q) s += v[i][q]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF, int ND> float run(double* d, int iters, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NF, ND><<<148, threads>>>(d, 10, 0.999, 1e-3, 0.999f, 1e-3f);
    cudaEventRecord(e0);
    probe<NF, ND><<<148, threads>>>(d, iters, 0.999, 1e-3, 0.999f, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    const int iters = 20000;
    for (int threads : {256, 512, 1024}) {
        const double groups = 8.0 * iters * (threads / 32.0 / 4.0);   // per scheduler
        auto cyc = [&](float ms) { return ms * 1e-3 * 1.965e9 / groups; };
        printf("threads %4d: cycles per group  DFMA %.2f | FFMA x1 %.2f x2 %.2f x4 %.2f | DFMA+1 FFMA %.2f  +2 %.2f  +4 %.2f\n", threads,
               cyc(run<0, 1>(d, iters, threads)), cyc(run<1, 0>(d, iters, threads)), cyc(run<2, 0>(d, iters, threads)), cyc(run<4, 0>(d, iters, threads)),
               cyc(run<1, 1>(d, iters, threads)), cyc(run<2, 1>(d, iters, threads)), cyc(run<4, 1>(d, iters, threads)));
    }
    return 0;
}
