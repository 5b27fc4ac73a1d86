// Third issue probe: which instruction classes overlap with the fp64 pipe?  One DFMA per group plus n instructions of
//   kind 0: LOP3 (ALU pipe)   kind 1: IMAD (FMA pipe)   kind 2: FSEL on a runtime predicate (ALU pipe)   kind 3: ISETP+SEL
// If a class overlaps, t(1 DFMA + n X) ~ max(2.2, cost of n X); if the fp64 instruction holds the issue port, ~ 2.2 + n.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int N, int ND>
__global__ void probe(double* out, int iters, double a, double b, int ka, int kb, float fa) {
    double x[8]; int v[8][4]; float f[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; for (int q = 0; q < 4; ++q) { v[i][q] = threadIdx.x * 7 + i + q; f[i][q] = v[i][q]; } }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (ND) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int q = 0; q < N; ++q) {
                if (KIND == 0) v[i][q] = (v[i][q] & ka) ^ kb;                       // LOP3
                if (KIND == 1) v[i][q] = v[i][q] * ka + kb;                         // IMAD
                if (KIND == 2) f[i][q] = (v[i][(q + 1) & 3] > it) ? f[i][q] : fa;   // ISETP (uniform-ish) + FSEL
                if (KIND == 3) v[i][q] = (v[i][q] > kb) ? ka : v[i][q] + 1;         // ISETP + SEL/IADD
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; for (int q = 0; q < 4; ++q) s += v[i][q] + f[i][q]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int N, int ND> float run(double* d, int iters, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<KIND, N, ND><<<148, threads>>>(d, 10, 0.999, 1e-3, 0x7ffffff3, 0x1235, 0.5f);
    cudaEventRecord(e0);
    probe<KIND, N, ND><<<148, threads>>>(d, iters, 0.999, 1e-3, 0x7ffffff3, 0x1235, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int KIND> void row(double* d, const char* name) {
    const int iters = 20000, threads = 512;
    const double groups = 8.0 * iters * (threads / 32.0 / 4.0);
    auto cyc = [&](float ms) { return ms * 1e-3 * 1.965e9 / groups; };
    printf("%-12s alone x1 %.2f x2 %.2f x4 %.2f | with 1 DFMA: +1 %.2f  +2 %.2f  +4 %.2f   (DFMA alone %.2f)\n", name,
           cyc(run<KIND, 1, 0>(d, iters, threads)), cyc(run<KIND, 2, 0>(d, iters, threads)), cyc(run<KIND, 4, 0>(d, iters, threads)),
           cyc(run<KIND, 1, 1>(d, iters, threads)), cyc(run<KIND, 2, 1>(d, iters, threads)), cyc(run<KIND, 4, 1>(d, iters, threads)),
           cyc(run<KIND, 0, 1>(d, iters, threads)));
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    printf("cycles per group of {1 DFMA, n X} per scheduler, 512 threads per SM, 8 independent chains per thread\n");
    row<0>(d, "LOP3");
    row<1>(d, "IMAD");
    row<2>(d, "ISETP+FSEL");
    row<3>(d, "ISETP+SEL");
    return 0;
}
