// Issue probe for Blackwell's packed fp32 arithmetic: does FFMA2 (fma.rn.f32x2) cost one issue slot or two, and does it
// overlap with ALU-pipe instructions?  Groups of {NF fp32 instructions of one kind, N LOP3}, cycles per group per scheduler.
//   kind 0: FFMA (scalar)   kind 1: FFMA2 (one instruction, two lanes of work)   kind 2: two scalar FFMA (same work as kind 1)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long x, unsigned long long y, unsigned long long z) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z)); return r;
}
template <int KIND, int N>
__global__ void probe(float* out, int iters, float a, float b, int ka, int kb) {
    float x[8], y[8]; unsigned long long p[8]; int v[8][4];
    const unsigned long long pa = (static_cast<unsigned long long>(__float_as_uint(a)) << 32) | __float_as_uint(a);
    const unsigned long long pb = (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(b);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = x[i] + 0.5f; p[i] = (static_cast<unsigned long long>(__float_as_uint(x[i])) << 32) | __float_as_uint(y[i]);
        for (int q = 0; q < 4; ++q) v[i][q] = threadIdx.x * 7 + i + q; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) x[i] = fmaf(x[i], a, b);
            if (KIND == 1) p[i] = ffma2(p[i], pa, pb);
            if (KIND == 2) { x[i] = fmaf(x[i], a, b); y[i] = fmaf(y[i], a, b); }
#pragma unroll
            for (int q = 0; q < N; ++q) v[i][q] = (v[i][q] & ka) ^ kb;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i] + y[i] + __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32)); for (int q = 0; q < 4; ++q) s += v[i][q]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int N> float run(float* d, int iters, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<KIND, N><<<148, threads>>>(d, 10, 0.999f, 1e-3f, 0x7ffffff3, 0x1235);
    cudaEventRecord(e0);
    probe<KIND, N><<<148, threads>>>(d, iters, 0.999f, 1e-3f, 0x7ffffff3, 0x1235);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int KIND> void row(float* d, const char* name) {
    const int iters = 20000, threads = 512;
    const double groups = 8.0 * iters * (threads / 32.0 / 4.0);
    auto cyc = [&](float ms) { return ms * 1e-3 * 1.965e9 / groups; };
    printf("%-14s alone %.2f | +1 LOP3 %.2f  +2 LOP3 %.2f  +4 LOP3 %.2f\n", name, cyc(run<KIND, 0>(d, iters, threads)),
           cyc(run<KIND, 1>(d, iters, threads)), cyc(run<KIND, 2>(d, iters, threads)), cyc(run<KIND, 4>(d, iters, threads)));
}
int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    printf("cycles per group per scheduler, 512 threads per SM, 8 independent chains per thread\n");
    row<0>(d, "FFMA");
    row<1>(d, "FFMA2");
    row<2>(d, "2 x FFMA");
    return 0;
}
