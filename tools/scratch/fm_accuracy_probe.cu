// Accuracy of the MUFU seeds and of the Newton / Halley refinements used by fm_rcp / fm_sqrt (fp64).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ double seed_rcp(double a) { double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a)); return x; }
__device__ double seed_rsq(double a) { double x; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a)); return x; }
__device__ double rcp_newton2(double a) { double x = seed_rcp(a); double e = fma(-a, x, 1.0); x = fma(x, e, x); e = fma(-a, x, 1.0); x = fma(x, e, x); return x; }
__device__ double rcp_halley(double a) { double x = seed_rcp(a); double e = fma(-a, x, 1.0); double t = fma(e, e, e); return fma(x, t, x); }
__device__ double sqrt_old(double a) {
    double y = seed_rsq(a); double g = a * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    const double d = fma(-g, g, a); g = fma(d, h, g); return g; }
__device__ double sqrt_halley(double a) {
    double y = seed_rsq(a); double g = a * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5); double t = fma(1.5 * r, r, r); return fma(g, t, g); }
__device__ double sqrt_halley_fix(double a) {
    double y = seed_rsq(a); double g = a * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5); double t = fma(1.5 * r, r, r); g = fma(g, t, g);
    const double d = fma(-g, g, a); return fma(d, h, g); }
__device__ double ulps(double got, double want) { return fabs(got - want) / (fabs(want) * 2.220446049250313e-16); }
__global__ void probe(double* out, int n) {
    unsigned long long s = 88172645463325252ull + 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x);
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double mant = 1.0 + (double)(s >> 11) * (1.0 / 9007199254740992.0);
        const int ex = (int)((s >> 3) % 80) - 40;
        const double a = ldexp(mant, ex);
        const double r = 1.0 / a, q = sqrt(a);
        m[0] = fmax(m[0], fabs(seed_rcp(a) * a - 1.0)); m[1] = fmax(m[1], fabs(seed_rsq(a) * seed_rsq(a) * a - 1.0));
        m[2] = fmax(m[2], ulps(rcp_newton2(a), r)); m[3] = fmax(m[3], ulps(rcp_halley(a), r));
        m[4] = fmax(m[4], ulps(sqrt_old(a), q)); m[5] = fmax(m[5], ulps(sqrt_halley(a), q)); m[6] = fmax(m[6], ulps(sqrt_halley_fix(a), q));
    }
    for (int j = 0; j < 8; ++j) out[(blockIdx.x * blockDim.x + threadIdx.x) * 8 + j] = m[j];
}
int main() {
    const int T = 148 * 256;
    double* d; cudaMalloc(&d, T * 8 * 8);
    probe<<<148, 256>>>(d, 20000);
    double* h = new double[T * 8]; cudaMemcpy(h, d, T * 8 * 8, cudaMemcpyDeviceToHost);
    double m[8] = {0};
    for (int i = 0; i < T; ++i) for (int j = 0; j < 8; ++j) m[j] = fmax(m[j], h[i * 8 + j]);
    printf("seed rel err: rcp %.3e rsqrt(y^2 a - 1) %.3e\nmax ulp: rcp newton2 %.3f  rcp halley %.3f | sqrt old %.3f  sqrt halley %.3f  sqrt halley+fix %.3f\n",
           m[0], m[1], m[2], m[3], m[4], m[5], m[6]);
    return 0;
}
