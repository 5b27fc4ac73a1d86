// hp_pair.cuh -- two adjacent columns per lane: the value type of the "wide" marching kernels (hp_march_pair.cuh).
//
// A lane of a wide kernel owns the columns (2 l, 2 l + 1) of its warp's 64-column strip and carries every quantity
// as a pair P2<R>{a, b}.  Everything a marching kernel pays once per ROW and lane -- ring slots and barriers,
// validity predicates, flag bookkeeping, addresses, the loop itself, the shuffles that carry faces between lanes --
// is then paid once per TWO cell-updates, the face between the two columns never leaves the thread, and a pair is
// loaded and stored as ONE 16-byte (fp64) / 8-byte (fp32) access.  In fp32 the arithmetic on a pair is Blackwell's
// packed f32x2 (fma/add/mul.rn.f32x2 -> FFMA2 / FADD2 / FMUL2 in SASS: one issue slot for both columns; measured with
// tools/scratch/f32x2_issue_probe.cu: a 1:1 mix of FFMA2 and ALU-pipe instructions runs at 1.5 cycles per pair where
// two scalar FFMAs take 2.0).  Every operation is the IEEE round-to-nearest one of the scalar code, component by
// component; in fp64 a pair is simply two doubles.
#pragma once

#include <cstdint>

#include "hp_fast_kernels.cuh"

namespace HP_NS {

template <class R> struct P2 { R a, b; };
struct B2 { bool a, b; };

template <class R> __device__ __forceinline__ P2<R> splat(R s) { return P2<R>{s, s}; }
__device__ __forceinline__ bool any(B2 m) { return m.a | m.b; }
__device__ __forceinline__ B2 operator!(B2 m) { return B2{!m.a, !m.b}; }
__device__ __forceinline__ B2 operator&(B2 x, B2 y) { return B2{x.a && y.a, x.b && y.b}; }
__device__ __forceinline__ B2 operator|(B2 x, B2 y) { return B2{x.a || y.a, x.b || y.b}; }
__device__ __forceinline__ B2 operator&(B2 x, bool y) { return B2{x.a && y, x.b && y}; }

// ---- fp64: two doubles (contraction is written out, so both precisions evaluate the same expression tree) ---------
__device__ __forceinline__ P2<double> operator+(P2<double> x, P2<double> y) { return {x.a + y.a, x.b + y.b}; }
__device__ __forceinline__ P2<double> operator-(P2<double> x, P2<double> y) { return {x.a - y.a, x.b - y.b}; }
__device__ __forceinline__ P2<double> operator*(P2<double> x, P2<double> y) { return {x.a * y.a, x.b * y.b}; }
__device__ __forceinline__ P2<double> fma2(P2<double> x, P2<double> y, P2<double> z) { return {fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)}; }

// ---- fp32: packed f32x2 (the register pair IS the 64-bit operand; packing and unpacking cost nothing) -----------------
__device__ __forceinline__ uint64_t pk2(P2<float> v) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.a), "f"(v.b));
    return r;
}
__device__ __forceinline__ P2<float> up2(uint64_t r) {
    P2<float> v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.a), "=f"(v.b) : "l"(r));
    return v;
}
__device__ __forceinline__ P2<float> operator+(P2<float> x, P2<float> y) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(x)), "l"(pk2(y)));
    return up2(r);
}
__device__ __forceinline__ P2<float> operator-(P2<float> x, P2<float> y) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(x)), "l"(pk2(y)));
    return up2(r);
}
__device__ __forceinline__ P2<float> operator*(P2<float> x, P2<float> y) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(x)), "l"(pk2(y)));
    return up2(r);
}
__device__ __forceinline__ P2<float> fma2(P2<float> x, P2<float> y, P2<float> z) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(x)), "l"(pk2(y)), "l"(pk2(z)));
    return up2(r);
}

// ---- component-wise pieces common to both ---------------------------------------------------------------------------
template <class R> __device__ __forceinline__ P2<R> operator-(P2<R> x) { return P2<R>{-x.a, -x.b}; }
template <class R> __device__ __forceinline__ P2<R> operator*(R s, P2<R> x) { return splat(s) * x; }
template <class R> __device__ __forceinline__ P2<R> pabs(P2<R> x) { return P2<R>{hp_abs(x.a), hp_abs(x.b)}; }
template <class R> __device__ __forceinline__ P2<R> pmax(P2<R> x, P2<R> y) { return P2<R>{fm_max(x.a, y.a), fm_max(x.b, y.b)}; }
template <class R> __device__ __forceinline__ P2<R> pmin(P2<R> x, P2<R> y) { return P2<R>{fm_min(x.a, y.a), fm_min(x.b, y.b)}; }
template <class R> __device__ __forceinline__ P2<R> sel(B2 m, P2<R> x, P2<R> y) { return P2<R>{m.a ? x.a : y.a, m.b ? x.b : y.b}; }
template <class R> __device__ __forceinline__ B2 operator<(P2<R> x, P2<R> y) { return B2{x.a < y.a, x.b < y.b}; }
template <class R> __device__ __forceinline__ B2 operator>(P2<R> x, P2<R> y) { return B2{x.a > y.a, x.b > y.b}; }
template <class R> __device__ __forceinline__ B2 operator<=(P2<R> x, P2<R> y) { return B2{x.a <= y.a, x.b <= y.b}; }
template <class R> __device__ __forceinline__ B2 operator<(P2<R> x, R s) { return B2{x.a < s, x.b < s}; }
template <class R> __device__ __forceinline__ B2 operator>(P2<R> x, R s) { return B2{x.a > s, x.b > s}; }
template <class R> __device__ __forceinline__ B2 operator<=(P2<R> x, R s) { return B2{x.a <= s, x.b <= s}; }
template <class R> __device__ __forceinline__ B2 operator==(P2<R> x, P2<R> y) { return B2{x.a == y.a, x.b == y.b}; }

// ---- elementary functions: one MUFU seed per component, the refinement on the pair ---------------------------------------
// (same steps as fm_rcp / fm_sqrt_pos / fm_rcbrt of hp_math.cuh and hp_fast_kernels.cuh)
__device__ __forceinline__ P2<double> prcp(P2<double> a) { return {fm_rcp(a.a), fm_rcp(a.b)}; }
__device__ __forceinline__ P2<float> prcp(P2<float> a) {
    P2<float> x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x.a) : "f"(a.a));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x.b) : "f"(a.b));
    const P2<float> e = fma2(-a, x, splat(1.0f));
    return fma2(x, e, x);
}
// sqrt of a strictly positive pair
__device__ __forceinline__ P2<double> psqrt_pos(P2<double> a) { return {fm_sqrt_pos(a.a), fm_sqrt_pos(a.b)}; }
__device__ __forceinline__ P2<float> psqrt_pos(P2<float> a) {
    P2<float> y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.a) : "f"(a.a));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.b) : "f"(a.b));
    const P2<float> g = a * y, h = 0.5f * y;
    return fma2(fma2(-g, g, a), h, g);
}
// a^(-1/3) of a strictly positive pair
__device__ __forceinline__ P2<double> prcbrt(P2<double> a) { return {fm_rcbrt(a.a), fm_rcbrt(a.b)}; }
__device__ __forceinline__ P2<float> prcbrt(P2<float> a) {
    // seed 2^(-lg2(a) / 3) from the two MUFU units (~2e-7), one Newton step y + y (1 - a y^3) / 3 on the pair
    const P2<float> y{fm_rcbrt_seed(a.a), fm_rcbrt_seed(a.b)};
    const P2<float> e = fma2(-(a * y), y * y, splat(1.0f));
    return fma2(y * e, splat(0.333333343f), y);
}

// a pair from shared memory / to global memory as one vector access
template <class R> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };
template <class R> __device__ __forceinline__ P2<R> ld_pair(const void* p) {
    const typename Vec2<R>::type v = *reinterpret_cast<const typename Vec2<R>::type*>(p);
    return P2<R>{v.x, v.y};
}
template <class R> __device__ __forceinline__ void st_pair(R* p, P2<R> v) {
    typename Vec2<R>::type w;
    w.x = v.a; w.y = v.b;
    *reinterpret_cast<typename Vec2<R>::type*>(p) = w;
}

}  // namespace HP_NS
