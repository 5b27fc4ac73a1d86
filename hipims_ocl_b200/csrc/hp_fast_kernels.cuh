// hp_fast_kernels.cuh -- the throughput path: TMA-staged tiles, shared faces, low-op-count maths.
//
// Included by hp_kernels.cu for the "fast" flavour only.  Same per-cell semantics as the
// reference-order kernels (hp_math.cuh), restructured for the B200:
//
//   * persistent CTAs (a multiple of the SM count), each looping over 64x8-cell tiles;
//   * the haloed tile of eta, qx, qy, zb arrives in shared memory through TMA
//     (cp.async.bulk.tensor.2d + mbarrier), double buffered, so the next tile streams in while
//     the current one is computed and every plane is read from HBM once per step;
//   * every cell FACE is solved once per step and shared by its two cells through shared
//     memory -- the reference solves each face twice.  Only the owner-relative
//     hydrostatic term differs between the two cells; it is added per cell in closed form;
//   * divisions and square roots are the cost of the reference's arithmetic on this machine
//     (4.8k instructions per cell-update measured, profiles/r01_v1_godunov_f64_4096.txt): the
//     velocities are formed once per cell, the HLLC star speed |a + du/4| needs no square root,
//     the two middle-state quotients share one reciprocal, the sign of S_M needs none, the bed
//     slope source and the z^2 part of the pressure flux cancel analytically, and the
//     remaining reciprocals / roots are MUFU seeds + Newton steps.  All of it is algebraically
//     identical to the reference; results agree to rounding (tests/test_cuda_parity.py).
#pragma once

#include <cuda.h>

#include "hp_math.cuh"

namespace HP_NS {

// max / min as one compare + select (fmax/fmin also canonicalise NaNs, which costs instructions)
template <class R> __device__ __forceinline__ R fm_max(R a, R b) { return a > b ? a : b; }
template <class R> __device__ __forceinline__ R fm_min(R a, R b) { return a < b ? a : b; }
// ... except in fp32, where FMNMX is a single instruction
template <> __device__ __forceinline__ float fm_max<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ float fm_min<float>(float a, float b) { return fminf(a, b); }
// max(v, 0) and max(a - b, 0): the non-negative part of a reconstructed depth.  The fp64 form is written on (a, b)
// because the compiler then tests a > b beside the subtraction instead of after it (measured: 5 % of the kernel)
template <class R> __device__ __forceinline__ R fm_pos(R v) { return v > R(0) ? v : R(0); }
template <> __device__ __forceinline__ float fm_pos<float>(float v) { return fmaxf(v, 0.0f); }
template <class R> __device__ __forceinline__ R fm_posdiff(R a, R b) { return (a - b > R(0)) ? (a - b) : R(0); }
template <> __device__ __forceinline__ float fm_posdiff<float>(float a, float b) { return fmaxf(a - b, 0.0f); }
// |v| < eps => 0: the reference's "round delta values to zero if small" (CLSchemeGodunov.clc:340-348)
template <class R> __device__ __forceinline__ R fm_chop(R v, R eps) { return hp_abs(v) < eps ? R(0) : v; }

// a * b + c in one rounding, written out where the compiler's contraction would share the product instead
__device__ __forceinline__ double hp_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float hp_fma(float a, float b, float c) { return fmaf(a, b, c); }

// ---- forms the compiler cannot turn into max.f64 / min.f64 ---------------------------------------------------------
// NVVM rewrites `v > 0 ? v : 0` into max.f64 and `a < c || b < c` into `min.f64(a, b) < c`; on this part either one
// becomes DSETP.MAX/MIN + five register moves + SEL + FSEL + LOP3 (NaN quieting) -- ten issue slots for what a compare
// and two selects do (profiles/r02_pluvial16384.txt: six of them per cell-update, ~5 % of the kernel's issue cycles).
// Non-negative part by the sign bit (-0 and negative values give +0 like the comparison form; no NaNs reach it):
__device__ __forceinline__ double fm_pos_s(double v) { return __double2hiint(v) < 0 ? 0.0 : v; }
__device__ __forceinline__ float fm_pos_s(float v) { return fmaxf(v, 0.0f); }
// a < b as an opaque predicate (kept out of the min/max pattern matcher)
__device__ __forceinline__ bool fm_lt_opaque(double a, double b) {
    int r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(r) : "d"(a), "d"(b));
    return r != 0;
}
__device__ __forceinline__ bool fm_lt_opaque(float a, float b) { return a < b; }
__device__ __forceinline__ bool fm_le_opaque(double a, double b) {
    int r;
    asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %1, %2;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(r) : "d"(a), "d"(b));
    return r != 0;
}
__device__ __forceinline__ bool fm_le_opaque(float a, float b) { return a <= b; }

// ---------------------------------------------------------------------------------------------
// TMA / mbarrier primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y) : "memory");
}

#ifndef HP_GODUNOV_STAGES
#define HP_GODUNOV_STAGES 1
#endif
#ifndef HP_F32_CTAS
#define HP_F32_CTAS 9
#endif

// ---------------------------------------------------------------------------------------------
// Tile geometry
// ---------------------------------------------------------------------------------------------
template <class R> struct Tile {
    static constexpr int TX = hp::kTmaTileX, TY = hp::kTmaTileY, NT = TX * TY / 2;   // two cells per thread
    // STAGES = 2: the next tile streams in during compute (16 warps/SM in fp64).  STAGES = 1: no ring,
    // shared memory then allows 24 warps/SM; the next tile is prefetched into L2 instead.
    static constexpr int STAGES = HP_GODUNOV_STAGES;
    static constexpr int CTAS_PER_SM = sizeof(R) == 8 ? (STAGES == 1 ? 6 : 4) : (STAGES == 1 ? HP_F32_CTAS : 8);
    // TMA wants the box to START on a 16-byte boundary of the inner dimension (measured on this
    // part: a start coordinate of x0-1 raises "illegal instruction") and its inner extent to be a
    // multiple of 16 bytes, so the halo columns are padded to CO = 16 / sizeof(R) on both sides.
    static constexpr int CO = 16 / int(sizeof(R));                // smem column of the tile's first cell
    static constexpr int BW = TX + 2 * CO;
    static constexpr int BH = TY + 2;
    static constexpr int PLANE_BYTES = (BW * BH * int(sizeof(R)) + 127) / 128 * 128;
    static constexpr int STAGE_BYTES = 4 * PLANE_BYTES;           // eta, qx, qy, zb
    static constexpr int NXF = (TX + 1) * TY, NYF = TX * (TY + 1);
    static constexpr int FX_BYTES = 3 * NXF * int(sizeof(R)), FY_BYTES = 3 * NYF * int(sizeof(R));
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 3 * PLANE_BYTES + FX_BYTES + FY_BYTES + 64;
};

struct TmaMaps { CUtensorMap eta, qx, qy, zb; };

// One shared face in the normal frame: state of the two sides -> core flux {m, n, t}.
// "core" = without the -g/2 z'^2 part of the pressure, which is owner specific and handled in
// closed form by the cell update (see header comment).
// sqrt for a STRICTLY positive argument: no zero guard (rsqrt(0) = inf would give 0 * inf)
__device__ __forceinline__ double fm_sqrt_pos(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y;
    const double h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    return fma(g, fma(1.5 * r, r, r), g);
}
__device__ __forceinline__ float fm_sqrt_pos(float a) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    const float g = a * y, h = 0.5f * y;
    return fmaf(fmaf(-g, g, a), h, g);
}
// sqrt(g h) for h >= 0 without a zero guard: the radicand is lifted by a denormal-free tiny constant inside the
// multiply-add, so h = 0 yields 1e-150 (fp32: 3e-19) instead of 0 * inf -- absorbed by every sum it enters
__device__ __forceinline__ double fm_celerity(double g, double h) { return fm_sqrt_pos(fma(g, h, 1.0e-300)); }
__device__ __forceinline__ float fm_celerity(float g, float h) { return fm_sqrt_pos(fmaf(g, h, 1.0e-37f)); }

template <class R, bool CACHED_CELERITY = true>
__device__ __forceinline__ Flux3<R> face_core_flux(const Params<R>& k, R etaL, R zL, R unL, R utL, R cL, R etaR, R zR, R unR,
                                                   R utR, R cR) {
    const R hg = R(0.5) * k.g;
    const R zmax = fm_max(zL, zR);
    const R dL = etaL - zmax, dR = etaR - zmax;
    if (__all_sync(__activemask(), dL > k.eps && dR > k.eps)) {
        // wet fast path: no clamps, no dry-side selects, no zero guard on the roots; the same operations the general
        // path below performs when both depths exceed the threshold.  Chosen by the lanes that are here together: a warp
        // with a single dry side takes the general path as a whole instead of executing both.
        const R aL = (CACHED_CELERITY && zmax == zL) ? cL : fm_sqrt_pos(k.g * dL);
        const R aR = (CACHED_CELERITY && zmax == zR) ? cR : fm_sqrt_pos(k.g * dR);
        const R qnL = dL * unL, qnR = dR * unR;
        const R as = hp_abs(R(0.5) * (aL + aR) + R(0.25) * (unL - unR));
        const R us = R(0.5) * (unL + unR) + aL - aR;
        const R sL = fm_min(unL - aL, us - as);
        const R sR = fm_max(unR + aR, us + as);
        const R FLn = unL * qnL + hg * dL * dL, FRn = unR * qnR + hg * dR * dR;
        if (sL >= R(0)) return Flux3<R>{qnL, FLn, qnL * utL};
        if (!(sR >= R(0))) return Flux3<R>{qnR, FRn, qnR * utR};
        const R inv = fm_rcp(sR - sL);
        const R ss = sL * sR;
        const R f1 = (sR * qnL - sL * qnR + ss * (dR - dL)) * inv;
        const R f2 = (sR * FLn - sL * FRn + ss * (qnR - qnL)) * inv;
        return Flux3<R>{f1, f2, f1 * (f1 >= R(0) ? utL : utR)};
    }
    const R hL = fm_pos_s(dL);
    const R hR = fm_pos_s(dR);
    const bool dryL = hL < k.eps, dryR = hR < k.eps;
    if (dryL && dryR) {
        const R hm = R(0.5) * (hL + hR);
        return Flux3<R>{R(0), hg * hm * hm, R(0)};
    }
    if (dryL) { unL = R(0); utL = R(0); }
    if (dryR) { unR = R(0); utR = R(0); }
    // celerity: the cell's own sqrt(g h) is reused whenever the face sits on the cell's own bed
    const R aL = (CACHED_CELERITY && zmax == zL) ? cL : fm_celerity(k.g, hL);
    const R aR = (CACHED_CELERITY && zmax == zR) ? cR : fm_celerity(k.g, hR);
    const R qnL = hL * unL, qnR = hR * unR;
    const R as = hp_abs(R(0.5) * (aL + aR) + R(0.25) * (unL - unR));     // sqrt(g h*) without the sqrt
    const R us = R(0.5) * (unL + unR) + aL - aR;
    const R sL = dryL ? unR - 2 * aR : fm_min(unL - aL, us - as);
    const R sR = dryR ? unL + 2 * aL : fm_max(unR + aR, us + as);
    const Flux3<R> FL{qnL, unL * qnL + hg * hL * hL, qnL * utL};
    const Flux3<R> FR{qnR, unR * qnR + hg * hR * hR, qnR * utR};
    if (sL >= R(0)) return FL;
    if (!(sR >= R(0))) return FR;
    const R inv = fm_rcp(sR - sL);
    const R ss = sL * sR;
    const R f1 = (sR * FL.m - sL * FR.m + ss * (hR - hL)) * inv;
    const R f2 = (sR * FL.n - sL * FR.n + ss * (qnR - qnL)) * inv;
    // The contact speed S_M = num / den has num = -f1 (sR - sL) and den = hR (unR - sR) - hL (unL - sL) < 0
    // whenever sL < unL and unR < sR (true by construction of the wave speeds), so sign(S_M) = sign(f1):
    // the tangential momentum is upwinded by the direction of the mass flux, no quotient needed.
    const bool smPos = f1 >= R(0);
    return Flux3<R>{f1, f2, f1 * (smPos ? utL : utR)};
}

// Owner-side bookkeeping of one face for the cell update: reconstructed bed seen by the owner,
// the neighbour side's reconstructed depth, and the stop-counter increments
// (CLSchemeGodunov.clc:83-137).
template <class R, bool ownIsLeft>
__device__ __forceinline__ void face_owner_terms(const Params<R>& k, R etaOwn, R zOwn, R unOwn, R ownQn, R etaNb, R zNb, R unNb,
                                                 R& bed, R& hNb, int& stop) {
    const R zmax = fm_max(zOwn, zNb);
    const R hOwn = fm_posdiff(etaOwn, zmax);
    hNb = fm_posdiff(etaNb, zmax);
    bed = fm_min(zmax, etaOwn);                                  // zmax - max(0, zmax - eta_own)
    if (fm_le_opaque(hOwn, k.eps) | fm_le_opaque(hNb, k.eps)) {           // only possible at wet/dry fronts
        const R hL = ownIsLeft ? hOwn : hNb, hR = ownIsLeft ? hNb : hOwn;
        const R unL = ownIsLeft ? unOwn : unNb, unR = ownIsLeft ? unNb : unOwn;
        if (ownIsLeft) { if (hL <= k.eps && ownQn > R(0)) ++stop; }
        else           { if (hR <= k.eps && ownQn < R(0)) ++stop; }
        if (hR <= k.eps && unL < R(0)) ++stop;
        if (hL <= k.eps && unR > R(0)) ++stop;
    }
}

// Point-implicit friction with one reciprocal per component; the "cannot reverse the flow" clamp
// of the reference (CLFriction.clc:51-65) can never bind because 2qx^2+qy^2 >= qx^2+qy^2.
template <class R> __device__ __forceinline__ void friction_fast(const Params<R>& k, R h, R rh, R& qx, R& qy, R n, R dt) {
    const R q2 = qx * qx + qy * qy;
    const R q = fm_sqrt(q2);
    if (fm_lt_opaque(h, k.eps) | fm_lt_opaque(q, k.eps)) return;
    const R A = dt * k.g * n * n * rh * rh * fm_rcbrt(h);                 // dt * Cf / h^2
    const R aq2 = A * q2;
    qx = qx - qx * aq2 * fm_rcp(q + A * (q2 + qx * qx));
    qy = qy - qy * aq2 * fm_rcp(q + A * (q2 + qy * qy));
}

// ---------------------------------------------------------------------------------------------
// Godunov step on TMA-staged tiles with shared faces.
// ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(Tile<R>::NT, Tile<R>::CTAS_PER_SM)
godunov_step_tma(const StepArgs a, const __grid_constant__ TmaMaps maps) {
    using T = Tile<R>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = smem_raw;
    // [stage0: eta qx qy zb][stage1: ...][u][v][c][FX][FY][barriers]
    R* const s_u = reinterpret_cast<R*>(base + T::STAGES * T::STAGE_BYTES);
    R* const s_v = reinterpret_cast<R*>(base + T::STAGES * T::STAGE_BYTES + T::PLANE_BYTES);
    R* const s_c = reinterpret_cast<R*>(base + T::STAGES * T::STAGE_BYTES + 2 * T::PLANE_BYTES);
    R* const s_fx = reinterpret_cast<R*>(base + T::STAGES * T::STAGE_BYTES + 3 * T::PLANE_BYTES);
    R* const s_fy = reinterpret_cast<R*>(base + T::STAGES * T::STAGE_BYTES + 3 * T::PLANE_BYTES + T::FX_BYTES);
    uint64_t* const s_bar = reinterpret_cast<uint64_t*>(base + T::STAGES * T::STAGE_BYTES + 3 * T::PLANE_BYTES + T::FX_BYTES + T::FY_BYTES);

    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const int tid = threadIdx.x;
    const R dt = read_timestep<R>(a.clock);
    const R inv_delta = fm_rcp(k.delta);
    const R hg = R(0.5) * k.g;

    const int tiles_x = (g.cols + T::TX - 1) / T::TX;
    const int tiles_y = (a.y1 - a.y0 + T::TY - 1) / T::TY;
    const int ntiles = tiles_x * tiles_y;
    const double inv_tiles_x = 1.0 / tiles_x;
    auto tile_row = [&](int tile) { return static_cast<int>((tile + 0.5) * inv_tiles_x); };   // exact for < 2^31 tiles

    const uint32_t bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile, int stage) {   // one thread: arm the barrier, launch the four plane loads
        const int ty = tile_row(tile), tx = tile - ty * tiles_x;
        const int x = tx * T::TX - T::CO, y = a.y0 + ty * T::TY - 1;
        const uint32_t bar = stage ? bar1 : bar0;
        const uint32_t dst = smem_u32(base + stage * T::STAGE_BYTES);
        mbar_expect_tx(bar, 4u * T::BW * T::BH * sizeof(R));
        tma_load_2d(dst + 0 * T::PLANE_BYTES, &maps.eta, x, y, bar);
        tma_load_2d(dst + 1 * T::PLANE_BYTES, &maps.qx, x, y, bar);
        tma_load_2d(dst + 2 * T::PLANE_BYTES, &maps.qy, x, y, bar);
        tma_load_2d(dst + 3 * T::PLANE_BYTES, &maps.zb, x, y, bar);
    };

    auto prefetch_l2 = [&](int tile) {        // warm L2 with the next tile's boxes
        const int ty = tile_row(tile), tx = tile - ty * tiles_x;
        const int x = tx * T::TX - T::CO, y = a.y0 + ty * T::TY - 1;
        tma_prefetch_2d(&maps.eta, x, y); tma_prefetch_2d(&maps.qx, x, y); tma_prefetch_2d(&maps.qy, x, y); tma_prefetch_2d(&maps.zb, x, y);
    };

    const View<R> s(a.src);
    const MutView<R> d(a.dst);
    const R* __restrict__ mann = static_cast<const R*>(a.manning);

    R ws = R(0);
    uint32_t phase0 = 0, phase1 = 0;
    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue(tile, 0);

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int stage = T::STAGES == 2 ? (it & 1) : 0;
        const int next = tile + gridDim.x;
        if (T::STAGES == 2) {
            if (tid == 0 && next < ntiles) issue(next, stage ^ 1);
        } else {
            if (tid == 0 && it > 0) issue(tile, 0);
            if (tid == 32 && next < ntiles) prefetch_l2(next);
        }
        if (stage == 0) { mbar_wait(bar0, phase0); phase0 ^= 1; } else { mbar_wait(bar1, phase1); phase1 ^= 1; }

        const R* const t_eta = reinterpret_cast<const R*>(base + stage * T::STAGE_BYTES);
        const R* const t_qx = reinterpret_cast<const R*>(base + stage * T::STAGE_BYTES + T::PLANE_BYTES);
        const R* const t_qy = reinterpret_cast<const R*>(base + stage * T::STAGE_BYTES + 2 * T::PLANE_BYTES);
        const R* const t_zb = reinterpret_cast<const R*>(base + stage * T::STAGE_BYTES + 3 * T::PLANE_BYTES);
        const int trow = tile_row(tile);
        const int x0 = (tile - trow * tiles_x) * T::TX, y0 = a.y0 + trow * T::TY;

        // point-wise planes (no halo): issue the global loads now, they are consumed in phase D
        R pre_emax[2], pre_mann[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int x = x0 + tid % T::TX, y = y0 + tid / T::TX + half * (T::TY / 2);
            const bool in = x < g.cols && y < a.y1;
            const size_t id = static_cast<size_t>(in ? y : a.y0) * g.pitch + (in ? x : 0);
            pre_emax[half] = s.emax[id];
            pre_mann[half] = k.friction ? mann[id] : R(0);
        }

        // ---- phase B: per-cell velocities and celerity for the tile and its halo ----------------
        for (int i = tid; i < (T::TX + 2) * T::BH; i += T::NT) {
            const int lx = i % (T::TX + 2) + T::CO - 1, ly = i / (T::TX + 2);
            const int o = ly * T::BW + lx;
            const R h = t_eta[o] - t_zb[o];
            const bool wet = !(h < k.eps);
            const R rh = wet ? fm_rcp(h) : R(0);
            s_u[o] = t_qx[o] * rh;
            s_v[o] = t_qy[o] * rh;
            s_c[o] = fm_celerity(k.g, fm_pos_s(h));
        }
        __syncthreads();

        // ---- phase C: every face of the tile once ------------------------------------------------
        if (dt > R(0)) {
            for (int f = tid; f < T::NXF + T::NYF; f += T::NT) {
                if (f < T::NXF) {
                    const int j = f / (T::TX + 1), i = f % (T::TX + 1);           // between local cells (i-1, j) and (i, j)
                    const int oL = (j + 1) * T::BW + i + T::CO - 1, oR = oL + 1;
                    const Flux3<R> F = face_core_flux(k, t_eta[oL], t_zb[oL], s_u[oL], s_v[oL], s_c[oL], t_eta[oR], t_zb[oR],
                                                      s_u[oR], s_v[oR], s_c[oR]);
                    s_fx[f] = F.m; s_fx[T::NXF + f] = F.n; s_fx[2 * T::NXF + f] = F.t;
                } else {
                    const int gq = f - T::NXF;
                    const int j = gq / T::TX, i = gq % T::TX;                     // between local cells (i, j-1) and (i, j)
                    const int oL = j * T::BW + i + T::CO, oR = oL + T::BW;
                    const Flux3<R> F = face_core_flux(k, t_eta[oL], t_zb[oL], s_v[oL], s_u[oL], s_c[oL], t_eta[oR], t_zb[oR],
                                                      s_v[oR], s_u[oR], s_c[oR]);
                    s_fy[gq] = F.m; s_fy[T::NYF + gq] = F.n; s_fy[2 * T::NYF + gq] = F.t;
                }
            }
        }
        __syncthreads();

        // ---- phase D: cell update, two cells per thread -------------------------------------------
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = tid % T::TX, j = tid / T::TX + half * (T::TY / 2);
            const int x = x0 + i, y = y0 + j;
            if (x >= g.cols || y >= a.y1) continue;
            const int o = (j + 1) * T::BW + i + T::CO;
            const size_t id = static_cast<size_t>(y) * g.pitch + x;
            const int gy = y + g.gy0;
            Cell<R> c{t_eta[o], pre_emax[half], t_qx[o], t_qy[o]};
            const R zb = t_zb[o];
            const R u = s_u[o], v = s_v[o];
            if (a.reduce_mode == hp::kReduceSrc) {
                const R h = c.eta - zb;
                if (h > k.eps10 && c.emax > R(-9999.0)) {
                    const R cc = s_c[o];
                    const R sp = k.simplified_speed ? cc : fm_max(hp_abs(u), hp_abs(v)) + cc;
                    ws = fm_max(sp, ws);
                }
            }
            bool wrote = false;
            const bool interior = x >= 1 && x <= g.cols - 2 && gy >= 1 && gy <= g.grows - 2;
            R rh_new = R(0), h_new = R(0);
            bool have_new = false;
            if (interior) {
                if (dt <= R(0)) {
                    wrote = !k.dt0_keep;                                 // CLSchemeGodunov.clc:201-206 (:477-478 with the quirk)
                } else if (c.emax <= R(-9999.0) || c.eta == R(-9999.0)) {
                    wrote = true;
                } else {
                    const int oN = o + T::BW, oS = o - T::BW, oE = o + 1, oW = o - 1;
                    const R etaN = t_eta[oN], etaS = t_eta[oS], etaE = t_eta[oE], etaW = t_eta[oW];
                    const R zN = t_zb[oN], zS = t_zb[oS], zE = t_zb[oE], zW = t_zb[oW];
                    const bool all_dry = c.eta - zb < k.eps && etaN - zN < k.eps && etaE - zE < k.eps && etaS - zS < k.eps &&
                                         etaW - zW < k.eps;                      // dry count of five, CLSchemeGodunov.clc:248-255
                    if (!all_dry) {
                        int stop = 0;
                        R bN, bS, bE, bW, hnN, hnS, hnE, hnW;
                        {
                            // all four faces wet on both sides (the usual case away from fronts): the reconstructed bed the
                            // owner sees IS the face's bed, the neighbour depth needs no clamp, no stop flag can be set
                            const R mN = fm_max(zb, zN), mS = fm_max(zb, zS), mE = fm_max(zb, zE), mW = fm_max(zb, zW);
                            const R hoN = c.eta - mN, hoS = c.eta - mS, hoE = c.eta - mE, hoW = c.eta - mW;
                            hnN = etaN - mN; hnS = etaS - mS; hnE = etaE - mE; hnW = etaW - mW;
                            const bool wet = hoN > k.eps && hoS > k.eps && hoE > k.eps && hoW > k.eps && hnN > k.eps && hnS > k.eps &&
                                             hnE > k.eps && hnW > k.eps;
                            if (__all_sync(__activemask(), wet)) {
                                bN = mN; bS = mS; bE = mE; bW = mW;
                            } else {
                                face_owner_terms<R, true>(k, c.eta, zb, v, c.qy, etaN, zN, s_v[oN], bN, hnN, stop);
                                face_owner_terms<R, false>(k, c.eta, zb, v, c.qy, etaS, zS, s_v[oS], bS, hnS, stop);
                                face_owner_terms<R, true>(k, c.eta, zb, u, c.qx, etaE, zE, s_u[oE], bE, hnE, stop);
                                face_owner_terms<R, false>(k, c.eta, zb, u, c.qx, etaW, zW, s_u[oW], bW, hnW, stop);
                            }
                        }
                        const int fe = j * (T::TX + 1) + i + 1, fw = fe - 1;       // x-faces east / west of (i, j)
                        const int fn = (j + 1) * T::TX + i, fs = fn - T::TX;       // y-faces north / south
                        const R mE = s_fx[fe], mW = s_fx[fw], mN = s_fy[fn], mS = s_fy[fs];
                        const R nE = s_fx[T::NXF + fe], nW = s_fx[T::NXF + fw], tE = s_fx[2 * T::NXF + fe], tW = s_fx[2 * T::NXF + fw];
                        const R nN = s_fy[T::NYF + fn], nS = s_fy[T::NYF + fs], tN = s_fy[2 * T::NYF + fn], tS = s_fy[2 * T::NYF + fs];
                        // flux divergence minus bed-slope source, hydrostatic z^2 terms cancelled analytically
                        R dEta = ((mE - mW) + (mN - mS)) * inv_delta;
                        R dQx = ((nE - nW) + (tN - tS) + hg * (bE - bW) * (hnE + hnW)) * inv_delta;
                        R dQy = ((tE - tW) + (nN - nS) + hg * (bN - bS) * (hnN + hnS)) * inv_delta;
                        if (stop > 0) { c.qx = R(0); c.qy = R(0); }
                        if (!(hp_abs(dEta) < k.eps)) c.eta = c.eta - dt * dEta;       // |D| < eps => 0 (CLSchemeGodunov.clc:340-348)
                        if (!(hp_abs(dQx) < k.eps)) c.qx = c.qx - dt * dQx;
                        if (!(hp_abs(dQy) < k.eps)) c.qy = c.qy - dt * dQy;
                        h_new = c.eta - zb;
                        if (!(h_new < k.eps)) { rh_new = fm_rcp(h_new); have_new = true; }
                        if (k.friction) friction_fast(k, h_new, rh_new, c.qx, c.qy, pre_mann[half], dt);
                        if (c.eta > c.emax && c.emax > R(-9990.0)) c.emax = c.eta;
                        if (h_new < k.eps) c.eta = zb;
                        wrote = true;
                    }
                }
                if (wrote) d.store(id, c);
            }
            if (a.reduce_mode == hp::kReduceDst) {
                if (!wrote) { c.eta = d.eta[id]; c.emax = d.emax[id]; c.qx = d.qx[id]; c.qy = d.qy[id]; have_new = false; }
                const R h = c.eta - zb;
                if (h > k.eps10 && c.emax > R(-9999.0)) {
                    const R cc = fm_sqrt(k.g * h);
                    R sp = cc;
                    if (!k.simplified_speed) {
                        const R rh = have_new ? rh_new : fm_rcp(h);
                        sp = fm_max(hp_abs(c.qx * rh), hp_abs(c.qy * rh)) + cc;
                    }
                    ws = fm_max(sp, ws);
                }
            }
        }
        // generic-proxy accesses to this stage are done; the next TMA write into it is async-proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    block_reduce_finalize<R>(ws, a, k);
}

// kernels with more than 48 KB of dynamic shared memory need the opt-in on every device they run on
constexpr int kMaxDevices = 64;
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

template <class R> static int launch_godunov_tma(const StepArgs& a_in, const TmaMaps& maps, int sm_count, cudaStream_t st) {
    using T = Tile<R>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(godunov_step_tma<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int tiles = ((a.grid.cols + T::TX - 1) / T::TX) * ((a.y1 - a.y0 + T::TY - 1) / T::TY);
    int grid = T::CTAS_PER_SM * sm_count;          // all resident CTAs of every SM, persistent
    if (grid > tiles) grid = tiles;
    a.total_ctas = grid;
    godunov_step_tma<R><<<grid, T::NT, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

// =============================================================================================
// MUSCL-Hancock, predictor + corrector fused, on TMA-staged tiles with shared faces.
//
// Phase A  TMA: eta, qx, qy, zb for the tile and a halo of two cells (the corrector needs the
//          neighbours' predictor output, which needs their neighbours); eta_max only matters as two
//          per-cell flags (boundary cell, "dry" neighbour) and is read with plain loads.
// Phase B  predictor for the tile + one halo cell: MINMOD slopes (no division), half-step evolve.
//          Kept per cell in shared memory: the evolved state and the eight slopes -- the reference
//          writes 4 face vectors per cell to global memory and reads 8 back
//          (src/Schemes/CLSchemeMUSCLHancock.clc:137-143, 600-635).
// Phase C  every face once: both sides' face estimates are rebuilt from (evolved state +- slope/2).
// Phase D  corrector; state ping-pongs and every owned cell is written (see step_mh_v1).
// =============================================================================================
template <class R> struct TileMH {
    static constexpr int TX = hp::kTmaTileXMH, TY = hp::kTmaTileY, NT = 256;
    static constexpr int CTAS_PER_SM = sizeof(R) == 8 ? 2 : 3;
    static constexpr int CO = 16 / int(sizeof(R));                // >= 2 halo columns, box starts 16-byte aligned
    static constexpr int BW = TX + 2 * CO, BH = TY + 4;
    static constexpr int PLANE_BYTES = (BW * BH * int(sizeof(R)) + 127) / 128 * 128;
    static constexpr int FLAG_BYTES = (BW * BH + 127) / 128 * 128;
    static constexpr int PW = TX + 2, PH = TY + 2;                // predictor cells: tile + one halo cell
    static constexpr int PPLANE = PW * PH;                        // elements per predictor plane
    static constexpr int NPRED = 11;                              // eta2 qx2 qy2 | sx(eta,h,qx,qy) | sy(eta,h,qx,qy)
    static constexpr int NXF = (TX + 1) * TY, NYF = TX * (TY + 1), NF = NXF + NYF;
    static constexpr int OFF_FLAGS = 4 * PLANE_BYTES;
    static constexpr int OFF_PRED = OFF_FLAGS + FLAG_BYTES;
    static constexpr int OFF_FLUX = OFF_PRED + (NPRED * PPLANE * int(sizeof(R)) + 127) / 128 * 128;
    static constexpr int OFF_BAR = OFF_FLUX + (3 * NF * int(sizeof(R)) + 127) / 128 * 128;
    static constexpr int SMEM_BYTES = OFF_BAR + 64;
};

// phi(r) a with r = b / a, phi = max(0, min(r, 1)) (CLSlopeLimiterMINMOD.clc:49-70, beta = 1): the argument of
// smaller magnitude when the signs agree, else 0.  A zero argument is picked by the magnitude test
// itself, so only the sign bits need comparing (integer pipe instead of an fp64 multiply + compare).
__device__ __forceinline__ double minmod(double a, double b) {
    const double t = fabs(b) < fabs(a) ? b : a;
    return ((__double2hiint(a) ^ __double2hiint(b)) < 0) ? 0.0 : t;
}
__device__ __forceinline__ float minmod(float a, float b) {
    const float t = fabsf(b) < fabsf(a) ? b : a;
    return ((__float_as_int(a) ^ __float_as_int(b)) < 0) ? 0.0f : t;
}
// ... with a switch: `off` is 0, or has its sign bit set to drop the slope altogether (a dry neighbour in that
// direction, CLSchemeMUSCLHancock.clc:301-320) -- it rides along in the sign test's LOP3 for free
__device__ __forceinline__ double minmod_sw(double a, double b, int off) {
    const double t = fabs(b) < fabs(a) ? b : a;
    return (((__double2hiint(a) ^ __double2hiint(b)) | off) < 0) ? 0.0 : t;
}
__device__ __forceinline__ float minmod_sw(float a, float b, int off) {
    const float t = fabsf(b) < fabsf(a) ? b : a;
    return (((__float_as_int(a) ^ __float_as_int(b)) | off) < 0) ? 0.0f : t;
}

template <class R>
__global__ void __launch_bounds__(TileMH<R>::NT, TileMH<R>::CTAS_PER_SM)
mh_step_tma(const StepArgs a, const __grid_constant__ TmaMaps maps) {
    using T = TileMH<R>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = smem_raw;
    const R* const t_eta = reinterpret_cast<const R*>(base);
    const R* const t_qx = reinterpret_cast<const R*>(base + T::PLANE_BYTES);
    const R* const t_qy = reinterpret_cast<const R*>(base + 2 * T::PLANE_BYTES);
    const R* const t_zb = reinterpret_cast<const R*>(base + 3 * T::PLANE_BYTES);
    unsigned char* const s_flag = base + T::OFF_FLAGS;            // bit0: eta_max <= -9998, bit1: eta_max < eps
    R* const s_p = reinterpret_cast<R*>(base + T::OFF_PRED);      // [NPRED][PH][PW]
    R* const s_f = reinterpret_cast<R*>(base + T::OFF_FLUX);      // [3][NF]
    uint64_t* const s_bar = reinterpret_cast<uint64_t*>(base + T::OFF_BAR);

    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const int tid = threadIdx.x;
    const R dt = read_timestep<R>(a.clock);
    const R inv_delta = fm_rcp(k.delta);
    const R hg = R(0.5) * k.g, half = R(0.5);

    const int tiles_x = (g.cols + T::TX - 1) / T::TX;
    const int tiles_y = (a.y1 - a.y0 + T::TY - 1) / T::TY;
    const int ntiles = tiles_x * tiles_y;
    const uint32_t bar = smem_u32(&s_bar[0]);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const View<R> s(a.src);
    const MutView<R> d(a.dst);
    const R* __restrict__ mann = static_cast<const R*>(a.manning);
    R ws = R(0);
    uint32_t phase = 0;

    const double inv_tiles_x = 1.0 / tiles_x;
    auto tile_origin = [&](int tile, int& x0, int& y0) {      // exact without an integer division
        const int ty = static_cast<int>((tile + 0.5) * inv_tiles_x);
        x0 = (tile - ty * tiles_x) * T::TX; y0 = a.y0 + ty * T::TY;
    };
    // eta_max of tile + halo 2 only matters as two flags per cell; the values are loaded one tile ahead
    constexpr int NEM = ((T::TX + 4) * T::BH + T::NT - 1) / T::NT;
    R em_next[NEM];
    auto load_emax = [&](int tile) {
        int x0, y0;
        tile_origin(tile, x0, y0);
#pragma unroll
        for (int r = 0; r < NEM; ++r) {
            const int i = tid + r * T::NT;
            const int x = x0 + i % (T::TX + 4) - 2, y = y0 + i / (T::TX + 4) - 2;
            const bool in = i < (T::TX + 4) * T::BH && x >= 0 && x < g.cols && y >= 0 && y < g.rows;
            em_next[r] = in ? s.emax[static_cast<size_t>(y) * g.pitch + x] : R(1);   // outside: neither flag
        }
    };
    if (static_cast<int>(blockIdx.x) < ntiles) load_emax(blockIdx.x);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int x0, y0;
        tile_origin(tile, x0, y0);
        const int next = tile + gridDim.x;
        if (tid == 0) {
            mbar_expect_tx(bar, 4u * T::BW * T::BH * sizeof(R));
            const uint32_t dst = smem_u32(base);
            tma_load_2d(dst + 0 * T::PLANE_BYTES, &maps.eta, x0 - T::CO, y0 - 2, bar);
            tma_load_2d(dst + 1 * T::PLANE_BYTES, &maps.qx, x0 - T::CO, y0 - 2, bar);
            tma_load_2d(dst + 2 * T::PLANE_BYTES, &maps.qy, x0 - T::CO, y0 - 2, bar);
            tma_load_2d(dst + 3 * T::PLANE_BYTES, &maps.zb, x0 - T::CO, y0 - 2, bar);
        }
        if (tid == 32 && next < ntiles) {          // warm L2 with the next tile's boxes
            int nx0, ny0;
            tile_origin(next, nx0, ny0);
            tma_prefetch_2d(&maps.eta, nx0 - T::CO, ny0 - 2); tma_prefetch_2d(&maps.qx, nx0 - T::CO, ny0 - 2);
            tma_prefetch_2d(&maps.qy, nx0 - T::CO, ny0 - 2); tma_prefetch_2d(&maps.zb, nx0 - T::CO, ny0 - 2);
        }
#pragma unroll
        for (int r = 0; r < NEM; ++r) {
            const int i = tid + r * T::NT;
            if (i < (T::TX + 4) * T::BH) {
                const R em = em_next[r];
                s_flag[(i / (T::TX + 4)) * T::BW + i % (T::TX + 4) + T::CO - 2] =
                    static_cast<unsigned char>((em <= R(-9998.0) ? 1 : 0) | (em < k.eps ? 2 : 0));
            }
        }
        R pre_emax[2], pre_mann[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int x = x0 + tid % T::TX, y = y0 + tid / T::TX + hf * (T::TY / 2);
            const bool in = x < g.cols && y < a.y1;
            const size_t id = static_cast<size_t>(in ? y : a.y0) * g.pitch + (in ? x : 0);
            pre_emax[hf] = s.emax[id];
            pre_mann[hf] = k.friction ? mann[id] : R(0);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        __syncthreads();   // flags visible

        // ---- phase B: predictor -------------------------------------------------------------------
        for (int i = tid; i < T::PPLANE; i += T::NT) {
            const int lx = i % T::PW, ly = i / T::PW;               // predictor-plane coordinates
            const int ci = lx - 1, cj = ly - 1;                     // tile-local cell coordinates
            const int o = (cj + 2) * T::BW + ci + T::CO;            // raw-plane offset
            const int x = x0 + ci, y = y0 + cj, gy = y + g.gy0;
            const R eta = t_eta[o], qx = t_qx[o], qy = t_qy[o], zb = t_zb[o];
            R e2 = eta, qx2 = qx, qy2 = qy;
            R sxE = R(0), sxH = R(0), sxQx = R(0), sxQy = R(0), syE = R(0), syH = R(0), syQx = R(0), syQy = R(0);
            const bool valid = dt > R(0) && x >= 1 && x <= g.cols - 2 && gy >= 1 && gy <= g.grows - 2 && y >= 1 && y <= g.rows - 2;
            const int oN = o + T::BW, oS = o - T::BW, oE = o + 1, oW = o - 1;
            const R h = eta - zb;
            if (valid && !(h < R(1E-5)) && !((s_flag[oN] | s_flag[oE] | s_flag[oS] | s_flag[oW]) & 1)) {
                const R etaE = t_eta[oE], etaW = t_eta[oW], etaN = t_eta[oN], etaS = t_eta[oS];
                const R hE = etaE - t_zb[oE], hW = etaW - t_zb[oW], hN = etaN - t_zb[oN], hS = etaS - t_zb[oS];
                if (!(hW < k.eps || hE < k.eps)) {
                    sxE = minmod(eta - etaW, etaE - eta); sxH = minmod(h - hW, hE - h);
                    sxQx = minmod(qx - t_qx[oW], t_qx[oE] - qx); sxQy = minmod(qy - t_qy[oW], t_qy[oE] - qy);
                }
                if (!(hS < k.eps || hN < k.eps)) {
                    syE = minmod(eta - etaS, etaN - eta); syH = minmod(h - hS, hN - h);
                    syQx = minmod(qx - t_qx[oS], t_qx[oN] - qx); syQy = minmod(qy - t_qy[oS], t_qy[oN] - qy);
                }
                // face estimates at the old time level and their analytic fluxes
                const R hEf = h + half * sxH, hWf = h - half * sxH, hNf = h + half * syH, hSf = h - half * syH;
                const R qxE = qx + half * sxQx, qxW = qx - half * sxQx, qyE = qy + half * sxQy, qyW = qy - half * sxQy;
                const R qxN = qx + half * syQx, qxS = qx - half * syQx, qyN = qy + half * syQy, qyS = qy - half * syQy;
                const R uE = hEf < k.eps ? R(0) : qxE * fm_rcp(hEf), uW = hWf < k.eps ? R(0) : qxW * fm_rcp(hWf);
                const R vN = hNf < k.eps ? R(0) : qyN * fm_rcp(hNf), vS = hSf < k.eps ? R(0) : qyS * fm_rcp(hSf);
                // flux divergence - bed-slope source; the hydrostatic parts collapse to g/2 (hE+hW) d(eta)
                R dEta = ((qxE - qxW) + (qyN - qyS)) * inv_delta;
                R dQx = (uE * qxE - uW * qxW + vN * qxN - vS * qxS + hg * sxE * (hEf + hWf)) * inv_delta;
                R dQy = (uE * qyE - uW * qyW + vN * qyN - vS * qyS + hg * syE * (hNf + hSf)) * inv_delta;
                dEta = fm_chop(dEta, k.eps); dQx = fm_chop(dQx, k.eps); dQy = fm_chop(dQy, k.eps);
                e2 = eta - half * dt * dEta; qx2 = qx - half * dt * dQx; qy2 = qy - half * dt * dQy;
            }
            s_p[0 * T::PPLANE + i] = e2;   s_p[1 * T::PPLANE + i] = qx2;  s_p[2 * T::PPLANE + i] = qy2;
            s_p[3 * T::PPLANE + i] = sxE;  s_p[4 * T::PPLANE + i] = sxH;  s_p[5 * T::PPLANE + i] = sxQx; s_p[6 * T::PPLANE + i] = sxQy;
            s_p[7 * T::PPLANE + i] = syE;  s_p[8 * T::PPLANE + i] = syH;  s_p[9 * T::PPLANE + i] = syQx; s_p[10 * T::PPLANE + i] = syQy;
        }
        __syncthreads();

        if (next < ntiles) load_emax(next);     // in flight during phases C and D

        // ---- phase C: every face once --------------------------------------------------------------
        if (dt > R(0)) {
            for (int f = tid; f < T::NF; f += T::NT) {
                const bool isx = f < T::NXF;
                const int gq = isx ? f : f - T::NXF;
                const int w = isx ? T::TX + 1 : T::TX;
                const int j = gq / w, i = gq - j * w;
                // x-face: cells (i-1, j) | (i, j);  y-face: cells (i, j-1) | (i, j)
                const int pL = isx ? (j + 1) * T::PW + i : j * T::PW + i + 1;
                const int pR = isx ? pL + 1 : pL + T::PW;
                const int oL = isx ? (j + 2) * T::BW + i - 1 + T::CO : (j + 1) * T::BW + i + T::CO;
                const int oR = isx ? oL + 1 : oL + T::BW;
                const int sb = isx ? 3 : 7;                         // slope block of this direction
                const R etaL = s_p[pL] + half * s_p[sb * T::PPLANE + pL], etaR = s_p[pR] - half * s_p[sb * T::PPLANE + pR];
                const R hfL = (s_p[pL] - t_zb[oL]) + half * s_p[(sb + 1) * T::PPLANE + pL];
                const R hfR = (s_p[pR] - t_zb[oR]) - half * s_p[(sb + 1) * T::PPLANE + pR];
                const R qxL = s_p[T::PPLANE + pL] + half * s_p[(sb + 2) * T::PPLANE + pL], qxR = s_p[T::PPLANE + pR] - half * s_p[(sb + 2) * T::PPLANE + pR];
                const R qyL = s_p[2 * T::PPLANE + pL] + half * s_p[(sb + 3) * T::PPLANE + pL], qyR = s_p[2 * T::PPLANE + pR] - half * s_p[(sb + 3) * T::PPLANE + pR];
                const R rL = hfL <= k.eps ? R(0) : fm_rcp(hfL), rR = hfR <= k.eps ? R(0) : fm_rcp(hfR);   // CLSchemeMUSCLHancock.clc:1140-1150
                const R uL = qxL * rL, vL = qyL * rL, uR = qxR * rR, vR = qyR * rR;
                const Flux3<R> F = isx ? face_core_flux<R, false>(k, etaL, etaL - hfL, uL, vL, R(0), etaR, etaR - hfR, uR, vR, R(0))
                                       : face_core_flux<R, false>(k, etaL, etaL - hfL, vL, uL, R(0), etaR, etaR - hfR, vR, uR, R(0));
                s_f[f] = F.m; s_f[T::NF + f] = F.n; s_f[2 * T::NF + f] = F.t;
            }
        }
        __syncthreads();

        // ---- phase D: corrector, two cells per thread ------------------------------------------------
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int i = tid % T::TX, j = tid / T::TX + hf * (T::TY / 2);
            const int x = x0 + i, y = y0 + j;
            if (x >= g.cols || y >= a.y1) continue;
            const int o = (j + 2) * T::BW + i + T::CO, p = (j + 1) * T::PW + i + 1;
            const size_t id = static_cast<size_t>(y) * g.pitch + x;
            const int gy = y + g.gy0;
            Cell<R> c{t_eta[o], pre_emax[hf], t_qx[o], t_qy[o]};
            const R zb = t_zb[o];
            const bool interior = x >= 2 && x <= g.cols - 3 && gy >= 2 && gy <= g.grows - 3;   // ring of two is frozen
            if (interior && dt > R(0) && !(c.emax <= R(-9999.0) || c.eta == R(-9999.0))) {
                int dry = (c.eta - zb < k.eps) ? 1 : 0;
                dry += (s_flag[o + T::BW] >> 1) + (s_flag[o + 1] >> 1) + (s_flag[o - T::BW] >> 1) + (s_flag[o - 1] >> 1);
                if (dry < 5) {
                    const R e2 = s_p[p], hc2 = e2 - zb;
                    const R sxE = s_p[3 * T::PPLANE + p], sxH = s_p[4 * T::PPLANE + p];
                    const R syE = s_p[7 * T::PPLANE + p], syH = s_p[8 * T::PPLANE + p];
                    int stop = 0;
                    bool front = false;
                    R bN, bS, bE, bW, hnN, hnS, hnE, hnW;
                    // owner-side terms of one face: own face estimate (etaO, hO) against the neighbour's
                    auto owner = [&](const bool ownIsLeft, const bool isx, const R etaO, const R hO, const int pn, const int on,
                                     R& bed, R& hNb) {
                        const int sb = isx ? 3 : 7;
                        const R sgn = ownIsLeft ? -half : half;        // the neighbour's facing estimate
                        const R etaN_ = s_p[pn] + sgn * s_p[sb * T::PPLANE + pn];
                        const R hN_ = (s_p[pn] - t_zb[on]) + sgn * s_p[(sb + 1) * T::PPLANE + pn];
                        const R zO = etaO - hO, zN_ = etaN_ - hN_;
                        const R zmax = fm_max(zO, zN_);
                        hNb = fm_posdiff(etaN_, zmax);
                        bed = fm_min(zmax, etaO);
                        front = front || (etaO - zmax <= k.eps) || (hNb <= k.eps);
                    };
                    // stop tests of one face (CLSchemeMUSCLHancock.clc:1172-1204); only reached at wet/dry fronts
                    auto owner_stop = [&](const bool ownIsLeft, const bool isx, const R etaO, const R hO, const int pn, const int on) {
                        const int sb = isx ? 3 : 7, qb = isx ? 1 : 2;
                        const R sgn = ownIsLeft ? -half : half;
                        const R etaN_ = s_p[pn] + sgn * s_p[sb * T::PPLANE + pn];
                        const R hN_ = (s_p[pn] - t_zb[on]) + sgn * s_p[(sb + 1) * T::PPLANE + pn];
                        const R zO = etaO - hO, zN_ = etaN_ - hN_;
                        const R zmax = fm_max(zO, zN_);
                        const R hOwn = fm_posdiff(etaO, zmax);
                        const R hNb = fm_posdiff(etaN_, zmax);
                        const R qO = s_p[qb * T::PPLANE + p] + (ownIsLeft ? half : -half) * s_p[(sb + qb + 1) * T::PPLANE + p];
                        const R qN_ = s_p[qb * T::PPLANE + pn] + sgn * s_p[(sb + qb + 1) * T::PPLANE + pn];
                        const R unO = hO <= k.eps ? R(0) : qO * fm_rcp(hO), unN = hN_ <= k.eps ? R(0) : qN_ * fm_rcp(hN_);
                        const R ownQn = isx ? c.qx : c.qy;
                        const R hL = ownIsLeft ? hOwn : hNb, hR = ownIsLeft ? hNb : hOwn;
                        const R unL = ownIsLeft ? unO : unN, unR = ownIsLeft ? unN : unO;
                        if (ownIsLeft) { if (hL <= k.eps && ownQn > R(0)) ++stop; }
                        else           { if (hR <= k.eps && ownQn < R(0)) ++stop; }
                        if (hR <= k.eps && unL < R(0)) ++stop;
                        if (hL <= k.eps && unR > R(0)) ++stop;
                    };
                    owner(true, false, e2 + half * syE, hc2 + half * syH, p + T::PW, o + T::BW, bN, hnN);
                    owner(false, false, e2 - half * syE, hc2 - half * syH, p - T::PW, o - T::BW, bS, hnS);
                    owner(true, true, e2 + half * sxE, hc2 + half * sxH, p + 1, o + 1, bE, hnE);
                    owner(false, true, e2 - half * sxE, hc2 - half * sxH, p - 1, o - 1, bW, hnW);
                    if (front) {
                        owner_stop(true, false, e2 + half * syE, hc2 + half * syH, p + T::PW, o + T::BW);
                        owner_stop(false, false, e2 - half * syE, hc2 - half * syH, p - T::PW, o - T::BW);
                        owner_stop(true, true, e2 + half * sxE, hc2 + half * sxH, p + 1, o + 1);
                        owner_stop(false, true, e2 - half * sxE, hc2 - half * sxH, p - 1, o - 1);
                    }
                    const int fe = j * (T::TX + 1) + i + 1, fw = fe - 1;
                    const int fn = T::NXF + (j + 1) * T::TX + i, fs = fn - T::TX;
                    R dEta = ((s_f[fe] - s_f[fw]) + (s_f[fn] - s_f[fs])) * inv_delta;
                    R dQx = ((s_f[T::NF + fe] - s_f[T::NF + fw]) + (s_f[2 * T::NF + fn] - s_f[2 * T::NF + fs]) +
                             hg * (bE - bW) * (hnE + hnW)) * inv_delta;
                    R dQy = ((s_f[2 * T::NF + fe] - s_f[2 * T::NF + fw]) + (s_f[T::NF + fn] - s_f[T::NF + fs]) +
                             hg * (bN - bS) * (hnN + hnS)) * inv_delta;
                    dEta = fm_chop(dEta, k.eps); dQx = fm_chop(dQx, k.eps); dQy = fm_chop(dQy, k.eps);
                    if (stop > 0) { c.qx = R(0); c.qy = R(0); }
                    c.eta = c.eta - dt * dEta; c.qx = c.qx - dt * dQx; c.qy = c.qy - dt * dQy;
                    const R h_new = c.eta - zb;
                    if (k.friction && !(h_new < k.eps)) friction_fast(k, h_new, fm_rcp(h_new), c.qx, c.qy, pre_mann[hf], dt);
                    if (h_new < k.eps) c.eta = zb;
                    if (c.eta > c.emax && c.emax > R(-9990.0)) c.emax = c.eta;
                }
            }
            d.store(id, c);
            if (a.reduce_mode != hp::kReduceNone) {
                const R h = c.eta - zb;
                if (h > k.eps10 && c.emax > R(-9999.0)) {
                    const R cc = fm_sqrt(k.g * h);
                    R sp = cc;
                    if (!k.simplified_speed) { const R rh = fm_rcp(h); sp = fm_max(hp_abs(c.qx * rh), hp_abs(c.qy * rh)) + cc; }
                    ws = fm_max(sp, ws);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    block_reduce_finalize<R>(ws, a, k);
}

template <class R> static int launch_mh_tma(const StepArgs& a_in, const TmaMaps& maps, int sm_count, cudaStream_t st) {
    using T = TileMH<R>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(mh_step_tma<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int tiles = ((a.grid.cols + T::TX - 1) / T::TX) * ((a.y1 - a.y0 + T::TY - 1) / T::TY);
    int grid = T::CTAS_PER_SM * sm_count;
    if (grid > tiles) grid = tiles;
    a.total_ctas = grid;
    mh_step_tma<R><<<grid, T::NT, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

}  // namespace HP_NS
