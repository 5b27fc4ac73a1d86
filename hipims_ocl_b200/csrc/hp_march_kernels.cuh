// hp_march_kernels.cuh -- "marching" step kernels: one warp owns a strip of 32 columns and walks it
// row by row, keeping everything that is reused between consecutive rows in registers.
//
// Why: the tile kernels (hp_fast_kernels.cuh) are bound by instruction issue, not by HBM
// (profiles/r01_*): per cell-update only ~36 % of their instructions are fp64 arithmetic, the rest
// is index arithmetic, shared-memory traffic for predictor / flux planes and CTA barriers between
// the phases.  Here
//   * raw rows arrive through TMA (cp.async.bulk.tensor.3d + mbarrier; ONE box = one row of the six
//     planes a step reads) into a per-warp ring of four rows, two rows ahead of the arithmetic;
//   * the y-direction never leaves the thread: the predictor of the row below, the flux through the
//     southern face and its owner terms are carried in registers from the previous row;
//   * the x-direction is exchanged between neighbouring lanes with warp shuffles: the east-side
//     face estimate goes one lane up, the solved face (flux, reconstructed bed, depth) comes back
//     one lane down -- every face is still solved exactly once;
//   * there is no CTA-level barrier and no shared-memory store in the loop: warps are autonomous,
//     a CTA is four warps on four adjacent strips so that the halo columns hit in L1/L2;
//   * work is split into equal runs of (strip group, row) units over a persistent grid that
//     exactly fills the SMs; each CTA takes several runs spread round-robin over the domain, so all
//     CTAs finish together (no tail wave) even where wet and dry regions cost differently.
//   * rows that cannot change (exactly dry stencils) are recognised by one warp vote and skipped.
// The lane at either strip edge only feeds its neighbour (halo lanes): 30 of 32 lanes update cells
// in fp64, 28 in fp32 (the TMA box must start on a 16-byte boundary).  MUSCL-Hancock needs raw
// values two columns out; those come straight from the box, which is wider than the warp.
// The per-cell arithmetic is the one of the tile kernels (same helper functions), so results are
// identical to them to the last bit where the operation order is the same.
#pragma once

#include "hp_fast_kernels.cuh"

namespace HP_NS {

struct TmaBlockMap { CUtensorMap block; };

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int p, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(p), "r"(bar) : "memory");
}

// Ring geometry.  ALT = false: the step reads buffer A (block planes 0..5 = eta qx qy emax zb n);
// ALT = true: buffer B (block planes 4..9 = zb n eta qx qy emax).
#ifndef HP_MARCH_RR
#define HP_MARCH_RR 4
#endif
#ifndef HP_MARCH_MH_RR
#define HP_MARCH_MH_RR 4
#endif
template <class R, int HALO, bool ALT, int RING = HP_MARCH_RR> struct March {
    static constexpr int NW = hp::kMarchWarps, RR = RING, NP = 6;              // RR: ring rows (power of two)
    static constexpr int A16 = 16 / int(sizeof(R));                       // elements per 16 bytes
    static constexpr int USE = hp::march_use(int(sizeof(R)), HALO);       // cells updated per warp row
    static constexpr int PADL = ((-HALO) % A16 + A16) % A16;              // box starts 16-byte aligned
    static constexpr int BW = hp::march_box_w(int(sizeof(R)), HALO);
    static constexpr int PLANE = BW * int(sizeof(R));                     // bytes between planes of one ring row
    static constexpr int ROW_TX = NP * PLANE;                             // bytes one TMA box delivers
    static constexpr int SLOT = (ROW_TX + 127) / 128 * 128;               // one ring row
    static constexpr int WARP_BYTES = RR * SLOT;
    static constexpr int SMEM_BYTES = NW * WARP_BYTES + NW * RR * 8;
    static constexpr int P0 = ALT ? 4 : 0;                                // first block plane of the box
    static constexpr int P_ETA = ALT ? 2 : 0, P_QX = ALT ? 3 : 1, P_QY = ALT ? 4 : 2, P_EMAX = ALT ? 5 : 3, P_ZB = ALT ? 0 : 4,
                         P_N = ALT ? 1 : 5;
    static_assert((USE % A16) == 0, "strip starts must keep the TMA box 16-byte aligned");
    static_assert(BW >= 32 + PADL && (BW * sizeof(R)) % 16 == 0, "box");
};

// One face in the normal frame, solved once for both cells: core flux (see face_core_flux), the
// common reconstructed bed, both reconstructed depths and the stop-counter increments of the two
// owners (CLSchemeGodunov.clc:83-137 / CLSchemeMUSCLHancock.clc:1172-1204; they can only be
// non-zero at a wet/dry front).  qOwnL / qOwnR: the owning CELLS' discharge normal to the face.
template <class R> struct FaceOut { R m, n, t, zmax, hL, hR; int stopL, stopR; };

// One face in the normal frame, solved once for both cells: core flux (see face_core_flux), the common reconstructed
// bed, both reconstructed depths and the stop-counter increments of the two owners (CLSchemeGodunov.clc:83-137 /
// CLSchemeMUSCLHancock.clc:1172-1204; they can only be non-zero at a wet/dry front).  qOwnL / qOwnR: the owning CELLS'
// discharge normal to the face.  WET FAST PATH in front: when both reconstructed depths exceed the dry threshold (one
// combined test) there are no stop flags, no dry-side selects, no clamps and the square roots need no zero guard --
// with identical results, the general path takes exactly these operations then.  Must be called by all 32 lanes.
template <class R, bool CACHED, class QOwnL, class QOwnR>
__device__ __forceinline__ void face_solve2(const Params<R>& k, R etaL, R zL, R unL, R utL, R aL_cached, R etaR, R zR, R unR, R utR,
                                            R aR_cached, QOwnL qOwnL, QOwnR qOwnR, FaceOut<R>& o) {
    const R hg = R(0.5) * k.g;
    const R zmax = fm_max(zL, zR);
    const R dL = etaL - zmax, dR = etaR - zmax;
    o.zmax = zmax;
    // warp-uniform choice (every lane reaches the faces together): a warp with a single dry side takes the general path
    // as a whole instead of executing both -- on thin films over rough terrain most warps are mixed
    if (__all_sync(0xffffffffu, dL > k.eps && dR > k.eps)) {
        o.hL = dL; o.hR = dR; o.stopL = 0; o.stopR = 0;
        const R aL = (CACHED && zmax == zL) ? aL_cached : fm_sqrt_pos(k.g * dL);
        const R aR = (CACHED && zmax == zR) ? aR_cached : fm_sqrt_pos(k.g * dR);
        const R qnL = dL * unL, qnR = dR * unR;
        const R as = hp_abs(R(0.5) * (aL + aR) + R(0.25) * (unL - unR));
        const R us = R(0.5) * (unL + unR) + aL - aR;
        const R sL = fm_min(unL - aL, us - as);
        const R sR = fm_max(unR + aR, us + as);
        const R FLn = unL * qnL + hg * dL * dL, FRn = unR * qnR + hg * dR * dR;
        if (sL >= R(0)) { o.m = qnL; o.n = FLn; o.t = qnL * utL; return; }
        if (!(sR >= R(0))) { o.m = qnR; o.n = FRn; o.t = qnR * utR; return; }
        const R inv = fm_rcp(sR - sL);
        const R ss = sL * sR;
        const R f1 = (sR * qnL - sL * qnR + ss * (dR - dL)) * inv;
        const R f2 = (sR * FLn - sL * FRn + ss * (qnR - qnL)) * inv;
        o.m = f1; o.n = f2; o.t = f1 * (f1 >= R(0) ? utL : utR);
        return;
    }
    // ---- general path: a dry side, stop flags (wet/dry fronts only) --------------------------------
    const R hL = fm_pos_s(dL), hR = fm_pos_s(dR);
    o.hL = hL; o.hR = hR;
    {
        // qOwnL() / qOwnR() are evaluated here only: the wet path above never needs the owners' raw discharge, so a
        // caller that can re-read it (from its ring) does not have to keep it in registers across its predictor
        int both = 0;
        if (hR <= k.eps && unL < R(0)) ++both;
        if (hL <= k.eps && unR > R(0)) ++both;
        o.stopL = both + ((hL <= k.eps && qOwnL() > R(0)) ? 1 : 0);
        o.stopR = both + ((hR <= k.eps && qOwnR() < R(0)) ? 1 : 0);
    }
    const bool dryL = hL < k.eps, dryR = hR < k.eps;
    if (dryL && dryR) {
        const R hm = R(0.5) * (hL + hR);
        o.m = R(0); o.n = hg * hm * hm; o.t = R(0);
        return;
    }
    if (dryL) { unL = R(0); utL = R(0); }
    if (dryR) { unR = R(0); utR = R(0); }
    const R aL = (CACHED && zmax == zL) ? aL_cached : fm_celerity(k.g, hL);
    const R aR = (CACHED && zmax == zR) ? aR_cached : fm_celerity(k.g, hR);
    const R qnL = hL * unL, qnR = hR * unR;
    const R as = hp_abs(R(0.5) * (aL + aR) + R(0.25) * (unL - unR));
    const R us = R(0.5) * (unL + unR) + aL - aR;
    const R sL = dryL ? unR - 2 * aR : fm_min(unL - aL, us - as);
    const R sR = dryR ? unL + 2 * aL : fm_max(unR + aR, us + as);
    const R FLn = unL * qnL + hg * hL * hL, FRn = unR * qnR + hg * hR * hR;
    if (sL >= R(0)) { o.m = qnL; o.n = FLn; o.t = qnL * utL; return; }
    if (!(sR >= R(0))) { o.m = qnR; o.n = FRn; o.t = qnR * utR; return; }
    const R inv = fm_rcp(sR - sL);
    const R ss = sL * sR;
    const R f1 = (sR * qnL - sL * qnR + ss * (hR - hL)) * inv;
    const R f2 = (sR * FLn - sL * FRn + ss * (qnR - qnL)) * inv;
    o.m = f1; o.n = f2; o.t = f1 * (f1 >= R(0) ? utL : utR);
}

template <class R> __device__ __forceinline__ R shfl_up1(R v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <class R> __device__ __forceinline__ R shfl_dn1(R v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// Cells the step leaves unwritten (frozen ring, all-dry stencil) still enter the CFL reduction with the value the
// destination buffer holds.  That read is a dependent global load in the middle of a row; issue a
// prefetch for it one row ahead wherever the cell is likely to stay unwritten.
template <class R> __device__ __forceinline__ void prefetch_dst(const MutView<R>& d, size_t id) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.eta + id));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.emax + id));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.qx + id));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(d.qy + id));
}

// =============================================================================================
// First-order Godunov, marching.  Carried per lane: the cell below (level, bed, velocities, celerity)
// and its southern face.  Cells whose stencil is dry stay unwritten, exactly like
// godunov_step_tma.
// =============================================================================================
template <class R> struct GodCell { R eta, zb, u, v, c; };

#ifndef HP_MARCH_GOD_CTAS64
#define HP_MARCH_GOD_CTAS64 4
#endif
#ifndef HP_MARCH_GOD_CTAS32
#define HP_MARCH_GOD_CTAS32 6
#endif
template <class R, bool ALT>
__global__ void __launch_bounds__(hp::kMarchWarps * 32, sizeof(R) == 8 ? HP_MARCH_GOD_CTAS64 : HP_MARCH_GOD_CTAS32)
godunov_step_march(const StepArgs a, const __grid_constant__ TmaBlockMap maps) {
    using T = March<R, 1, ALT>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* const ring = smem_raw + warp * T::WARP_BYTES;
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t bar_u = smem_u32(smem_raw + T::NW * T::WARP_BYTES) + warp * T::RR * 8;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < T::RR; ++r) mbar_init(bar_u + 8 * r, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const R dt = read_timestep<R>(a.clock);
    const R inv_delta = fm_rcp(k.delta);
    const R hg = R(0.5) * k.g;
    const MutView<R> d(a.dst);
    const bool stepping = dt > R(0);

    const int nrows = a.y1 - a.y0;
    const int nstrips = (g.cols + T::USE - 1) / T::USE;
    const int ngroups = (nstrips + T::NW - 1) / T::NW;
    const long long units = static_cast<long long>(ngroups) * nrows;
    // A CTA works on `march_runs` equal runs of units, taken round-robin from the whole domain (run r belongs to CTA
    // r mod grid): every CTA gets the same amount of work AND a sample of the domain, so wet and dry regions even out.
    const long long total_runs = static_cast<long long>(gridDim.x) * a.march_runs;
    int run = 0;
    long long u = units * blockIdx.x / total_runs, u1 = units * (blockIdx.x + 1) / total_runs;

    const int lc = (lane + T::PADL) * int(sizeof(R));
    const int lw = (lane > 0 ? lane - 1 + T::PADL : T::PADL) * int(sizeof(R));
    auto ld = [&](int row_off, int plane, int col_off) -> R {
        return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + col_off);
    };
    const bool lane_owns = lane >= 1 && lane < 1 + T::USE;

    R ws = R(0);
    uint32_t ph = 0;

    for (;;) {
        if (u >= u1) {
            if (++run >= a.march_runs) break;
            const long long r = static_cast<long long>(run) * gridDim.x + blockIdx.x;
            u = units * r / total_runs; u1 = units * (r + 1) / total_runs;
            continue;
        }
        const int grp = static_cast<int>(u / nrows);
        const int ya = a.y0 + static_cast<int>(u - static_cast<long long>(grp) * nrows);
        const long long gend = static_cast<long long>(grp + 1) * nrows;
        const int yb = ya + static_cast<int>((u1 < gend ? u1 : gend) - u);
        u += yb - ya;
        const int strip = grp * T::NW + warp;
        if (strip >= nstrips) continue;

        const int X0 = strip * T::USE - 1;               // column of lane 0
        const int x = X0 + lane;
        const int rs = ya - 1;                            // first raw row of this run
        const int NR = yb - ya + 2;                       // raw rows 0 .. NR-1; rows 1 .. NR-2 are updated
        const bool x_interior = x >= 1 && x <= g.cols - 2;
        const bool x_store = lane_owns && x < g.cols;
        auto issue_row = [&](int j) {
            const uint32_t bar = bar_u + 8 * (j & (T::RR - 1));
            mbar_expect_tx(bar, uint32_t(T::ROW_TX));
            tma_load_3d(ring_u + (j & (T::RR - 1)) * T::SLOT, &maps.block, X0 - T::PADL, rs + j, T::P0, bar);
        };
        auto wait_row = [&](int j) {
            const int s = j & (T::RR - 1);
            mbar_wait(bar_u + 8 * s, (ph >> s) & 1u);
            ph ^= 1u << s;
        };
        auto derive = [&](int row_off, GodCell<R>& o, R& qx, R& qy) {      // phase B of the tile kernel
            o.eta = ld(row_off, T::P_ETA, lc); o.zb = ld(row_off, T::P_ZB, lc);
            qx = ld(row_off, T::P_QX, lc); qy = ld(row_off, T::P_QY, lc);
            const R h = o.eta - o.zb;
            const R rh = !(h < k.eps) ? fm_rcp(h) : R(0);
            o.u = qx * rh; o.v = qy * rh;
            o.c = fm_celerity(k.g, fm_pos_s(h));
        };
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < T::RR; ++j) if (j < NR) issue_row(j);
        }
        wait_row(0);

        GodCell<R> P;
        R p_qx, p_qy;
        derive(0, P, p_qx, p_qy);
        R sM = R(0), sN = R(0), sT = R(0), sZ = R(0), sH = R(0);   // southern face of row j-1
        int sStop = 0;
        bool dry_s = true;                                          // dryness of the cell below row j-1

        for (int j = 1; j < NR; ++j) {
            const int y = rs + j;
            const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT, o_c = (j & (T::RR - 1)) * T::SLOT;
            wait_row(j);
            GodCell<R> C;
            R c_qx, c_qy;
            derive(o_c, C, c_qx, c_qy);
            const bool dry_p = P.eta - P.zb < k.eps, dry_c = C.eta - C.zb < k.eps;
            if (a.reduce_mode == hp::kReduceDst && x_store && j + 1 < NR) {      // row y is updated in the next trip
                const int gy = y + g.gy0;
                if (!x_interior || gy < 1 || gy > g.grows - 2 || (dry_c && dry_p)) prefetch_dst(d, static_cast<size_t>(y) * g.pitch + x);
            }

            FaceOut<R> fy;
            if (stepping) face_solve2<R, true>(k, P.eta, P.zb, P.v, P.u, P.c, C.eta, C.zb, C.v, C.u, C.c, [&] { return p_qy; }, [&] { return c_qy; }, fy);
            else { fy.m = fy.n = fy.t = fy.zmax = fy.hL = fy.hR = R(0); fy.stopL = fy.stopR = 0; }

            if (j >= 2) {
                const int yc = y - 1, gyc = yc + g.gy0;
                Cell<R> c{P.eta, ld(o_m, T::P_EMAX, lc), p_qx, p_qy};
                const R zb = P.zb;
                // rows y-2, y-1, y dry in every lane: the stencil of every cell of row y-1 is dry, the reference returns
                // without writing (CLSchemeGodunov.clc:248-255) -- no face of that row is needed
                const bool skip = stepping && __all_sync(FULL, dry_s && dry_p && dry_c && !(c.emax <= R(-9999.0) || c.eta == R(-9999.0)));
                if (a.reduce_mode == hp::kReduceSrc && x_store) {
                    const R h = c.eta - zb;
                    if (h > k.eps10 && c.emax > R(-9999.0)) {
                        const R sp = k.simplified_speed ? P.c : fm_max(hp_abs(P.u), hp_abs(P.v)) + P.c;
                        ws = fm_max(sp, ws);
                    }
                }
                bool wrote = false;
                R rh_new = R(0);
                bool have_new = false;
                if (!skip) {
                // west face of row y-1: the cell of lane-1 against the own cell
                const R wU = shfl_up1(P.u), wV = shfl_up1(P.v), wC = shfl_up1(P.c);
                const unsigned drym = __ballot_sync(FULL, dry_p);
                FaceOut<R> fx;
                if (stepping) face_solve2<R, true>(k, ld(o_m, T::P_ETA, lw), ld(o_m, T::P_ZB, lw), wU, wV, wC, P.eta, P.zb, P.u, P.v, P.c,
                                            [&] { return ld(o_m, T::P_QX, lw); }, [&] { return p_qx; }, fx);
                else { fx.m = fx.n = fx.t = fx.zmax = fx.hL = fx.hR = R(0); fx.stopL = fx.stopR = 0; }
                const R eM = shfl_dn1(fx.m), eN = shfl_dn1(fx.n), eT = shfl_dn1(fx.t), eZ = shfl_dn1(fx.zmax), eH = shfl_dn1(fx.hR);
                const int eStop = __shfl_down_sync(FULL, fx.stopL, 1);
                if (x_interior && gyc >= 1 && gyc <= g.grows - 2) {                  // frozen outer ring
                    if (!stepping) {
                        wrote = !k.dt0_keep;                                            // CLSchemeGodunov.clc:201-206 (:477-478)
                    } else if (c.emax <= R(-9999.0) || c.eta == R(-9999.0)) {
                        wrote = true;                                                   // disabled cell: copied through
                    } else {
                        const bool dry_e = (drym >> ((lane + 1) & 31)) & 1u, dry_w = (drym >> ((lane + 31) & 31)) & 1u;
                        const bool all_dry = dry_p && dry_c && dry_s && dry_e && dry_w; // CLSchemeGodunov.clc:248-255
                        if (!all_dry) {
                            const R bN = fm_min(fy.zmax, c.eta), bS = fm_min(sZ, c.eta), bE = fm_min(eZ, c.eta), bW = fm_min(fx.zmax, c.eta);
                            const int stop = fy.stopL + sStop + fx.stopR + eStop;
                            const R dEta = ((eM - fx.m) + (fy.m - sM)) * inv_delta;
                            const R dQx = ((eN - fx.n) + (fy.t - sT) + hg * (bE - bW) * (eH + fx.hL)) * inv_delta;
                            const R dQy = ((eT - fx.t) + (fy.n - sN) + hg * (bN - bS) * (fy.hR + sH)) * inv_delta;
                            if (stop > 0) { c.qx = R(0); c.qy = R(0); }
                            if (!(hp_abs(dEta) < k.eps)) c.eta = c.eta - dt * dEta;   // |D| < eps => 0 (CLSchemeGodunov.clc:340-348)
                            if (!(hp_abs(dQx) < k.eps)) c.qx = c.qx - dt * dQx;
                            if (!(hp_abs(dQy) < k.eps)) c.qy = c.qy - dt * dQy;
                            const R h_new = c.eta - zb;
                            if (!(h_new < k.eps)) { rh_new = fm_rcp(h_new); have_new = true; }
                            if (k.friction) friction_fast(k, h_new, rh_new, c.qx, c.qy, ld(o_m, T::P_N, lc), dt);
                            if (c.eta > c.emax && c.emax > R(-9990.0)) c.emax = c.eta;
                            if (h_new < k.eps) c.eta = zb;
                            wrote = true;
                        }
                    }
                }
                }
                if (x_store) {
                    const size_t id = static_cast<size_t>(yc) * g.pitch + x;
                    if (wrote) d.store(id, c);
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
            }
            sM = fy.m; sN = fy.n; sT = fy.t; sZ = fy.zmax; sH = fy.hL; sStop = fy.stopR;
            dry_s = dry_p;
            P = C; p_qx = c_qx; p_qy = c_qy;

            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && j - 1 + T::RR < NR) issue_row(j - 1 + T::RR);
        }
    }
    block_reduce_finalize<R>(ws, a, k);
}

// =============================================================================================
// Partial inertial scheme, marching.  A face's discharge is stored at its northern / eastern cell
// (CLSchemeInertial.clc:141-152) and both cells of a face evaluate the same flux -- except for the
// Manning coefficient, which each takes from ITSELF (:107-110).  So a face is evaluated once up to
// the friction denominator, which is then applied per owner (once more only where n differs).
// =============================================================================================
template <class R> struct InFace { R num, A, qmax; bool wet; };

// calculateInertialFlux (CLSchemeInertial.clc:335-378) without the owner's Manning coefficient:
// q(n) = clamp((prev - g dt h S) / (1 + g dt n^2 |prev| / h^(7/3)), +-0.8 h sqrt(g h)), 0 if h < eps
template <class R>
__device__ __forceinline__ InFace<R> inertial_face(const Params<R>& k, R gdt, R prev, R etaUp, R zUp, R etaDown, R zDown, R inv_delta) {
    InFace<R> f;
    const R h = fm_max(etaDown, etaUp) - fm_max(zUp, zDown);
    f.wet = !(h < k.eps);
    f.num = R(0); f.A = R(0); f.qmax = R(0);
    if (f.wet) {
        const R rh = fm_rcp(h);
        f.num = prev - gdt * h * ((etaDown - etaUp) * inv_delta);
        f.A = gdt * hp_abs(prev) * rh * rh * fm_rcbrt(h);
        f.qmax = R(0.8) * h * fm_sqrt(k.g * h);
    }
    return f;
}
template <class R> __device__ __forceinline__ R inertial_q(const InFace<R>& f, R n) {
    if (!f.wet) return R(0);
    const R q = f.num * fm_rcp(R(1.0) + f.A * (n * n));
    return q > f.qmax ? f.qmax : (q < -f.qmax ? -f.qmax : q);
}

#ifndef HP_MARCH_INE_CTAS64
#define HP_MARCH_INE_CTAS64 6
#endif
template <class R, bool ALT>
__global__ void __launch_bounds__(hp::kMarchWarps * 32, sizeof(R) == 8 ? HP_MARCH_INE_CTAS64 : 8)
inertial_step_march(const StepArgs a, const __grid_constant__ TmaBlockMap maps) {
    using T = March<R, 1, ALT>;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* const ring = smem_raw + warp * T::WARP_BYTES;
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t bar_u = smem_u32(smem_raw + T::NW * T::WARP_BYTES) + warp * T::RR * 8;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < T::RR; ++r) mbar_init(bar_u + 8 * r, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const R dt = read_timestep<R>(a.clock);
    const R inv_delta = fm_rcp(k.delta);
    const R gdt = k.g * dt;
    const MutView<R> d(a.dst);
    const bool stepping = dt > R(0);

    const int nrows = a.y1 - a.y0;
    const int nstrips = (g.cols + T::USE - 1) / T::USE;
    const int ngroups = (nstrips + T::NW - 1) / T::NW;
    const long long units = static_cast<long long>(ngroups) * nrows;
    // A CTA works on `march_runs` equal runs of units, taken round-robin from the whole domain (run r belongs to CTA
    // r mod grid): every CTA gets the same amount of work AND a sample of the domain, so wet and dry regions even out.
    const long long total_runs = static_cast<long long>(gridDim.x) * a.march_runs;
    int run = 0;
    long long u = units * blockIdx.x / total_runs, u1 = units * (blockIdx.x + 1) / total_runs;

    const int lc = (lane + T::PADL) * int(sizeof(R));
    const int lw = (lane > 0 ? lane - 1 + T::PADL : T::PADL) * int(sizeof(R));
    auto ld = [&](int row_off, int plane, int col_off) -> R {
        return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + col_off);
    };
    const bool lane_owns = lane >= 1 && lane < 1 + T::USE;

    R ws = R(0);
    uint32_t ph = 0;

    for (;;) {
        if (u >= u1) {
            if (++run >= a.march_runs) break;
            const long long r = static_cast<long long>(run) * gridDim.x + blockIdx.x;
            u = units * r / total_runs; u1 = units * (r + 1) / total_runs;
            continue;
        }
        const int grp = static_cast<int>(u / nrows);
        const int ya = a.y0 + static_cast<int>(u - static_cast<long long>(grp) * nrows);
        const long long gend = static_cast<long long>(grp + 1) * nrows;
        const int yb = ya + static_cast<int>((u1 < gend ? u1 : gend) - u);
        u += yb - ya;
        const int strip = grp * T::NW + warp;
        if (strip >= nstrips) continue;

        const int X0 = strip * T::USE - 1;               // column of lane 0
        const int x = X0 + lane;
        const int rs = ya - 1;                            // first raw row of this run
        const int NR = yb - ya + 2;                       // raw rows 0 .. NR-1; rows 1 .. NR-2 are updated
        const bool x_interior = x >= 1 && x <= g.cols - 2;
        const bool x_store = lane_owns && x < g.cols;
        auto issue_row = [&](int j) {
            const uint32_t bar = bar_u + 8 * (j & (T::RR - 1));
            mbar_expect_tx(bar, uint32_t(T::ROW_TX));
            tma_load_3d(ring_u + (j & (T::RR - 1)) * T::SLOT, &maps.block, X0 - T::PADL, rs + j, T::P0, bar);
        };
        auto wait_row = [&](int j) {
            const int s = j & (T::RR - 1);
            mbar_wait(bar_u + 8 * s, (ph >> s) & 1u);
            ph ^= 1u << s;
        };
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < T::RR; ++j) if (j < NR) issue_row(j);
        }
        wait_row(0);

        R p_eta = ld(0, T::P_ETA, lc), p_zb = ld(0, T::P_ZB, lc);       // the cell below the face being formed
        R p_n = ld(0, T::P_N, lc);
        R qS = R(0);                                                      // flux through the southern face of row j-1, own n
        bool dry_s = true;

        for (int j = 1; j < NR; ++j) {
            const int y = rs + j;
            const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT, o_c = (j & (T::RR - 1)) * T::SLOT;
            wait_row(j);
            const R c_eta = ld(o_c, T::P_ETA, lc), c_zb = ld(o_c, T::P_ZB, lc), c_qy = ld(o_c, T::P_QY, lc), c_n = ld(o_c, T::P_N, lc);
            const bool dry_p = p_eta - p_zb < k.eps, dry_c = c_eta - c_zb < k.eps;
            if (a.reduce_mode == hp::kReduceDst && x_store && j + 1 < NR) {      // row y is updated in the next trip
                const int gy = y + g.gy0;
                if (!x_interior || gy < 1 || gy > g.grows - 2 || !stepping || (dry_c && dry_p)) prefetch_dst(d, static_cast<size_t>(y) * g.pitch + x);
            }
            // face between rows y-1 (down) and y (up); its discharge is stored in row y
            InFace<R> fy{R(0), R(0), R(0), false};
            if (stepping) fy = inertial_face<R>(k, gdt, c_qy, c_eta, c_zb, p_eta, p_zb, inv_delta);
            const R qN = inertial_q(fy, p_n);                              // as the cell below sees it (CLSchemeInertial.clc:107)
            const R qS_next = (c_n == p_n) ? qN : inertial_q(fy, c_n);     // as the cell above sees it (:109)

            if (j >= 2) {
                const int yc = y - 1, gyc = yc + g.gy0;
                const R p_qx = ld(o_m, T::P_QX, lc), p_qy = ld(o_m, T::P_QY, lc);
                // west face of row y-1: own cell is "up", the cell of lane-1 "down"; the discharge is the own qx
                const R w_n = ld(o_m, T::P_N, lw);
                InFace<R> fx{R(0), R(0), R(0), false};
                if (stepping) fx = inertial_face<R>(k, gdt, p_qx, p_eta, p_zb, ld(o_m, T::P_ETA, lw), ld(o_m, T::P_ZB, lw), inv_delta);
                const R qW = inertial_q(fx, p_n);                          // :110
                const R qE_for_west = (w_n == p_n) ? qW : inertial_q(fx, w_n);   // the same face as lane-1's eastern one (:108)
                const R qE = shfl_dn1(qE_for_west);
                const unsigned drym = __ballot_sync(FULL, dry_p);

                Cell<R> c{p_eta, ld(o_m, T::P_EMAX, lc), p_qx, p_qy};
                if (a.reduce_mode == hp::kReduceSrc && x_store) {
                    const R h = c.eta - p_zb;
                    if (h > k.eps10 && c.emax > R(-9999.0)) {
                        const R cc = fm_sqrt(k.g * h);
                        R sp = cc;
                        if (!k.simplified_speed) { const R rh = fm_rcp(h); sp = fm_max(hp_abs(c.qx * rh), hp_abs(c.qy * rh)) + cc; }
                        ws = fm_max(sp, ws);
                    }
                }
                bool wrote = false;
                if (x_interior && gyc >= 1 && gyc <= g.grows - 2 && stepping) {         // dt <= 0 returns first (:60-61)
                    if (c.emax <= R(-9999.0) || c.eta == R(-9999.0)) {
                        wrote = true;                                                   // disabled cell: copied through
                    } else {
                        const bool dry_e = (drym >> ((lane + 1) & 31)) & 1u, dry_w = (drym >> ((lane + 31) & 31)) & 1u;
                        if (!(dry_p && dry_c && dry_s && dry_e && dry_w)) {             // :92-99
                            c.qx = qW; c.qy = qS;                                       // :141-142
                            c.eta = c.eta + dt * ((qE - qW + qN - qS) * inv_delta);     // :145-152
                            if (c.eta > c.emax) c.emax = c.eta;
                            if (c.eta - p_zb < k.eps) c.eta = p_zb;
                            wrote = true;
                        }
                    }
                }
                if (x_store) {
                    const size_t id = static_cast<size_t>(yc) * g.pitch + x;
                    if (wrote) d.store(id, c);
                    if (a.reduce_mode == hp::kReduceDst) {
                        if (!wrote) { c.eta = d.eta[id]; c.emax = d.emax[id]; c.qx = d.qx[id]; c.qy = d.qy[id]; }
                        const R h = c.eta - p_zb;
                        if (h > k.eps10 && c.emax > R(-9999.0)) {
                            const R cc = fm_sqrt(k.g * h);
                            R sp = cc;
                            if (!k.simplified_speed) { const R rh = fm_rcp(h); sp = fm_max(hp_abs(c.qx * rh), hp_abs(c.qy * rh)) + cc; }
                            ws = fm_max(sp, ws);
                        }
                    }
                }
            }
            qS = qS_next; dry_s = dry_p;
            p_eta = c_eta; p_zb = c_zb; p_n = c_n;

            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && j - 1 + T::RR < NR) issue_row(j - 1 + T::RR);
        }
    }
    block_reduce_finalize<R>(ws, a, k);
}

// runs per CTA: pieces of at least 64 rows (a run pays two or three start-up rows), at most 32
static int march_runs(const StepArgs& a, int use, int nw, int grid) {
    const int nrows = a.y1 - a.y0;
    const int nstrips = (a.grid.cols + use - 1) / use, ngroups = (nstrips + nw - 1) / nw;
    const long long per_cta = static_cast<long long>(ngroups) * nrows / grid;
    const long long k = per_cta / 64;
    return static_cast<int>(k < 1 ? 1 : (k > 32 ? 32 : k));
}
static int march_grid(const StepArgs& a, int use, int nw, int ctas_per_sm, int sm_count) {
    const int nrows = a.y1 - a.y0;
    const int nstrips = (a.grid.cols + use - 1) / use, ngroups = (nstrips + nw - 1) / nw;
    const long long units = static_cast<long long>(ngroups) * nrows;
    const int min_rows = nrows < 8 ? nrows : 8;                      // amortise the three start-up rows of a run
    long long grid = (units + min_rows - 1) / min_rows;
    const long long cap = static_cast<long long>(ctas_per_sm) * sm_count;
    if (grid > cap) grid = cap;
    return static_cast<int>(grid < 1 ? 1 : grid);
}

template <class R> static int launch_godunov_march(const StepArgs& a_in, const TmaBlockMap& maps, int alt, int sm_count, cudaStream_t st) {
    using T = March<R, 1, false>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(godunov_step_march<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        cudaFuncSetAttribute(godunov_step_march<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int grid = march_grid(a, T::USE, T::NW, sizeof(R) == 8 ? HP_MARCH_GOD_CTAS64 : HP_MARCH_GOD_CTAS32, sm_count);
    a.total_ctas = grid; a.march_runs = march_runs(a, T::USE, T::NW, grid);
    if (alt) godunov_step_march<R, true><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    else godunov_step_march<R, false><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

template <class R> static int launch_inertial_march(const StepArgs& a_in, const TmaBlockMap& maps, int alt, int sm_count, cudaStream_t st) {
    using T = March<R, 1, false>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(inertial_step_march<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        cudaFuncSetAttribute(inertial_step_march<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int grid = march_grid(a, T::USE, T::NW, sizeof(R) == 8 ? HP_MARCH_INE_CTAS64 : 8, sm_count);
    a.total_ctas = grid; a.march_runs = march_runs(a, T::USE, T::NW, grid);
    if (alt) inertial_step_march<R, true><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    else inertial_step_march<R, false><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

}  // namespace HP_NS
