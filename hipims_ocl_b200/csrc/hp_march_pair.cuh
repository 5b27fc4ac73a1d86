// hp_march_pair.cuh -- "wide" marching kernels: one warp owns a strip of 64 columns, TWO adjacent columns per lane.
//
// Same skeleton as hp_march_kernels.cuh (rows through a per-warp TMA ring, y-direction carried in registers,
// x-direction between lanes by shuffles, runs of rows dealt round-robin to a persistent grid), with every quantity a
// pair P2<R> (hp_pair.cuh).  Why: the one-column kernels are bound by instruction issue, and a third to a half of what
// they issue is paid per ROW AND LANE, not per cell -- ring slots, barrier waits, TMA issue, validity predicates,
// addresses, flags, loop control, shuffles (profiles/r02_*: 373 warp instructions per 32 cell-updates for the fp32
// inertial kernel, 94 of them floating-point arithmetic).  With two columns per lane all of that is paid once per
// two cell-updates, the face between the lane's own columns needs no shuffle, pairs are loaded and stored as single
// 16- / 8-byte accesses, and in fp32 the arithmetic itself is packed (FFMA2 / FADD2 / FMUL2).
//
// Strip geometry (both precisions): lane l holds the columns X0 + 2 l and X0 + 2 l + 1, X0 = 60 strip - 2.  Lanes
// 1 .. 30 update their 60 cells; lane 0 and lane 31 only supply neighbours (raw values and, for lane 31's first
// column, the face it shares with the last updated cell), so 60 of 64 columns are updated per warp row (one-column
// kernels: 30 of 32 in fp64, 28 of 32 in fp32).  The TMA box is one row of 64 (fp64) / 68 (fp32) columns x 6 planes
// and starts on a 16-byte boundary; a lane's pair is 16- / 8-byte aligned in it.
#pragma once

#include "hp_march_kernels.cuh"
#include "hp_pair.cuh"

namespace HP_NS {

template <class R, bool ALT, int RING = 4> struct Wide {
    static constexpr int NW = hp::kMarchWarps, RR = RING, NP = 6;
    static constexpr int USE = hp::kWideUse;                               // cells updated per warp row
    static constexpr int PADL = sizeof(R) == 8 ? 0 : 2;                    // box columns in front of lane 0's first column
    static constexpr int BW = hp::wide_box_w(int(sizeof(R)));
    static constexpr int PLANE = BW * int(sizeof(R));
    static constexpr int ROW_TX = NP * PLANE;
    static constexpr int SLOT = (ROW_TX + 127) / 128 * 128;
    static constexpr int WARP_BYTES = RR * SLOT;
    static constexpr int SMEM_BYTES = NW * WARP_BYTES + NW * RR * 8;
    static constexpr int P0 = ALT ? 4 : 0;
    static constexpr int P_ETA = ALT ? 2 : 0, P_QX = ALT ? 3 : 1, P_QY = ALT ? 4 : 2, P_EMAX = ALT ? 5 : 3, P_ZB = ALT ? 0 : 4,
                         P_N = ALT ? 1 : 5;
    static_assert(BW >= 64 + PADL && (BW * sizeof(R)) % 16 == 0 && (USE * sizeof(R)) % 16 == 0 && (PADL * sizeof(R)) % 8 == 0, "box");
    static_assert(((2 - PADL) * int(sizeof(R))) % 16 == 0, "the box of strip 0 starts at column -2 - PADL: 16-byte aligned");
};

// Wave speed of stored cells for the CFL reduction (CLDynamicTimestep.clc:81-110), component by component.
template <class R>
__device__ __forceinline__ R wide_speed(const Params<R>& k, R eta, R emax, R qx, R qy, R zb) {
    const R h = eta - zb;
    if (!(h > k.eps10 && emax > R(-9999.0))) return R(0);
    const R cc = fm_sqrt(k.g * h);
    if (k.simplified_speed) return cc;
    const R rh = fm_rcp(h);
    return fm_max(hp_abs(qx * rh), hp_abs(qy * rh)) + cc;
}

// ... of a pair, straight-line: components outside `ok` ride along on a depth of one and contribute nothing
template <class R>
__device__ __forceinline__ R wide_speed2(const Params<R>& k, P2<R> eta, P2<R> emax, P2<R> qx, P2<R> qy, P2<R> zb, B2 ok, R ws) {
    const P2<R> h = eta - zb;
    ok = ok & (h > k.eps10) & (emax > R(-9999.0));
    const P2<R> hs = sel(ok, h, splat(R(1)));
    P2<R> sp = psqrt_pos(k.g * hs);
    if (!k.simplified_speed) {
        const P2<R> rh = prcp(hs);
        sp = pmax(pabs(qx * rh), pabs(qy * rh)) + sp;
    }
    sp = sel(ok, sp, splat(R(0)));
    return fm_max(fm_max(sp.a, sp.b), ws);
}

// =============================================================================================
// Partial inertial scheme (see inertial_step_march for the face / owner split of the Manning coefficient).
// =============================================================================================
// No per-lane branches around the face arithmetic: the wide kernels are bound by fixed-latency dependencies at their 16-24
// warps per SM, and every divergent branch (BSSY / BSYNC, a scheduling fence for ptxas) costs more than the work it
// skips only when the WHOLE warp skips it.  A dry component rides along on a depth of one and is selected away; what
// remains is one warp vote per face row (measured: fp32 83.8 -> 91.4, fp64 52.7 -> 55.9 G cell-updates/s).
template <class R> struct InFace2 { P2<R> num, A, qmax; B2 wet; };

// Must be called by all 32 lanes.
template <class R>
__device__ __forceinline__ InFace2<R> inertial_face2(const Params<R>& k, R gdt, P2<R> prev, P2<R> etaUp, P2<R> zUp, P2<R> etaDown,
                                                    P2<R> zDown, R inv_delta) {
    InFace2<R> f;
    const P2<R> h = pmax(etaDown, etaUp) - pmax(zUp, zDown);
    f.wet = !(h < k.eps);
    f.num = splat(R(0)); f.A = splat(R(0)); f.qmax = splat(R(0));
    if (__any_sync(0xffffffffu, any(f.wet))) {
        const P2<R> hs = sel(f.wet, h, splat(R(1)));
        const P2<R> rh = prcp(hs);
        f.num = fma2(-(gdt * hs), (etaDown - etaUp) * splat(inv_delta), prev);
        f.A = gdt * pabs(prev) * rh * rh * prcbrt(hs);
        f.qmax = (R(0.8) * hs) * psqrt_pos(k.g * hs);
    }
    return f;
}
template <class R> __device__ __forceinline__ P2<R> inertial_q2(const InFace2<R>& f, P2<R> n) {
    const P2<R> q = f.num * prcp(fma2(f.A, n * n, splat(R(1))));
    const P2<R> c = pmax(pmin(q, f.qmax), -f.qmax);
    return sel(f.wet, c, splat(R(0)));
}

#ifndef HP_WIDE_INE_CTAS64
#define HP_WIDE_INE_CTAS64 4
#endif
#ifndef HP_WIDE_INE_CTAS32
#define HP_WIDE_INE_CTAS32 6
#endif
template <class R, bool ALT>
__global__ void __launch_bounds__(hp::kMarchWarps * 32, sizeof(R) == 8 ? HP_WIDE_INE_CTAS64 : HP_WIDE_INE_CTAS32)
inertial_step_wide(const StepArgs a, const __grid_constant__ TmaBlockMap maps) {
    using T = Wide<R, ALT>;
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
    const long long total_runs = static_cast<long long>(gridDim.x) * a.march_runs;
    int run = 0;
    long long u = units * blockIdx.x / total_runs, u1 = units * (blockIdx.x + 1) / total_runs;

    // byte offsets inside a plane row of the box: the lane's pair, and the single column west of it (lane 0 has none
    // in the box and is given its own first column -- it only supplies neighbours)
    const int lc = (2 * lane + T::PADL) * int(sizeof(R));
    const int lw = lane > 0 ? lc - int(sizeof(R)) : lc;
    auto ld2 = [&](int row_off, int plane) -> P2<R> { return ld_pair<R>(ring + row_off + plane * T::PLANE + lc); };
    auto ldw = [&](int row_off, int plane) -> R { return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + lw); };
    const bool lane_owns = lane >= 1 && lane <= 30;

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

        const int X0 = strip * T::USE - 2;               // first column of lane 0
        const int xa = X0 + 2 * lane;                     // the lane's columns: xa, xa + 1
        const int rs = ya - 1;                            // first raw row of this run
        const int NR = yb - ya + 2;                       // raw rows 0 .. NR-1; rows 1 .. NR-2 are updated
        const B2 x_interior{xa >= 1 && xa <= g.cols - 2, xa + 1 >= 1 && xa + 1 <= g.cols - 2};
        const B2 x_store{lane_owns && xa < g.cols, lane_owns && xa + 1 < g.cols};
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

        P2<R> p_eta = ld2(0, T::P_ETA), p_zb = ld2(0, T::P_ZB), p_n = ld2(0, T::P_N);   // the cells below the faces being formed
        P2<R> qS = splat(R(0));                                                          // flux through their southern faces, own n
        B2 dry_s{true, true};

        for (int j = 1; j < NR; ++j) {
            const int y = rs + j;
            const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT, o_c = (j & (T::RR - 1)) * T::SLOT;
            wait_row(j);
            const P2<R> c_eta = ld2(o_c, T::P_ETA), c_zb = ld2(o_c, T::P_ZB), c_qy = ld2(o_c, T::P_QY), c_n = ld2(o_c, T::P_N);
            const B2 dry_p = (p_eta - p_zb) < k.eps, dry_c = (c_eta - c_zb) < k.eps;
            if (a.reduce_mode == hp::kReduceDst && any(x_store) && j + 1 < NR) {      // row y is updated in the next trip
                const int gy = y + g.gy0;
                if (!(x_interior.a && x_interior.b) || gy < 1 || gy > g.grows - 2 || !stepping || any(dry_c & dry_p))
                    prefetch_dst(d, static_cast<size_t>(y) * g.pitch + xa);          // the pair shares its 32-byte sectors
            }
            // faces between rows y-1 (down) and y (up); their discharge is stored in row y
            // (dt <= 0: the arithmetic runs on, nothing is written -- see `rows_ok` below)
            const InFace2<R> fy = inertial_face2<R>(k, gdt, c_qy, c_eta, c_zb, p_eta, p_zb, inv_delta);
            const P2<R> qN = inertial_q2(fy, p_n);                         // as the cells below see them (CLSchemeInertial.clc:107)
            P2<R> qS_next = qN;                                            // as the cells above see them (:109)
            if (__any_sync(FULL, c_n.a != p_n.a || c_n.b != p_n.b)) qS_next = sel(c_n == p_n, qN, inertial_q2(fy, c_n));

            if (j >= 2) {
                const int yc = y - 1, gyc = yc + g.gy0;
                const P2<R> p_qx = ld2(o_m, T::P_QX), p_qy = ld2(o_m, T::P_QY);
                // west faces of row y-1: the own cell is "up", its western neighbour "down" -- the last column of lane-1
                // (from the box) for the first column, the lane's own first column for the second; the discharge is the own qx
                const P2<R> w_eta{ldw(o_m, T::P_ETA), p_eta.a}, w_zb{ldw(o_m, T::P_ZB), p_zb.a}, w_n{ldw(o_m, T::P_N), p_n.a};
                const InFace2<R> fx = inertial_face2<R>(k, gdt, p_qx, p_eta, p_zb, w_eta, w_zb, inv_delta);
                const P2<R> qW = inertial_q2(fx, p_n);                     // :110
                P2<R> qE_for_west = qW;                                    // the same faces as the western neighbours see them (:108)
                if (__any_sync(FULL, w_n.a != p_n.a || w_n.b != p_n.b)) qE_for_west = sel(w_n == p_n, qW, inertial_q2(fx, w_n));
                // east faces: the second column's western face for the first column, lane+1's first for the second
                const P2<R> qE{qE_for_west.b, shfl_dn1(qE_for_west.a)};
                const unsigned drym_a = __ballot_sync(FULL, dry_p.a), drym_b = __ballot_sync(FULL, dry_p.b);
                const B2 dry_w{((drym_b >> ((lane + 31) & 31)) & 1u) != 0, dry_p.a};
                const B2 dry_e{dry_p.b, ((drym_a >> ((lane + 1) & 31)) & 1u) != 0};

                P2<R> eta = p_eta, emax = ld2(o_m, T::P_EMAX), qx = p_qx, qy = p_qy;
                if (a.reduce_mode == hp::kReduceSrc) ws = wide_speed2(k, eta, emax, qx, qy, p_zb, x_store, ws);
                const bool rows_ok = gyc >= 1 && gyc <= g.grows - 2 && stepping;        // dt <= 0 returns first (:60-61)
                const B2 disabled = (emax <= R(-9999.0)) | (eta == splat(R(-9999.0)));
                const B2 all_dry = dry_p & dry_c & dry_s & dry_e & dry_w;               // :92-99
                const B2 upd = (x_interior & !disabled & !all_dry) & rows_ok;
                const B2 wrote = (x_interior & (disabled | !all_dry)) & rows_ok;        // a disabled cell is copied through
                {
                    const P2<R> n_eta = fma2(splat(dt), (((qE - qW) + qN) - qS) * splat(inv_delta), eta);       // :145-152
                    const P2<R> n_emax = sel(n_eta > emax, n_eta, emax);
                    const P2<R> f_eta = sel((n_eta - p_zb) < k.eps, p_zb, n_eta);
                    eta = sel(upd, f_eta, eta); emax = sel(upd, n_emax, emax);
                    qx = sel(upd, qW, qx); qy = sel(upd, qS, qy);                       // :141-142
                }
                if (any(x_store)) {
                    const size_t id = static_cast<size_t>(yc) * g.pitch + xa;
                    if (wrote.a && wrote.b && x_store.b) {
                        st_pair(d.eta + id, eta); st_pair(d.emax + id, emax); st_pair(d.qx + id, qx); st_pair(d.qy + id, qy);
                    } else {
                        if (wrote.a && x_store.a) { d.eta[id] = eta.a; d.emax[id] = emax.a; d.qx[id] = qx.a; d.qy[id] = qy.a; }
                        if (wrote.b && x_store.b) { d.eta[id + 1] = eta.b; d.emax[id + 1] = emax.b; d.qx[id + 1] = qx.b; d.qy[id + 1] = qy.b; }
                    }
                    if (a.reduce_mode == hp::kReduceDst) {
                        // cells left unwritten enter the reduction with what the destination holds
                        if (!wrote.a && x_store.a) { eta.a = d.eta[id]; emax.a = d.emax[id]; qx.a = d.qx[id]; qy.a = d.qy[id]; }
                        if (!wrote.b && x_store.b) { eta.b = d.eta[id + 1]; emax.b = d.emax[id + 1]; qx.b = d.qx[id + 1]; qy.b = d.qy[id + 1]; }
                        ws = wide_speed2(k, eta, emax, qx, qy, p_zb, x_store, ws);
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

template <class R> static int launch_inertial_wide(const StepArgs& a_in, const TmaBlockMap& maps, int alt, int sm_count, cudaStream_t st) {
    using T = Wide<R, false>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(inertial_step_wide<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        cudaFuncSetAttribute(inertial_step_wide<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int grid = march_grid(a, T::USE, T::NW, sizeof(R) == 8 ? HP_WIDE_INE_CTAS64 : HP_WIDE_INE_CTAS32, sm_count);
    a.total_ctas = grid; a.march_runs = march_runs(a, T::USE, T::NW, grid);
    if (alt) inertial_step_wide<R, true><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    else inertial_step_wide<R, false><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}


// =============================================================================================
// MUSCL-Hancock, two columns per lane.  Same row structure as mh_step_march2 (hp_march_mh.cuh): predictor of row y,
// face between rows y-1 and y, corrector of row y-1, x-faces of row y reduced to three sums, hand-over -- nothing is
// rotated, what crosses to the next trip is the northern face estimate, the sums minus the southern flux, the
// southern bed / depth and a stop count, all as pairs.  Of a lane's two x-faces only ONE needs the neighbouring lane
// (the face between its first column and lane-1's second); the other lies between its own columns.  Lane 0's second
// column and lane 31's first are halo columns (predictor only), their outer columns supply raw values.
// =============================================================================================
struct I2 { int a, b; };
template <class R> struct FaceOut2 { P2<R> m, n, t, zmax, hL, hR; I2 stopL, stopR; };

template <class R> __device__ __forceinline__ P2<R> ppos(P2<R> v) { return P2<R>{fm_pos_s(v.a), fm_pos_s(v.b)}; }
__device__ __forceinline__ P2<double> pcelerity(double g, P2<double> h) { return psqrt_pos(fma2(splat(g), h, splat(1.0e-300))); }
__device__ __forceinline__ P2<float> pcelerity(float g, P2<float> h) { return psqrt_pos(fma2(splat(g), h, splat(1.0e-37f))); }
template <class R> __device__ __forceinline__ P2<R> pminmod_sw(P2<R> x, P2<R> y, I2 off) {
    return P2<R>{minmod_sw(x.a, y.a, off.a), minmod_sw(x.b, y.b, off.b)};
}
template <class R> __device__ __forceinline__ P2<R> pshfl_up_b(P2<R> v) { return P2<R>{shfl_up1(v.b), v.a}; }     // the western neighbours' values
template <class R> __device__ __forceinline__ P2<R> pshfl_dn_a(P2<R> v) { return P2<R>{v.b, shfl_dn1(v.a)}; }     // the eastern neighbours' values

// Two faces in the normal frame, solved once each for the cells on either side (face_solve2 of hp_march_kernels.cuh,
// component by component).  The wet fast path is taken when every face of the warp's row is wet on both sides; the
// supercritical exits and the dry cases are selects behind one warp vote each.  Must be called by all 32 lanes.
// CACHED (first-order callers, whose face values are cell values): cL / cR are the celerities of the two cells on their own
// beds, sqrt(g (eta - z)); a side whose bed IS the face's bed in all 64 columns -- flat or uniformly sloping ground, one
// warp vote per side -- takes them instead of a new root.
template <class R, bool CACHED, class QOwnL, class QOwnR>
__device__ __forceinline__ void face_solve_pair_c(const Params<R>& k, P2<R> etaL, P2<R> zL, P2<R> unL, P2<R> utL, P2<R> cL, P2<R> etaR,
                                                  P2<R> zR, P2<R> unR, P2<R> utR, P2<R> cR, QOwnL qOwnL, QOwnR qOwnR, FaceOut2<R>& o) {
    constexpr unsigned FULL = 0xffffffffu;
    const R hg = R(0.5) * k.g;
    const P2<R> zero = splat(R(0));
    const P2<R> zmax = pmax(zL, zR);
    const P2<R> dL = etaL - zmax, dR = etaR - zmax;
    o.zmax = zmax;
    const bool ownL = CACHED && __all_sync(FULL, zmax.a == zL.a && zmax.b == zL.b);
    const bool ownR = CACHED && __all_sync(FULL, zmax.a == zR.a && zmax.b == zR.b);
    const B2 wet = (dL > k.eps) & (dR > k.eps);
    if (__all_sync(FULL, wet.a && wet.b)) {
        o.hL = dL; o.hR = dR; o.stopL = I2{0, 0}; o.stopR = I2{0, 0};
        const P2<R> aL = ownL ? cL : psqrt_pos(k.g * dL), aR = ownR ? cR : psqrt_pos(k.g * dR);
        const P2<R> qnL = dL * unL, qnR = dR * unR;
        const P2<R> as = pabs(fma2(splat(R(0.25)), unL - unR, R(0.5) * (aL + aR)));
        const P2<R> us = fma2(splat(R(0.5)), unL + unR, aL) - aR;
        const P2<R> sL = pmin(unL - aL, us - as);
        const P2<R> sR = pmax(unR + aR, us + as);
        const P2<R> FLn = fma2(unL, qnL, (hg * dL) * dL), FRn = fma2(unR, qnR, (hg * dR) * dR);
        const P2<R> inv = prcp(sR - sL);
        const P2<R> ss = sL * sR;
        const P2<R> f1 = fma2(ss, dR - dL, fma2(sR, qnL, -(sL * qnR))) * inv;
        const P2<R> f2 = fma2(ss, qnR - qnL, fma2(sR, FLn, -(sL * FRn))) * inv;
        const B2 supL = !(sL < R(0)), supR = sR < R(0);
        if (__all_sync(FULL, !(supL.a || supL.b || supR.a || supR.b))) {
            o.m = f1; o.n = f2; o.t = f1 * sel(!(f1 < R(0)), utL, utR);
        } else {                                        // supercritical somewhere: the upwind state's own flux
            o.m = sel(supL, qnL, sel(supR, qnR, f1));
            o.n = sel(supL, FLn, sel(supR, FRn, f2));
            o.t = o.m * sel(supL, utL, sel(supR, utR, sel(!(f1 < R(0)), utL, utR)));
        }
        return;
    }
    // ---- general path: a dry side, stop flags (wet/dry fronts only) --------------------------------
    const P2<R> hL = ppos(dL), hR = ppos(dR);
    o.hL = hL; o.hR = hR;
    const B2 lowL = hL <= k.eps, lowR = hR <= k.eps;
    {
        const I2 both{((lowR.a && unL.a < R(0)) ? 1 : 0) + ((lowL.a && unR.a > R(0)) ? 1 : 0),
                      ((lowR.b && unL.b < R(0)) ? 1 : 0) + ((lowL.b && unR.b > R(0)) ? 1 : 0)};
        o.stopL = both; o.stopR = both;
        if (__any_sync(FULL, any(lowL | lowR))) {        // the owners' raw discharge is read here only
            const P2<R> qL = qOwnL(), qR = qOwnR();
            o.stopL = I2{both.a + ((lowL.a && qL.a > R(0)) ? 1 : 0), both.b + ((lowL.b && qL.b > R(0)) ? 1 : 0)};
            o.stopR = I2{both.a + ((lowR.a && qR.a < R(0)) ? 1 : 0), both.b + ((lowR.b && qR.b < R(0)) ? 1 : 0)};
        }
    }
    const B2 dryL = hL < k.eps, dryR = hR < k.eps;
    const B2 dd = dryL & dryR;
    const P2<R> hm = R(0.5) * (hL + hR);
    const P2<R> ddn = (hg * hm) * hm;
    if (__all_sync(FULL, dd.a && dd.b)) { o.m = zero; o.n = ddn; o.t = zero; return; }
    unL = sel(dryL, zero, unL); utL = sel(dryL, zero, utL);
    unR = sel(dryR, zero, unR); utR = sel(dryR, zero, utR);
    const P2<R> aL = ownL ? cL : pcelerity(k.g, hL), aR = ownR ? cR : pcelerity(k.g, hR);
    const P2<R> qnL = hL * unL, qnR = hR * unR;
    const P2<R> as = pabs(fma2(splat(R(0.25)), unL - unR, R(0.5) * (aL + aR)));
    const P2<R> us = fma2(splat(R(0.5)), unL + unR, aL) - aR;
    const P2<R> sL = sel(dryL, fma2(splat(R(-2)), aR, unR), pmin(unL - aL, us - as));
    const P2<R> sR = sel(dryR, fma2(splat(R(2)), aL, unL), pmax(unR + aR, us + as));
    const P2<R> FLn = fma2(unL, qnL, (hg * hL) * hL), FRn = fma2(unR, qnR, (hg * hR) * hR);
    const P2<R> inv = prcp(sR - sL);
    const P2<R> ss = sL * sR;
    const P2<R> f1 = fma2(ss, hR - hL, fma2(sR, qnL, -(sL * qnR))) * inv;
    const P2<R> f2 = fma2(ss, qnR - qnL, fma2(sR, FLn, -(sL * FRn))) * inv;
    const B2 supL = !(sL < R(0)), supR = sR < R(0);
    const P2<R> m = sel(supL, qnL, sel(supR, qnR, f1));
    o.m = sel(dd, zero, m);
    o.n = sel(dd, ddn, sel(supL, FLn, sel(supR, FRn, f2)));
    o.t = sel(dd, zero, m * sel(supL, utL, sel(supR, utR, sel(!(f1 < R(0)), utL, utR))));
}

template <class R, class QOwnL, class QOwnR>
__device__ __forceinline__ void face_solve_pair(const Params<R>& k, P2<R> etaL, P2<R> zL, P2<R> unL, P2<R> utL, P2<R> etaR, P2<R> zR,
                                                P2<R> unR, P2<R> utR, QOwnL qOwnL, QOwnR qOwnR, FaceOut2<R>& o) {
    face_solve_pair_c<R, false>(k, etaL, zL, unL, utL, etaL, etaR, zR, unR, utR, etaR, qOwnL, qOwnR, o);
}

// Point-implicit friction on a pair (friction_fast component by component); `go` masks the components it applies to.
template <class R> __device__ __forceinline__ void friction_pair(const Params<R>& k, P2<R> h, P2<R> rh, P2<R>& qx, P2<R>& qy, P2<R> n, R dt, B2 go) {
    const P2<R> sx = qx * qx, sy = qy * qy;
    const P2<R> q2 = sx + sy;
    const P2<R> q = pcelerity(R(1), q2);                 // sqrt(q2), a zero discharge giving a harmless tiny root
    go = go & !(h < k.eps) & !(q < k.eps);
    const P2<R> A = (dt * k.g) * n * n * rh * rh * prcbrt(h);             // dt * Cf / h^2
    const P2<R> aq2 = A * q2;
    const P2<R> nx = fma2(-(qx * aq2), prcp(fma2(A, q2 + sx, q)), qx);
    const P2<R> ny = fma2(-(qy * aq2), prcp(fma2(A, q2 + sy, q)), qy);
    qx = sel(go, nx, qx); qy = sel(go, ny, qy);
}

#ifndef HP_WIDE_MH_CTAS64
#define HP_WIDE_MH_CTAS64 2
#endif
#ifndef HP_WIDE_MH_CTAS32
#define HP_WIDE_MH_CTAS32 4
#endif
#ifndef HP_WIDE_MH_WARPS
#define HP_WIDE_MH_WARPS 4
#endif

template <class R, bool ALT>
__global__ void __launch_bounds__(hp::kMarchWarps * 32, sizeof(R) == 8 ? HP_WIDE_MH_CTAS64 : HP_WIDE_MH_CTAS32)
mh_step_wide(const StepArgs a, const __grid_constant__ TmaBlockMap maps) {
    using T = Wide<R, ALT>;
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
    const P2<R> half = splat(R(0.5)), mhalf = splat(R(-0.5)), zero = splat(R(0));
    const R hdt = R(0.5) * dt;
    const MutView<R> d(a.dst);
    const bool stepping = dt > R(0);

    const int nrows = a.y1 - a.y0;
    const int nstrips = (g.cols + T::USE - 1) / T::USE;
    const int ngroups = (nstrips + T::NW - 1) / T::NW;
    const long long units = static_cast<long long>(ngroups) * nrows;
    const long long total_runs = static_cast<long long>(gridDim.x) * a.march_runs;
    int run = 0;
    long long u = units * blockIdx.x / total_runs, u1 = units * (blockIdx.x + 1) / total_runs;

    // byte offsets inside a plane row of the box: the lane's pair, the column west of it and the column east of it
    // (lane 0 / lane 31 have none in the box and are given one of their own: their outer columns only supply raw values)
    constexpr int SZ = int(sizeof(R));
    const int lc = (2 * lane + T::PADL) * SZ;
    const int lw = lane > 0 ? lc - SZ : lc;
    const int le = lane < 31 ? lc + 2 * SZ : lc + SZ;
    auto ld2 = [&](int row_off, int plane) -> P2<R> { return ld_pair<R>(ring + row_off + plane * T::PLANE + lc); };
    auto ldw = [&](int row_off, int plane) -> R { return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + lw); };
    auto lde = [&](int row_off, int plane) -> R { return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + le); };
    // re-read of raw values the predictor has overwritten in registers (only wet/dry fronts ask for them)
    auto ld2_again = [&](int row_off, int plane) -> P2<R> {
        const volatile R* p = reinterpret_cast<const volatile R*>(ring + row_off + plane * T::PLANE + lc);
        return P2<R>{p[0], p[1]};
    };
    auto ldw_again = [&](int row_off, int plane) -> R { return *reinterpret_cast<const volatile R*>(ring + row_off + plane * T::PLANE + lw); };
    auto flag_of = [&](R em) -> int { return (em <= R(-9998.0) ? 1 : 0) | (em < k.eps ? 2 : 0); };
    auto flags_of = [&](P2<R> em) -> I2 { return I2{flag_of(em.a), flag_of(em.b)}; };
    const bool lane_owns = lane >= 1 && lane <= 30;
    // wave speed of a stored cell for the CFL reduction (CLDynamicTimestep.clc:81-110); rh = 1/h where the caller has it
    auto speed_of = [&](R h, R qx, R qy, bool have_rh, R rh) -> R {
        const R cc = fm_sqrt_pos(k.g * h);
        if (k.simplified_speed) return cc;
        if (!have_rh) rh = fm_rcp(h);
        return fm_max(hp_abs(qx * rh), hp_abs(qy * rh)) + cc;
    };

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

        const int X0 = strip * T::USE - 2;               // first column of lane 0
        const int xa = X0 + 2 * lane, xb = xa + 1;        // the lane's columns
        const int rs = ya - 2;                            // first raw row of this run
        const int J = yb - ya + 2;                        // raw rows 0 .. J+1, predictor rows 1 .. J, updated rows 2 .. J-1
        const B2 x_store{lane_owns && xa < g.cols, lane_owns && xb < g.cols};
        const B2 x_valid{xa >= 1 && xa <= g.cols - 2, xb >= 1 && xb <= g.cols - 2};      // predictor runs on 1 .. cols-2
        const B2 x_interior{xa >= 2 && xa <= g.cols - 3, xb >= 2 && xb <= g.cols - 3};   // the ring of two is frozen
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
        // one row of the destination: whole pairs where both columns are stored, single columns at the domain's edge
        auto store_row = [&](int yrow, P2<R> eta, P2<R> emax, P2<R> qx, P2<R> qy) {
            const size_t id = static_cast<size_t>(yrow) * g.pitch + xa;
            if (x_store.b) {
                st_pair(d.eta + id, eta); st_pair(d.emax + id, emax); st_pair(d.qx + id, qx); st_pair(d.qy + id, qy);
            } else if (x_store.a) {
                d.eta[id] = eta.a; d.emax[id] = emax.a; d.qx[id] = qx.a; d.qy[id] = qy.a;
            }
        };
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < T::RR; ++j) if (j <= J + 1) issue_row(j);
        }
        wait_row(0);
        wait_row(1);

        if (!stepping) {
            // dt <= 0: the reference's kernels return (CLSchemeMUSCLHancock.clc:62-63, 581-582); the ping-pong copies the
            // state through and the reduction sees it
            for (int j = 1; j <= J; ++j) {
                const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT;
                wait_row(j + 1);
                if (j >= 3 && x_store.a) {
                    const P2<R> eta = ld2(o_m, T::P_ETA), emax = ld2(o_m, T::P_EMAX), qx = ld2(o_m, T::P_QX), qy = ld2(o_m, T::P_QY);
                    store_row(rs + j - 1, eta, emax, qx, qy);
                    if (a.reduce_mode != hp::kReduceNone) {
                        const P2<R> h = eta - ld2(o_m, T::P_ZB);
                        if (h.a > k.eps10 && emax.a > R(-9999.0)) ws = fm_max(speed_of(h.a, qx.a, qy.a, false, R(0)), ws);
                        if (x_store.b && h.b > k.eps10 && emax.b > R(-9999.0)) ws = fm_max(speed_of(h.b, qx.b, qy.b, false, R(0)), ws);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && j - 1 + T::RR <= J + 1) issue_row(j - 1 + T::RR);
            }
            continue;
        }

        // ---- state carried from row to row; each item is dead before its successor is computed ---------------
        I2 f_m2{0, 0}, f_m1 = flags_of(ld2(0 * T::SLOT, T::P_EMAX)), f_c = flags_of(ld2(1 * T::SLOT, T::P_EMAX));
        I2 f_ew_prev{0, 0};                                  // bit1 flags of the x-neighbours of the previous row: W | E<<2
        P2<R> Le = zero, Lh = zero, Lun = zero, Lut = zero;  // northern face estimate of the previous row: eta, depth, v, u
        P2<R> Aeta = zero, Aqx = zero, Aqy = zero;           // x-face sums of the previous row minus its southern flux
        P2<R> bS = zero, sH = zero;                          // its southern face: reconstructed bed (owner side), neighbour depth
        I2 cStop{0, 0};                                      // its stop count so far (west, east, south)
        bool dr_m2 = false, dr_m1 = false, dr_c = false;     // rows j-2, j-1, j exactly dry and at rest in all 64 columns
        bool in_skip = false;                                // the previous trip copied its row through

        for (int j = 1; j <= J; ++j) {
            const int y = rs + j, gy = y + g.gy0;
            const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT, o_c = (j & (T::RR - 1)) * T::SLOT,
                      o_p = ((j + 1) & (T::RR - 1)) * T::SLOT;
            wait_row(j + 1);
            const P2<R> eta = ld2(o_c, T::P_ETA), qx = ld2(o_c, T::P_QX), qy = ld2(o_c, T::P_QY), zb = ld2(o_c, T::P_ZB);
            // (tested on every fourth row, and on every row while the rows below are dry: wet regions pay almost nothing)
            dr_c = false;
            P2<R> emax_c = zero;
            if (dr_m1 || (j & 3) == 0) {
                emax_c = ld2(o_c, T::P_EMAX);
                dr_c = __all_sync(FULL, eta.a == zb.a && qx.a == R(0) && qy.a == R(0) && !(eta.a > emax_c.a) &&
                                            eta.b == zb.b && qx.b == R(0) && qy.b == R(0) && !(eta.b > emax_c.b));
            }

            if (dr_m2 && dr_m1 && dr_c) {
                // rows y-2, y-1, y exactly dry in every column (so j >= 3): the cells of row y-1 cannot change; copy them
                // and keep the one flag bit a dry row can have (see mh_step_march2)
                if (x_store.a) store_row(y - 1, ld2(o_m, T::P_ETA), ld2(o_m, T::P_EMAX), ld2(o_m, T::P_QX), ld2(o_m, T::P_QY));
                f_m2 = f_m1; f_m1 = I2{emax_c.a < k.eps ? 2 : 0, emax_c.b < k.eps ? 2 : 0};
                in_skip = true;
            } else {
                if (in_skip) {
                    // leaving the copy-through mode: row y-1 is exactly dry -- its northern face estimate is its level with no
                    // depth and no velocity, its sums are empty; f_m1 / f_m2 were kept, the rest of the flags is re-derived
                    in_skip = false;
                    Le = ld2(o_m, T::P_ETA); Lh = zero; Lun = zero; Lut = zero;
                    Aeta = zero; Aqx = zero; Aqy = zero; bS = Le; sH = zero; cStop = I2{0, 0};
                    f_c = flags_of(ld2(o_c, T::P_EMAX));
                    const int wA = __shfl_up_sync(FULL, f_m1.b, 1), eB = __shfl_down_sync(FULL, f_m1.a, 1);
                    f_ew_prev = I2{(wA >> 1) | ((f_m1.b >> 1) << 2), (f_m1.a >> 1) | ((eB >> 1) << 2)};
                }
                const I2 f_p = flags_of(ld2(o_p, T::P_EMAX));
                // flags of the x-neighbours of row y: of a lane's two columns each is the other's neighbour on one side
                const I2 f_w{__shfl_up_sync(FULL, f_c.b, 1), f_c.a}, f_e{f_c.b, __shfl_down_sync(FULL, f_c.a, 1)};
                // ---- predictor of row y (CLSchemeMUSCLHancock.clc:301-382) ---------------------------
                P2<R> ce = eta, cqx = qx, cqy = qy;
                P2<R> sxE = zero, sxH = zero, sxQx = zero, sxQy = zero, syE = zero, syH = zero, syQx = zero, syQy = zero;
                {
                    const bool valid_y = gy >= 1 && gy <= g.grows - 2 && y >= 1 && y <= g.rows - 2;
                    const P2<R> h = eta - zb;
                    const B2 pred{valid_y && x_valid.a && !(h.a < R(1E-5)) && !((f_p.a | f_e.a | f_m1.a | f_w.a) & 1),
                                  valid_y && x_valid.b && !(h.b < R(1E-5)) && !((f_p.b | f_e.b | f_m1.b | f_w.b) & 1)};
                    if (__any_sync(FULL, any(pred))) {             // a warp vote, not a per-lane branch (see inertial_face2)
                        const P2<R> etaE{eta.b, lde(o_c, T::P_ETA)}, etaW{ldw(o_c, T::P_ETA), eta.a};
                        const P2<R> etaN = ld2(o_p, T::P_ETA), etaS = ld2(o_m, T::P_ETA);
                        const P2<R> hE = etaE - P2<R>{zb.b, lde(o_c, T::P_ZB)}, hW = etaW - P2<R>{ldw(o_c, T::P_ZB), zb.a};
                        const P2<R> hN = etaN - ld2(o_p, T::P_ZB), hS = etaS - ld2(o_m, T::P_ZB);
                        const P2<R> qxE{qx.b, lde(o_c, T::P_QX)}, qxW{ldw(o_c, T::P_QX), qx.a};
                        const P2<R> qyE{qy.b, lde(o_c, T::P_QY)}, qyW{ldw(o_c, T::P_QY), qy.a};
                        // a dry neighbour drops the slopes of its direction (:301-320); the switch rides in the limiter's sign test
                        const I2 xoff{(fm_lt_opaque(hW.a, k.eps) | fm_lt_opaque(hE.a, k.eps)) ? int(0x80000000u) : 0,
                                      (fm_lt_opaque(hW.b, k.eps) | fm_lt_opaque(hE.b, k.eps)) ? int(0x80000000u) : 0};
                        const I2 yoff{(fm_lt_opaque(hS.a, k.eps) | fm_lt_opaque(hN.a, k.eps)) ? int(0x80000000u) : 0,
                                      (fm_lt_opaque(hS.b, k.eps) | fm_lt_opaque(hN.b, k.eps)) ? int(0x80000000u) : 0};
                        sxE = pminmod_sw(eta - etaW, etaE - eta, xoff); sxH = pminmod_sw(h - hW, hE - h, xoff);
                        sxQx = pminmod_sw(qx - qxW, qxE - qx, xoff); sxQy = pminmod_sw(qy - qyW, qyE - qy, xoff);
                        syE = pminmod_sw(eta - etaS, etaN - eta, yoff); syH = pminmod_sw(h - hS, hN - h, yoff);
                        syQx = pminmod_sw(qx - ld2(o_m, T::P_QX), ld2(o_p, T::P_QX) - qx, yoff);
                        syQy = pminmod_sw(qy - ld2(o_m, T::P_QY), ld2(o_p, T::P_QY) - qy, yoff);
                        // a component outside `pred` keeps zero slopes and its raw state, like a lane that skips this block
                        sxE = sel(pred, sxE, zero); sxH = sel(pred, sxH, zero); sxQx = sel(pred, sxQx, zero); sxQy = sel(pred, sxQy, zero);
                        syE = sel(pred, syE, zero); syH = sel(pred, syH, zero); syQx = sel(pred, syQx, zero); syQy = sel(pred, syQy, zero);
                        // Face depths h +- s/2 with |s| <= |h - h_neighbour| and both >= 0: never below h/2 >= 5e-6 (a component
                        // outside `pred` is given a depth of one), so the reference's zero-velocity guard (:333-346) cannot fire.
                        const P2<R> hp = sel(pred, h, splat(R(1)));
                        const P2<R> hEf = fma2(half, sxH, hp), hWf = fma2(mhalf, sxH, hp), hNf = fma2(half, syH, hp), hSf = fma2(mhalf, syH, hp);
                        const P2<R> qxEf = fma2(half, sxQx, qx), qxWf = fma2(mhalf, sxQx, qx), qyEf = fma2(half, sxQy, qy), qyWf = fma2(mhalf, sxQy, qy);
                        const P2<R> qxNf = fma2(half, syQx, qx), qxSf = fma2(mhalf, syQx, qx), qyNf = fma2(half, syQy, qy), qySf = fma2(mhalf, syQy, qy);
                        const P2<R> uE = qxEf * prcp(hEf), uW = qxWf * prcp(hWf);
                        const P2<R> vN = qyNf * prcp(hNf), vS = qySf * prcp(hSf);
                        const P2<R> dEta = ((qxEf - qxWf) + (qyNf - qySf)) * splat(inv_delta);
                        const P2<R> dQx = fma2(hg * sxE, hEf + hWf, fma2(-vS, qxSf, fma2(vN, qxNf, fma2(-uW, qxWf, uE * qxEf)))) * splat(inv_delta);
                        const P2<R> dQy = fma2(hg * syE, hNf + hSf, fma2(-vS, qySf, fma2(vN, qyNf, fma2(-uW, qyWf, uE * qyEf)))) * splat(inv_delta);
                        ce = sel(pred & !(pabs(dEta) < k.eps), fma2(splat(-hdt), dEta, eta), eta);          // |D| < eps => 0 (:363-371)
                        cqx = sel(pred & !(pabs(dQx) < k.eps), fma2(splat(-hdt), dQx, qx), qx);
                        cqy = sel(pred & !(pabs(dQy) < k.eps), fma2(splat(-hdt), dQy, qy), qy);
                    }
                }
                const P2<R> ch = ce - zb;

                // ---- faces between rows y-1 (left, carried) and y (right); normal = y.  At j == 1 there is no row below:
                // the result only feeds values that are dead before anything reads them ----------------------
                FaceOut2<R> fy;
                const P2<R> etaR = fma2(mhalf, syE, ce);                  // southern face estimate of row y
                {
                    const P2<R> hfR = fma2(mhalf, syH, ch);
                    const P2<R> rR = sel(hfR <= k.eps, zero, prcp(hfR));                          // :1140-1150
                    face_solve_pair<R>(k, Le, Le - Lh, Lun, Lut, etaR, etaR - hfR, fma2(mhalf, syQy, cqy) * rR, fma2(mhalf, syQx, cqx) * rR,
                                       [&] { return ld2_again(o_m, T::P_QY); }, [&] { return ld2_again(o_c, T::P_QY); }, fy);
                }
                if (j >= 3) {
                    // ---- corrector of row y-1 (CLSchemeMUSCLHancock.clc:596-800) -----------------
                    const int gyc = gy - 1;
                    P2<R> c_eta = ld2(o_m, T::P_ETA), c_emax = ld2(o_m, T::P_EMAX), c_qx = ld2(o_m, T::P_QX), c_qy = ld2(o_m, T::P_QY);
                    const P2<R> pzb = ld2(o_m, T::P_ZB);
                    const bool rows_ok = gyc >= 2 && gyc <= g.grows - 3;
                    const B2 live = x_interior & !((c_emax <= R(-9999.0)) | (c_eta == splat(R(-9999.0))));
                    const I2 dry{((c_eta.a - pzb.a < k.eps) ? 1 : 0) + (f_c.a >> 1) + (f_m2.a >> 1) + (f_ew_prev.a & 1) + (f_ew_prev.a >> 2),
                                 ((c_eta.b - pzb.b < k.eps) ? 1 : 0) + (f_c.b >> 1) + (f_m2.b >> 1) + (f_ew_prev.b & 1) + (f_ew_prev.b >> 2)};
                    const B2 upd{rows_ok && live.a && dry.a < 5, rows_ok && live.b && dry.b < 5};
                    if (__any_sync(FULL, any(upd))) {
                        const P2<R> bN = pmin(fy.zmax, Le);
                        const P2<R> dEta = (Aeta + fy.m) * splat(inv_delta);
                        const P2<R> dQx = (Aqx + fy.t) * splat(inv_delta);
                        const P2<R> dQy = fma2(hg * (bN - bS), fy.hR + sH, Aqy + fy.n) * splat(inv_delta);
                        const B2 stop{cStop.a + fy.stopL.a > 0, cStop.b + fy.stopL.b > 0};
                        P2<R> n_qx = sel(stop, zero, c_qx), n_qy = sel(stop, zero, c_qy);
                        P2<R> n_eta = sel(!(pabs(dEta) < k.eps), fma2(splat(-dt), dEta, c_eta), c_eta);   // |D| < eps => 0 (:741-749)
                        n_qx = sel(!(pabs(dQx) < k.eps), fma2(splat(-dt), dQx, n_qx), n_qx);
                        n_qy = sel(!(pabs(dQy) < k.eps), fma2(splat(-dt), dQy, n_qy), n_qy);
                        const P2<R> h_new = n_eta - pzb;
                        const B2 wet_new = !(h_new < k.eps);
                        const P2<R> hs = sel(wet_new, h_new, splat(R(1)));
                        const P2<R> rh_new = prcp(hs);
                        if (k.friction) friction_pair(k, hs, rh_new, n_qx, n_qy, ld2(o_m, T::P_N), dt, wet_new);
                        n_eta = sel(wet_new, n_eta, pzb);
                        const P2<R> n_emax = sel((n_eta > c_emax) & (c_emax > R(-9990.0)), n_eta, c_emax);
                        c_eta = sel(upd, n_eta, c_eta); c_emax = sel(upd, n_emax, c_emax);
                        c_qx = sel(upd, n_qx, c_qx); c_qy = sel(upd, n_qy, c_qy);
                    }
                    if (x_store.a) {
                        store_row(y - 1, c_eta, c_emax, c_qx, c_qy);
                        if (a.reduce_mode != hp::kReduceNone) ws = wide_speed2(k, c_eta, c_emax, c_qx, c_qy, pzb, x_store, ws);
                    }
                }

                // ---- x-faces of row y, reduced at once to the three sums its corrector needs ------------------
                P2<R> Xeta = zero, Xqx = zero, Xqy = zero;
                I2 xStop{0, 0};
                if (j >= 2 && j < J) {
                    // the east-side estimates meet the west side of the next column: lane-1's second column for the first
                    // column (one lane up), the lane's own first column for the second
                    const P2<R> xe_eta = fma2(half, sxE, ce), xe_h = fma2(half, sxH, ch);
                    const P2<R> xe_r = sel(xe_h <= k.eps, zero, prcp(xe_h));
                    const P2<R> xe_u = fma2(half, sxQx, cqx) * xe_r, xe_v = fma2(half, sxQy, cqy) * xe_r;
                    const P2<R> etaL = pshfl_up_b(xe_eta), hfL = pshfl_up_b(xe_h), uL = pshfl_up_b(xe_u), vL = pshfl_up_b(xe_v);
                    const P2<R> xw_eta = fma2(mhalf, sxE, ce), hfR = fma2(mhalf, sxH, ch);
                    const P2<R> rR = sel(hfR <= k.eps, zero, prcp(hfR));
                    FaceOut2<R> fx;
                    face_solve_pair<R>(k, etaL, etaL - hfL, uL, vL, xw_eta, xw_eta - hfR, fma2(mhalf, sxQx, cqx) * rR, fma2(mhalf, sxQy, cqy) * rR,
                                       [&] { const P2<R> q = ld2_again(o_c, T::P_QX); return P2<R>{ldw_again(o_c, T::P_QX), q.a}; },
                                       [&] { return ld2_again(o_c, T::P_QX); }, fx);
                    // the east faces: the second column's western face for the first column, lane+1's first for the second
                    const P2<R> eM = pshfl_dn_a(fx.m), eN = pshfl_dn_a(fx.n), eT = pshfl_dn_a(fx.t), eZ = pshfl_dn_a(fx.zmax), eH = pshfl_dn_a(fx.hR);
                    const I2 eStop{fx.stopL.b, __shfl_down_sync(FULL, fx.stopL.a, 1)};
                    const P2<R> bE = pmin(eZ, xe_eta), bW = pmin(fx.zmax, xw_eta);
                    Xeta = eM - fx.m;
                    Xqx = fma2(hg * (bE - bW), eH + fx.hL, eN - fx.n);
                    Xqy = eT - fx.t;
                    xStop = I2{fx.stopR.a + eStop.a, fx.stopR.b + eStop.b};
                }

                // ---- hand over to the next row (everything carried is dead by now) -----------------------------
                Aeta = Xeta - fy.m; Aqx = Xqx - fy.t; Aqy = Xqy - fy.n;
                bS = pmin(fy.zmax, etaR); sH = fy.hL; cStop = I2{xStop.a + fy.stopR.a, xStop.b + fy.stopR.b};
                Le = fma2(half, syE, ce); Lh = fma2(half, syH, ch);
                const P2<R> rL = sel(Lh <= k.eps, zero, prcp(Lh));
                Lun = fma2(half, syQy, cqy) * rL; Lut = fma2(half, syQx, cqx) * rL;
                f_ew_prev = I2{(f_w.a >> 1) | ((f_e.a >> 1) << 2), (f_w.b >> 1) | ((f_e.b >> 1) << 2)};
                f_m2 = f_m1; f_m1 = f_c; f_c = f_p;
            }
            dr_m2 = dr_m1; dr_m1 = dr_c;

            // row j-1 is dead: refill its ring slot with row j-1+RR
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && j - 1 + T::RR <= J + 1) issue_row(j - 1 + T::RR);
        }
    }
    block_reduce_finalize<R>(ws, a, k);
}

template <class R> static int launch_mh_wide(const StepArgs& a_in, const TmaBlockMap& maps, int alt, int sm_count, cudaStream_t st) {
    using T = Wide<R, false>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(mh_step_wide<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        cudaFuncSetAttribute(mh_step_wide<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int grid = march_grid(a, T::USE, T::NW, sizeof(R) == 8 ? HP_WIDE_MH_CTAS64 : HP_WIDE_MH_CTAS32, sm_count);
    a.total_ctas = grid; a.march_runs = march_runs(a, T::USE, T::NW, grid);
    if (alt) mh_step_wide<R, true><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    else mh_step_wide<R, false><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

// =============================================================================================
// First-order Godunov, two columns per lane.  Same row structure as godunov_step_march (hp_march_kernels.cuh): the
// cells of the row below (level, bed, velocities, celerity) and their southern faces are carried as pairs; of a lane's
// two west faces only the first needs the neighbouring lane.  Cells whose stencil is dry stay unwritten (SURVEY.md
// Q2), exactly like godunov_step_tma.
// =============================================================================================
template <class R> struct GodPair { P2<R> eta, zb, u, v, c; };

#ifndef HP_WIDE_GOD_CTAS64
#define HP_WIDE_GOD_CTAS64 3
#endif
#ifndef HP_WIDE_GOD_CTAS32
#define HP_WIDE_GOD_CTAS32 4
#endif
template <class R, bool ALT>
__global__ void __launch_bounds__(hp::kMarchWarps * 32, sizeof(R) == 8 ? HP_WIDE_GOD_CTAS64 : HP_WIDE_GOD_CTAS32)
godunov_step_wide(const StepArgs a, const __grid_constant__ TmaBlockMap maps) {
    using T = Wide<R, ALT>;
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
    const P2<R> zero = splat(R(0));
    const MutView<R> d(a.dst);
    const bool stepping = dt > R(0);

    const int nrows = a.y1 - a.y0;
    const int nstrips = (g.cols + T::USE - 1) / T::USE;
    const int ngroups = (nstrips + T::NW - 1) / T::NW;
    const long long units = static_cast<long long>(ngroups) * nrows;
    const long long total_runs = static_cast<long long>(gridDim.x) * a.march_runs;
    int run = 0;
    long long u = units * blockIdx.x / total_runs, u1 = units * (blockIdx.x + 1) / total_runs;

    constexpr int SZ = int(sizeof(R));
    const int lc = (2 * lane + T::PADL) * SZ;
    const int lw = lane > 0 ? lc - SZ : lc;
    auto ld2 = [&](int row_off, int plane) -> P2<R> { return ld_pair<R>(ring + row_off + plane * T::PLANE + lc); };
    auto ldw = [&](int row_off, int plane) -> R { return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + lw); };
    // re-read of the raw discharge (only wet/dry fronts ask for it)
    auto ld2_again = [&](int row_off, int plane) -> P2<R> {
        const volatile R* p = reinterpret_cast<const volatile R*>(ring + row_off + plane * T::PLANE + lc);
        return P2<R>{p[0], p[1]};
    };
    auto ldw_again = [&](int row_off, int plane) -> R { return *reinterpret_cast<const volatile R*>(ring + row_off + plane * T::PLANE + lw); };
    const bool lane_owns = lane >= 1 && lane <= 30;

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

        const int X0 = strip * T::USE - 2;               // first column of lane 0
        const int xa = X0 + 2 * lane;                     // the lane's columns: xa, xa + 1
        const int rs = ya - 1;                            // first raw row of this run
        const int NR = yb - ya + 2;                       // raw rows 0 .. NR-1; rows 1 .. NR-2 are updated
        const B2 x_interior{xa >= 1 && xa <= g.cols - 2, xa + 1 >= 1 && xa + 1 <= g.cols - 2};
        const B2 x_store{lane_owns && xa < g.cols, lane_owns && xa + 1 < g.cols};
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
        auto derive = [&](int row_off, GodPair<R>& o, B2& dry) {          // phase B of the tile kernel
            o.eta = ld2(row_off, T::P_ETA); o.zb = ld2(row_off, T::P_ZB);
            const P2<R> qx = ld2(row_off, T::P_QX), qy = ld2(row_off, T::P_QY);
            const P2<R> h = o.eta - o.zb;
            dry = h < k.eps;
            const P2<R> rh = sel(dry, zero, prcp(sel(dry, splat(R(1)), h)));
            o.u = qx * rh; o.v = qy * rh;
            o.c = pcelerity(k.g, ppos(h));
        };
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < T::RR; ++j) if (j < NR) issue_row(j);
        }
        wait_row(0);

        GodPair<R> P;
        B2 dry_p;
        derive(0, P, dry_p);
        P2<R> sM = zero, sN = zero, sT = zero, sZ = zero, sH = zero;       // southern faces of row j-1
        I2 sStop{0, 0};
        B2 dry_s{true, true};                                               // dryness of the cells below row j-1

        for (int j = 1; j < NR; ++j) {
            const int y = rs + j;
            const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT, o_c = (j & (T::RR - 1)) * T::SLOT;
            wait_row(j);
            GodPair<R> C;
            B2 dry_c;
            derive(o_c, C, dry_c);
            if (a.reduce_mode == hp::kReduceDst && any(x_store) && j + 1 < NR) {      // row y is updated in the next trip
                const int gy = y + g.gy0;
                if (!(x_interior.a && x_interior.b) || gy < 1 || gy > g.grows - 2 || any(dry_c & dry_p))
                    prefetch_dst(d, static_cast<size_t>(y) * g.pitch + xa);          // the pair shares its 32-byte sectors
            }

            // faces between rows y-1 (left) and y (right); normal = y
            FaceOut2<R> fy;
            face_solve_pair_c<R, true>(k, P.eta, P.zb, P.v, P.u, P.c, C.eta, C.zb, C.v, C.u, C.c,
                                       [&] { return ld2_again(o_m, T::P_QY); }, [&] { return ld2_again(o_c, T::P_QY); }, fy);

            if (j >= 2) {
                const int yc = y - 1, gyc = yc + g.gy0;
                P2<R> eta = P.eta, emax = ld2(o_m, T::P_EMAX), qx = ld2(o_m, T::P_QX), qy = ld2(o_m, T::P_QY);
                const P2<R> zb = P.zb;
                const B2 disabled = (emax <= R(-9999.0)) | (eta == splat(R(-9999.0)));
                // rows y-2, y-1, y dry in every column: the stencil of every cell of row y-1 is dry, the reference returns
                // without writing (CLSchemeGodunov.clc:248-255) -- no face of that row is needed
                const bool skip = stepping && __all_sync(FULL, dry_s.a && dry_s.b && dry_p.a && dry_p.b && dry_c.a && dry_c.b &&
                                                                   !(disabled.a || disabled.b));
                if (a.reduce_mode == hp::kReduceSrc && any(x_store)) {
                    const B2 ok = x_store & ((eta - zb) > k.eps10) & (emax > R(-9999.0));
                    const P2<R> sp = sel(ok, k.simplified_speed ? P.c : pmax(pabs(P.u), pabs(P.v)) + P.c, zero);
                    ws = fm_max(fm_max(sp.a, sp.b), ws);
                }
                B2 wrote{false, false};
                if (!skip) {
                    // west faces of row y-1: lane-1's second column against the own first, the own first against the own second
                    const P2<R> w_eta{ldw(o_m, T::P_ETA), P.eta.a}, w_zb{ldw(o_m, T::P_ZB), P.zb.a};
                    const P2<R> w_u = pshfl_up_b(P.u), w_v = pshfl_up_b(P.v), w_c = pshfl_up_b(P.c);
                    FaceOut2<R> fx;
                    face_solve_pair_c<R, true>(k, w_eta, w_zb, w_u, w_v, w_c, P.eta, P.zb, P.u, P.v, P.c,
                                               [&] { const P2<R> q = ld2_again(o_m, T::P_QX); return P2<R>{ldw_again(o_m, T::P_QX), q.a}; },
                                               [&] { return ld2_again(o_m, T::P_QX); }, fx);
                    // east faces: the second column's western face for the first column, lane+1's first for the second
                    const P2<R> eM = pshfl_dn_a(fx.m), eN = pshfl_dn_a(fx.n), eT = pshfl_dn_a(fx.t), eZ = pshfl_dn_a(fx.zmax), eH = pshfl_dn_a(fx.hR);
                    const I2 eStop{fx.stopL.b, __shfl_down_sync(FULL, fx.stopL.a, 1)};
                    const unsigned drym_a = __ballot_sync(FULL, dry_p.a), drym_b = __ballot_sync(FULL, dry_p.b);
                    const B2 dry_w{((drym_b >> ((lane + 31) & 31)) & 1u) != 0, dry_p.a};
                    const B2 dry_e{dry_p.b, ((drym_a >> ((lane + 1) & 31)) & 1u) != 0};
                    const bool rows_ok = gyc >= 1 && gyc <= g.grows - 2;                    // frozen outer ring
                    const B2 all_dry = dry_p & dry_c & dry_s & dry_e & dry_w;               // CLSchemeGodunov.clc:248-255
                    const B2 upd = (x_interior & !disabled & !all_dry) & (rows_ok && stepping);
                    // dt <= 0 copies the state through (:201-206; not with gts_cacheEnabled's rule, :477-478), and so does a
                    // disabled cell
                    wrote = stepping ? (x_interior & (disabled | !all_dry)) & rows_ok : x_interior & (rows_ok && !k.dt0_keep);
                    if (__any_sync(FULL, any(upd))) {
                        const P2<R> bN = pmin(fy.zmax, eta), bS = pmin(sZ, eta), bE = pmin(eZ, eta), bW = pmin(fx.zmax, eta);
                        const B2 stop{fy.stopL.a + sStop.a + fx.stopR.a + eStop.a > 0, fy.stopL.b + sStop.b + fx.stopR.b + eStop.b > 0};
                        const P2<R> dEta = ((eM - fx.m) + (fy.m - sM)) * splat(inv_delta);
                        const P2<R> dQx = fma2(hg * (bE - bW), eH + fx.hL, (eN - fx.n) + (fy.t - sT)) * splat(inv_delta);
                        const P2<R> dQy = fma2(hg * (bN - bS), fy.hR + sH, (eT - fx.t) + (fy.n - sN)) * splat(inv_delta);
                        P2<R> n_qx = sel(stop, zero, qx), n_qy = sel(stop, zero, qy);
                        P2<R> n_eta = sel(!(pabs(dEta) < k.eps), fma2(splat(-dt), dEta, eta), eta);   // |D| < eps => 0 (:340-348)
                        n_qx = sel(!(pabs(dQx) < k.eps), fma2(splat(-dt), dQx, n_qx), n_qx);
                        n_qy = sel(!(pabs(dQy) < k.eps), fma2(splat(-dt), dQy, n_qy), n_qy);
                        const P2<R> h_new = n_eta - zb;
                        const B2 wet_new = !(h_new < k.eps);
                        const P2<R> hs = sel(wet_new, h_new, splat(R(1)));
                        if (k.friction) friction_pair(k, hs, prcp(hs), n_qx, n_qy, ld2(o_m, T::P_N), dt, wet_new);
                        const P2<R> n_emax = sel((n_eta > emax) & (emax > R(-9990.0)), n_eta, emax);
                        n_eta = sel(wet_new, n_eta, zb);
                        eta = sel(upd, n_eta, eta); emax = sel(upd, n_emax, emax);
                        qx = sel(upd, n_qx, qx); qy = sel(upd, n_qy, qy);
                    }
                }
                if (any(x_store)) {
                    const size_t id = static_cast<size_t>(yc) * g.pitch + xa;
                    if (wrote.a && wrote.b && x_store.b) {
                        st_pair(d.eta + id, eta); st_pair(d.emax + id, emax); st_pair(d.qx + id, qx); st_pair(d.qy + id, qy);
                    } else {
                        if (wrote.a && x_store.a) { d.eta[id] = eta.a; d.emax[id] = emax.a; d.qx[id] = qx.a; d.qy[id] = qy.a; }
                        if (wrote.b && x_store.b) { d.eta[id + 1] = eta.b; d.emax[id + 1] = emax.b; d.qx[id + 1] = qx.b; d.qy[id + 1] = qy.b; }
                    }
                    if (a.reduce_mode == hp::kReduceDst) {
                        // cells left unwritten enter the reduction with what the destination holds
                        if (!wrote.a && x_store.a) { eta.a = d.eta[id]; emax.a = d.emax[id]; qx.a = d.qx[id]; qy.a = d.qy[id]; }
                        if (!wrote.b && x_store.b) { eta.b = d.eta[id + 1]; emax.b = d.emax[id + 1]; qx.b = d.qx[id + 1]; qy.b = d.qy[id + 1]; }
                        ws = wide_speed2(k, eta, emax, qx, qy, zb, x_store, ws);
                    }
                }
            }
            sM = fy.m; sN = fy.n; sT = fy.t; sZ = fy.zmax; sH = fy.hL; sStop = fy.stopR;
            dry_s = dry_p; dry_p = dry_c;
            P = C;

            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && j - 1 + T::RR < NR) issue_row(j - 1 + T::RR);
        }
    }
    block_reduce_finalize<R>(ws, a, k);
}

template <class R> static int launch_godunov_wide(const StepArgs& a_in, const TmaBlockMap& maps, int alt, int sm_count, cudaStream_t st) {
    using T = Wide<R, false>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(godunov_step_wide<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        cudaFuncSetAttribute(godunov_step_wide<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int grid = march_grid(a, T::USE, T::NW, sizeof(R) == 8 ? HP_WIDE_GOD_CTAS64 : HP_WIDE_GOD_CTAS32, sm_count);
    a.total_ctas = grid; a.march_runs = march_runs(a, T::USE, T::NW, grid);
    if (alt) godunov_step_wide<R, true><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    else godunov_step_wide<R, false><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

}  // namespace HP_NS
