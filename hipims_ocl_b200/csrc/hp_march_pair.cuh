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

// =============================================================================================
// Partial inertial scheme (see inertial_step_march for the face / owner split of the Manning coefficient).
// =============================================================================================
template <class R> struct InFace2 { P2<R> num, A, qmax; B2 wet; };

template <class R>
__device__ __forceinline__ InFace2<R> inertial_face2(const Params<R>& k, R gdt, P2<R> prev, P2<R> etaUp, P2<R> zUp, P2<R> etaDown,
                                                    P2<R> zDown, R inv_delta) {
    InFace2<R> f;
    const P2<R> h = pmax(etaDown, etaUp) - pmax(zUp, zDown);
    f.wet = !(h < k.eps);
    f.num = splat(R(0)); f.A = splat(R(0)); f.qmax = splat(R(0));
    if (any(f.wet)) {
        // a dry component rides along on a depth of one and is zeroed by inertial_q2
        const P2<R> hs = sel(f.wet, h, splat(R(1)));
        const P2<R> rh = prcp(hs);
        f.num = fma2(-(gdt * hs), (etaDown - etaUp) * splat(inv_delta), prev);
        f.A = gdt * pabs(prev) * rh * rh * prcbrt(hs);
        f.qmax = (R(0.8) * hs) * psqrt_pos(k.g * hs);
    }
    return f;
}
template <class R> __device__ __forceinline__ P2<R> inertial_q2(const InFace2<R>& f, P2<R> n) {
    if (!any(f.wet)) return splat(R(0));
    const P2<R> q = f.num * prcp(fma2(f.A, n * n, splat(R(1))));
    const P2<R> c = pmax(pmin(q, f.qmax), -f.qmax);
    return sel(f.wet, c, splat(R(0)));
}

#ifndef HP_WIDE_INE_CTAS64
#define HP_WIDE_INE_CTAS64 4
#endif
#ifndef HP_WIDE_INE_CTAS32
#define HP_WIDE_INE_CTAS32 5
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
            InFace2<R> fy{splat(R(0)), splat(R(0)), splat(R(0)), B2{false, false}};
            if (stepping) fy = inertial_face2<R>(k, gdt, c_qy, c_eta, c_zb, p_eta, p_zb, inv_delta);
            const P2<R> qN = inertial_q2(fy, p_n);                         // as the cells below see them (CLSchemeInertial.clc:107)
            P2<R> qS_next = qN;                                            // as the cells above see them (:109)
            if (c_n.a != p_n.a || c_n.b != p_n.b) qS_next = sel(c_n == p_n, qN, inertial_q2(fy, c_n));

            if (j >= 2) {
                const int yc = y - 1, gyc = yc + g.gy0;
                const P2<R> p_qx = ld2(o_m, T::P_QX), p_qy = ld2(o_m, T::P_QY);
                // west faces of row y-1: the own cell is "up", its western neighbour "down" -- the last column of lane-1
                // (from the box) for the first column, the lane's own first column for the second; the discharge is the own qx
                const P2<R> w_eta{ldw(o_m, T::P_ETA), p_eta.a}, w_zb{ldw(o_m, T::P_ZB), p_zb.a}, w_n{ldw(o_m, T::P_N), p_n.a};
                InFace2<R> fx{splat(R(0)), splat(R(0)), splat(R(0)), B2{false, false}};
                if (stepping) fx = inertial_face2<R>(k, gdt, p_qx, p_eta, p_zb, w_eta, w_zb, inv_delta);
                const P2<R> qW = inertial_q2(fx, p_n);                     // :110
                P2<R> qE_for_west = qW;                                    // the same faces as the western neighbours see them (:108)
                if (w_n.a != p_n.a || w_n.b != p_n.b) qE_for_west = sel(w_n == p_n, qW, inertial_q2(fx, w_n));
                // east faces: the second column's western face for the first column, lane+1's first for the second
                const P2<R> qE{qE_for_west.b, shfl_dn1(qE_for_west.a)};
                const unsigned drym_a = __ballot_sync(FULL, dry_p.a), drym_b = __ballot_sync(FULL, dry_p.b);
                const B2 dry_w{((drym_b >> ((lane + 31) & 31)) & 1u) != 0, dry_p.a};
                const B2 dry_e{dry_p.b, ((drym_a >> ((lane + 1) & 31)) & 1u) != 0};

                P2<R> eta = p_eta, emax = ld2(o_m, T::P_EMAX), qx = p_qx, qy = p_qy;
                if (a.reduce_mode == hp::kReduceSrc) {
                    if (x_store.a) ws = fm_max(wide_speed(k, eta.a, emax.a, qx.a, qy.a, p_zb.a), ws);
                    if (x_store.b) ws = fm_max(wide_speed(k, eta.b, emax.b, qx.b, qy.b, p_zb.b), ws);
                }
                B2 wrote{false, false};
                if (gyc >= 1 && gyc <= g.grows - 2 && stepping) {                       // dt <= 0 returns first (:60-61)
                    const B2 disabled = (emax <= R(-9999.0)) | (eta == splat(R(-9999.0)));
                    const B2 all_dry = dry_p & dry_c & dry_s & dry_e & dry_w;           // :92-99
                    const B2 upd = x_interior & !disabled & !all_dry;
                    wrote = x_interior & (disabled | !all_dry);                         // a disabled cell is copied through
                    if (any(upd)) {
                        const P2<R> n_eta = fma2(splat(dt), (((qE - qW) + qN) - qS) * splat(inv_delta), eta);   // :145-152
                        P2<R> n_emax = sel(n_eta > emax, n_eta, emax);
                        const P2<R> f_eta = sel((n_eta - p_zb) < k.eps, p_zb, n_eta);
                        eta = sel(upd, f_eta, eta); emax = sel(upd, n_emax, emax);
                        qx = sel(upd, qW, qx); qy = sel(upd, qS, qy);                   // :141-142
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
                        if (x_store.a) ws = fm_max(wide_speed(k, eta.a, emax.a, qx.a, qy.a, p_zb.a), ws);
                        if (x_store.b) ws = fm_max(wide_speed(k, eta.b, emax.b, qx.b, qy.b, p_zb.b), ws);
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

}  // namespace HP_NS
