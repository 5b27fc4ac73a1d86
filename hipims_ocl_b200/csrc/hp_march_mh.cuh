// hp_march_mh.cuh -- marching MUSCL-Hancock kernel, second generation ("instruction diet").
//
// Same skeleton as the other marching kernels (hp_march_kernels.cuh: one warp per 32-column strip, rows through a
// per-warp TMA ring, x-direction by shuffles, y-direction in registers), restructured around what the profile of the
// first generation showed (profiles/r01_march_mh_f64_4096.txt: 1084 warp instructions per 32 cell-updates, only 406
// of them fp64 arithmetic; every instruction costs an issue slot and an fp64 one 2.2 of them):
//
//   * NOTHING IS ROTATED.  The first generation carried the whole predictor of the previous row (evolved state +
//     eight slopes) plus the southern face to the next row and paid ~40 register moves per row for it.  Here a row's
//     x-faces are solved in the same trip as its predictor and reduced at once to three partial sums; what crosses to
//     the next row is the northern face estimate (4 values), those sums minus the southern flux (3), the southern bed
//     and depth (2) and a stop count -- and every one of them is dead before its successor is computed, so the
//     compiler updates them in place;
//   * WET FAST PATH in the face solver: when both reconstructed depths exceed the dry threshold (one combined test)
//     there are no stop flags, no dry-side selects, no clamps and the square roots need no zero guard;
//   * the |D| < eps => 0 chop of the reference (CLSchemeMUSCLHancock.clc:363-371, 741-749) is a predicated update
//     instead of two selects per component;
//   * the predictor's `face depth < eps => zero velocity` tests are gone: in the second-order branch the limited
//     face depth is at least half the cell depth (>= 5e-6), so they can never fire;
//   * the reciprocal of the new depth is shared by friction and the CFL wave speed.
// Results are those of the first generation up to the order of a few additions.
#pragma once

#include "hp_march_kernels.cuh"

namespace HP_NS {

#ifndef HP_MH2_CTAS64
#define HP_MH2_CTAS64 4
#endif
#ifndef HP_MH2_CTAS32
#define HP_MH2_CTAS32 6
#endif

template <class R, bool ALT>
__global__ void __launch_bounds__(hp::kMarchWarps * 32, sizeof(R) == 8 ? HP_MH2_CTAS64 : HP_MH2_CTAS32)
mh_step_march2(const StepArgs a, const __grid_constant__ TmaBlockMap maps) {
    using T = March<R, 1, ALT, HP_MARCH_MH_RR>;
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
    const R hg = R(0.5) * k.g, half = R(0.5);
    const R hdt = half * dt;
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

    // per-lane byte offsets of the own column and its x-neighbours inside a plane row of the box
    const int lc = (lane + T::PADL) * int(sizeof(R));
    const int lw = lc - int(sizeof(R)), le = lc + int(sizeof(R));
    static_assert(T::PADL >= 1 && T::BW >= 33 + T::PADL, "the box must hold one raw column beyond either edge lane");
    auto ld = [&](int row_off, int plane, int col_off) -> R {
        return *reinterpret_cast<const R*>(ring + row_off + plane * T::PLANE + col_off);
    };
    auto flags_of = [&](R em) -> int { return (em <= R(-9998.0) ? 1 : 0) | (em < k.eps ? 2 : 0); };
    // bit 0 of flags_of alone, for the columns beyond the edge lanes (their bit 1 only enters the corrector of a halo
    // lane, which stores nothing).  -9998.0 has a zero low word: the test is one unsigned compare of the high word
    auto disabled_at = [&](int row_off, int col_off) -> int {
        if constexpr (sizeof(R) == 8)
            return *reinterpret_cast<const unsigned*>(ring + row_off + T::P_EMAX * T::PLANE + col_off + 4) >= 0xC0C38700u ? 1 : 0;
        else
            return *reinterpret_cast<const R*>(ring + row_off + T::P_EMAX * T::PLANE + col_off) <= R(-9998.0) ? 1 : 0;
    };
    // re-read of a raw value the predictor has overwritten in registers (only wet/dry fronts ask for it)
    auto ld_again = [&](int row_off, int plane, int col_off) -> R {
        return *reinterpret_cast<const volatile R*>(ring + row_off + plane * T::PLANE + col_off);
    };
    const bool lane_owns = lane >= 1 && lane < 1 + T::USE;
    // wave speed of a stored cell for the CFL reduction (CLDynamicTimestep.clc:81-110); rh = 1/h if the caller has it
    auto speed_of = [&](R h, R qx, R qy, R rh) -> R {
        const R cc = fm_sqrt_pos(k.g * h);
        return k.simplified_speed ? cc : fm_max(hp_abs(qx * rh), hp_abs(qy * rh)) + cc;
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

        const int X0 = strip * T::USE - 1;               // column of lane 0
        const int x = X0 + lane;
        const int rs = ya - 2;                            // first raw row of this run
        const int J = yb - ya + 2;                        // raw rows 0 .. J+1, predictor rows 1 .. J, updated rows 2 .. J-1
        const bool x_store = lane_owns && x < g.cols;
        const bool x_valid = x >= 1 && x <= g.cols - 2;           // predictor runs on 1 .. cols-2 (CLSchemeMUSCLHancock.clc:54-58)
        const bool x_interior = x >= 2 && x <= g.cols - 3;        // the ring of two is frozen (:569-577)
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
        // the previous run's rows are all consumed; order its generic reads before the async writes
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
                if (j >= 3 && x_store) {
                    const Cell<R> c{ld(o_m, T::P_ETA, lc), ld(o_m, T::P_EMAX, lc), ld(o_m, T::P_QX, lc), ld(o_m, T::P_QY, lc)};
                    d.store(static_cast<size_t>(rs + j - 1) * g.pitch + x, c);
                    if (a.reduce_mode != hp::kReduceNone) {
                        const R h = c.eta - ld(o_m, T::P_ZB, lc);
                        if (h > k.eps10 && c.emax > R(-9999.0)) ws = fm_max(speed_of(h, c.qx, c.qy, k.simplified_speed ? R(0) : fm_rcp(h)), ws);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && j - 1 + T::RR <= J + 1) issue_row(j - 1 + T::RR);
            }
            continue;
        }

        // ---- state carried from row to row; each item is dead before its successor is computed ---------------
        int f_m2 = 0, f_m1 = flags_of(ld(0 * T::SLOT, T::P_EMAX, lc)), f_c = flags_of(ld(1 * T::SLOT, T::P_EMAX, lc));
        int f_ew_prev = 0;                                   // bit1 flags of the x-neighbours of the previous row: W | E<<2
        R Le = R(0), Lh = R(0), Lun = R(0), Lut = R(0);      // northern face estimate of the previous row: eta, depth, v, u
        R Aeta = R(0), Aqx = R(0), Aqy = R(0);               // x-face sums of the previous row minus its southern flux
        R bS = R(0), sH = R(0);                              // its southern face: reconstructed bed (owner side), neighbour depth
        int cStop = 0;                                       // its stop count so far (west, east, south)
        // Rows in which every lane is EXACTLY dry and at rest (eta == zb, q == 0, eta_max not below eta) -- most of a
        // flood model's domain.  Such a cell falls back to first order with zero slopes, a face between two of them
        // carries no flux at all, and a cell whose whole stencil is like that cannot change: the row is copied through
        // (identical to what the full arithmetic produces, at a fraction of its instructions).
        bool dr_m2 = false, dr_m1 = false, dr_c = false;     // rows j-2, j-1, j (warp-uniform)

        bool in_skip = false;                                // the previous trip copied its row through
        for (int j = 1; j <= J; ++j) {
            const int y = rs + j, gy = y + g.gy0;
            const int o_m = ((j - 1) & (T::RR - 1)) * T::SLOT, o_c = (j & (T::RR - 1)) * T::SLOT,
                      o_p = ((j + 1) & (T::RR - 1)) * T::SLOT;
            wait_row(j + 1);
            const R eta = ld(o_c, T::P_ETA, lc), qx = ld(o_c, T::P_QX, lc), qy = ld(o_c, T::P_QY, lc), zb = ld(o_c, T::P_ZB, lc);
            // (tested on every fourth row, and on every row while the rows below are dry: wet regions pay almost nothing)
            dr_c = false;
            R emax_c = R(0);
            if (dr_m1 || (j & 3) == 0) {
                emax_c = ld(o_c, T::P_EMAX, lc);
                dr_c = __all_sync(FULL, eta == zb && qx == R(0) && qy == R(0) && !(eta > emax_c));
            }

            if (dr_m2 && dr_m1 && dr_c) {
                // rows y-2, y-1, y exactly dry in every lane (so j >= 3): the cell (x, y-1) cannot change.  A trip in this
                // mode does the least possible: copy the row, keep the one flag bit a dry row can have (eta_max < eps; a
                // disabled cell never has eta == zb).  Everything else is rebuilt when the mode is left.
                if (x_store)
                    d.store(static_cast<size_t>(y - 1) * g.pitch + x,
                            Cell<R>{ld(o_m, T::P_ETA, lc), ld(o_m, T::P_EMAX, lc), ld(o_m, T::P_QX, lc), ld(o_m, T::P_QY, lc)});
                f_m2 = f_m1; f_m1 = emax_c < k.eps ? 2 : 0;          // flags of rows y-1 and y for the next trip
                in_skip = true;
            } else {
                if (in_skip) {
                    // leaving the copy-through mode: row y-1 is exactly dry -- its northern face estimate is its level with no
                    // depth and no velocity, its sums are empty; f_m1 / f_m2 were kept, the rest of the flags is re-derived
                    in_skip = false;
                    Le = ld(o_m, T::P_ETA, lc); Lh = R(0); Lun = R(0); Lut = R(0);
                    Aeta = R(0); Aqx = R(0); Aqy = R(0); bS = Le; sH = R(0); cStop = 0;
                    f_c = flags_of(ld(o_c, T::P_EMAX, lc));
                    f_ew_prev = (__shfl_up_sync(FULL, f_m1, 1) >> 1) | ((__shfl_down_sync(FULL, f_m1, 1) >> 1) << 2);
                }
                const int f_p = flags_of(ld(o_p, T::P_EMAX, lc));
                int f_w = __shfl_up_sync(FULL, f_c, 1), f_e = __shfl_down_sync(FULL, f_c, 1);
                if (lane == 0) f_w = disabled_at(o_c, lw);                      // the columns beyond the edge lanes have no lane
                if (lane == 31) f_e = disabled_at(o_c, le);
                // ---- predictor of row y (CLSchemeMUSCLHancock.clc:301-382) ---------------------------
                R ce = eta, cqx = qx, cqy = qy;
                R sxE = R(0), sxH = R(0), sxQx = R(0), sxQy = R(0), syE = R(0), syH = R(0), syQx = R(0), syQy = R(0);
                {
                    const bool valid = x_valid && gy >= 1 && gy <= g.grows - 2 && y >= 1 && y <= g.rows - 2;
                    const R h = eta - zb;
                    if (valid && !(h < R(1E-5)) && !((f_p | f_e | f_m1 | f_w) & 1)) {
                        const R etaE = ld(o_c, T::P_ETA, le), etaW = ld(o_c, T::P_ETA, lw), etaN = ld(o_p, T::P_ETA, lc), etaS = ld(o_m, T::P_ETA, lc);
                        const R hE = etaE - ld(o_c, T::P_ZB, le), hW = etaW - ld(o_c, T::P_ZB, lw);
                        const R hN = etaN - ld(o_p, T::P_ZB, lc), hS = etaS - ld(o_m, T::P_ZB, lc);
                        // a dry neighbour drops the slopes of its direction (:301-320): no branch, the switch rides in the
                        // limiter's sign test (where this block runs at all, both directions are almost always kept)
                        const int xoff = (fm_lt_opaque(hW, k.eps) | fm_lt_opaque(hE, k.eps)) ? int(0x80000000u) : 0;
                        const int yoff = (fm_lt_opaque(hS, k.eps) | fm_lt_opaque(hN, k.eps)) ? int(0x80000000u) : 0;
                        sxE = minmod_sw(eta - etaW, etaE - eta, xoff); sxH = minmod_sw(h - hW, hE - h, xoff);
                        sxQx = minmod_sw(qx - ld(o_c, T::P_QX, lw), ld(o_c, T::P_QX, le) - qx, xoff);
                        sxQy = minmod_sw(qy - ld(o_c, T::P_QY, lw), ld(o_c, T::P_QY, le) - qy, xoff);
                        syE = minmod_sw(eta - etaS, etaN - eta, yoff); syH = minmod_sw(h - hS, hN - h, yoff);
                        syQx = minmod_sw(qx - ld(o_m, T::P_QX, lc), ld(o_p, T::P_QX, lc) - qx, yoff);
                        syQy = minmod_sw(qy - ld(o_m, T::P_QY, lc), ld(o_p, T::P_QY, lc) - qy, yoff);
                        // Face depths h +- s/2 with |s| <= |h - h_neighbour| and both >= 0: never below h/2 >= 5e-6, so the
                        // reference's `face depth < VERY_SMALL => zero velocity` (:333-346) cannot fire here.
                        // v +- s/2 as two multiply-adds: s/2 is exact, so each is the correctly rounded v +- s/2 of the
                        // product-then-sum form, in two fp64 operations per pair instead of three
                        const R hEf = hp_fma(half, sxH, h), hWf = hp_fma(-half, sxH, h), hNf = hp_fma(half, syH, h), hSf = hp_fma(-half, syH, h);
                        const R qxE = hp_fma(half, sxQx, qx), qxW = hp_fma(-half, sxQx, qx), qyE = hp_fma(half, sxQy, qy), qyW = hp_fma(-half, sxQy, qy);
                        const R qxN = hp_fma(half, syQx, qx), qxS = hp_fma(-half, syQx, qx), qyN = hp_fma(half, syQy, qy), qyS = hp_fma(-half, syQy, qy);
                        const R uE = qxE * fm_rcp(hEf), uW = qxW * fm_rcp(hWf);
                        const R vN = qyN * fm_rcp(hNf), vS = qyS * fm_rcp(hSf);
                        const R dEta = ((qxE - qxW) + (qyN - qyS)) * inv_delta;
                        const R dQx = (uE * qxE - uW * qxW + vN * qxN - vS * qxS + hg * sxE * (hEf + hWf)) * inv_delta;
                        const R dQy = (uE * qyE - uW * qyW + vN * qyN - vS * qyS + hg * syE * (hNf + hSf)) * inv_delta;
                        if (!(hp_abs(dEta) < k.eps)) ce = eta - hdt * dEta;          // |D| < eps => 0 (:363-371)
                        if (!(hp_abs(dQx) < k.eps)) cqx = qx - hdt * dQx;
                        if (!(hp_abs(dQy) < k.eps)) cqy = qy - hdt * dQy;
                    }
                }
                const R ch = ce - zb;

                FaceOut<R> fy;
                fy.m = R(0); fy.n = R(0); fy.t = R(0); fy.zmax = R(0); fy.hL = R(0); fy.hR = R(0); fy.stopL = 0; fy.stopR = 0;
                const R etaR = hp_fma(-half, syE, ce);           // southern face estimate of row y
                if (j >= 2) {
                    // ---- face between rows y-1 (left, carried) and y (right); normal = y ----------------------
                    const R hfR = hp_fma(-half, syH, ch);
                    const R rR = hfR <= k.eps ? R(0) : fm_rcp(hfR);                              // :1140-1150
                    face_solve2<R, false>(k, Le, Le - Lh, Lun, Lut, R(0), etaR, etaR - hfR, hp_fma(-half, syQy, cqy) * rR,
                                          hp_fma(-half, syQx, cqx) * rR, R(0), [&] { return ld_again(o_m, T::P_QY, lc); },
                                          [&] { return ld_again(o_c, T::P_QY, lc); }, fy);
                    if (j >= 3) {
                        // ---- corrector of row y-1 (CLSchemeMUSCLHancock.clc:596-800) -----------------
                        const int gyc = gy - 1;
                        Cell<R> c{ld(o_m, T::P_ETA, lc), ld(o_m, T::P_EMAX, lc), ld(o_m, T::P_QX, lc), ld(o_m, T::P_QY, lc)};
                        const R pzb = ld(o_m, T::P_ZB, lc);
                        R rh_new = R(0);
                        bool have_rh = false;
                        if (x_interior && gyc >= 2 && gyc <= g.grows - 3 && !(c.emax <= R(-9999.0) || c.eta == R(-9999.0))) {
                            int dry = (c.eta - pzb < k.eps) ? 1 : 0;
                            dry += (f_c >> 1) + (f_m2 >> 1) + (f_ew_prev & 1) + (f_ew_prev >> 2);
                            if (dry < 5) {
                                const R bN = fm_min(fy.zmax, Le);
                                const R dEta = (Aeta + fy.m) * inv_delta;
                                const R dQx = (Aqx + fy.t) * inv_delta;
                                const R dQy = (Aqy + fy.n + hg * (bN - bS) * (fy.hR + sH)) * inv_delta;
                                if (cStop + fy.stopL > 0) { c.qx = R(0); c.qy = R(0); }
                                if (!(hp_abs(dEta) < k.eps)) c.eta = c.eta - dt * dEta;   // |D| < eps => 0 (:741-749)
                                if (!(hp_abs(dQx) < k.eps)) c.qx = c.qx - dt * dQx;
                                if (!(hp_abs(dQy) < k.eps)) c.qy = c.qy - dt * dQy;
                                const R h_new = c.eta - pzb;
                                if (!(h_new < k.eps)) {
                                    rh_new = fm_rcp(h_new); have_rh = true;
                                    if (k.friction) friction_fast(k, h_new, rh_new, c.qx, c.qy, ld(o_m, T::P_N, lc), dt);
                                } else {
                                    c.eta = pzb;
                                }
                                if (c.eta > c.emax && c.emax > R(-9990.0)) c.emax = c.eta;
                            }
                        }
                        if (x_store) {
                            d.store(static_cast<size_t>(y - 1) * g.pitch + x, c);
                            if (a.reduce_mode != hp::kReduceNone) {
                                const R h = c.eta - pzb;
                                if (h > k.eps10 && c.emax > R(-9999.0)) {
                                    const R rh = k.simplified_speed ? R(0) : (have_rh ? rh_new : fm_rcp(h));
                                    ws = fm_max(speed_of(h, c.qx, c.qy, rh), ws);
                                }
                            }
                        }
                    }
                }

                // ---- x-faces of row y, reduced at once to the three sums its corrector needs ------------------
                R Xeta = R(0), Xqx = R(0), Xqy = R(0);
                int xStop = 0;
                if (j >= 2 && j < J) {
                    // the east-side estimate goes one lane up and meets that lane's west side
                    const R xe_eta = hp_fma(half, sxE, ce), xe_h = hp_fma(half, sxH, ch);
                    const R xe_r = xe_h <= k.eps ? R(0) : fm_rcp(xe_h);
                    const R xe_u = hp_fma(half, sxQx, cqx) * xe_r, xe_v = hp_fma(half, sxQy, cqy) * xe_r;
                    const R etaL = shfl_up1(xe_eta), hfL = shfl_up1(xe_h), uL = shfl_up1(xe_u), vL = shfl_up1(xe_v);
                    const R xw_eta = hp_fma(-half, sxE, ce), hfR = hp_fma(-half, sxH, ch);
                    const R rR = hfR <= k.eps ? R(0) : fm_rcp(hfR);
                    FaceOut<R> fx;
                    face_solve2<R, false>(k, etaL, etaL - hfL, uL, vL, R(0), xw_eta, xw_eta - hfR, hp_fma(-half, sxQx, cqx) * rR,
                                          hp_fma(-half, sxQy, cqy) * rR, R(0), [&] { return ld_again(o_c, T::P_QX, lw); },
                                          [&] { return ld_again(o_c, T::P_QX, lc); }, fx);
                    // the east face comes back from lane+1
                    const R eM = shfl_dn1(fx.m), eN = shfl_dn1(fx.n), eT = shfl_dn1(fx.t), eZ = shfl_dn1(fx.zmax), eH = shfl_dn1(fx.hR);
                    const int eStop = __shfl_down_sync(FULL, fx.stopL, 1);
                    const R bE = fm_min(eZ, xe_eta), bW = fm_min(fx.zmax, xw_eta);
                    Xeta = eM - fx.m;
                    Xqx = (eN - fx.n) + hg * (bE - bW) * (eH + fx.hL);
                    Xqy = eT - fx.t;
                    xStop = fx.stopR + eStop;
                }

                // ---- hand over to the next row (everything carried is dead by now) -----------------------------
                Aeta = Xeta - fy.m; Aqx = Xqx - fy.t; Aqy = Xqy - fy.n;
                bS = fm_min(fy.zmax, etaR); sH = fy.hL; cStop = xStop + fy.stopR;
                Le = hp_fma(half, syE, ce); Lh = hp_fma(half, syH, ch);
                const R rL = Lh <= k.eps ? R(0) : fm_rcp(Lh);
                Lun = hp_fma(half, syQy, cqy) * rL; Lut = hp_fma(half, syQx, cqx) * rL;
                f_ew_prev = (f_w >> 1) | ((f_e >> 1) << 2);
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

template <class R> static int launch_mh_march2(const StepArgs& a_in, const TmaBlockMap& maps, int alt, int sm_count, cudaStream_t st) {
    using T = March<R, 1, false, HP_MARCH_MH_RR>;
    StepArgs a = a_in;
    if (a.y1 <= a.y0) return 0;
    static bool configured[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        cudaFuncSetAttribute(mh_step_march2<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        cudaFuncSetAttribute(mh_step_march2<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
        configured[dev] = true;
    }
    const int grid = march_grid(a, T::USE, T::NW, sizeof(R) == 8 ? HP_MH2_CTAS64 : HP_MH2_CTAS32, sm_count);
    a.total_ctas = grid; a.march_runs = march_runs(a, T::USE, T::NW, grid);
    if (alt) mh_step_march2<R, true><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    else mh_step_march2<R, false><<<grid, T::NW * 32, T::SMEM_BYTES, st>>>(a, maps);
    return 1;
}

}  // namespace HP_NS
