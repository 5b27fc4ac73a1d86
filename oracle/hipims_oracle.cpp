/*
 * hipims_oracle.cpp -- CPU restatement of the HiPIMS explicit cell-update path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * executor: it is imported only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product
 * (hipims_ocl_b200/, include/) never links or calls it.
 *
 * Parity status: PINNED AGAINST THE REFERENCE ITSELF.  The reference ships no
 * golden vectors, so the pin is the reference's own
 * kernel sources compiled here through oracle/ref_shim (-> oracle/_ref/*.so);
 * tests/test_oracle_vs_reference.py asserts this restatement is bit-identical
 * to them (both built with -ffp-contract=off), and tests/golden/ holds vectors
 * generated from that reference build so the check travels to the GPU box.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference).  The arithmetic order of every expression follows the
 * reference so that results are bit-identical without FMA contraction; the
 * code structure (templates, structs, loops) is ours.
 *
 * Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/build_oracle.py)
 */
#include <algorithm>
#include <cmath>
#include <cstdint>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "sim_driver.h"

namespace {

using hpo::Vec4;

enum Dir { N = 0, E = 1, S = 2, W = 3 };  // src/Domain/Cartesian/CLDomainCartesian.clh:32-35

/* Per-run constants; literals are converted to the working precision exactly as
 * -cl-single-precision-constant would (src/OpenCL/Executors/COCLProgram.cpp:69-73). */
template <class R> struct Consts {
    R g, eps, eps10, delta, courant, end_time, fixed_dt;
    int64_t cols, rows;
    bool dt0_keep;  // gts_cacheEnabled returns before any write when dt <= 0 (src/Schemes/CLSchemeGodunov.clc:477-478)
    explicit Consts(const hpo_config& c)
        : g(R(9.81)),  // src/OpenCL/Executors/CLUniversalHeader.clh:33
          eps(static_cast<R>(c.very_small)), eps10(static_cast<R>(c.quite_small)), delta(static_cast<R>(c.delta)),
          courant(static_cast<R>(c.courant)), end_time(static_cast<R>(c.end_time)),
          fixed_dt(static_cast<R>(c.fixed_dt)), cols(c.cols), rows(c.rows),
          dt0_keep((c.quirks & HPO_QUIRK_GODUNOV_DT0_KEEP) != 0) {}
};

/* One side of a Riemann problem: the reference's 8-vector {eta,h,qx,qy,u,v,zb,-}. */
template <class R> struct Side { R eta, h, qx, qy, zb; };
template <class R> struct Flux { R m, fx, fy; };

inline void set_threads(const hpo_config& c) {
#ifdef _OPENMP
    if (c.threads > 0) omp_set_num_threads(c.threads);
#else
    (void)c;
#endif
}

/* ------------------------------------------------------------------------------------------
 * HLLC approximate Riemann solver.  src/Solvers/CLSolverHLLC.clc:27-248
 * ---------------------------------------------------------------------------------------- */
template <class R> Flux<R> hllc(const Consts<R>& k, int dir, const Side<R>& L, const Side<R>& Rr) {
    const R g = k.g, half = R(0.5);
    const unsigned dx = (dir == N || dir == S) ? 0u : 1u, dy = 1u - dx;  // :42
    const R fdx = static_cast<R>(dx), fdy = static_cast<R>(dy);

    if (L.h < k.eps && Rr.h < k.eps) {  // :45-61 both sides dry
        const R s = L.eta + Rr.eta;
        const R p = (s / 2) * (s / 2) - L.zb * s;
        return Flux<R>{R(0), fdx * half * g * p, fdy * half * g * p};
    }

    const R uL = L.h < k.eps ? R(0) : L.qx / L.h, vL = L.h < k.eps ? R(0) : L.qy / L.h;          // :87-88
    const R uR = Rr.h < k.eps ? R(0) : Rr.qx / Rr.h, vR = Rr.h < k.eps ? R(0) : Rr.qy / Rr.h;    // :91-92

    const R velL = fdx * uL + fdy * vL, velR = fdx * uR + fdy * vR;                              // :96-99
    const R disL = fdx * L.qx + fdy * L.qy, disR = fdx * Rr.qx + fdy * Rr.qy;                    // :100-103
    const R aL = std::sqrt(g * L.h), aR = std::sqrt(g * Rr.h);                                   // :104-107

    const R aAvg = (aL + aR) / 2;                                                                // :123
    const R hStar = ((aAvg + (velL - velR) / 4) * (aAvg + (velL - velR) / 4)) / g;               // :124
    const R uStar = (velL + velR) / 2 + aL - aR;                                                 // :125
    const R aStar = std::sqrt(g * hStar);                                                        // :126

    R sL, sR;
    if (L.h < k.eps) sL = velR - 2 * aR;                                                         // :129-134
    else sL = ((velL - aL) > (uStar - aStar)) ? (uStar - aStar) : (velL - aL);
    if (Rr.h < k.eps) sR = velL + 2 * aL;                                                        // :135-140
    else sR = ((velR + aR) < (uStar + aStar)) ? (uStar + aStar) : (velR + aR);
    const R sM = (sL * Rr.h * (velR - sR) - sR * L.h * (velL - sL)) /
                 (Rr.h * (velR - sR) - L.h * (velL - sL));                                       // :141-142

    /* note the LEFT bed is used for both sides, :154-155 */
    const Flux<R> FL{disL, velL * L.qx + fdx * half * g * (L.eta * L.eta - 2 * L.zb * L.eta),
                     velL * L.qy + fdy * half * g * (L.eta * L.eta - 2 * L.zb * L.eta)};
    const Flux<R> FR{disR, velR * Rr.qx + fdx * half * g * (Rr.eta * Rr.eta - 2 * L.zb * Rr.eta),
                     velR * Rr.qy + fdy * half * g * (Rr.eta * Rr.eta - 2 * L.zb * Rr.eta)};

    const bool left = sL >= R(0);                                                                // :174-177
    const bool mid1 = sL < R(0) && sR >= R(0) && sM >= R(0);
    const bool mid2 = sL < R(0) && sR >= R(0) && !mid1;
    const bool right = !left && !mid1 && !mid2;
    if (left) return FL;
    if (right) return FR;

    const R fmL = fdx * FL.fx + fdy * FL.fy, fmR = fdx * FR.fx + fdy * FR.fy;                    // :200-201
    const R f1 = (sR * FL.m - sL * FR.m + sL * sR * (Rr.eta - L.eta)) / (sR - sL);               // :202
    const R f2 = (sR * fmL - sL * fmR + sL * sR * (disR - disL)) / (sR - sL);                    // :203
    if (mid1) return Flux<R>{f1, fdx * f2 + fdy * f1 * uL, fdx * f1 * vL + fdy * f2};            // :206-214
    return Flux<R>{f1, fdx * f2 + fdy * f1 * uR, fdx * f1 * vR + fdy * f2};                      // :216-224
}

/* ------------------------------------------------------------------------------------------
 * Shared tail of both reconstructions: depth-positive states above the higher bed, the
 * owner-relative vertical shift and the stop counter.
 *   Godunov  src/Schemes/CLSchemeGodunov.clc:83-158
 *   MH       src/Schemes/CLSchemeMUSCLHancock.clc:1154-1229
 * etaL/etaR, zL/zR and the velocities are the "initial values" each variant starts from;
 * ownQ is the owning CELL's discharge in the direction normal to the face.
 * ---------------------------------------------------------------------------------------- */
template <class R>
int reconstruct_tail(const Consts<R>& k, int dir, R etaL, R zL, R uL, R vL, R etaR, R zR, R uR, R vR, R ownQ,
                     Side<R>& oL, Side<R>& oR) {
    const R zmax = zL > zR ? zL : zR;
    R shift = zmax - (dir < S ? etaL : etaR);
    if (shift < R(0)) shift = R(0);

    oL.h = (etaL - zmax > R(0)) ? (etaL - zmax) : R(0);
    oL.eta = oL.h + zmax; oL.qx = oL.h * uL; oL.qy = oL.h * vL;
    oR.h = (etaR - zmax > R(0)) ? (etaR - zmax) : R(0);
    oR.eta = oR.h + zmax; oR.qx = oR.h * uR; oR.qy = oR.h * vR;

    /* stop counter; the velocity zeroing the reference also does here has no effect on the
     * flux because the solver recomputes u,v from q/h (CLSolverHLLC.clc:87-92). */
    int stop = 0;
    const R nL = (dir == N || dir == S) ? vL : uL, nR = (dir == N || dir == S) ? vR : uR;
    if (dir == N || dir == E) { if (oL.h <= k.eps && ownQ > R(0)) ++stop; }
    else                      { if (oR.h <= k.eps && ownQ < R(0)) ++stop; }
    if (oR.h <= k.eps && nL < R(0)) ++stop;
    if (oL.h <= k.eps && nR > R(0)) ++stop;

    oL.zb = zmax - shift; oR.zb = zmax - shift;
    oL.eta -= shift; oR.eta -= shift;
    return stop;
}

/* Godunov-type reconstruction from cell states.  src/Schemes/CLSchemeGodunov.clc:27-159 */
template <class R>
int reconstruct_godunov(const Consts<R>& k, int dir, const Vec4<R>& cL, R zL, const Vec4<R>& cR, R zR, Side<R>& oL,
                        Side<R>& oR) {
    const R hL = cL.x - zL, hR = cR.x - zR;                                                      // :39-40
    const R uL = hL < k.eps ? R(0) : cL.z / hL, vL = hL < k.eps ? R(0) : cL.w / hL;              // :49-50
    const R uR = hR < k.eps ? R(0) : cR.z / hR, vR = hR < k.eps ? R(0) : cR.w / hR;              // :58-59
    const Vec4<R>& own = (dir < S) ? cL : cR;
    const R ownQ = (dir == N || dir == S) ? own.w : own.z;
    return reconstruct_tail(k, dir, cL.x, zL, uL, vL, cR.x, zR, uR, vR, ownQ, oL, oR);
}

/* MUSCL-Hancock reconstruction from face estimates {eta,h,qx,qy}.
 * src/Schemes/CLSchemeMUSCLHancock.clc:1119-1230 */
template <class R>
int reconstruct_mh(const Consts<R>& k, int dir, const Vec4<R>& own, const Vec4<R>& eL, const Vec4<R>& eR, Side<R>& oL,
                   Side<R>& oR) {
    const R uL = eL.y <= k.eps ? R(0) : eL.z / eL.y, vL = eL.y <= k.eps ? R(0) : eL.w / eL.y;    // :1140-1141
    const R uR = eR.y <= k.eps ? R(0) : eR.z / eR.y, vR = eR.y <= k.eps ? R(0) : eR.w / eR.y;    // :1149-1150
    const R zL = eL.x - eL.y, zR = eR.x - eR.y;                                                  // :1142,1151
    const R ownQ = (dir == N || dir == S) ? own.w : own.z;
    return reconstruct_tail(k, dir, eL.x, zL, uL, vL, eR.x, zR, uR, vR, ownQ, oL, oR);
}

/* Point-implicit Manning friction.  src/Schemes/CLFriction.clc:26-72 */
template <class R> void implicit_friction(const Consts<R>& k, Vec4<R>& c, R zb, R n, R dt) {
    const R q = std::sqrt(c.z * c.z + c.w * c.w);
    const R h = c.x - zb;
    if (h < k.eps || q < k.eps) return;
    const R cf = (k.g * n * n) / std::pow(h, R(1.0) / R(3.0));
    const R sfx = (-cf / (h * h)) * c.z * q;
    const R sfy = (-cf / (h * h)) * c.w * q;
    const R ddx = R(1.0) + dt * (cf / (h * h)) * (2 * (c.z * c.z) + (c.w * c.w)) / q;
    const R ddy = R(1.0) + dt * (cf / (h * h)) * ((c.z * c.z) + 2 * (c.w * c.w)) / q;
    R fx = sfx / ddx, fy = sfy / ddy;
    if (c.z >= R(0)) { if (fx < -c.z / dt) fx = -c.z / dt; } else { if (fx > -c.z / dt) fx = -c.z / dt; }
    if (c.w >= R(0)) { if (fy < -c.w / dt) fy = -c.w / dt; } else { if (fy > -c.w / dt) fy = -c.w / dt; }
    c.z = c.z + dt * fx;
    c.w = c.w + dt * fy;
}

template <class R> inline R chop(R v, R eps) {  // "round delta values to zero if small"
    if ((v > R(0) && v < eps) || (v < R(0) && v > -eps)) return R(0);
    return v;
}

/* ------------------------------------------------------------------------------------------
 * First-order Godunov cell update.  src/Schemes/CLSchemeGodunov.clc:164-384
 * ---------------------------------------------------------------------------------------- */
template <class R>
void godunov_cell(const Consts<R>& k, bool friction, int64_t x, int64_t y, R dt, const R* bed, const Vec4<R>* src,
                  Vec4<R>* dst, const R* manning) {
    const int64_t id = y * k.cols + x;
    if (dt <= R(0)) {
        if (!k.dt0_keep) { dst[id] = src[id]; return; }                                          // gts_cacheDisabled, :201-206
        /* gts_cacheEnabled: the disabled-cell copy comes first (:449-454), then the return without a write (:477-478) */
        const Vec4<R> c0 = src[id];
        if (c0.y <= R(-9999.0) || c0.x == R(-9999.0)) dst[id] = c0;
        return;
    }
    Vec4<R> c = src[id];
    const R zb = bed[id];
    if (c.y <= R(-9999.0) || c.x == R(-9999.0)) { dst[id] = c; return; }                         // :214-218

    Vec4<R> cN = src[id + k.cols], cE = src[id + 1], cS = src[id - k.cols], cW = src[id - 1];
    R zN = bed[id + k.cols], zE = bed[id + 1], zS = bed[id - k.cols], zW = bed[id - 1];

    int dry = 0;                                                                                 // :248-255
    if (c.x - zb < k.eps) ++dry;
    if (cN.x - zN < k.eps) ++dry;
    if (cE.x - zE < k.eps) ++dry;
    if (cS.x - zS < k.eps) ++dry;
    if (cW.x - zW < k.eps) ++dry;
    if (dry >= 5) return;

    Side<R> l, r;
    int stop = 0;
    stop += reconstruct_godunov(k, N, c, zb, cN, zN, l, r);  cN.x = r.eta; zN = r.zb;            // :259-277
    const Flux<R> fN = hllc(k, N, l, r);
    stop += reconstruct_godunov(k, S, cS, zS, c, zb, l, r);  cS.x = l.eta; zS = l.zb;            // :280-291
    const Flux<R> fS = hllc(k, S, l, r);
    stop += reconstruct_godunov(k, E, c, zb, cE, zE, l, r);  cE.x = r.eta; zE = r.zb;            // :294-305
    const Flux<R> fE = hllc(k, E, l, r);
    stop += reconstruct_godunov(k, W, cW, zW, c, zb, l, r);  cW.x = l.eta; zW = l.zb;            // :308-319
    const Flux<R> fW = hllc(k, W, l, r);

    const R srcX = -1 * k.g * ((cE.x + cW.x) / 2) * ((zE - zW) / k.delta);                       // :324
    const R srcY = -1 * k.g * ((cN.x + cS.x) / 2) * ((zN - zS) / k.delta);                       // :325

    R dEta = (fE.m - fW.m) / k.delta + (fN.m - fS.m) / k.delta - R(0);                           // :328-336
    R dQx = (fE.fx - fW.fx) / k.delta + (fN.fx - fS.fx) / k.delta - srcX;
    R dQy = (fE.fy - fW.fy) / k.delta + (fN.fy - fS.fy) / k.delta - srcY;
    dEta = chop(dEta, k.eps); dQx = chop(dQx, k.eps); dQy = chop(dQy, k.eps);                    // :340-348

    if (stop > 0) { c.z = R(0); c.w = R(0); }                                                    // :351-355
    c.x = c.x - dt * dEta; c.z = c.z - dt * dQx; c.w = c.w - dt * dQy;                           // :358-360
    if (friction) implicit_friction(k, c, zb, manning[id], dt);                                  // :362-372
    if (c.x > c.y && c.y > R(-9990.0)) c.y = c.x;                                                // :375-376
    if (c.x - zb < k.eps) c.x = zb;                                                              // :379-380
    dst[id] = c;
}

/* ------------------------------------------------------------------------------------------
 * MINMOD limited slopes.  src/Schemes/Limiters/CLSlopeLimiterMINMOD.clc:26-70
 * ---------------------------------------------------------------------------------------- */
template <class R> inline R limited_slope(R l, R c, R r) {
    const R a = c - l, b = r - c, beta = R(1.0);
    const R ratio = (std::fabs(a) <= R(0)) ? R(0) : (b / a);
    return std::fmax(std::fmax(R(0), std::fmin(beta * ratio, R(1.0))), std::fmin(ratio, beta)) * a;
}
template <class R>
Vec4<R> slope_limiter(const Consts<R>& k, const Vec4<R>& l, const Vec4<R>& c, const Vec4<R>& r, R zl, R zc, R zr) {
    if ((l.x - zl) < k.eps || (r.x - zr) < k.eps) return Vec4<R>{0, 0, 0, 0};
    return Vec4<R>{limited_slope(l.x, c.x, r.x), limited_slope(l.x - zl, c.x - zc, r.x - zr),
                   limited_slope(l.z, c.z, r.z), limited_slope(l.w, c.w, r.w)};
}

/* Face value {eta,h,qx,qy} = cell + coef*slope.  CLSchemeMUSCLHancock.clc:389-403 */
template <class R> inline Vec4<R> face_extrapolate(R zb, const Vec4<R>& c, const Vec4<R>& s, R coef) {
    return Vec4<R>{c.x + coef * s.x, (c.x - zb) + coef * s.y, c.z + coef * s.z, c.w + coef * s.w};
}
/* Analytic flux of a face state.  CLSchemeMUSCLHancock.clc:420-471 */
template <class R> inline Flux<R> flux_x(const Consts<R>& k, const Vec4<R>& f) {
    const R u = f.y < k.eps ? R(0) : f.z / f.y;
    return Flux<R>{f.z, u * f.z + R(0.5) * k.g * ((f.x * f.x) - 2 * (f.x - f.y) * f.x), u * f.w};
}
template <class R> inline Flux<R> flux_y(const Consts<R>& k, const Vec4<R>& f) {
    const R v = f.y < k.eps ? R(0) : f.w / f.y;
    return Flux<R>{f.w, v * f.z, v * f.w + R(0.5) * k.g * ((f.x * f.x) - 2 * (f.x - f.y) * f.x)};
}

/* MUSCL-Hancock predictor for one cell.  CLSchemeMUSCLHancock.clc:301-382 (+ :476-526 evolve) */
template <class R>
void mh_predict(const Consts<R>& k, R dt, Vec4<R> c, const Vec4<R>& cN, const Vec4<R>& cE, const Vec4<R>& cS,
                const Vec4<R>& cW, R zb, R zN, R zE, R zS, R zW, Vec4<R>& oN, Vec4<R>& oE, Vec4<R>& oS, Vec4<R>& oW) {
    const bool first_order = (c.x - zb < R(1E-5)) || cN.y <= R(-9998.0) || cE.y <= R(-9998.0) ||
                             cS.y <= R(-9998.0) || cW.y <= R(-9998.0);                           // :325-330
    c.y = c.x - zb;                                                                              // :333
    oN = oE = oS = oW = c;
    if (first_order) return;

    const Vec4<R> sx = slope_limiter(k, cW, c, cE, zW, zb, zE);                                  // :343-346
    const Vec4<R> sy = slope_limiter(k, cS, c, cN, zS, zb, zN);
    oN = face_extrapolate(zb, c, sy, R(+0.5)); oE = face_extrapolate(zb, c, sx, R(+0.5));        // :349-352
    oS = face_extrapolate(zb, c, sy, R(-0.5)); oW = face_extrapolate(zb, c, sx, R(-0.5));
    const Flux<R> fN = flux_y(k, oN), fE = flux_x(k, oE), fS = flux_y(k, oS), fW = flux_x(k, oW);  // :355-358

    /* half-step evolve, :476-526 */
    const R srcX = -1 * k.g * ((oE.x + oW.x) / 2) * (((oE.x - oE.y) - (oW.x - oW.y)) / k.delta);
    const R srcY = -1 * k.g * ((oN.x + oS.x) / 2) * (((oN.x - oN.y) - (oS.x - oS.y)) / k.delta);
    R dEta = (fE.m - fW.m) / k.delta + (fN.m - fS.m) / k.delta - R(0);
    R dQx = (fE.fx - fW.fx) / k.delta + (fN.fx - fS.fx) / k.delta - srcX;
    R dQy = (fE.fy - fW.fy) / k.delta + (fN.fy - fS.fy) / k.delta - srcY;
    dEta = chop(dEta, k.eps); dQx = chop(dQx, k.eps); dQy = chop(dQy, k.eps);
    c.x = c.x - R(0.5) * dt * dEta; c.z = c.z - R(0.5) * dt * dQx; c.w = c.w - R(0.5) * dt * dQy;

    oN = face_extrapolate(zb, c, sy, R(+0.5)); oE = face_extrapolate(zb, c, sx, R(+0.5));        // :376-379
    oS = face_extrapolate(zb, c, sy, R(-0.5)); oW = face_extrapolate(zb, c, sx, R(-0.5));
}

/* Stage-1 kernel wrapper.  CLSchemeMUSCLHancock.clc:28-152 */
template <class R>
void mh_stage1_cell(const Consts<R>& k, int64_t x, int64_t y, R dt, const R* bed, const Vec4<R>* st, Vec4<R>* fN,
                    Vec4<R>* fE, Vec4<R>* fS, Vec4<R>* fW) {
    if (dt <= R(0)) return;                                                                      // :66-67
    const int64_t id = y * k.cols + x;
    const Vec4<R> c = st[id], cN = st[id + k.cols], cE = st[id + 1], cS = st[id - k.cols], cW = st[id - 1];
    if (c.y <= R(-9999.0) && cN.y <= R(-9999.0) && cE.y <= R(-9999.0) && cS.y <= R(-9999.0) &&
        cW.y <= R(-9999.0)) return;                                                              // :92-97
    Vec4<R> oN, oE, oS, oW;
    mh_predict(k, dt, c, cN, cE, cS, cW, bed[id], bed[id + k.cols], bed[id + 1], bed[id - k.cols], bed[id - 1], oN, oE,
               oS, oW);
    fN[id] = oN; fE[id] = oE; fS[id] = oS; fW[id] = oW;
}

/* Stage-2 (corrector) kernel, in place.  CLSchemeMUSCLHancock.clc:533-801 */
template <class R>
void mh_stage2_cell(const Consts<R>& k, bool friction, int64_t x, int64_t y, R dt, Vec4<R>* st, const R* bed,
                    const R* manning, const Vec4<R>* fN, const Vec4<R>* fE, const Vec4<R>* fS, const Vec4<R>* fW) {
    if (dt <= R(0)) return;                                                                      // :576-577
    const int64_t id = y * k.cols + x;
    Vec4<R> c = st[id];
    const R zb = bed[id];
    if (c.y <= R(-9999.0) || c.x == R(-9999.0)) return;                                          // :593-594
    const int64_t iN = id + k.cols, iE = id + 1, iS = id - k.cols, iW = id - 1;

    int dry = 0;
    if (c.x - zb < k.eps) ++dry;                                                                 // :596-597
    /* neighbours count as dry on their eta_MAX, :633-634 */
    if (st[iN].y < k.eps) ++dry;
    if (st[iE].y < k.eps) ++dry;
    if (st[iS].y < k.eps) ++dry;
    if (st[iW].y < k.eps) ++dry;
    if (dry >= 5) return;                                                                        // :638

    Side<R> l, r;
    int stop = 0;
    /* internal face = own cell's, external = neighbour's opposite face, :582-583 */
    stop += reconstruct_mh(k, N, c, fN[id], fS[iN], l, r);  const R eN = r.eta, zN = r.zb;       // :642-655
    const Flux<R> xN = hllc(k, N, l, r);
    stop += reconstruct_mh(k, E, c, fE[id], fW[iE], l, r);  const R eE = r.eta, zE = r.zb;       // :658-671
    const Flux<R> xE = hllc(k, E, l, r);
    stop += reconstruct_mh(k, S, c, fN[iS], fS[id], l, r);  const R eS = l.eta, zS = l.zb;       // :674-687
    const Flux<R> xS = hllc(k, S, l, r);
    stop += reconstruct_mh(k, W, c, fE[iW], fW[id], l, r);  const R eW = l.eta, zW = l.zb;       // :690-703
    const Flux<R> xW = hllc(k, W, l, r);

    const R srcX = -1 * k.g * ((eE + eW) / 2) * ((zE - zW) / k.delta);                           // :708
    const R srcY = -1 * k.g * ((eN + eS) / 2) * ((zN - zS) / k.delta);                           // :709
    R dEta = (xE.m - xW.m) / k.delta + (xN.m - xS.m) / k.delta - R(0);                           // :712-720
    R dQx = (xE.fx - xW.fx) / k.delta + (xN.fx - xS.fx) / k.delta - srcX;
    R dQy = (xE.fy - xW.fy) / k.delta + (xN.fy - xS.fy) / k.delta - srcY;
    dEta = chop(dEta, k.eps); dQx = chop(dQx, k.eps); dQy = chop(dQy, k.eps);                    // :723-731

    if (stop > 0) { c.w = R(0); c.z = R(0); }                                                    // :734-738
    c.x = c.x - dt * dEta; c.z = c.z - dt * dQx; c.w = c.w - dt * dQy;                           // :743-745
    if (friction) implicit_friction(k, c, zb, manning[id], dt);                                  // :779-789
    if (c.x - zb < k.eps) c.x = zb;                                                              // :792-793
    if (c.x > c.y && c.y > R(-9990.0)) c.y = c.x;                                                // :796-797
    st[id] = c;
}

/* ------------------------------------------------------------------------------------------
 * Partial-inertial scheme.  src/Schemes/CLSchemeInertial.clc:27-163, 335-378
 * ---------------------------------------------------------------------------------------- */
template <class R>
R inertial_flux(const Consts<R>& k, R n, R dt, R prev, R etaUp, R zUp, R etaDown, R zDown) {
    const R froude = R(0.8);  // CLSchemeInertial.clh:24
    R q = R(0);
    const R h = std::fmax(etaDown, etaUp) - std::max(zUp, zDown);                                // :346
    const R slope = (etaDown - etaUp) / k.delta;                                                 // :347
    q = (prev - (k.g * h * dt * slope)) /
        (R(1.0) + k.g * h * dt * n * n * std::fabs(prev) / std::pow(h, R(10.0) / R(3.0)));       // :350-352
    if (q > R(0) && ((std::fabs(q) / h) / std::sqrt(k.g * h)) > froude) q = h * std::sqrt(k.g * h) * froude;
    if (q < R(0) && ((std::fabs(q) / h) / std::sqrt(k.g * h)) > froude) q = R(0) - h * std::sqrt(k.g * h) * froude;
    if (h < k.eps) q = R(0);                                                                     // :373-374
    return q;
}

template <class R>
void inertial_cell(const Consts<R>& k, int64_t x, int64_t y, R dt, const R* bed, const Vec4<R>* src, Vec4<R>* dst,
                   const R* manning) {
    if (dt <= R(0)) return;                                                                      // :60-61
    const int64_t id = y * k.cols + x;
    Vec4<R> c = src[id];
    const R zb = bed[id], n = manning[id];
    if (c.y <= R(-9999.0) || c.x == R(-9999.0)) { dst[id] = c; return; }                         // :69-73
    const Vec4<R> cN = src[id + k.cols], cE = src[id + 1], cS = src[id - k.cols], cW = src[id - 1];
    const R zN = bed[id + k.cols], zE = bed[id + 1], zS = bed[id - k.cols], zW = bed[id - 1];
    int dry = 0;                                                                                 // :92-99
    if (c.x - zb < k.eps) ++dry;
    if (cN.x - zN < k.eps) ++dry;
    if (cE.x - zE < k.eps) ++dry;
    if (cS.x - zS < k.eps) ++dry;
    if (cW.x - zW < k.eps) ++dry;
    if (dry >= 5) return;

    const R qN = inertial_flux(k, n, dt, cN.w, cN.x, zN, c.x, zb);                               // :103-111
    const R qE = inertial_flux(k, n, dt, cE.z, cE.x, zE, c.x, zb);                               // :113-121
    const R qS = inertial_flux(k, n, dt, c.w, c.x, zb, cS.x, zS);                                // :123-131
    const R qW = inertial_flux(k, n, dt, c.z, c.x, zb, cW.x, zW);                                // :133-141
    c.z = qW; c.w = qS;                                                                          // :143-144
    const R dEta = (qE - qW + qN - qS) / k.delta;                                                // :147-148
    c.x = c.x + dt * dEta;                                                                       // :151
    if (c.x > c.y) c.y = c.x;                                                                    // :154-155
    if (c.x - zb < k.eps) c.x = zb;                                                              // :158-159
    dst[id] = c;
}

/* ------------------------------------------------------------------------------------------
 * The policy class handed to hpo::Sim.
 * ---------------------------------------------------------------------------------------- */
template <class R> struct OracleKernels {
    static void configure(const hpo_config& c, size_t) { set_threads(c); }

    /* every scheme kernel freezes the outer ring: CLSchemeGodunov.clc:183-187 etc. */
    template <class F> static void interior(const hpo_config& c, int ring, F&& f) {
#pragma omp parallel for schedule(static)
        for (int64_t y = ring; y < c.rows - ring; ++y)
            for (int64_t x = ring; x < c.cols - ring; ++x) f(x, y);
    }

    static void gts(const hpo_config& c, const R* dt, const R* bed, const Vec4<R>* src, Vec4<R>* dst, const R* mn) {
        const Consts<R> k(c); const R t = *dt; const bool fr = c.friction != 0;
        interior(c, 1, [&](int64_t x, int64_t y) { godunov_cell(k, fr, x, y, t, bed, src, dst, mn); });
    }
    static void ine(const hpo_config& c, const R* dt, const R* bed, const Vec4<R>* src, Vec4<R>* dst, const R* mn) {
        const Consts<R> k(c); const R t = *dt;
        interior(c, 1, [&](int64_t x, int64_t y) { inertial_cell(k, x, y, t, bed, src, dst, mn); });
    }
    static void mch_1st(const hpo_config& c, const R* dt, const R* bed, const Vec4<R>* st, Vec4<R>* fN, Vec4<R>* fE,
                        Vec4<R>* fS, Vec4<R>* fW) {
        const Consts<R> k(c); const R t = *dt;
        interior(c, 1, [&](int64_t x, int64_t y) { mh_stage1_cell(k, x, y, t, bed, st, fN, fE, fS, fW); });
    }
    static void mch_2nd(const hpo_config& c, const R* dt, Vec4<R>* st, const R* bed, const R* mn, const Vec4<R>* fN,
                        const Vec4<R>* fE, const Vec4<R>* fS, const Vec4<R>* fW) {
        const Consts<R> k(c); const R t = *dt; const bool fr = c.friction != 0;
        /* ring of two is frozen, CLSchemeMUSCLHancock.clc:569-573.  The in-place update is safe
         * cell-by-cell: neighbours are read only for eta_max (Q8), see DESIGN.md. */
        interior(c, 2, [&](int64_t x, int64_t y) { mh_stage2_cell(k, fr, x, y, t, st, bed, mn, fN, fE, fS, fW); });
    }

    /* Stage 1 of the CFL reduction: per-worker strided maximum of the wave speed.
     * src/Schemes/CLDynamicTimestep.clc:166-249 (inertial: TIMESTEP_SIMPLIFIED,
     * src/Schemes/CLSchemeInertial.clh:25) */
    static void reduce(const hpo_config& c, const Vec4<R>* st, const R* bed, R* red, size_t workers) {
        const Consts<R> k(c);
        const int64_t cells = c.cols * c.rows;
        const bool simplified = c.scheme == HPO_SCHEME_INERTIAL;
#pragma omp parallel for schedule(static)
        for (int64_t wk = 0; wk < static_cast<int64_t>(workers); ++wk) {
            R best = R(0);
            for (int64_t id = wk; id < cells; id += static_cast<int64_t>(workers)) {
                const Vec4<R> s = st[id];
                const R h = s.x - bed[id];
                R speed = R(0);
                if (h > k.eps10 && s.y > R(-9999.0)) {
                    R vx, vy;
                    if (!simplified) {
                        vx = s.z / h; vy = s.w / h;
                        if (vx < R(0)) vx = -vx;
                        if (vy < R(0)) vy = -vy;
                        vx += std::sqrt(k.g * h); vy += std::sqrt(k.g * h);
                    } else { vx = std::sqrt(k.g * h); vy = std::sqrt(k.g * h); }
                    speed = (vx < vy) ? vy : vx;
                }
                if (speed > best) best = speed;
            }
            red[wk] = best;
        }
    }

    /* Stage 2 + device-resident time controller.  CLDynamicTimestep.clc:27-146 */
    static void advance(const hpo_config& c, R* time, R* timestep, R* hydro, const R* red, size_t workers,
                        const R* target, R* batch, uint32_t* ok, uint32_t* skipped) {
        const Consts<R> k(c);
        R t = *time, dt = std::fmax(R(0), *timestep), th = *hydro;
        const R sync = *target;
        t += dt; *batch += dt;
        if (dt > R(0)) ++*ok; else ++*skipped;
        if (th > R(1.0)) th = dt; else th += dt;                                                 // :61-66
        if (c.dynamic) {
            R vmax = R(0);
            for (size_t i = 0; i < workers; ++i) if (red[i] > vmax) vmax = red[i];               // :73-80
            R tmin = k.delta / vmax;                                                             // :84
            if (t < R(1.0) && tmin < R(1E-10)) tmin = R(1E-10);                                  // :85-86
            dt = k.courant * tmin;                                                               // :89
        } else {
            dt = k.fixed_dt;                                                                     // :94
        }
        if (dt > R(0) && dt < R(1E-10)) dt = R(1E-10);                                           // :112-113
        if ((t + dt) >= sync) {                                                                  // :118-124
            if (sync - t > k.eps) dt = sync - t;
            if (sync - t <= k.eps) dt = -dt;
        }
        if (t < R(60.0) && dt > R(0.1)) dt = R(0.1);                                             // :128-129
        if ((t + dt) > k.end_time) dt = k.end_time - t;                                          // :132-133
        if (dt > R(15.0)) dt = R(15.0);                                                          // :136-137
        *time = t; *timestep = dt; *hydro = th;
    }

    /* Timestep refresh after a sync point.  CLDynamicTimestep.clc:255-317 */
    static void update_timestep(const hpo_config& c, const R* time, R* timestep, const R* red, size_t workers,
                                const R* target, R* batch) {
        const Consts<R> k(c);
        const R t = *time, original = std::fabs(*timestep), sync = *target;
        R dt = R(0);
        if (c.dynamic) {
            R vmax = R(0);
            for (size_t i = 0; i < workers; ++i) if (red[i] > vmax) vmax = red[i];
            R tmin = k.delta / vmax;
            if (t < R(1.0) && tmin < R(1E-10)) tmin = R(1E-10);
            dt = k.courant * tmin;
        }
        dt = std::fmin(dt, original);                                                            // :299
        *batch = *batch - original + dt;                                                         // :300
        if (t < R(60.0) && dt > R(0.1)) dt = R(0.1);                                             // :303-304
        if ((t + dt) >= sync) dt = std::fmax(R(0), sync - t);                                    // :307-308
        if (dt > R(15.0)) dt = R(15.0);                                                          // :311-312
        *timestep = dt;
    }

    /* Uniform rain / loss.  src/Boundaries/CLBoundaries.clc:130-184 */
    static void bdy_uniform(const hpo_config& c, const hpo::BdyUniformConf<R>* conf, const R* series, const R* time,
                            const R* timestep, const R* hydro, Vec4<R>* st, const R* bed, const R*, int64_t gx,
                            int64_t gy) {
        const R t = *time, real_dt = *timestep, acc = *hydro;
        if (acc < R(1.0) || real_dt <= R(0)) return;                                             // :165-166
        if (t >= conf->TimeseriesLength) return;                                                 // :168
        const uint64_t step = static_cast<uint64_t>(std::floor(t / conf->TimeseriesInterval));   // :172
        const R rate = series[2 * step + 1];
        const int64_t x1 = std::min<int64_t>(gx, c.cols - 1), y1 = std::min<int64_t>(gy, c.rows - 1);
#pragma omp parallel for schedule(static)
        for (int64_t y = 1; y < y1; ++y)
            for (int64_t x = 1; x < x1; ++x) {
                Vec4<R>& s = st[y * c.cols + x];
                if (s.y <= R(-9999.0)) continue;
                if (conf->Definition == 0) s.x += rate / R(3600000.0) * acc;                     // :176-177
                if (conf->Definition == 1) s.x = std::max(bed[y * c.cols + x], s.x - rate / R(3600000.0) * acc);  // :179-180
            }
    }

    /* Gridded rain / mass flux.  src/Boundaries/CLBoundaries.clc:186-246 */
    static void bdy_gridded(const hpo_config& c, const hpo::BdyGriddedConf<R>* conf, const R* series, const R* time,
                            const R*, const R* hydro, Vec4<R>* st, const R*, const R*, int64_t gx, int64_t gy) {
        const R t = *time, acc = *hydro, delta = static_cast<R>(c.delta);
        if (acc < R(1.0)) return;                                                                // :224-225
        uint64_t step = static_cast<uint64_t>(std::floor(t / conf->TimeseriesInterval));         // :228-229
        if (step >= conf->TimeseriesEntries) step = conf->TimeseriesEntries;
        const int64_t x1 = std::min<int64_t>(gx, c.cols - 1), y1 = std::min<int64_t>(gy, c.rows - 1);
#pragma omp parallel for schedule(static)
        for (int64_t y = 1; y < y1; ++y)
            for (int64_t x = 1; x < x1; ++x) {
                Vec4<R>& s = st[y * c.cols + x];
                if (s.y <= R(-9999.0) || s.x == R(-9999.0)) continue;                            // :220-221
                const R col = std::floor(((static_cast<R>(x) * delta) - conf->GridOffsetX) / conf->GridResolution);
                const R row = std::floor(((static_cast<R>(y) * delta) - conf->GridOffsetY) / conf->GridResolution);
                const uint64_t cell = (conf->GridRows * conf->GridCols) * step +
                                      (conf->GridCols * static_cast<uint64_t>(row)) + static_cast<uint64_t>(col);
                const R rate = series[cell];
                if (conf->Definition == 0) s.x += rate / R(3600000.0) * acc;                     // :238-239
                if (conf->Definition == 2) s.x += rate / (delta * delta) * acc;                  // :241-242
            }
    }

    /* Point / edge cells with an imposed level, discharge or volume.  CLBoundaries.clc:23-128 */
    static void bdy_cell(const hpo_config& c, const hpo::BdyCellConf<R>* conf, const uint64_t* rel, const R* series,
                         const R* time, const R* timestep, const R*, Vec4<R>* st, const R* bed, const R*, int64_t) {
        const Consts<R> k(c);
        const R t = *time, dt = *timestep;
        if (t >= conf->TimeseriesLength || dt <= R(0)) return;                                   // :40-41
        const uint64_t base = static_cast<uint64_t>(std::floor(t / conf->TimeseriesInterval));   // :43-44
        const Vec4<R> a = reinterpret_cast<const Vec4<R>*>(series)[base];
        const Vec4<R> b = reinterpret_cast<const Vec4<R>*>(series)[base + 1];
        const R w = std::fmod(t, conf->TimeseriesInterval) / conf->TimeseriesInterval;           // :52
        for (uint64_t i = 0; i < conf->RelationCount; ++i) {
            Vec4<R> ts{a.x + (b.x - a.x) * w, a.y + (b.y - a.y) * w, a.z + (b.z - a.z) * w, a.w + (b.w - a.w) * w};
            const uint64_t id = rel[i];
            Vec4<R> s = st[id];
            const R zb = bed[id];
            if (conf->DefinitionDepth == 2) {                                                    // :55-61 depth
                s.x = zb + ts.y;
            } else if (conf->DefinitionDepth == 1) {                                             // :62-68 fsl
                s.x = std::fmax(zb, ts.y);
            } else if (std::fabs(ts.z) > k.eps || std::fabs(ts.w) > k.eps || conf->DefinitionDischarge == 3) {
                R depth = (std::fabs(ts.z) * dt) / k.delta + (std::fabs(ts.w) * dt) / k.delta;   // :79
                R critical = std::fmax(std::pow(std::pow(ts.z, R(2)) / k.g, R(1.0) / R(3.0)),
                                       std::pow(std::pow(ts.w, R(2)) / k.g, R(1.0) / R(3.0)));    // :81
                if (conf->DefinitionDischarge == 3) {                                            // :85-93 volume
                    depth = (std::fabs(ts.z) * dt) / (k.delta * k.delta);
                    critical = R(0); ts.z = R(0); ts.w = R(0);
                }
                s.x = std::fmax(zb + critical, s.x + depth);                                     // :95
            }
            if (conf->DefinitionDischarge == 1) s.z = ts.z;                                      // :103-111
            else if (conf->DefinitionDischarge == 2) s.z = ts.z * (s.x - zb);
            if (conf->DefinitionDischarge == 1) s.w = ts.w;                                      // :113-121
            else if (conf->DefinitionDischarge == 2) s.w = ts.w * (s.x - zb);
            st[id] = s;
        }
    }
};

}  // namespace

HPO_DEFINE_API(hpo_f64_, double, OracleKernels<double>)
HPO_DEFINE_API(hpo_f32_, float, OracleKernels<float>)
