// hp_math.cuh -- per-cell arithmetic of the explicit shallow-water update, as device functions.
//
// Written from the numerical specification in SURVEY.md section 10 / DESIGN.md; each function
// names the reference kernel whose result it must reproduce (paths relative to the reference
// root).  Everything is templated on the working precision R (double | float) -- the
// reference switches every cl_double typedef to float for single precision
// (src/OpenCL/Executors/COCLProgram.cpp:381-399).
//
// Faces are solved in a face-normal frame (n = normal, t = tangential component), so the
// reference's multiplications by the 0/1 direction vector (src/Solvers/CLSolverHLLC.clc:42)
// disappear; for finite inputs that is exact, not an approximation.
//
// The translation unit is compiled twice (see hp_kernels.cu): once with -fmad=false
// ("strict": every operation rounded as written, bit-comparable with the IEEE evaluation of
// the reference's expressions) and once with FMA contraction ("fast").
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "hp_kernels.cuh"

namespace HP_NS {

template <class R> struct Params {
    R g;          // GRAVITY 9.81 (src/OpenCL/Executors/CLUniversalHeader.clh:33)
    R eps;        // VERY_SMALL
    R eps10;      // QUITE_SMALL
    R delta;      // DOMAIN_DELTAX
    R courant;    // COURANT_NUMBER
    R end_time;   // SCHEME_ENDTIME
    R fixed_dt;   // TIMESTEP_FIXED
    int dynamic;  // TIMESTEP_DYNAMIC
    int friction; // FRICTION_ENABLED
    int simplified_speed;  // TIMESTEP_SIMPLIFIED (inertial program, CLSchemeInertial.clh:25)
    int dt0_keep;          // gts_cacheEnabled's rule for dt <= 0: return before any write (CLSchemeGodunov.clc:477-478)
};

using hp::Clock;  // device-resident clock record, see hp_kernels.cuh

template <class R> struct Cell { R eta, emax, qx, qy; };
template <class R> struct Flux3 { R m, n, t; };  // mass, normal momentum, tangential momentum

template <class R> __device__ __forceinline__ R hp_sqrt(R v);
template <> __device__ __forceinline__ double hp_sqrt<double>(double v) { return sqrt(v); }
template <> __device__ __forceinline__ float hp_sqrt<float>(float v) { return sqrtf(v); }
template <class R> __device__ __forceinline__ R hp_pow(R a, R b);
template <> __device__ __forceinline__ double hp_pow<double>(double a, double b) { return pow(a, b); }
template <> __device__ __forceinline__ float hp_pow<float>(float a, float b) { return powf(a, b); }
template <class R> __device__ __forceinline__ R hp_abs(R v) { return v < R(0) ? -v : v; }
template <> __device__ __forceinline__ double hp_abs<double>(double v) { return fabs(v); }
template <> __device__ __forceinline__ float hp_abs<float>(float v) { return fabsf(v); }
template <class R> __device__ __forceinline__ R hp_fmax(R a, R b);
template <> __device__ __forceinline__ double hp_fmax<double>(double a, double b) { return fmax(a, b); }
template <> __device__ __forceinline__ float hp_fmax<float>(float a, float b) { return fmaxf(a, b); }
template <class R> __device__ __forceinline__ R hp_fmin(R a, R b);
template <> __device__ __forceinline__ double hp_fmin<double>(double a, double b) { return fmin(a, b); }
template <> __device__ __forceinline__ float hp_fmin<float>(float a, float b) { return fminf(a, b); }
template <class R> __device__ __forceinline__ R hp_floor(R v);
template <> __device__ __forceinline__ double hp_floor<double>(double v) { return floor(v); }
template <> __device__ __forceinline__ float hp_floor<float>(float v) { return floorf(v); }
template <class R> __device__ __forceinline__ R hp_fmod(R a, R b);
template <> __device__ __forceinline__ double hp_fmod<double>(double a, double b) { return fmod(a, b); }
template <> __device__ __forceinline__ float hp_fmod<float>(float a, float b) { return fmodf(a, b); }

#ifndef HP_FLAVOUR_STRICT
// ---------------------------------------------------------------------------------------------
// Low-op-count elementary functions (full working precision to ~1 ulp, no slow paths).
// ---------------------------------------------------------------------------------------------
#ifndef HP_FM_HALLEY
#define HP_FM_HALLEY 1
#endif
__device__ __forceinline__ double fm_rcp(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
#if HP_FM_HALLEY
    // the seed is good to 1e-6 (measured, tools/scratch/fm_accuracy_probe.cu); one third-order step x (1 + e + e^2)
    // reaches 1e-18: three dependent FMAs instead of four, result within 1 ulp of the IEEE quotient
    const double e = fma(-a, x, 1.0);
    return fma(x, fma(e, e, e), x);
#else
    double e = fma(-a, x, 1.0); x = fma(x, e, x);
    e = fma(-a, x, 1.0); x = fma(x, e, x);
    return x;
#endif
}
__device__ __forceinline__ float fm_rcp(float a) {
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(a));
    const float e = fmaf(-a, x, 1.0f);
    return fmaf(x, e, x);
}
// sqrt for a >= 0 (returns 0 for a == 0)
__device__ __forceinline__ double fm_sqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
#if HP_FM_HALLEY
    // sqrt(a) = g (1 - 2r)^(-1/2) = g (1 + r + 3/2 r^2 + O(r^3)) with r = 1/2 - g h ~ 1e-6: one third-order step,
    // within 1 ulp of the IEEE root (six fp64 operations instead of ten, five dependent instead of eight)
    const double r = fma(-g, h, 0.5);
    g = fma(g, fma(1.5 * r, r, r), g);
#else
    double r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5); g = fma(g, r, g); h = fma(h, r, h);
    const double d = fma(-g, g, a);
    g = fma(d, h, g);
#endif
    return a > 0.0 ? g : 0.0;
}
__device__ __forceinline__ float fm_sqrt(float a) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    float g = a * y;
    const float h = 0.5f * y;
    const float d = fmaf(-g, g, a);
    g = fmaf(d, h, g);
    return a > 0.0f ? g : 0.0f;
}
// a^(-1/3) for a > 0: single-precision seed (MUFU lg2/ex2) refined in double precision
// 2^(-lg2(a) / 3) straight from the MUFU units: the callers' arguments are depths above the dry threshold (1e-10) and
// far below 2^126, so neither lg2's denormal scaling nor ex2's range reduction (what exp2f adds) can be needed
__device__ __forceinline__ float fm_rcbrt_seed(float a) {
    float l, y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(a));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(-0.333333343f * l));
    return y;
}
__device__ __forceinline__ double fm_rcbrt(double a) {
    double y = static_cast<double>(fm_rcbrt_seed(static_cast<float>(a)));
#if HP_FM_HALLEY
    // the seed is good to ~2e-7 (MUFU lg2 / ex2); with e = 1 - a y^3, a^(-1/3) = y (1 - e)^(-1/3) =
    // y (1 + e/3 + 2/9 e^2 + O(e^3)): one third-order step, six operations instead of nine
    const double e = fma(-(a * y), y * y, 1.0);
    return fma(y * e, fma(e, 0.22222222222222222, 0.33333333333333333), y);
#else
    double t = y * y * y;
    y = y * fma(-0.33333333333333333 * a, t, 1.3333333333333333);
    t = y * y * y;
    y = y * fma(-0.33333333333333333 * a, t, 1.3333333333333333);
    return y;
#endif
}
__device__ __forceinline__ float fm_rcbrt(float a) { return rcbrtf(a); }

#endif  // !HP_FLAVOUR_STRICT

// ---------------------------------------------------------------------------------------------
// HLLC approximate Riemann solver in the face-normal frame.
// Must reproduce riemannSolver(), src/Solvers/CLSolverHLLC.clc:27-248.
// `zb` is the common (reconstructed) bed of both sides; the reference reads the LEFT bed for
// both (:154-155) and reconstruction always makes the two equal.
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ Flux3<R> hllc(const Params<R>& k, R etaL, R hL, R qnL, R qtL, R etaR, R hR, R qnR, R qtR,
                                         R zb) {
    const R hg = R(0.5) * k.g;
    if (hL < k.eps && hR < k.eps) {                                  // :45-61
        const R s = etaL + etaR;
        const R p = (s / 2) * (s / 2) - zb * s;
        return Flux3<R>{R(0), hg * p, R(0)};
    }
    const R unL = hL < k.eps ? R(0) : qnL / hL, utL = hL < k.eps ? R(0) : qtL / hL;   // :87-92
    const R unR = hR < k.eps ? R(0) : qnR / hR, utR = hR < k.eps ? R(0) : qtR / hR;
    const R aL = hp_sqrt(k.g * hL), aR = hp_sqrt(k.g * hR);                            // :104-107
    const R aAvg = (aL + aR) / 2;                                                      // :123
    const R hs = ((aAvg + (unL - unR) / 4) * (aAvg + (unL - unR) / 4)) / k.g;          // :124
    const R us = (unL + unR) / 2 + aL - aR;                                            // :125
    const R as = hp_sqrt(k.g * hs);                                                    // :126
    R sL, sR;
    if (hL < k.eps) sL = unR - 2 * aR;                                                 // :129-134
    else sL = ((unL - aL) > (us - as)) ? (us - as) : (unL - aL);
    if (hR < k.eps) sR = unL + 2 * aL;                                                 // :135-140
    else sR = ((unR + aR) < (us + as)) ? (us + as) : (unR + aR);

    const Flux3<R> FL{qnL, unL * qnL + hg * (etaL * etaL - 2 * zb * etaL), unL * qtL};  // :146-157
    const Flux3<R> FR{qnR, unR * qnR + hg * (etaR * etaR - 2 * zb * etaR), unR * qtR};
    if (sL >= R(0)) return FL;                                                         // :174-198
    if (!(sR >= R(0))) return FR;

    const R sM = (sL * hR * (unR - sR) - sR * hL * (unL - sL)) / (hR * (unR - sR) - hL * (unL - sL));  // :141-142
    const R f1 = (sR * FL.m - sL * FR.m + sL * sR * (etaR - etaL)) / (sR - sL);        // :202
    const R f2 = (sR * FL.n - sL * FR.n + sL * sR * (qnR - qnL)) / (sR - sL);          // :203
    return Flux3<R>{f1, f2, f1 * ((sM >= R(0)) ? utL : utR)};                          // :206-224
}

// ---------------------------------------------------------------------------------------------
// Non-negative hydrostatic reconstruction at one face + HLLC.  L/R are the two sides in the
// direction of increasing coordinate; `ownIsLeft` says which side owns the update (the
// reference takes the vertical shift relative to the owner, CLSchemeGodunov.clc:85, so the two
// cells sharing a face do NOT see the same momentum flux).
// Reproduces reconstructInterface() tails of src/Schemes/CLSchemeGodunov.clc:83-158 and
// src/Schemes/CLSchemeMUSCLHancock.clc:1154-1229 followed by riemannSolver().
//   eta*, z*   : level and bed the reconstruction starts from
//   un*, ut*   : velocities normal / tangential to the face
//   ownQn      : the owning CELL's discharge normal to the face (stop test)
// Outputs: flux, the neighbour side's reconstructed level and bed (for the bed-slope source),
// and the stop counter increment.
// ---------------------------------------------------------------------------------------------
template <class R, bool ownIsLeft>
__device__ __forceinline__ Flux3<R> face_flux(const Params<R>& k, R etaL, R zL, R unL, R utL, R etaR, R zR, R unR,
                                              R utR, R ownQn, R& neighEta, R& neighZ, int& stop) {
    const R zmax = zL > zR ? zL : zR;
    R shift = zmax - (ownIsLeft ? etaL : etaR);
    if (shift < R(0)) shift = R(0);
    const R hL = (etaL - zmax > R(0)) ? (etaL - zmax) : R(0);
    const R hR = (etaR - zmax > R(0)) ? (etaR - zmax) : R(0);
    if (ownIsLeft) { if (hL <= k.eps && ownQn > R(0)) ++stop; }
    else           { if (hR <= k.eps && ownQn < R(0)) ++stop; }
    if (hR <= k.eps && unL < R(0)) ++stop;
    if (hL <= k.eps && unR > R(0)) ++stop;
    const R zb = zmax - shift;
    const R eL = (hL + zmax) - shift, eR = (hR + zmax) - shift;
    neighEta = ownIsLeft ? eR : eL;
    neighZ = zb;
    return hllc(k, eL, hL, hL * unL, hL * utL, eR, hR, hR * unR, hR * utR, zb);
}

// Point-implicit Manning friction; reproduces implicitFriction(), src/Schemes/CLFriction.clc:26-72.
template <class R> __device__ __forceinline__ void implicit_friction(const Params<R>& k, R eta, R& qx, R& qy, R zb, R n, R dt) {
    const R q = hp_sqrt(qx * qx + qy * qy);
    const R h = eta - zb;
    if (h < k.eps || q < k.eps) return;
    const R cf = (k.g * n * n) / hp_pow(h, R(1.0) / R(3.0));
    const R sfx = (-cf / (h * h)) * qx * q;
    const R sfy = (-cf / (h * h)) * qy * q;
    const R ddx = R(1.0) + dt * (cf / (h * h)) * (2 * (qx * qx) + (qy * qy)) / q;
    const R ddy = R(1.0) + dt * (cf / (h * h)) * ((qx * qx) + 2 * (qy * qy)) / q;
    R fx = sfx / ddx, fy = sfy / ddy;
    if (qx >= R(0)) { if (fx < -qx / dt) fx = -qx / dt; } else { if (fx > -qx / dt) fx = -qx / dt; }
    if (qy >= R(0)) { if (fy < -qy / dt) fy = -qy / dt; } else { if (fy > -qy / dt) fy = -qy / dt; }
    qx = qx + dt * fx;
    qy = qy + dt * fy;
}

template <class R> __device__ __forceinline__ R chop(R v, R eps) {
    return ((v > R(0) && v < eps) || (v < R(0) && v > -eps)) ? R(0) : v;
}

// Wave speed of one cell for the CFL reduction; reproduces the body of tst_Reduce,
// src/Schemes/CLDynamicTimestep.clc:185-223.
template <class R> __device__ __forceinline__ R wave_speed(const Params<R>& k, R eta, R emax, R qx, R qy, R zb) {
    const R h = eta - zb;
    if (!(h > k.eps10 && emax > R(-9999.0))) return R(0);
    const R c = hp_sqrt(k.g * h);
    if (k.simplified_speed) return c;
    R vx = qx / h, vy = qy / h;
    if (vx < R(0)) vx = -vx;
    if (vy < R(0)) vy = -vy;
    vx += c; vy += c;
    return (vx < vy) ? vy : vx;
}

// ---------------------------------------------------------------------------------------------
// First-order Godunov update of one cell from its 5-point stencil.
// Reproduces gts_cacheDisabled after its loads, src/Schemes/CLSchemeGodunov.clc:248-383.
// Returns false when the reference returns without writing (all-dry stencil, :255).
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ bool godunov_update(const Params<R>& k, R dt, Cell<R>& c, R zb, R mann, R etaN, R qxN, R qyN,
                                               R zN, R etaE, R qxE, R qyE, R zE, R etaS, R qxS, R qyS, R zS, R etaW,
                                               R qxW, R qyW, R zW) {
    int dry = 0;
    const R h = c.eta - zb, hN = etaN - zN, hE = etaE - zE, hS = etaS - zS, hW = etaW - zW;
    if (h < k.eps) ++dry;
    if (hN < k.eps) ++dry;
    if (hE < k.eps) ++dry;
    if (hS < k.eps) ++dry;
    if (hW < k.eps) ++dry;
    if (dry >= 5) return false;

    // velocities as reconstructInterface() forms them (:49-50, :58-59)
    const R u = h < k.eps ? R(0) : c.qx / h, v = h < k.eps ? R(0) : c.qy / h;
    const R uN = hN < k.eps ? R(0) : qxN / hN, vN = hN < k.eps ? R(0) : qyN / hN;
    const R uE = hE < k.eps ? R(0) : qxE / hE, vE = hE < k.eps ? R(0) : qyE / hE;
    const R uS = hS < k.eps ? R(0) : qxS / hS, vS = hS < k.eps ? R(0) : qyS / hS;
    const R uW = hW < k.eps ? R(0) : qxW / hW, vW = hW < k.eps ? R(0) : qyW / hW;

    int stop = 0;
    R eN, bN, eE, bE, eS, bS, eW, bW;
    // north / south faces: normal = y, tangential = x
    const Flux3<R> fN = face_flux<R, true>(k, c.eta, zb, v, u, etaN, zN, vN, uN, c.qy, eN, bN, stop);
    const Flux3<R> fS = face_flux<R, false>(k, etaS, zS, vS, uS, c.eta, zb, v, u, c.qy, eS, bS, stop);
    // east / west faces: normal = x, tangential = y
    const Flux3<R> fE = face_flux<R, true>(k, c.eta, zb, u, v, etaE, zE, uE, vE, c.qx, eE, bE, stop);
    const Flux3<R> fW = face_flux<R, false>(k, etaW, zW, uW, vW, c.eta, zb, u, v, c.qx, eW, bW, stop);

    const R srcX = -1 * k.g * ((eE + eW) / 2) * ((bE - bW) / k.delta);                       // :324
    const R srcY = -1 * k.g * ((eN + eS) / 2) * ((bN - bS) / k.delta);                       // :325
    R dEta = (fE.m - fW.m) / k.delta + (fN.m - fS.m) / k.delta - R(0);                       // :328-336
    R dQx = (fE.n - fW.n) / k.delta + (fN.t - fS.t) / k.delta - srcX;
    R dQy = (fE.t - fW.t) / k.delta + (fN.n - fS.n) / k.delta - srcY;
    dEta = chop(dEta, k.eps); dQx = chop(dQx, k.eps); dQy = chop(dQy, k.eps);                // :340-348

    if (stop > 0) { c.qx = R(0); c.qy = R(0); }                                              // :351-355
    c.eta = c.eta - dt * dEta; c.qx = c.qx - dt * dQx; c.qy = c.qy - dt * dQy;               // :358-360
    if (k.friction) implicit_friction(k, c.eta, c.qx, c.qy, zb, mann, dt);                   // :362-372
    if (c.eta > c.emax && c.emax > R(-9990.0)) c.emax = c.eta;                               // :375-376
    if (c.eta - zb < k.eps) c.eta = zb;                                                      // :379-380
    return true;
}

// ---------------------------------------------------------------------------------------------
// MUSCL-Hancock pieces.
// ---------------------------------------------------------------------------------------------
// MINMOD-limited slope; reproduces calculateLimitedSlope(),
// src/Schemes/Limiters/CLSlopeLimiterMINMOD.clc:49-70 (beta = 1).
template <class R> __device__ __forceinline__ R limited_slope(R l, R c, R r) {
    const R a = c - l, b = r - c;
    const R ratio = (hp_abs(a) <= R(0)) ? R(0) : (b / a);
    return hp_fmax(hp_fmax(R(0), hp_fmin(R(1.0) * ratio, R(1.0))), hp_fmin(ratio, R(1.0))) * a;
}

// A face estimate {eta, h, qx, qy} (the reference's cl_double4 per face, CLSchemeMUSCLHancock.clh)
template <class R> struct FaceState { R eta, h, qx, qy; };
template <class R> struct Faces { FaceState<R> n, e, s, w; };

// Predictor for one cell: limited slopes, face extrapolation, half-step evolve, re-extrapolation.
// Reproduces mch_1st(), src/Schemes/CLSchemeMUSCLHancock.clc:301-382 (+ :389-526).
// emaxN..emaxW are the neighbours' eta_max (boundary-cell test, :326-329).
template <class R>
__device__ __forceinline__ Faces<R> mh_predict(const Params<R>& k, R dt, R eta, R qx, R qy, R zb, R etaN, R qxN, R qyN,
                                               R zN, R emaxN, R etaE, R qxE, R qyE, R zE, R emaxE, R etaS, R qxS,
                                               R qyS, R zS, R emaxS, R etaW, R qxW, R qyW, R zW, R emaxW) {
    Faces<R> f;
    const R h = eta - zb;
    f.n = f.e = f.s = f.w = FaceState<R>{eta, h, qx, qy};
    if (h < R(1E-5) || emaxN <= R(-9998.0) || emaxE <= R(-9998.0) || emaxS <= R(-9998.0) || emaxW <= R(-9998.0))
        return f;                                                                                   // :325-340

    R sxE = R(0), sxH = R(0), sxQx = R(0), sxQy = R(0), syE = R(0), syH = R(0), syQx = R(0), syQy = R(0);
    if (!((etaW - zW) < k.eps || (etaE - zE) < k.eps)) {                                            // CLSlopeLimiterMINMOD.clc:38-39
        sxE = limited_slope(etaW, eta, etaE); sxH = limited_slope(etaW - zW, eta - zb, etaE - zE);
        sxQx = limited_slope(qxW, qx, qxE);   sxQy = limited_slope(qyW, qy, qyE);
    }
    if (!((etaS - zS) < k.eps || (etaN - zN) < k.eps)) {
        syE = limited_slope(etaS, eta, etaN); syH = limited_slope(etaS - zS, eta - zb, etaN - zN);
        syQx = limited_slope(qxS, qx, qxN);   syQy = limited_slope(qyS, qy, qyN);
    }
    const R ph = R(0.5), mh = R(-0.5);
    // extrapolation at the current time level, :349-352 / :389-403
    FaceState<R> N{eta + ph * syE, (eta - zb) + ph * syH, qx + ph * syQx, qy + ph * syQy};
    FaceState<R> E{eta + ph * sxE, (eta - zb) + ph * sxH, qx + ph * sxQx, qy + ph * sxQy};
    FaceState<R> S{eta + mh * syE, (eta - zb) + mh * syH, qx + mh * syQx, qy + mh * syQy};
    FaceState<R> W{eta + mh * sxE, (eta - zb) + mh * sxH, qx + mh * sxQx, qy + mh * sxQy};
    // analytic fluxes of the face states, :420-471
    const R hg = R(0.5) * k.g;
    const R vN = N.h < k.eps ? R(0) : N.qy / N.h, vS = S.h < k.eps ? R(0) : S.qy / S.h;
    const R uE = E.h < k.eps ? R(0) : E.qx / E.h, uW = W.h < k.eps ? R(0) : W.qx / W.h;
    const R fNm = N.qy, fNx = vN * N.qx, fNy = vN * N.qy + hg * ((N.eta * N.eta) - 2 * (N.eta - N.h) * N.eta);
    const R fSm = S.qy, fSx = vS * S.qx, fSy = vS * S.qy + hg * ((S.eta * S.eta) - 2 * (S.eta - S.h) * S.eta);
    const R fEm = E.qx, fEx = uE * E.qx + hg * ((E.eta * E.eta) - 2 * (E.eta - E.h) * E.eta), fEy = uE * E.qy;
    const R fWm = W.qx, fWx = uW * W.qx + hg * ((W.eta * W.eta) - 2 * (W.eta - W.h) * W.eta), fWy = uW * W.qy;
    // half-step evolve, :476-526
    const R srcX = -1 * k.g * ((E.eta + W.eta) / 2) * (((E.eta - E.h) - (W.eta - W.h)) / k.delta);
    const R srcY = -1 * k.g * ((N.eta + S.eta) / 2) * (((N.eta - N.h) - (S.eta - S.h)) / k.delta);
    R dEta = (fEm - fWm) / k.delta + (fNm - fSm) / k.delta - R(0);
    R dQx = (fEx - fWx) / k.delta + (fNx - fSx) / k.delta - srcX;
    R dQy = (fEy - fWy) / k.delta + (fNy - fSy) / k.delta - srcY;
    dEta = chop(dEta, k.eps); dQx = chop(dQx, k.eps); dQy = chop(dQy, k.eps);
    const R eta2 = eta - R(0.5) * dt * dEta, qx2 = qx - R(0.5) * dt * dQx, qy2 = qy - R(0.5) * dt * dQy;
    // re-extrapolate from the evolved state with the same slopes, :376-379
    f.n = FaceState<R>{eta2 + ph * syE, (eta2 - zb) + ph * syH, qx2 + ph * syQx, qy2 + ph * syQy};
    f.e = FaceState<R>{eta2 + ph * sxE, (eta2 - zb) + ph * sxH, qx2 + ph * sxQx, qy2 + ph * sxQy};
    f.s = FaceState<R>{eta2 + mh * syE, (eta2 - zb) + mh * syH, qx2 + mh * syQx, qy2 + mh * syQy};
    f.w = FaceState<R>{eta2 + mh * sxE, (eta2 - zb) + mh * sxH, qx2 + mh * sxQx, qy2 + mh * sxQy};
    return f;
}

// Corrector for one cell from its own four face estimates and the four facing estimates of its
// neighbours.  Reproduces mch_2nd_cacheNone after its loads,
// src/Schemes/CLSchemeMUSCLHancock.clc:596-800, with reconstructInterface() :1119-1230.
// `dryNeighbours` = number of neighbours whose eta_max < VERY_SMALL (:633-634).
// Returns false when the reference returns without writing (:638).
template <class R>
__device__ __forceinline__ bool mh_correct(const Params<R>& k, R dt, Cell<R>& c, R zb, R mann, const Faces<R>& own,
                                           const FaceState<R>& nS /* north cell's south face */,
                                           const FaceState<R>& eW, const FaceState<R>& sN, const FaceState<R>& wE,
                                           int dryNeighbours) {
    int dry = dryNeighbours;
    if (c.eta - zb < k.eps) ++dry;                                                                  // :596-597
    if (dry >= 5) return false;                                                                     // :638

    auto un = [&](const FaceState<R>& f, R q) { return f.h <= k.eps ? R(0) : q / f.h; };            // :1140-1150
    int stop = 0;
    R eN, bN, eE, bE, eS, bS, eW_, bW;
    // N: left = own north face, right = north cell's south face; normal = y
    const Flux3<R> fN = face_flux<R, true>(k, own.n.eta, own.n.eta - own.n.h, un(own.n, own.n.qy), un(own.n, own.n.qx),
                                           nS.eta, nS.eta - nS.h, un(nS, nS.qy), un(nS, nS.qx), c.qy, eN, bN, stop);
    const Flux3<R> fE = face_flux<R, true>(k, own.e.eta, own.e.eta - own.e.h, un(own.e, own.e.qx), un(own.e, own.e.qy),
                                           eW.eta, eW.eta - eW.h, un(eW, eW.qx), un(eW, eW.qy), c.qx, eE, bE, stop);
    const Flux3<R> fS = face_flux<R, false>(k, sN.eta, sN.eta - sN.h, un(sN, sN.qy), un(sN, sN.qx), own.s.eta,
                                            own.s.eta - own.s.h, un(own.s, own.s.qy), un(own.s, own.s.qx), c.qy, eS, bS,
                                            stop);
    const Flux3<R> fW = face_flux<R, false>(k, wE.eta, wE.eta - wE.h, un(wE, wE.qx), un(wE, wE.qy), own.w.eta,
                                            own.w.eta - own.w.h, un(own.w, own.w.qx), un(own.w, own.w.qy), c.qx, eW_, bW,
                                            stop);

    const R srcX = -1 * k.g * ((eE + eW_) / 2) * ((bE - bW) / k.delta);                             // :708
    const R srcY = -1 * k.g * ((eN + eS) / 2) * ((bN - bS) / k.delta);                              // :709
    R dEta = (fE.m - fW.m) / k.delta + (fN.m - fS.m) / k.delta - R(0);                              // :712-720
    R dQx = (fE.n - fW.n) / k.delta + (fN.t - fS.t) / k.delta - srcX;
    R dQy = (fE.t - fW.t) / k.delta + (fN.n - fS.n) / k.delta - srcY;
    dEta = chop(dEta, k.eps); dQx = chop(dQx, k.eps); dQy = chop(dQy, k.eps);                       // :723-731
    if (stop > 0) { c.qx = R(0); c.qy = R(0); }                                                     // :734-738
    c.eta = c.eta - dt * dEta; c.qx = c.qx - dt * dQx; c.qy = c.qy - dt * dQy;                      // :743-745
    if (k.friction) implicit_friction(k, c.eta, c.qx, c.qy, zb, mann, dt);                          // :779-789
    if (c.eta - zb < k.eps) c.eta = zb;                                                             // :792-793
    if (c.eta > c.emax && c.emax > R(-9990.0)) c.emax = c.eta;                                      // :796-797
    return true;
}

// ---------------------------------------------------------------------------------------------
// Partial-inertial flux; reproduces calculateInertialFlux(), src/Schemes/CLSchemeInertial.clc:335-378.
// ---------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ R inertial_flux(const Params<R>& k, R n, R dt, R prev, R etaUp, R zUp, R etaDown, R zDown) {
    const R froude = R(0.8);
    const R h = hp_fmax(etaDown, etaUp) - (zUp < zDown ? zDown : zUp);
    const R slope = (etaDown - etaUp) / k.delta;
    R q = (prev - (k.g * h * dt * slope)) /
          (R(1.0) + k.g * h * dt * n * n * hp_abs(prev) / hp_pow(h, R(10.0) / R(3.0)));
    if (q > R(0) && ((hp_abs(q) / h) / hp_sqrt(k.g * h)) > froude) q = h * hp_sqrt(k.g * h) * froude;
    if (q < R(0) && ((hp_abs(q) / h) / hp_sqrt(k.g * h)) > froude) q = R(0) - h * hp_sqrt(k.g * h) * froude;
    if (h < k.eps) q = R(0);
    return q;
}

#ifndef HP_FLAVOUR_STRICT
// Same flux with h^(-7/3) = (1/h)^2 * rcbrt(h) instead of pow(h, 10/3), one reciprocal for the
// quotient and the Froude limiter written as a clamp |q| <= 0.8 h sqrt(g h).
template <class R>
__device__ __forceinline__ R inertial_flux_fast(const Params<R>& k, R n, R dt, R prev, R etaUp, R zUp, R etaDown, R zDown, R inv_delta) {
    const R h = hp_fmax(etaDown, etaUp) - (zUp < zDown ? zDown : zUp);
    if (h < k.eps) return R(0);
    const R slope = (etaDown - etaUp) * inv_delta;
    const R rh = fm_rcp(h);
    const R gdt = k.g * dt;
    const R q = (prev - gdt * h * slope) * fm_rcp(R(1.0) + gdt * n * n * hp_abs(prev) * rh * rh * fm_rcbrt(h));
    const R qmax = R(0.8) * h * fm_sqrt(k.g * h);
    return q > qmax ? qmax : (q < -qmax ? -qmax : q);
}
#endif

// Reproduces ine_cacheDisabled after its loads, src/Schemes/CLSchemeInertial.clc:92-162.
template <class R>
__device__ __forceinline__ bool inertial_update(const Params<R>& k, R dt, Cell<R>& c, R zb, R mann, R etaN, R qyN, R zN,
                                                R etaE, R qxE, R zE, R etaS, R zS, R etaW, R zW) {
    int dry = 0;
    if (c.eta - zb < k.eps) ++dry;
    if (etaN - zN < k.eps) ++dry;
    if (etaE - zE < k.eps) ++dry;
    if (etaS - zS < k.eps) ++dry;
    if (etaW - zW < k.eps) ++dry;
    if (dry >= 5) return false;
#ifdef HP_FLAVOUR_STRICT
    const R qN = inertial_flux(k, mann, dt, qyN, etaN, zN, c.eta, zb);
    const R qE = inertial_flux(k, mann, dt, qxE, etaE, zE, c.eta, zb);
    const R qS = inertial_flux(k, mann, dt, c.qy, c.eta, zb, etaS, zS);
    const R qW = inertial_flux(k, mann, dt, c.qx, c.eta, zb, etaW, zW);
    c.qx = qW; c.qy = qS;
    const R dEta = (qE - qW + qN - qS) / k.delta;
#else
    const R inv_delta = fm_rcp(k.delta);
    const R qN = inertial_flux_fast(k, mann, dt, qyN, etaN, zN, c.eta, zb, inv_delta);
    const R qE = inertial_flux_fast(k, mann, dt, qxE, etaE, zE, c.eta, zb, inv_delta);
    const R qS = inertial_flux_fast(k, mann, dt, c.qy, c.eta, zb, etaS, zS, inv_delta);
    const R qW = inertial_flux_fast(k, mann, dt, c.qx, c.eta, zb, etaW, zW, inv_delta);
    c.qx = qW; c.qy = qS;
    const R dEta = (qE - qW + qN - qS) * inv_delta;
#endif
    c.eta = c.eta + dt * dEta;
    if (c.eta > c.emax) c.emax = c.eta;
    if (c.eta - zb < k.eps) c.eta = zb;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Time controller; reproduces tst_Advance_Normal, src/Schemes/CLDynamicTimestep.clc:27-146.
// `vmax` is the reduced maximum wave speed (the scan over TIMESTEP_WORKERS entries, :73-80).
// ---------------------------------------------------------------------------------------------
template <class R> __device__ __forceinline__ void advance_clock(const Params<R>& k, Clock<R>& ck, R vmax) {
    R t = ck.time, dt = hp_fmax(R(0), ck.timestep), th = ck.time_hydro;
    const R sync = ck.time_target;
    t += dt;
    ck.batch_timesteps += dt;
    if (dt > R(0)) ++ck.batch_successful; else ++ck.batch_skipped;
    if (th > R(1.0)) th = dt; else th += dt;                                           // :61-66
    if (k.dynamic) {
        R tmin = k.delta / vmax;                                                       // :84
        if (t < R(1.0) && tmin < R(1E-10)) tmin = R(1E-10);                            // :85-86
        dt = k.courant * tmin;                                                         // :89
    } else {
        dt = k.fixed_dt;                                                               // :94
    }
    if (dt > R(0) && dt < R(1E-10)) dt = R(1E-10);                                     // :112-113
    if ((t + dt) >= sync) {                                                            // :118-124
        if (sync - t > k.eps) dt = sync - t;
        if (sync - t <= k.eps) dt = -dt;
    }
    if (t < R(60.0) && dt > R(0.1)) dt = R(0.1);                                       // :128-129
    if ((t + dt) > k.end_time) dt = k.end_time - t;                                    // :132-133
    if (dt > R(15.0)) dt = R(15.0);                                                    // :136-137
    ck.time = t; ck.timestep = dt; ck.time_hydro = th;
}

// Reproduces tst_UpdateTimestep, src/Schemes/CLDynamicTimestep.clc:255-317.
template <class R> __device__ __forceinline__ void update_timestep_clock(const Params<R>& k, Clock<R>& ck, R vmax) {
    const R t = ck.time, original = hp_abs(ck.timestep), sync = ck.time_target;
    R dt = R(0);
    if (k.dynamic) {
        R tmin = k.delta / vmax;
        if (t < R(1.0) && tmin < R(1E-10)) tmin = R(1E-10);
        dt = k.courant * tmin;
    }
    dt = hp_fmin(dt, original);
    ck.batch_timesteps = ck.batch_timesteps - original + dt;
    if (t < R(60.0) && dt > R(0.1)) dt = R(0.1);
    if ((t + dt) >= sync) dt = hp_fmax(R(0), sync - t);
    if (dt > R(15.0)) dt = R(15.0);
    ck.timestep = dt;
}

}  // namespace HP_NS
