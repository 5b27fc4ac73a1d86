// hp_kernels.cu -- sm_100a kernels of the explicit shallow-water hot path.
//
// Compiled twice by build.py:
//   -DHP_NS=hp_strict -fmad=false   every operation rounded as written
//   -DHP_NS=hp_fast                 FMA contraction on
// Both flavours export a hp::KernelTable (strict_kernels() / fast_kernels()).
//
// Kernels in this file ("v1"): one thread per cell, neighbours through plain (read-only path)
// global loads.  The TMA-staged persistent variants live in hp_kernels_tma.cuh and share the
// per-cell arithmetic of hp_math.cuh.
//
// What one iteration launches (reference: src/Schemes/CSchemeGodunov.cpp:1617-1666):
//   reference                      here
//   bdy_* kernels                  bdy_* kernels (grid-stride, early exit on the device clock)
//   gts_/ine_/mch_1st+mch_2nd      ONE step kernel ...
//   tst_Reduce (2nd pass)          ... which also reduces the wave speed (warp shuffle ->
//   tst_Advance_Normal (1 item)        shared memory -> one atomicMax per CTA) and whose last
//                                      CTA runs the time controller
#ifndef HP_NS
#error "compile with -DHP_NS=hp_strict or -DHP_NS=hp_fast"
#endif

#include "hp_math.cuh"

namespace HP_NS {

using hp::BdyCellArgs;
using hp::BdyGriddedArgs;
using hp::BdyUniformArgs;
using hp::Grid;
using hp::ParamsD;
using hp::Planes;
using hp::StepArgs;

template <class R> __device__ __forceinline__ Params<R> make_params(const ParamsD& p) {
    Params<R> k;
    k.g = R(9.81);
    k.eps = static_cast<R>(p.eps); k.eps10 = static_cast<R>(p.eps10); k.delta = static_cast<R>(p.delta);
    k.courant = static_cast<R>(p.courant); k.end_time = static_cast<R>(p.end_time);
    k.fixed_dt = static_cast<R>(p.fixed_dt);
    k.dynamic = p.dynamic; k.friction = p.friction; k.simplified_speed = p.simplified_speed; k.dt0_keep = p.dt0_keep;
    return k;
}

// Non-negative reals keep their order when read as unsigned bit patterns.
__device__ __forceinline__ unsigned long long speed_bits(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long speed_bits(float v) { return (unsigned long long)__float_as_uint(v); }
template <class R> __device__ __forceinline__ R bits_speed(unsigned long long b);
template <> __device__ __forceinline__ double bits_speed<double>(unsigned long long b) { return __longlong_as_double((long long)b); }
template <> __device__ __forceinline__ float bits_speed<float>(unsigned long long b) { return __uint_as_float((unsigned int)b); }

template <class R> __device__ __forceinline__ R read_timestep(const void* clock) {
    return *reinterpret_cast<const volatile R*>(&reinterpret_cast<const Clock<R>*>(clock)->timestep);
}

// Stage 1 + 2 of the CFL reduction inside the step kernel: warp shuffle, shared memory, one
// atomicMax per CTA; the last CTA to arrive runs the time controller (tst_Advance_Normal) so the
// timestep never leaves the device and no second pass over the state is needed.
template <class R> __device__ __forceinline__ void block_reduce_finalize(R ws, const StepArgs& a, const Params<R>& k) {
    __shared__ R s_max[32];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int lane = tid & 31, wid = tid >> 5, nwarps = (nthreads + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const R other = __shfl_xor_sync(0xffffffffu, ws, o); ws = other > ws ? other : ws; }
    if (lane == 0) s_max[wid] = ws;
    __syncthreads();
    if (wid != 0) return;
    ws = lane < nwarps ? s_max[lane] : R(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const R other = __shfl_xor_sync(0xffffffffu, ws, o); ws = other > ws ? other : ws; }
    if (lane != 0) return;
    if (a.reduce_mode != hp::kReduceNone && ws > R(0)) atomicMax(a.max_bits, speed_bits(ws));
    if (!a.finalize) return;
    __threadfence();
    const unsigned int arrived = atomicAdd(a.ticket, 1u);
    if (arrived != a.total_ctas - 1u) return;
    __threadfence();
    const R vmax = bits_speed<R>(atomicExch(a.max_bits, 0ull));
    Clock<R>* ck = reinterpret_cast<Clock<R>*>(a.clock);
    Clock<R> c = *ck;
    advance_clock(k, c, vmax);
    *ck = c;
    *a.ticket = 0u;
}

template <class R> struct View {
    const R *eta, *emax, *qx, *qy;
    __device__ explicit View(const Planes& p)
        : eta(static_cast<const R*>(p.eta)), emax(static_cast<const R*>(p.emax)), qx(static_cast<const R*>(p.qx)),
          qy(static_cast<const R*>(p.qy)) {}
};
template <class R> struct MutView {
    R *eta, *emax, *qx, *qy;
    __device__ explicit MutView(const Planes& p)
        : eta(static_cast<R*>(p.eta)), emax(static_cast<R*>(p.emax)), qx(static_cast<R*>(p.qx)), qy(static_cast<R*>(p.qy)) {}
    __device__ __forceinline__ void store(size_t id, const Cell<R>& c) const {
        eta[id] = c.eta; emax[id] = c.emax; qx[id] = c.qx; qy[id] = c.qy;
    }
};

constexpr int kTileX = 32, kTileY = 8;

// ---------------------------------------------------------------------------------------------
// Godunov (SCHEME 0) and inertial (SCHEME 2) step, src -> dst.  Rows [y0, y1) of the local grid.
// Cells that the reference leaves unwritten stay unwritten.
// ---------------------------------------------------------------------------------------------
template <class R, int SCHEME>
__global__ void __launch_bounds__(kTileX* kTileY) step_pingpong_v1(const StepArgs a) {
    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = a.y0 + blockIdx.y * kTileY + threadIdx.y;
    R ws = R(0);
    if (x < g.cols && y < a.y1) {
        const R dt = read_timestep<R>(a.clock);
        const View<R> s(a.src);
        const MutView<R> d(a.dst);
        const R* __restrict__ bed = static_cast<const R*>(a.bed);
        const size_t id = static_cast<size_t>(y) * g.pitch + x;
        const int gy = y + g.gy0;
        Cell<R> c{s.eta[id], s.emax[id], s.qx[id], s.qy[id]};
        const R zb = bed[id];
        if (a.reduce_mode == hp::kReduceSrc) ws = wave_speed(k, c.eta, c.emax, c.qx, c.qy, zb);
        bool wrote = false;
        const bool interior = x >= 1 && x <= g.cols - 2 && gy >= 1 && gy <= g.grows - 2;  // frozen outer ring
        if (interior) {
            if (SCHEME == 0 && dt <= R(0)) {
                wrote = !k.dt0_keep;                                 // CLSchemeGodunov.clc:201-206 (:477-478 with the quirk)
            } else if (dt <= R(0)) {
                wrote = false;                                       // CLSchemeInertial.clc:60-61
            } else if (c.emax <= R(-9999.0) || c.eta == R(-9999.0)) {
                wrote = true;                                        // disabled cell: copied through
            } else {
                const size_t iN = id + g.pitch, iS = id - g.pitch, iE = id + 1, iW = id - 1;
                if (SCHEME == 0) {
                    wrote = godunov_update(k, dt, c, zb, static_cast<const R*>(a.manning)[id],
                                           s.eta[iN], s.qx[iN], s.qy[iN], bed[iN], s.eta[iE], s.qx[iE], s.qy[iE], bed[iE],
                                           s.eta[iS], s.qx[iS], s.qy[iS], bed[iS], s.eta[iW], s.qx[iW], s.qy[iW], bed[iW]);
                } else {
                    wrote = inertial_update(k, dt, c, zb, static_cast<const R*>(a.manning)[id], s.eta[iN], s.qy[iN], bed[iN],
                                            s.eta[iE], s.qx[iE], bed[iE], s.eta[iS], bed[iS], s.eta[iW], bed[iW]);
                }
            }
            if (wrote) d.store(id, c);
        }
        if (a.reduce_mode == hp::kReduceDst) {
            if (!wrote) { c.eta = d.eta[id]; c.emax = d.emax[id]; c.qx = d.qx[id]; c.qy = d.qy[id]; }
            ws = wave_speed(k, c.eta, c.emax, c.qx, c.qy, zb);
        }
    }
    block_reduce_finalize<R>(ws, a, k);
}

// ---------------------------------------------------------------------------------------------
// MUSCL-Hancock, predictor and corrector fused (SCHEME 1), src -> dst.
// Phase 1: the CTA evaluates the predictor for its tile plus one halo cell all round and keeps
// the 4 x {eta,h,qx,qy} face estimates in shared memory -- the reference writes them to four
// global buffers (128 B/cell) and reads eight back (src/Schemes/CLSchemeMUSCLHancock.clc:137-143,
// 600-635).  Phase 2: corrector.  The reference updates in place; with the fused kernel the halo
// reads of neighbouring CTAs make that unsafe, so the state ping-pongs and every owned cell is
// written (cells the reference leaves unchanged are copied through).
// ---------------------------------------------------------------------------------------------
constexpr int kMhX = kTileX + 2, kMhY = kTileY + 2;

template <class R>
__global__ void __launch_bounds__(kTileX* kTileY) step_mh_v1(const StepArgs a) {
    __shared__ R s_face[16][kMhY][kMhX];  // [face*4 + component][y][x]; faces N,E,S,W
    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const View<R> s(a.src);
    const R* __restrict__ bed = static_cast<const R*>(a.bed);
    const R dt = read_timestep<R>(a.clock);
    const int x0 = blockIdx.x * kTileX, y0 = a.y0 + blockIdx.y * kTileY;
    const int tid = threadIdx.y * kTileX + threadIdx.x;

    if (dt > R(0)) {
        for (int i = tid; i < kMhX * kMhY; i += kTileX * kTileY) {
            const int lx = i % kMhX, ly = i / kMhX;
            const int x = x0 + lx - 1, y = y0 + ly - 1, gy = y + g.gy0;
            // stage 1 runs on cells 1..N-2 (CLSchemeMUSCLHancock.clc:54-58) whose stencil we hold
            if (x < 1 || x > g.cols - 2 || gy < 1 || gy > g.grows - 2 || y < 1 || y > g.rows - 2) continue;
            const size_t id = static_cast<size_t>(y) * g.pitch + x;
            const size_t iN = id + g.pitch, iS = id - g.pitch, iE = id + 1, iW = id - 1;
            const Faces<R> f = mh_predict(k, dt, s.eta[id], s.qx[id], s.qy[id], bed[id],
                                          s.eta[iN], s.qx[iN], s.qy[iN], bed[iN], s.emax[iN],
                                          s.eta[iE], s.qx[iE], s.qy[iE], bed[iE], s.emax[iE],
                                          s.eta[iS], s.qx[iS], s.qy[iS], bed[iS], s.emax[iS],
                                          s.eta[iW], s.qx[iW], s.qy[iW], bed[iW], s.emax[iW]);
            s_face[0][ly][lx] = f.n.eta;  s_face[1][ly][lx] = f.n.h;  s_face[2][ly][lx] = f.n.qx;  s_face[3][ly][lx] = f.n.qy;
            s_face[4][ly][lx] = f.e.eta;  s_face[5][ly][lx] = f.e.h;  s_face[6][ly][lx] = f.e.qx;  s_face[7][ly][lx] = f.e.qy;
            s_face[8][ly][lx] = f.s.eta;  s_face[9][ly][lx] = f.s.h;  s_face[10][ly][lx] = f.s.qx; s_face[11][ly][lx] = f.s.qy;
            s_face[12][ly][lx] = f.w.eta; s_face[13][ly][lx] = f.w.h; s_face[14][ly][lx] = f.w.qx; s_face[15][ly][lx] = f.w.qy;
        }
    }
    __syncthreads();

    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    R ws = R(0);
    if (x < g.cols && y < a.y1) {
        const MutView<R> d(a.dst);
        const size_t id = static_cast<size_t>(y) * g.pitch + x;
        const int gy = y + g.gy0;
        Cell<R> c{s.eta[id], s.emax[id], s.qx[id], s.qy[id]};
        const R zb = bed[id];
        const bool interior = x >= 2 && x <= g.cols - 3 && gy >= 2 && gy <= g.grows - 3;  // ring of two is frozen
        if (interior && dt > R(0) && !(c.emax <= R(-9999.0) || c.eta == R(-9999.0))) {
            const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
            auto face = [&](int f, int yy, int xx) {
                return FaceState<R>{s_face[f * 4 + 0][yy][xx], s_face[f * 4 + 1][yy][xx], s_face[f * 4 + 2][yy][xx],
                                    s_face[f * 4 + 3][yy][xx]};
            };
            Faces<R> own{face(0, ly, lx), face(1, ly, lx), face(2, ly, lx), face(3, ly, lx)};
            int dryN = 0;  // neighbours are "dry" on eta_max, CLSchemeMUSCLHancock.clc:633-634
            if (s.emax[id + g.pitch] < k.eps) ++dryN;
            if (s.emax[id + 1] < k.eps) ++dryN;
            if (s.emax[id - g.pitch] < k.eps) ++dryN;
            if (s.emax[id - 1] < k.eps) ++dryN;
            mh_correct(k, dt, c, zb, static_cast<const R*>(a.manning)[id], own, face(2, ly + 1, lx), face(3, ly, lx + 1),
                       face(0, ly - 1, lx), face(1, ly, lx - 1), dryN);
        }
        d.store(id, c);
        if (a.reduce_mode != hp::kReduceNone) ws = wave_speed(k, c.eta, c.emax, c.qx, c.qy, zb);
    }
    block_reduce_finalize<R>(ws, a, k);
}

// ---------------------------------------------------------------------------------------------
// Stand-alone CFL pieces (sync points, multi-GPU split): tst_Reduce, tst_Advance_Normal,
// tst_UpdateTimestep (src/Schemes/CLDynamicTimestep.clc:27-146, 166-249, 255-317).
// ---------------------------------------------------------------------------------------------
template <class R> __global__ void __launch_bounds__(256) reduce_only_kernel(const StepArgs a) {
    const Params<R> k = make_params<R>(a.params);
    const Grid g = a.grid;
    const View<R> s(a.src);
    const R* __restrict__ bed = static_cast<const R*>(a.bed);
    R ws = R(0);
    const long long n = static_cast<long long>(a.y1 - a.y0) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = a.y0 + static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols);
        const size_t id = static_cast<size_t>(y) * g.pitch + x;
        const R v = wave_speed(k, s.eta[id], s.emax[id], s.qx[id], s.qy[id], bed[id]);
        ws = v > ws ? v : ws;
    }
    block_reduce_finalize<R>(ws, a, k);
}

template <class R, bool UPDATE_ONLY> __global__ void clock_kernel(const StepArgs a) {
    const Params<R> k = make_params<R>(a.params);
    const R vmax = bits_speed<R>(atomicExch(a.max_bits, 0ull));
    Clock<R>* ck = reinterpret_cast<Clock<R>*>(a.clock);
    Clock<R> c = *ck;
    if (UPDATE_ONLY) update_timestep_clock(k, c, vmax); else advance_clock(k, c, vmax);
    *ck = c;
}

// ---------------------------------------------------------------------------------------------
// Row strips over peer memory (see PeerBox in hp_kernels.cuh): push my edge rows into the neighbours' halo rows, publish
// my wave-speed maximum to every strip, wait for theirs, run the time controller.  Replaces the reference's
// CDomainLink::pullFromBuffer -> sendOverMPI -> pushToBuffer (src/Domain/Links/CDomainLink.cpp:168-270) and the
// MPI_Allreduce of CMPIManager::reduceTimeData (src/MPI/CMPIManager.cpp:837-889) -- and this library's own NCCL calls
// (hp_comm.cpp) -- by stores over NVLink from one kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void peer_hello_kernel(const hp::PeerArgs p) {
    const int r = threadIdx.x;
    if (r < p.world) { st_sys(&p.box[r]->hello[p.rank], 1ull); __threadfence_system(); }
}

// Stream-ordered barrier over all strips: what follows on this stream (the first cell update after an upload) starts only
// when every strip has finished what preceded it on its own stream (its upload) -- a neighbour's edge rows must not land
// in halo rows an upload is still going to overwrite.
__global__ void peer_barrier_kernel(const hp::PeerArgs p) {
    __shared__ unsigned long long s_epoch;
    hp::PeerBox* const mine = p.box[p.rank];
    if (threadIdx.x == 0) s_epoch = mine->bar_epoch + 1ull;
    __syncthreads();
    const unsigned long long e = s_epoch;
    const int r = threadIdx.x;
    if (r < p.world) {
        __threadfence_system();
        st_sys(&p.box[r]->bar[p.rank], e);
        const long long t0 = clock64();
        while (ld_sys(&mine->bar[r]) < e) {
            if (clock64() - t0 > hp::kPeerSpinCycles) { mine->error = 1u; break; }
            __nanosleep(64);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) mine->bar_epoch = e;
}

template <class R, bool UPDATE_ONLY> __global__ void __launch_bounds__(256) peer_exchange_kernel(const hp::PeerArgs p, const StepArgs a) {
    __shared__ unsigned int s_last;
    __shared__ unsigned long long s_it;
    hp::PeerBox* const mine = p.box[p.rank];
    // ---- halo rows: 16-byte stores straight into the neighbours' memory, all CTAs --------------------------------
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        const size_t vecs = p.bytes[n] / 16;
        if (vecs == 0) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4* __restrict__ s = static_cast<const uint4*>(p.src[n][q]);
            uint4* __restrict__ d = static_cast<uint4*>(p.dst[n][q]);
            for (size_t i = tid; i < vecs; i += nthreads) d[i] = s[i];
        }
    }
    __threadfence_system();                          // my stores are visible system-wide before this CTA counts as arrived
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int arrived = atomicAdd(&mine->ticket, 1u);
        s_last = arrived == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- the last CTA: publish, wait for the peers, advance the clock ------------------------------------------------
    __threadfence();                                 // everything the other CTAs stored before they arrived
    if (threadIdx.x == 0) {
        mine->ticket = 0u;
        s_it = mine->iter + 1ull;
    }
    __syncthreads();
    const unsigned long long it = s_it;
    const int r = threadIdx.x;
    if (r < p.world) {
        // (every lane reads the same local word; the exchange of lane 0 below resets it)
        const unsigned long long bits = *reinterpret_cast<volatile unsigned long long*>(a.max_bits);
        st_sys(&p.box[r]->vmax[it & 1ull][p.rank], bits);
        __threadfence_system();                      // halo rows (fenced above) and vmax before the signal
        st_sys(&p.box[r]->sig[p.rank], it);
        const long long t0 = clock64();
        while (ld_sys(&mine->sig[r]) < it) {
            if (clock64() - t0 > hp::kPeerSpinCycles) { mine->error = 1u; break; }
            __nanosleep(64);
        }
        __threadfence_system();                      // acquire: strip r's halo rows and vmax
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    unsigned long long best = 0ull;
    for (int q = 0; q < p.world; ++q) { const unsigned long long v = ld_sys(&mine->vmax[it & 1ull][q]); best = v > best ? v : best; }
    *a.max_bits = 0ull;
    const Params<R> k = make_params<R>(a.params);
    Clock<R>* ck = reinterpret_cast<Clock<R>*>(a.clock);
    Clock<R> c = *ck;
    if (UPDATE_ONLY) update_timestep_clock(k, c, bits_speed<R>(best)); else advance_clock(k, c, bits_speed<R>(best));
    *ck = c;
    mine->iter = it;
}

// ---------------------------------------------------------------------------------------------
// Boundary kernels; reproduce src/Boundaries/CLBoundaries.clc.  Grid-stride with a small fixed
// grid: most iterations they only read the device clock and leave (the hydrological
// accumulator gates rain to about once per simulated second).
// ---------------------------------------------------------------------------------------------
template <class R> __global__ void __launch_bounds__(256) bdy_uniform_kernel(const BdyUniformArgs a) {
    const Clock<R> ck = *reinterpret_cast<const Clock<R>*>(a.clock);
    const R acc = ck.time_hydro;
    if (acc < R(1.0) || ck.timestep <= R(0)) return;                       // CLBoundaries.clc:165-166
    if (ck.time >= static_cast<R>(a.length)) return;                       // :168
    unsigned long long step = static_cast<unsigned long long>(hp_floor(ck.time / static_cast<R>(a.interval)));
    // the reference trusts TimeseriesLength <= (entries - 1) * interval (only true for evenly spaced series) and reads
    // past its buffer otherwise; stay inside the series instead
    if (step >= a.entries) step = a.entries - 1;
    const R rate = static_cast<const R*>(a.series)[2 * step + 1];          // :172-173
    const Grid g = a.grid;
    R* __restrict__ eta = static_cast<R*>(a.state.eta);
    const R* __restrict__ emax = static_cast<const R*>(a.state.emax);
    const R* __restrict__ bed = static_cast<const R*>(a.bed);
    const int x1 = min(a.cover_x, g.cols - 1), gy1 = min(a.cover_y, g.grows - 1);
    const long long n = static_cast<long long>(g.rows) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols), gy = y + g.gy0;
        if (x < 1 || x >= x1 || gy < 1 || gy >= gy1) continue;             // :148-152 and launch coverage (Q6)
        const size_t id = static_cast<size_t>(y) * g.pitch + x;
        if (emax[id] <= R(-9999.0)) continue;                              // :168
        R e = eta[id];
        if (a.definition == 0u) e += rate / R(3600000.0) * acc;            // :176-177
        if (a.definition == 1u) { const R lowered = e - rate / R(3600000.0) * acc; const R z = bed[id]; e = z < lowered ? lowered : z; }  // :179-180
        eta[id] = e;
    }
}

template <class R> __global__ void __launch_bounds__(256) bdy_gridded_kernel(const BdyGriddedArgs a) {
    const Clock<R> ck = *reinterpret_cast<const Clock<R>*>(a.clock);
    const R acc = ck.time_hydro;
    if (acc < R(1.0)) return;                                              // CLBoundaries.clc:224-225
    unsigned long long step = static_cast<unsigned long long>(hp_floor(ck.time / static_cast<R>(a.interval)));
    if (step >= a.entries) step = a.entries;                               // :228-229
    const Grid g = a.grid;
    R* __restrict__ eta = static_cast<R*>(a.state.eta);
    const R* __restrict__ emax = static_cast<const R*>(a.state.emax);
    const R* __restrict__ series = static_cast<const R*>(a.series);
    const R delta = static_cast<R>(a.delta), res = static_cast<R>(a.resolution);
    const R offx = static_cast<R>(a.offset_x), offy = static_cast<R>(a.offset_y);
    const int x1 = min(a.cover_x, g.cols - 1), gy1 = min(a.cover_y, g.grows - 1);
    const long long n = static_cast<long long>(g.rows) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols), gy = y + g.gy0;
        if (x < 1 || x >= x1 || gy < 1 || gy >= gy1) continue;
        const size_t id = static_cast<size_t>(y) * g.pitch + x;
        R e = eta[id];
        if (emax[id] <= R(-9999.0) || e == R(-9999.0)) continue;          // :220-221
        const R col = hp_floor(((static_cast<R>(x) * delta) - offx) / res);    // :231-232
        const R row = hp_floor(((static_cast<R>(gy) * delta) - offy) / res);
        const unsigned long long cell = (a.grows * a.gcols) * step + (a.gcols * static_cast<unsigned long long>(row)) +
                                        static_cast<unsigned long long>(col);
        const R rate = series[cell];
        if (a.definition == 0ull) e += rate / R(3600000.0) * acc;          // :238-239
        if (a.definition == 2ull) e += rate / (delta * delta) * acc;       // :241-242
        eta[id] = e;
    }
}

template <class R> __global__ void __launch_bounds__(128) bdy_cell_kernel(const BdyCellArgs a) {
    const Params<R> k = make_params<R>(a.params);
    const Clock<R> ck = *reinterpret_cast<const Clock<R>*>(a.clock);
    const R t = ck.time, dt = ck.timestep;
    if (t >= static_cast<R>(a.length) || dt <= R(0)) return;              // CLBoundaries.clc:40-41
    const R interval = static_cast<R>(a.interval);
    unsigned long long base = static_cast<unsigned long long>(hp_floor(t / interval));         // :43-44
    if (base >= a.entries) base = a.entries - 1;   // unevenly spaced series: stay inside the buffer (entry `entries` is zero padding)
    const R* __restrict__ ts = static_cast<const R*>(a.series);
    const R w = hp_fmod(t, interval) / interval;                           // :52
    const R tsDepth = ts[4 * base + 1] + (ts[4 * base + 5] - ts[4 * base + 1]) * w;
    const R tsQx0 = ts[4 * base + 2] + (ts[4 * base + 6] - ts[4 * base + 2]) * w;
    const R tsQy0 = ts[4 * base + 3] + (ts[4 * base + 7] - ts[4 * base + 3]) * w;
    R* __restrict__ eta = static_cast<R*>(a.state.eta);
    R* __restrict__ qx = static_cast<R*>(a.state.qx);
    R* __restrict__ qy = static_cast<R*>(a.state.qy);
    const R* __restrict__ bed = static_cast<const R*>(a.bed);
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.count;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const long long id = a.relations[i];
        if (id < 0) continue;                                              // not held by this strip
        R tsQx = tsQx0, tsQy = tsQy0;
        R e = eta[id];
        const R zb = bed[id];
        if (a.def_depth == 2u) {                                           // :55-61
            e = zb + tsDepth;
        } else if (a.def_depth == 1u) {                                    // :62-68
            e = hp_fmax(zb, tsDepth);
        } else if (hp_abs(tsQx) > k.eps || hp_abs(tsQy) > k.eps || a.def_discharge == 3u) {    // :74-76
            R depth = (hp_abs(tsQx) * dt) / k.delta + (hp_abs(tsQy) * dt) / k.delta;            // :79
            R critical = hp_fmax(hp_pow(hp_pow(tsQx, R(2)) / k.g, R(1.0) / R(3.0)),
                                 hp_pow(hp_pow(tsQy, R(2)) / k.g, R(1.0) / R(3.0)));             // :81
            if (a.def_discharge == 3u) {                                   // :85-93
                depth = (hp_abs(tsQx) * dt) / (k.delta * k.delta);
                critical = R(0); tsQx = R(0); tsQy = R(0);
            }
            e = hp_fmax(zb + critical, e + depth);                         // :95
        }
        eta[id] = e;
        if (a.def_discharge == 1u) { qx[id] = tsQx; qy[id] = tsQy; }                             // :103-117
        else if (a.def_discharge == 2u) { qx[id] = tsQx * (e - zb); qy[id] = tsQy * (e - zb); }  // :108-121
    }
}

// ---------------------------------------------------------------------------------------------
// Host layout (array of {eta, eta_max, qx, qy}) <-> device planes.
// ---------------------------------------------------------------------------------------------
template <class R> struct Vec4T { R x, y, z, w; };

template <class R> __global__ void __launch_bounds__(256) aos_to_soa_kernel(const Vec4T<R>* __restrict__ aos, Planes p, Grid g,
                                                                             int row0, int nrows) {
    const MutView<R> d(p);
    const long long n = static_cast<long long>(nrows) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = row0 + static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols);
        const Vec4T<R> v = aos[i];
        d.store(static_cast<size_t>(y) * g.pitch + x, Cell<R>{v.x, v.y, v.z, v.w});
    }
}
template <class R> __global__ void __launch_bounds__(256) soa_to_aos_kernel(Planes p, Vec4T<R>* __restrict__ aos, Grid g, int row0,
                                                                             int nrows) {
    const View<R> s(p);
    const long long n = static_cast<long long>(nrows) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = row0 + static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols);
        const size_t id = static_cast<size_t>(y) * g.pitch + x;
        aos[i] = Vec4T<R>{s.eta[id], s.emax[id], s.qx[id], s.qy[id]};
    }
}
template <class R> __global__ void __launch_bounds__(256) copy_plane_rows_kernel(const R* __restrict__ dense, R* __restrict__ plane,
                                                                                  Grid g, int row0, int nrows) {
    const long long n = static_cast<long long>(nrows) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = row0 + static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols);
        plane[static_cast<size_t>(y) * g.pitch + x] = dense[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Output rasters: the per-cell derivation of CRasterDataset::domainToRaster
// (src/Datasets/CRasterDataset.cpp:180-267), in double like the reference's host loop, written
// in raster order (north row first, :270-280).  Value codes: src/Datasets/CRasterDataset.h:33-46.
// The executor always takes this kernel from the strict flavour (no FMA contraction), so the result
// is bit-identical to the host arithmetic of the reference.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double derive_output(int value, double eta, double emax, double qx, double qy, double bed, double res,
                                                double nodata) {
    const double depth = eta - bed;
    switch (value) {
    case 10: return (emax < bed + 1E-8 || bed > 9999.0) ? nodata : emax;                        // kMaxFSL :187-196
    case 2: return (eta < bed + 1E-8 || bed > 9999.0) ? nodata : eta;                           // kFreeSurfaceLevel :197-206
    case 9: { const double d = fmax(0.0, emax - bed); return (d < 1E-8 || d <= -9990.0 || d >= 9999.0) ? nodata : d; }   // kMaxDepth :207-213
    case 1: { const double d = fmax(0.0, depth); return d < 1E-8 ? nodata : d; }                // kDepth :214-220
    case 5: return qx * res;                                                                    // kDischargeX :221-226
    case 6: return qy * res;                                                                    // kDischargeY :227-232
    case 3: return depth > 1E-8 ? qx / depth : nodata;                                          // kVelocityX :233-242
    case 4: return depth > 1E-8 ? qy / depth : nodata;                                          // kVelocityY :243-252
    case 11: { const double u = qx / depth, v = qy / depth;                                     // kFroudeNumber :253-266
               return depth > 1E-8 ? sqrt(u * u + v * v) / sqrt(9.81 * depth) : nodata; }
    default: return nodata;                                                                     // :183 (row pre-filled with -9999)
    }
}

template <class R> __global__ void __launch_bounds__(256) derive_raster_kernel(Planes p, const R* __restrict__ bed, double* __restrict__ out,
                                                                                Grid g, int row_end, int nrows, int value, double res,
                                                                                double nodata) {
    const View<R> s(p);
    const long long n = static_cast<long long>(nrows) * g.cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / g.cols), x = static_cast<int>(i % g.cols);      // r = 0 is the northernmost row
        const size_t id = static_cast<size_t>(row_end - 1 - r) * g.pitch + x;
        out[i] = derive_output(value, static_cast<double>(s.eta[id]), static_cast<double>(s.emax[id]), static_cast<double>(s.qx[id]),
                               static_cast<double>(s.qy[id]), static_cast<double>(bed[id]), res, nodata);
    }
}

}  // namespace HP_NS

#ifndef HP_FLAVOUR_STRICT
#include "hp_fast_kernels.cuh"
#include "hp_march_kernels.cuh"
#include "hp_march_mh.cuh"
#include "hp_march_pair.cuh"
#endif

namespace HP_NS {

// ---------------------------------------------------------------------------------------------
// Launch interface
// ---------------------------------------------------------------------------------------------
#ifndef HP_FLAVOUR_STRICT
static_assert(sizeof(TmaMaps) == sizeof(hp::TmaMapsPOD), "descriptor block layout");
static_assert(Tile<double>::BW == hp::tma_box_w(8, 1) && Tile<float>::BW == hp::tma_box_w(4, 1) && Tile<double>::BH == hp::tma_box_h(1), "tile box");
static_assert(TileMH<double>::BW == hp::tma_box_w(8, 2) && TileMH<float>::BW == hp::tma_box_w(4, 2) && TileMH<double>::BH == hp::tma_box_h(2), "tile box");
static int launch_step_tma(int scheme, int real_bytes, const StepArgs& a, const hp::TmaMapsPOD* maps, int sm_count, cudaStream_t st) {
    const TmaMaps& m = *reinterpret_cast<const TmaMaps*>(maps);
    if (scheme == 0) return real_bytes == 8 ? launch_godunov_tma<double>(a, m, sm_count, st) : launch_godunov_tma<float>(a, m, sm_count, st);
    if (scheme == 1) return real_bytes == 8 ? launch_mh_tma<double>(a, m, sm_count, st) : launch_mh_tma<float>(a, m, sm_count, st);
    return -1;
}
static_assert(sizeof(TmaBlockMap) == sizeof(hp::TmaMaps6POD), "descriptor block layout");
// which (scheme, precision) pairs have a two-columns-per-lane kernel
// MUSCL-Hancock in fp64 stays on the one-column kernel: two columns need 228 registers, i.e. 8 warps per SM, and the
// fixed-latency stalls of so few warps cost more than the 15 % fewer instructions save (23.0 against 25.2 G
// cell-updates/s on the wet dam break; 22.4 G with 168 registers, 12 warps and spills) -- profiles/r02_wide_mh_f64.txt
#ifndef HP_WIDE_MH64
#define HP_WIDE_MH64 0
#endif
#ifndef HP_WIDE_MH32
#define HP_WIDE_MH32 1
#endif
// mode: 0 = the default kernel of (scheme, precision), 1 = one column per lane, 2 = two columns per lane wherever such a
// kernel exists
// Godunov (only with HP_OPT_MARCH_GODUNOV; the default kernel of the scheme is the tile kernel): in fp64 two columns need
// 218 registers -- 168 with spills at 12 warps per SM run at 32.2 G cell-updates/s on the 4096^2 dam break where the
// one-column kernel reaches 35.7 G; in fp32 (121 registers, 16 warps) the two-column kernel is the faster one, 51.6 against
// 46.5 G (tiles: 48.1 G)
#ifndef HP_WIDE_GOD64
#define HP_WIDE_GOD64 0
#endif
#ifndef HP_WIDE_GOD32
#define HP_WIDE_GOD32 1
#endif
static bool use_wide(int scheme, int real_bytes, int mode) {
    if (mode == 1 || scheme < 0 || scheme > 2) return false;
    if (mode == 2 || scheme == 2) return true;
    if (scheme == 0) return real_bytes == 8 ? HP_WIDE_GOD64 != 0 : HP_WIDE_GOD32 != 0;
    return real_bytes == 8 ? HP_WIDE_MH64 != 0 : HP_WIDE_MH32 != 0;
}
static int march_box_w(int scheme, int real_bytes, int mode) {
    return use_wide(scheme, real_bytes, mode) ? hp::wide_box_w(real_bytes) : hp::march_box_w(real_bytes, 1);
}
static int launch_step_march(int scheme, int real_bytes, const StepArgs& a, const hp::TmaMaps6POD* maps, int alt_bits, int sm_count, cudaStream_t st) {
    const TmaBlockMap& m = *reinterpret_cast<const TmaBlockMap*>(maps);
    const int alt = alt_bits & 1;
    if (use_wide(scheme, real_bytes, (alt_bits >> 1) & 3)) {
        if (scheme == 0) return real_bytes == 8 ? launch_godunov_wide<double>(a, m, alt, sm_count, st) : launch_godunov_wide<float>(a, m, alt, sm_count, st);
        if (scheme == 1) return real_bytes == 8 ? launch_mh_wide<double>(a, m, alt, sm_count, st) : launch_mh_wide<float>(a, m, alt, sm_count, st);
        if (scheme == 2) return real_bytes == 8 ? launch_inertial_wide<double>(a, m, alt, sm_count, st) : launch_inertial_wide<float>(a, m, alt, sm_count, st);
    }
    if (scheme == 0) return real_bytes == 8 ? launch_godunov_march<double>(a, m, alt, sm_count, st) : launch_godunov_march<float>(a, m, alt, sm_count, st);
    if (scheme == 1) return real_bytes == 8 ? launch_mh_march2<double>(a, m, alt, sm_count, st) : launch_mh_march2<float>(a, m, alt, sm_count, st);
    if (scheme == 2) return real_bytes == 8 ? launch_inertial_march<double>(a, m, alt, sm_count, st) : launch_inertial_march<float>(a, m, alt, sm_count, st);
    return -1;
}
#endif

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}
static int stride_grid(long long n, int threads, int ctas_per_sm) {
    const long long want = (n + threads - 1) / threads;
    const long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;   // multiples of the SM count
    return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

static int launch_step(int scheme, int real_bytes, const StepArgs& a_in, cudaStream_t st) {
    StepArgs a = a_in;
    const dim3 block(kTileX, kTileY);
    const dim3 grid((a.grid.cols + kTileX - 1) / kTileX, (a.y1 - a.y0 + kTileY - 1) / kTileY);
    if (a.y1 <= a.y0) return 0;
    a.total_ctas = grid.x * grid.y;
    if (real_bytes == 8) {
        if (scheme == 0) step_pingpong_v1<double, 0><<<grid, block, 0, st>>>(a);
        else if (scheme == 2) step_pingpong_v1<double, 2><<<grid, block, 0, st>>>(a);
        else step_mh_v1<double><<<grid, block, 0, st>>>(a);
    } else {
        if (scheme == 0) step_pingpong_v1<float, 0><<<grid, block, 0, st>>>(a);
        else if (scheme == 2) step_pingpong_v1<float, 2><<<grid, block, 0, st>>>(a);
        else step_mh_v1<float><<<grid, block, 0, st>>>(a);
    }
    return 1;
}

static int launch_reduce_only(int real_bytes, const StepArgs& a_in, cudaStream_t st) {
    StepArgs a = a_in;
    const long long n = static_cast<long long>(a.y1 - a.y0) * a.grid.cols;
    const int grid = stride_grid(n, 256, 8);
    a.total_ctas = grid; a.finalize = 0; a.reduce_mode = hp::kReduceSrc;
    if (real_bytes == 8) reduce_only_kernel<double><<<grid, dim3(256, 1), 0, st>>>(a);
    else reduce_only_kernel<float><<<grid, dim3(256, 1), 0, st>>>(a);
    return 1;
}
static int launch_advance(int real_bytes, const StepArgs& a, cudaStream_t st) {
    if (real_bytes == 8) clock_kernel<double, false><<<1, 1, 0, st>>>(a); else clock_kernel<float, false><<<1, 1, 0, st>>>(a);
    return 1;
}
static int launch_update_timestep(int real_bytes, const StepArgs& a, cudaStream_t st) {
    if (real_bytes == 8) clock_kernel<double, true><<<1, 1, 0, st>>>(a); else clock_kernel<float, true><<<1, 1, 0, st>>>(a);
    return 1;
}
static int launch_peer_exchange(int real_bytes, const hp::PeerArgs& p, const StepArgs& a, int update_only, cudaStream_t st) {
    // enough CTAs to keep the links busy (a few MB at most), never more than the copy has 16-byte pieces for
    const unsigned long long vecs = (p.bytes[0] + p.bytes[1]) / 16 * 4;
    int grid = static_cast<int>((vecs + 1023) / 1024);
    grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
    if (real_bytes == 8) {
        if (update_only) peer_exchange_kernel<double, true><<<grid, 256, 0, st>>>(p, a); else peer_exchange_kernel<double, false><<<grid, 256, 0, st>>>(p, a);
    } else {
        if (update_only) peer_exchange_kernel<float, true><<<grid, 256, 0, st>>>(p, a); else peer_exchange_kernel<float, false><<<grid, 256, 0, st>>>(p, a);
    }
    return 1;
}
static int launch_peer_hello(const hp::PeerArgs& p, cudaStream_t st) { peer_hello_kernel<<<1, 32, 0, st>>>(p); return 1; }
static int launch_peer_barrier(const hp::PeerArgs& p, cudaStream_t st) { peer_barrier_kernel<<<1, 32, 0, st>>>(p); return 1; }
static int launch_bdy_uniform(int real_bytes, const BdyUniformArgs& a, cudaStream_t st) {
    const int grid = stride_grid(static_cast<long long>(a.grid.rows) * a.grid.cols, 256, 8);
    if (real_bytes == 8) bdy_uniform_kernel<double><<<grid, 256, 0, st>>>(a); else bdy_uniform_kernel<float><<<grid, 256, 0, st>>>(a);
    return 1;
}
static int launch_bdy_gridded(int real_bytes, const BdyGriddedArgs& a, cudaStream_t st) {
    const int grid = stride_grid(static_cast<long long>(a.grid.rows) * a.grid.cols, 256, 8);
    if (real_bytes == 8) bdy_gridded_kernel<double><<<grid, 256, 0, st>>>(a); else bdy_gridded_kernel<float><<<grid, 256, 0, st>>>(a);
    return 1;
}
static int launch_bdy_cell(int real_bytes, const BdyCellArgs& a, cudaStream_t st) {
    if (a.count == 0) return 0;
    const int grid = stride_grid(static_cast<long long>(a.count), 128, 4);
    if (real_bytes == 8) bdy_cell_kernel<double><<<grid, 128, 0, st>>>(a); else bdy_cell_kernel<float><<<grid, 128, 0, st>>>(a);
    return 1;
}
static int launch_aos_to_soa(int real_bytes, const void* aos, Planes dst, Grid g, int row0, int nrows, cudaStream_t st) {
    const int grid = stride_grid(static_cast<long long>(nrows) * g.cols, 256, 8);
    if (real_bytes == 8) aos_to_soa_kernel<double><<<grid, 256, 0, st>>>(static_cast<const Vec4T<double>*>(aos), dst, g, row0, nrows);
    else aos_to_soa_kernel<float><<<grid, 256, 0, st>>>(static_cast<const Vec4T<float>*>(aos), dst, g, row0, nrows);
    return 1;
}
static int launch_soa_to_aos(int real_bytes, Planes src, void* aos, Grid g, int row0, int nrows, cudaStream_t st) {
    const int grid = stride_grid(static_cast<long long>(nrows) * g.cols, 256, 8);
    if (real_bytes == 8) soa_to_aos_kernel<double><<<grid, 256, 0, st>>>(src, static_cast<Vec4T<double>*>(aos), g, row0, nrows);
    else soa_to_aos_kernel<float><<<grid, 256, 0, st>>>(src, static_cast<Vec4T<float>*>(aos), g, row0, nrows);
    return 1;
}
static int launch_copy_plane_rows(int real_bytes, const void* dense, void* plane, Grid g, int row0, int nrows, cudaStream_t st) {
    const int grid = stride_grid(static_cast<long long>(nrows) * g.cols, 256, 8);
    if (real_bytes == 8) copy_plane_rows_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(dense), static_cast<double*>(plane), g, row0, nrows);
    else copy_plane_rows_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dense), static_cast<float*>(plane), g, row0, nrows);
    return 1;
}

static int launch_derive_raster(int real_bytes, Planes src, const void* bed, double* out, Grid g, int row_end, int nrows, int value,
                                double resolution, double nodata, cudaStream_t st) {
    const int grid = stride_grid(static_cast<long long>(nrows) * g.cols, 256, 8);
    if (real_bytes == 8) derive_raster_kernel<double><<<grid, 256, 0, st>>>(src, static_cast<const double*>(bed), out, g, row_end, nrows, value, resolution, nodata);
    else derive_raster_kernel<float><<<grid, 256, 0, st>>>(src, static_cast<const float*>(bed), out, g, row_end, nrows, value, resolution, nodata);
    return 1;
}

static const hp::KernelTable g_table = {
#ifndef HP_FLAVOUR_STRICT
    launch_step_tma, launch_step_march, march_box_w,
#else
    nullptr, nullptr, nullptr,
#endif
    launch_step, launch_reduce_only, launch_advance, launch_update_timestep, launch_peer_exchange, launch_peer_hello, launch_peer_barrier, launch_bdy_uniform, launch_bdy_gridded,
    launch_bdy_cell, launch_aos_to_soa, launch_soa_to_aos, launch_copy_plane_rows, launch_derive_raster,
};

}  // namespace HP_NS

namespace hp {
#define HP_CAT2(a, b) a##b
#define HP_CAT(a, b) HP_CAT2(a, b)
#ifdef HP_FLAVOUR_STRICT
const KernelTable& strict_kernels() { return HP_NS::g_table; }
#else
const KernelTable& fast_kernels() { return HP_NS::g_table; }
#endif
}  // namespace hp
