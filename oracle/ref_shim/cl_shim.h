/*
 * cl_shim.h -- just enough OpenCL C 1.1 on top of C++17 for the reference's kernel
 * sources (/root/reference/src/**.clc, .clh) to compile unmodified with g++ and run
 * serially / under OpenMP on the host.  TEST INFRASTRUCTURE ONLY (see ../hpo_api.h).
 *
 * oracle/build_ref.py prefixes this header to the concatenated reference sources (in
 * the order the reference's prepareCode() uses) after three mechanical rewrites:
 *   (T)(a, b, ...)  vector constructors            -> shim_make<T>(a, b, ...)
 *   __attribute__((reqd_work_group_size(..)))      -> nothing
 * Nothing of the reference is stored in this repository; the generated translation unit
 * lives in a temp directory and only the resulting library lands in oracle/_ref/.
 *
 * The JIT "#define" surface is mapped onto globals so that one
 * library serves any domain size: DOMAIN_COLS -> shim_cols etc.  Preprocessor switches
 * (TIMESTEP_DYNAMIC | TIMESTEP_FIXED, FRICTION_ENABLED) select the library variant.
 */
#ifndef HPO_CL_SHIM_H
#define HPO_CL_SHIM_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>

#ifndef SHIM_REAL
#define SHIM_REAL double
#endif

/* ---- scalar aliases ------------------------------------------------------------------- */
/* OpenCL long/ulong are 64-bit, as are LP64 long/unsigned long: no remapping needed.
 * (glibc's <sys/types.h> already provides identical ushort/uint/ulong typedefs.) */
typedef unsigned char uchar;
typedef unsigned short int ushort;
typedef unsigned int uint;
typedef unsigned long int ulong;
static_assert(sizeof(long) == 8, "LP64 expected");

/* ---- vectors: plain aggregates so both (T){a,b} and shim_make<T>(a,b) work ------------- */
template <class T> struct shim_v2 {
    union { struct { T x, y; }; struct { T S0, S1; }; struct { T s0, s1; }; T s[2]; };
};
template <class T> struct shim_v4 {
    union { struct { T x, y, z, w; }; struct { T S0, S1, S2, S3; }; struct { T s0, s1, s2, s3; }; T s[4]; };
};
template <class T> struct shim_v8 {
    union { struct { T S0, S1, S2, S3, S4, S5, S6, S7; }; struct { T s0, s1, s2, s3, s4, s5, s6, s7; }; T s[8]; };
};

#define SHIM_VEC_OPS(V, N)                                                                               \
    template <class T> inline V<T> operator+(const V<T>& a, const V<T>& b) {                             \
        V<T> r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b.s[i]; return r; }                        \
    template <class T> inline V<T> operator-(const V<T>& a, const V<T>& b) {                             \
        V<T> r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b.s[i]; return r; }                        \
    template <class T> inline V<T> operator*(const V<T>& a, const V<T>& b) {                             \
        V<T> r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b.s[i]; return r; }                        \
    template <class T, class S> inline V<T> operator*(S k, const V<T>& a) {                              \
        V<T> r; for (int i = 0; i < N; ++i) r.s[i] = static_cast<T>(k) * a.s[i]; return r; }             \
    template <class T, class S> inline V<T> operator*(const V<T>& a, S k) {                              \
        V<T> r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * static_cast<T>(k); return r; }
SHIM_VEC_OPS(shim_v2, 2)
SHIM_VEC_OPS(shim_v4, 4)
SHIM_VEC_OPS(shim_v8, 8)

template <class V, class... A> inline V shim_make(A... a) {
    V v; typedef decltype(v.s[0] + 0) elem_t;
    const elem_t tmp[] = {static_cast<elem_t>(a)...};
    for (unsigned i = 0; i < sizeof...(A); ++i) v.s[i] = tmp[i];
    return v;
}

#define SHIM_VEC_TYPES(base, T)                                                                          \
    typedef shim_v2<T> base##2; typedef shim_v4<T> base##4; typedef shim_v8<T> base##8;
SHIM_VEC_TYPES(int, int)
SHIM_VEC_TYPES(uint, unsigned int)
typedef shim_v2<int64_t> long2;  typedef shim_v4<int64_t> long4;  typedef shim_v8<int64_t> long8;
typedef shim_v2<uint64_t> ulong2; typedef shim_v4<uint64_t> ulong4; typedef shim_v8<uint64_t> ulong8;
SHIM_VEC_TYPES(float, float)
SHIM_VEC_TYPES(double, double)

/* what COCLProgram::getExtensionHeader emits (src/OpenCL/Executors/COCLProgram.cpp:381-399) */
typedef SHIM_REAL cl_double;
typedef shim_v2<SHIM_REAL> cl_double2;
typedef shim_v4<SHIM_REAL> cl_double4;
typedef shim_v8<SHIM_REAL> cl_double8;

/* ---- address spaces / qualifiers ------------------------------------------------------- */
#define __kernel
#define __global
#define __constant
#define __private
/* Work-group local memory: one copy per host thread, alive across the work-items a thread runs one after the other
 * (every __local in the reference is a function-scope array). */
#define __local static thread_local
#define restrict __restrict__
#define CLK_LOCAL_MEM_FENCE 0
/* Barriers.  The kernels normally run with work-groups of ONE item, where a barrier is nothing.  The kernels that stage
 * a tile in local memory (gts_cacheEnabled, ine_cacheEnabled: one barrier, and everything before it only LOADS into local
 * memory) are run a work-group at a time in two passes (ref_unit.inc: ndrange2_tiles): pass 1 takes every work-item up
 * to the barrier, where it leaves the kernel; pass 2 runs every work-item from the top with the barrier a no-op -- the
 * repeated loads store the same values again. */
extern thread_local int shim_barrier_leaves;
struct shim_barrier_hit {};
inline void barrier(int) { if (shim_barrier_leaves) throw shim_barrier_hit(); }

/* ---- work-item functions: one work-item per call; work-groups of one unless ndrange2_tiles says otherwise ---------- */
extern thread_local int64_t shim_gid[3], shim_lid[3], shim_lsize[3], shim_grp[3];
extern thread_local int shim_tiled;          /* inside ndrange2_tiles: real work-groups */
extern int64_t shim_gsize[3];
inline int64_t get_global_id(int d) { return shim_gid[d]; }
inline int64_t get_global_size(int d) { return shim_gsize[d]; }
inline int64_t get_local_id(int d) { return shim_lid[d]; }
inline int64_t get_local_size(int d) { return shim_lsize[d]; }
inline int64_t get_group_id(int d) { return shim_tiled ? shim_grp[d] : shim_gid[d]; }

/* ---- maths built-ins -------------------------------------------------------------------- */
using std::sqrt; using std::pow; using std::fabs; using std::floor; using std::fmod;
using std::fmax; using std::fmin; using std::max; using std::min; using std::trunc;
inline SHIM_REAL pown(SHIM_REAL v, int n) { SHIM_REAL r = 1; for (int i = 0; i < n; ++i) r *= v; return r; }

/* ---- JIT constants -> globals (names on the right are defined in ref_unit.inc) ---------- */
extern int64_t shim_cols, shim_rows, shim_cells;
extern unsigned int shim_workers;
extern cl_double shim_delta, shim_very_small, shim_quite_small, shim_courant, shim_endtime, shim_fixed_dt;
#define DOMAIN_COLS shim_cols
#define DOMAIN_ROWS shim_rows
#define DOMAIN_CELLCOUNT shim_cells
#define DOMAIN_DELTAX shim_delta
#define DOMAIN_DELTAY shim_delta
#define VERY_SMALL shim_very_small
#define QUITE_SMALL shim_quite_small
#define COURANT_NUMBER shim_courant
#define SCHEME_ENDTIME shim_endtime
#define SCHEME_OUTPUTTIME shim_endtime
#define TIMESTEP_WORKERS shim_workers
#define TIMESTEP_GROUPSIZE 1
#ifdef SHIM_TIMESTEP_FIXED
#define TIMESTEP_FIXED shim_fixed_dt
#else
#define TIMESTEP_DYNAMIC 1
#endif
#ifdef SHIM_FRICTION
#define FRICTION_ENABLED 1
#define FRICTION_IN_FLUX_KERNEL 1
#endif
#define REQD_WG_SIZE_FULL_TS
#define REQD_WG_SIZE_HALF_TS
#define REQD_WG_SIZE_LINE
/* local-tile extents of the cached kernel variants (the work-group size ndrange2_tiles runs them with) */
#define GTS_DIM1 16
#define GTS_DIM2 16
#define MCH_STG1_DIM1 16
#define MCH_STG1_DIM2 16
#define INE_DIM1 16
#define INE_DIM2 16
#define MEM_SEPARATE_FACES 1

#endif
