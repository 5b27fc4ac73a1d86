/*
 * hpo_api.h -- C interface shared by the two CPU checkers under oracle/:
 *
 *   liboracle.so            our own restatement of the reference kernels
 *                           (oracle/hipims_oracle.cpp), both precisions.
 *   _ref/ref_<variant>.so   the reference's own .clc/.clh sources compiled
 *                           through oracle/ref_shim/cl_shim.h (one scheme
 *                           program, one precision, per library).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (hipims_ocl_b200/,
 * include/) may include, link or call anything declared here; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Array layouts are the reference's host layouts (src/Domain/CDomain.h:28-33,
 * src/Schemes/CSchemeGodunov.cpp:832-845): cell state is an array of
 * {eta, eta_max, qx, qy} 4-vectors of `real`, bed and manning are arrays of
 * `real`; row-major with row 0 the southern edge
 * (src/Domain/Cartesian/CLDomainCartesian.clc:26-30).
 */
#ifndef HPO_API_H
#define HPO_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { HPO_SCHEME_GODUNOV = 0, HPO_SCHEME_MUSCL_HANCOCK = 1, HPO_SCHEME_INERTIAL = 2 };

/* Reference quirks that the driver can reproduce. */
enum {
    HPO_QUIRK_REDUCE_BUFFER_A   = 1u << 0, /* Q1: tst_Reduce always reads "Cell states"      */
    HPO_QUIRK_BDY_COVERAGE      = 1u << 1, /* Q6: bdy_Uniform/Gridded global = floor(n/8)*8   */
    HPO_QUIRK_MH_NO_BOUNDARIES  = 1u << 2, /* Q4: MUSCL-Hancock never applies boundaries      */
    HPO_QUIRK_GODUNOV_DT0_KEEP  = 1u << 3  /* gts_cacheEnabled's rule for dt <= 0: no write   */
};

typedef struct hpo_config {
    int64_t  cols, rows;
    double   delta;          /* DOMAIN_DELTAX == DOMAIN_DELTAY                      */
    double   very_small;     /* VERY_SMALL (dryThreshold)                           */
    double   quite_small;    /* QUITE_SMALL = very_small * 10 (computed in double)  */
    double   courant;        /* COURANT_NUMBER                                      */
    double   end_time;       /* SCHEME_ENDTIME                                      */
    double   fixed_dt;       /* TIMESTEP_FIXED when dynamic == 0                    */
    double   initial_dt;     /* first timestep (reference default 0.001)            */
    int32_t  scheme;         /* HPO_SCHEME_*                                        */
    int32_t  dynamic;        /* TIMESTEP_DYNAMIC                                    */
    int32_t  friction;       /* FRICTION_ENABLED && FRICTION_IN_FLUX_KERNEL         */
    uint32_t quirks;         /* HPO_QUIRK_*                                         */
    int32_t  threads;        /* OpenMP threads, 0 = runtime default                 */
    int32_t  real_bytes;     /* 4 or 8 (ref libraries only accept their own)        */
} hpo_config;

typedef struct hpo_stats {
    double   time, timestep, time_hydro, time_target, batch_timesteps;
    uint32_t batch_successful, batch_skipped;
    uint32_t use_alternate;  /* 1 when the next source buffer is "Cell states (alternate)" */
    uint32_t pad;
} hpo_stats;

/* bdy kinds and definitions follow src/Boundaries/CLBoundaries.clh:31-52 */
typedef struct hpo_bdy_uniform {
    uint32_t entries; uint32_t definition;     /* 0 rain-intensity, 1 loss-rate */
    double   interval, length;
} hpo_bdy_uniform;

typedef struct hpo_bdy_gridded {
    double   interval, resolution, offset_x, offset_y;
    uint64_t entries, definition, rows, cols;  /* 0 rain-intensity, 2 mass-flux */
} hpo_bdy_gridded;

typedef struct hpo_bdy_cell {
    uint64_t entries;
    double   interval, length;
    uint64_t relations;
    uint32_t def_depth, def_discharge;
} hpo_bdy_cell;

#define HPO_DECLARE(P)                                                                          \
    void*  P##create(const hpo_config* cfg);                                                    \
    void   P##destroy(void* sim);                                                               \
    void   P##upload(void* sim, const void* states, const void* bed, const void* manning);      \
    void   P##download(void* sim, void* states);                                                \
    void   P##download_both(void* sim, void* states_a, void* states_b);                         \
    void   P##set_target(void* sim, double t);                                                  \
    void   P##set_clock(void* sim, double time, double timestep, double time_hydro);            \
    int    P##add_uniform(void* sim, const hpo_bdy_uniform* c, const double* series_tv);        \
    int    P##add_gridded(void* sim, const hpo_bdy_gridded* c, const double* series);           \
    int    P##add_cell(void* sim, const hpo_bdy_cell* c, const uint64_t* relations,             \
                       const double* series_tdxy);                                              \
    void   P##iterate(void* sim, int n);                                                        \
    void   P##update_timestep(void* sim);                                                       \
    void   P##reset_counters(void* sim);                                                        \
    void   P##stats(void* sim, hpo_stats* out);                                                 \
    /* single kernels on caller-owned arrays, for kernel-by-kernel cross checks */              \
    void   P##k_gts(const hpo_config*, const void* dt, const void* bed, const void* src,        \
                    void* dst, const void* manning);                                            \
    void   P##k_ine(const hpo_config*, const void* dt, const void* bed, const void* src,        \
                    void* dst, const void* manning);                                            \
    void   P##k_mch_1st(const hpo_config*, const void* dt, const void* bed, const void* state,  \
                        void* fN, void* fE, void* fS, void* fW);                                \
    void   P##k_mch_2nd(const hpo_config*, const void* dt, void* state, const void* bed,        \
                        const void* manning, const void* fN, const void* fE, const void* fS,    \
                        const void* fW);                                                        \
    double P##k_reduce(const hpo_config*, const void* state, const void* bed);                  \
    /* tst_Advance_Normal on a caller-owned clock {time, timestep, time_hydro, target, batch}  \
       (reals) + {successful, skipped} (uint32) with an already reduced maximum wave speed */   \
    void   P##k_advance(const hpo_config*, void* clock_reals5, uint32_t* counters2, double vmax);

HPO_DECLARE(hpo_f64_)
HPO_DECLARE(hpo_f32_)
HPO_DECLARE(hpo_ref_)

#ifdef __cplusplus
}
#endif
#endif
