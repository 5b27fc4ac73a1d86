/*
 * sim_driver.h -- host-side sequencing shared by both CPU checkers (TEST
 * INFRASTRUCTURE ONLY, see hpo_api.h).
 *
 * Restates, in our own words, what the reference's host does around the
 * kernels on the hot path:
 *   - buffer set           src/Schemes/CSchemeGodunov.cpp:789-893  (two state buffers
 *                          uploaded with the same initial data, :1060-1061), MH face
 *                          buffers src/Schemes/CSchemeMUSCLHancock.cpp:468-495
 *   - reduction geometry   src/Schemes/CSchemeGodunov.cpp:656-658
 *   - one iteration        src/Schemes/CSchemeGodunov.cpp:1617-1666 (Godunov, inertial)
 *                          src/Schemes/CSchemeMUSCLHancock.cpp:646-680 (MUSCL-Hancock)
 *   - ping-pong toggling   src/Schemes/CSchemeGodunov.cpp:1287-1301
 *   - boundary launches    src/Boundaries/CBoundaryUniform.cpp:294-295 etc.
 *
 * The class is a template over a `Kernels` policy so that the very same
 * sequencing drives our restated kernels (hipims_oracle.cpp) and the reference's
 * own kernel sources compiled through the shim (ref_shim/ref_unit.inc).
 */
#ifndef HPO_SIM_DRIVER_H
#define HPO_SIM_DRIVER_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "hpo_api.h"

namespace hpo {

template <class R> struct Vec4 { R x, y, z, w; };

/* Device-side configuration records, field order as src/Boundaries/CLBoundaries.clh:54-82. */
template <class R> struct BdyCellConf {
    uint64_t TimeseriesEntries; R TimeseriesInterval; R TimeseriesLength;
    uint64_t RelationCount; uint32_t DefinitionDepth; uint32_t DefinitionDischarge;
};
template <class R> struct BdyGriddedConf {
    R TimeseriesInterval, GridResolution, GridOffsetX, GridOffsetY;
    uint64_t TimeseriesEntries, Definition, GridRows, GridCols;
};
template <class R> struct BdyUniformConf {
    uint32_t TimeseriesEntries; R TimeseriesInterval; R TimeseriesLength; uint32_t Definition;
};

inline size_t reduction_workers(int64_t cells) {
    /* ceil((cells / 200) / 512) * 512, src/Schemes/CSchemeGodunov.cpp:60,656-658 */
    const double wavefronts = 200.0, wg = 512.0;
    return static_cast<size_t>(std::ceil((static_cast<double>(cells) / wavefronts) / wg) * wg);
}

template <class R, class Kernels> class Sim {
  public:
    explicit Sim(const hpo_config& c) : cfg(c) {
        cells = static_cast<size_t>(c.cols) * static_cast<size_t>(c.rows);
        stateA.assign(cells, Vec4<R>{0, 0, 0, 0});
        stateB = stateA;
        bed.assign(cells, R(0));
        manning.assign(cells, R(0));
        if (c.scheme == HPO_SCHEME_MUSCL_HANCOCK)
            for (auto& f : faces) f.assign(cells, Vec4<R>{0, 0, 0, 0});
        workers = reduction_workers(static_cast<int64_t>(cells));
        reduction.assign(workers, R(0));
        time = R(0); timestep = static_cast<R>(c.initial_dt); time_hydro = R(0); time_target = R(0);
        batch_timesteps = R(0); batch_successful = 0; batch_skipped = 0;
        use_alternate = false;
    }

    void upload(const void* s, const void* b, const void* m) {
        std::memcpy(stateA.data(), s, cells * sizeof(Vec4<R>));
        std::memcpy(stateB.data(), s, cells * sizeof(Vec4<R>));
        std::memcpy(bed.data(), b, cells * sizeof(R));
        std::memcpy(manning.data(), m, cells * sizeof(R));
    }
    /* readDomainAll reads the next source buffer, CSchemeGodunov.cpp:1671-1679 */
    void download(void* s) const {
        const auto& cur = (use_alternate && cfg.scheme != HPO_SCHEME_MUSCL_HANCOCK) ? stateB : stateA;
        std::memcpy(s, cur.data(), cells * sizeof(Vec4<R>));
    }
    void download_both(void* a, void* b) const {
        std::memcpy(a, stateA.data(), cells * sizeof(Vec4<R>));
        std::memcpy(b, stateB.data(), cells * sizeof(Vec4<R>));
    }

    int add_uniform(const hpo_bdy_uniform& c, const double* tv) {
        Boundary b; b.kind = 0;
        b.uniform = BdyUniformConf<R>{c.entries, static_cast<R>(c.interval), static_cast<R>(c.length), c.definition};
        b.series.resize(2 * static_cast<size_t>(c.entries));
        for (size_t i = 0; i < b.series.size(); ++i) b.series[i] = static_cast<R>(tv[i]);
        bdys.push_back(std::move(b));
        return static_cast<int>(bdys.size()) - 1;
    }
    int add_gridded(const hpo_bdy_gridded& c, const double* v) {
        Boundary b; b.kind = 1;
        b.gridded = BdyGriddedConf<R>{static_cast<R>(c.interval), static_cast<R>(c.resolution), static_cast<R>(c.offset_x),
                                      static_cast<R>(c.offset_y), c.entries, c.definition, c.rows, c.cols};
        /* the kernel may index frame `entries` (one past the end, CLBoundaries.clc:229): keep a zero frame there */
        const size_t n = static_cast<size_t>(c.rows * c.cols * c.entries);
        b.series.assign(n + static_cast<size_t>(c.rows * c.cols), R(0));
        for (size_t i = 0; i < n; ++i) b.series[i] = static_cast<R>(v[i]);
        bdys.push_back(std::move(b));
        return static_cast<int>(bdys.size()) - 1;
    }
    int add_cell(const hpo_bdy_cell& c, const uint64_t* rel, const double* tdxy) {
        Boundary b; b.kind = 2;
        b.cell = BdyCellConf<R>{c.entries, static_cast<R>(c.interval), static_cast<R>(c.length), c.relations, c.def_depth,
                                c.def_discharge};
        /* the kernel reads entry base+1 (CLBoundaries.clc:44,49): keep one padding entry */
        b.series.assign(4 * (static_cast<size_t>(c.entries) + 1), R(0));
        for (size_t i = 0; i < 4 * static_cast<size_t>(c.entries); ++i) b.series[i] = static_cast<R>(tdxy[i]);
        b.relations.assign(rel, rel + c.relations);
        bdys.push_back(std::move(b));
        return static_cast<int>(bdys.size()) - 1;
    }

    void apply_boundaries(Vec4<R>* state) {
        int64_t gx = cfg.cols, gy = cfg.rows;
        if (cfg.quirks & HPO_QUIRK_BDY_COVERAGE) { gx = (cfg.cols / 8) * 8; gy = (cfg.rows / 8) * 8; }
        for (auto& b : bdys) {
            if (b.kind == 0)
                Kernels::bdy_uniform(cfg, &b.uniform, b.series.data(), &time, &timestep, &time_hydro, state, bed.data(),
                                     manning.data(), gx, gy);
            else if (b.kind == 1)
                Kernels::bdy_gridded(cfg, &b.gridded, b.series.data(), &time, &timestep, &time_hydro, state, bed.data(),
                                     manning.data(), gx, gy);
            else
                Kernels::bdy_cell(cfg, &b.cell, b.relations.data(), b.series.data(), &time, &timestep, &time_hydro, state,
                                  bed.data(), manning.data(),
                                  (static_cast<int64_t>(b.cell.RelationCount) / 8 + 1) * 8);
        }
    }

    void iterate_once() {
        Kernels::configure(cfg, workers);
        if (cfg.scheme == HPO_SCHEME_MUSCL_HANCOCK) {
            if (!(cfg.quirks & HPO_QUIRK_MH_NO_BOUNDARIES)) apply_boundaries(stateA.data());
            Kernels::mch_1st(cfg, &timestep, bed.data(), stateA.data(), faces[0].data(), faces[1].data(), faces[2].data(),
                             faces[3].data());
            Kernels::mch_2nd(cfg, &timestep, stateA.data(), bed.data(), manning.data(), faces[0].data(), faces[1].data(),
                             faces[2].data(), faces[3].data());
            if (cfg.dynamic) Kernels::reduce(cfg, stateA.data(), bed.data(), reduction.data(), workers);
        } else {
            Vec4<R>* src = use_alternate ? stateB.data() : stateA.data();
            Vec4<R>* dst = use_alternate ? stateA.data() : stateB.data();
            apply_boundaries(src);
            if (cfg.scheme == HPO_SCHEME_GODUNOV)
                Kernels::gts(cfg, &timestep, bed.data(), src, dst, manning.data());
            else
                Kernels::ine(cfg, &timestep, bed.data(), src, dst, manning.data());
            if (cfg.dynamic) {
                const Vec4<R>* red_src = (cfg.quirks & HPO_QUIRK_REDUCE_BUFFER_A) ? stateA.data() : dst;
                Kernels::reduce(cfg, red_src, bed.data(), reduction.data(), workers);
            }
        }
        Kernels::advance(cfg, &time, &timestep, &time_hydro, reduction.data(), workers, &time_target, &batch_timesteps,
                         &batch_successful, &batch_skipped);
        use_alternate = !use_alternate;
    }

    void iterate(int n) { for (int i = 0; i < n; ++i) iterate_once(); }

    /* tst_Reduce + tst_UpdateTimestep on the next source buffer, CSchemeGodunov.cpp:1191-1196 */
    void update_timestep() {
        Kernels::configure(cfg, workers);
        Kernels::reduce(cfg, stateA.data(), bed.data(), reduction.data(), workers);
        Kernels::update_timestep(cfg, &time, &timestep, reduction.data(), workers, &time_target, &batch_timesteps);
    }
    void reset_counters() { batch_timesteps = R(0); batch_successful = 0; batch_skipped = 0; }

    void stats(hpo_stats* o) const {
        o->time = time; o->timestep = timestep; o->time_hydro = time_hydro; o->time_target = time_target;
        o->batch_timesteps = batch_timesteps; o->batch_successful = batch_successful; o->batch_skipped = batch_skipped;
        o->use_alternate = use_alternate ? 1u : 0u; o->pad = 0;
    }

    hpo_config cfg;
    size_t cells = 0, workers = 0;
    std::vector<Vec4<R>> stateA, stateB, faces[4];
    std::vector<R> bed, manning, reduction;
    R time, timestep, time_hydro, time_target, batch_timesteps;
    uint32_t batch_successful, batch_skipped;
    bool use_alternate;

  private:
    struct Boundary {
        int kind = 0;
        BdyUniformConf<R> uniform{}; BdyGriddedConf<R> gridded{}; BdyCellConf<R> cell{};
        std::vector<R> series; std::vector<uint64_t> relations;
    };
    std::vector<Boundary> bdys;
};

}  // namespace hpo

/* Emits the extern "C" entry points of hpo_api.h for one (prefix, real, kernels) triple. */
#define HPO_DEFINE_API(P, R, KERNELS)                                                                                       \
    extern "C" {                                                                                                            \
    void* P##create(const hpo_config* c) { return new hpo::Sim<R, KERNELS>(*c); }                                           \
    void P##destroy(void* s) { delete static_cast<hpo::Sim<R, KERNELS>*>(s); }                                              \
    void P##upload(void* s, const void* st, const void* b, const void* m) {                                                 \
        static_cast<hpo::Sim<R, KERNELS>*>(s)->upload(st, b, m); }                                                          \
    void P##download(void* s, void* st) { static_cast<hpo::Sim<R, KERNELS>*>(s)->download(st); }                            \
    void P##download_both(void* s, void* a, void* b) { static_cast<hpo::Sim<R, KERNELS>*>(s)->download_both(a, b); }        \
    void P##set_target(void* s, double t) { static_cast<hpo::Sim<R, KERNELS>*>(s)->time_target = static_cast<R>(t); }       \
    void P##set_clock(void* s, double t, double dt, double th) {                                                            \
        auto* p = static_cast<hpo::Sim<R, KERNELS>*>(s);                                                                    \
        p->time = static_cast<R>(t); p->timestep = static_cast<R>(dt); p->time_hydro = static_cast<R>(th); }                \
    int P##add_uniform(void* s, const hpo_bdy_uniform* c, const double* tv) {                                               \
        return static_cast<hpo::Sim<R, KERNELS>*>(s)->add_uniform(*c, tv); }                                                \
    int P##add_gridded(void* s, const hpo_bdy_gridded* c, const double* v) {                                                \
        return static_cast<hpo::Sim<R, KERNELS>*>(s)->add_gridded(*c, v); }                                                 \
    int P##add_cell(void* s, const hpo_bdy_cell* c, const uint64_t* rel, const double* v) {                                 \
        return static_cast<hpo::Sim<R, KERNELS>*>(s)->add_cell(*c, rel, v); }                                               \
    void P##iterate(void* s, int n) { static_cast<hpo::Sim<R, KERNELS>*>(s)->iterate(n); }                                  \
    void P##update_timestep(void* s) { static_cast<hpo::Sim<R, KERNELS>*>(s)->update_timestep(); }                          \
    void P##reset_counters(void* s) { static_cast<hpo::Sim<R, KERNELS>*>(s)->reset_counters(); }                            \
    void P##stats(void* s, hpo_stats* o) { static_cast<hpo::Sim<R, KERNELS>*>(s)->stats(o); }                               \
    void P##k_gts(const hpo_config* c, const void* dt, const void* bed, const void* src, void* dst, const void* mn) {       \
        KERNELS::configure(*c, hpo::reduction_workers(c->cols * c->rows));                                                  \
        KERNELS::gts(*c, static_cast<const R*>(dt), static_cast<const R*>(bed), static_cast<const hpo::Vec4<R>*>(src),      \
                     static_cast<hpo::Vec4<R>*>(dst), static_cast<const R*>(mn)); }                                         \
    void P##k_ine(const hpo_config* c, const void* dt, const void* bed, const void* src, void* dst, const void* mn) {       \
        KERNELS::configure(*c, hpo::reduction_workers(c->cols * c->rows));                                                  \
        KERNELS::ine(*c, static_cast<const R*>(dt), static_cast<const R*>(bed), static_cast<const hpo::Vec4<R>*>(src),      \
                     static_cast<hpo::Vec4<R>*>(dst), static_cast<const R*>(mn)); }                                         \
    void P##k_mch_1st(const hpo_config* c, const void* dt, const void* bed, const void* st, void* fN, void* fE, void* fS,   \
                      void* fW) {                                                                                           \
        KERNELS::configure(*c, hpo::reduction_workers(c->cols * c->rows));                                                  \
        KERNELS::mch_1st(*c, static_cast<const R*>(dt), static_cast<const R*>(bed), static_cast<const hpo::Vec4<R>*>(st),   \
                         static_cast<hpo::Vec4<R>*>(fN), static_cast<hpo::Vec4<R>*>(fE), static_cast<hpo::Vec4<R>*>(fS),    \
                         static_cast<hpo::Vec4<R>*>(fW)); }                                                                 \
    void P##k_mch_2nd(const hpo_config* c, const void* dt, void* st, const void* bed, const void* mn, const void* fN,       \
                      const void* fE, const void* fS, const void* fW) {                                                     \
        KERNELS::configure(*c, hpo::reduction_workers(c->cols * c->rows));                                                  \
        KERNELS::mch_2nd(*c, static_cast<const R*>(dt), static_cast<hpo::Vec4<R>*>(st), static_cast<const R*>(bed),         \
                         static_cast<const R*>(mn), static_cast<const hpo::Vec4<R>*>(fN),                                   \
                         static_cast<const hpo::Vec4<R>*>(fE), static_cast<const hpo::Vec4<R>*>(fS),                        \
                         static_cast<const hpo::Vec4<R>*>(fW)); }                                                           \
    double P##k_reduce(const hpo_config* c, const void* st, const void* bed) {                                              \
        const size_t w = hpo::reduction_workers(c->cols * c->rows);                                                         \
        KERNELS::configure(*c, w);                                                                                          \
        std::vector<R> red(w, R(0));                                                                                        \
        KERNELS::reduce(*c, static_cast<const hpo::Vec4<R>*>(st), static_cast<const R*>(bed), red.data(), w);               \
        R m = R(0); for (size_t i = 0; i < w; ++i) if (red[i] > m) m = red[i];                                              \
        return static_cast<double>(m); }                                                                                    \
    void P##k_advance(const hpo_config* c, void* clock, uint32_t* counters, double vmax) {                                  \
        KERNELS::configure(*c, 1);                                                                                          \
        R* k = static_cast<R*>(clock);                                                                                      \
        R red = static_cast<R>(vmax);                                                                                       \
        KERNELS::advance(*c, &k[0], &k[1], &k[2], &red, 1, &k[3], &k[4], &counters[0], &counters[1]); }                     \
    }

#endif
