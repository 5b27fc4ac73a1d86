// hp_executor.cu -- implementation of the C ABI declared in include/hipims_cuda.h.
//
// Owns what the reference's scheme classes own through the OpenCL wrappers: the device buffers
// (src/Schemes/CSchemeGodunov.cpp:789-893), the kernel launch sequence of one iteration
// (:1617-1666; MUSCL-Hancock src/Schemes/CSchemeMUSCLHancock.cpp:646-680) and the data movement of
// prepareSimulation / readDomainAll / CDomainLink.  One CUDA stream per executor; batches of
// iterations are replayed as CUDA graphs so that small domains are not launch-bound (the
// reference blocks the host once per batch, CSchemeGodunov.cpp:1337-1341 -- so do we).
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/hipims_cuda.h"
#include "hp_comm.h"
#include "hp_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define HP_CUDA(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t e__ = (expr);                                                                                  \
        if (e__ != cudaSuccess)                                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? HP_ERR_OOM : HP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                              \
    } while (0)

template <class R> hp::Clock<R> make_clock(double t, double dt, double th, double target) {
    hp::Clock<R> c;
    c.time = static_cast<R>(t); c.timestep = static_cast<R>(dt); c.time_hydro = static_cast<R>(th);
    c.time_target = static_cast<R>(target); c.batch_timesteps = R(0); c.batch_successful = 0; c.batch_skipped = 0;
    return c;
}

struct Boundary {
    int kind = 0;  // 0 uniform, 1 gridded, 2 cell
    hp_bdy_uniform uniform{}; hp_bdy_gridded gridded{}; hp_bdy_cell cell{};
    void* series = nullptr; long long* relations = nullptr; unsigned long long count = 0;
};

}  // namespace

// Every entry point takes the handle's lock: the reference's main thread polls and reads back while the scheme's
// worker thread enqueues batches (src/Schemes/CSchemeGodunov.cpp:1116-1141, CScheme.h:137-139).  Recursive because
// entry points call each other (destroy from a failed create, ...).  Lock order: scheme, then executor.
struct hp_executor {
    std::recursive_mutex mu;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaDeviceProp prop{};
};

struct hp_scheme {
    std::recursive_mutex mu;
    hp_executor* ex = nullptr;
    hp_scheme_config cfg{};
    hp::Grid grid{};
    hp::ParamsD params{};
    const hp::KernelTable* K = nullptr;
    size_t rb = 8, plane_bytes = 0;
    // the ten planes live in ONE allocation, in the order  A.eta A.qx A.qy A.emax | zb n | B.eta B.qx B.qy B.emax,
    // so that a single 3-D TMA box of six consecutive planes holds everything a step reads (buffer A: planes
    // 0..5, buffer B: planes 4..9)
    char* block = nullptr;
    hp::Planes A{}, B{};
    void *bed = nullptr, *manning = nullptr, *clock = nullptr;
    unsigned long long* max_bits = nullptr;
    unsigned int* ticket = nullptr;
    void* staging = nullptr; size_t staging_bytes = 0; int staging_rows = 0;
    std::vector<Boundary> bdys;
    bool use_alt = false;
    uint64_t iterations = 0, launches = 0;
    // graphs: [0] one pair of iterations (A->B, B->A), [1] kGraphPairs pairs
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    int graph_launches[2] = {0, 0};
    bool use_tma = false, use_march = false;
    hp::TmaMapsPOD maps_a{}, maps_b{};   // descriptors with buffer A / buffer B as the source
    hp::TmaMaps6POD march_map{};         // bytes[0]: 3-D descriptor over the ten-plane block
    hp::Comm* comm = nullptr;
    int world = 1;
    bool small_strip = true;             // decided from rank-independent quantities in hp_scheme_attach_comm
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_edges = nullptr, ev_halo = nullptr;
    // row strips over peer memory (hp_scheme_attach_peers): my mailbox, the peers' mapped buffers, and the exchange
    // kernel's arguments for the two ping-pong directions
    hp::PeerBox* peer_box = nullptr;
    bool peers = false;
    int peer_rank = 0;
    hp::PeerArgs peer_args[2]{};          // [alt]
    std::vector<void*> peer_mapped;       // cudaIpcOpenMemHandle results to close
    // hp_scheme_strip_timing: per-phase device times of the strip iteration (direct launches while enabled)
    bool strip_timing = false;
    cudaEvent_t ev_t[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double t_phase[5] = {0, 0, 0, 0, 0};
    uint64_t t_count = 0;
};

namespace {

constexpr int kGraphPairs = 8;
constexpr long long kSmallStripCells = 32ll << 20;   // strips up to this size: unsplit step, graph replay
constexpr int kCommSpareSMs = 4;     // SMs left to NCCL during the interior launch of a strip

void drop_graphs(hp_scheme* s) {
    for (auto& g : s->graph_exec) { if (g) cudaGraphExecDestroy(g); g = nullptr; }
}

hp::Planes& src_planes(hp_scheme* s, bool alt) { return alt ? s->B : s->A; }
hp::Planes& dst_planes(hp_scheme* s, bool alt) { return alt ? s->A : s->B; }

hp::StepArgs base_args(hp_scheme* s, bool alt) {
    hp::StepArgs a{};
    a.src = src_planes(s, alt); a.dst = dst_planes(s, alt);
    a.bed = s->bed; a.manning = s->manning; a.clock = s->clock; a.max_bits = s->max_bits; a.ticket = s->ticket;
    a.grid = s->grid; a.params = s->params; a.y0 = s->grid.own_y0; a.y1 = s->grid.own_y1;
    a.reduce_mode = hp::kReduceNone; a.finalize = 1; a.total_ctas = 0;
    return a;
}

int apply_boundaries(hp_scheme* s, const hp::Planes& state) {
    int n = 0;
    const auto& g = s->grid;
    const bool q6 = (s->cfg.quirks & HP_QUIRK_BDY_COVERAGE) != 0;
    const int cover_x = q6 ? (g.cols / 8) * 8 : g.cols, cover_y = q6 ? (g.grows / 8) * 8 : g.grows;
    for (const auto& b : s->bdys) {
        if (b.kind == 0) {
            hp::BdyUniformArgs a{state, s->bed, s->clock, b.series, g, b.uniform.entries, b.uniform.definition,
                                 b.uniform.interval, b.uniform.length, cover_x, cover_y};
            n += s->K->bdy_uniform(static_cast<int>(s->rb), a, s->ex->stream);
        } else if (b.kind == 1) {
            hp::BdyGriddedArgs a{state, s->clock, b.series, g, b.gridded.interval, b.gridded.resolution, b.gridded.offset_x,
                                 b.gridded.offset_y, s->cfg.delta, b.gridded.entries, b.gridded.definition, b.gridded.rows,
                                 b.gridded.cols, cover_x, cover_y};
            n += s->K->bdy_gridded(static_cast<int>(s->rb), a, s->ex->stream);
        } else {
            hp::BdyCellArgs a{state, s->bed, s->clock, b.series, b.relations, g, s->params, b.cell.entries, b.count,
                              b.cell.interval, b.cell.length, b.cell.def_depth, b.cell.def_discharge};
            n += s->K->bdy_cell(static_cast<int>(s->rb), a, s->ex->stream);
        }
    }
    return n;
}

// width mode of the marching kernels (KernelTable::step_march): 0 default, 1 one column per lane, 2 two columns
static int march_width_mode(const hp_scheme* s) {
    return (s->cfg.options & HP_OPT_NARROW_MARCH) ? 1 : ((s->cfg.options & HP_OPT_WIDE_MARCH) ? 2 : 0);
}

// One iteration = CSchemeGodunov::scheduleIteration / CSchemeMUSCLHancock::scheduleIteration.
int enqueue_iteration(hp_scheme* s, bool alt, int* launched) {
    const int rb = static_cast<int>(s->rb);
    cudaStream_t st = s->ex->stream;
    hp::StepArgs a = base_args(s, alt);
    int n = 0;
    const bool mh = s->cfg.scheme == HP_SCHEME_MUSCL_HANCOCK;
    if (!(mh && (s->cfg.quirks & HP_QUIRK_MH_NO_BOUNDARIES))) n += apply_boundaries(s, a.src);
    if (s->cfg.dynamic_timestep) {
        if (mh || !(s->cfg.quirks & HP_QUIRK_REDUCE_BUFFER_A)) a.reduce_mode = hp::kReduceDst;
        else a.reduce_mode = alt ? hp::kReduceDst : hp::kReduceSrc;   // Q1: always buffer A
    }
    // `spare_sms`: the persistent kernels fill every SM; the interior launch of a strip leaves a few SMs free so that
    // the NCCL send/recv kernels of the halo exchange can run beside it instead of after it
    auto step = [&](const hp::StepArgs& args, int spare_sms = 0) {
        const int sms = s->ex->prop.multiProcessorCount - spare_sms > 0 ? s->ex->prop.multiProcessorCount - spare_sms : 1;
        if (s->use_march)
            return s->K->step_march(static_cast<int>(s->cfg.scheme), rb, args, &s->march_map,
                                    (alt ? 1 : 0) | (march_width_mode(s) << 1), sms, st);
        if (s->use_tma) return s->K->step_tma(static_cast<int>(s->cfg.scheme), rb, args, alt ? &s->maps_b : &s->maps_a, sms, st);
        return s->K->step(static_cast<int>(s->cfg.scheme), rb, args, st);
    };
    // per-phase device times of a strip iteration (hp_scheme_strip_timing; never inside a graph capture)
    const bool timing = s->strip_timing && s->comm != nullptr;
    auto mark = [&](int i) { if (timing) cudaEventRecord(s->ev_t[i], st); };
    auto collect = [&]() {
        if (!timing) return;
        cudaEventSynchronize(s->ev_t[5]);
        for (int i = 0; i < 5; ++i) { float ms = 0; cudaEventElapsedTime(&ms, s->ev_t[i], s->ev_t[i + 1]); s->t_phase[i] += ms; }
        ++s->t_count;
    };
    if (s->peers) {
        // Row strips over peer memory: the cell update of all owned rows, then ONE kernel that stores the edge rows into
        // the neighbours' halo rows, exchanges the wave-speed maximum through the strips' mailboxes and runs the time
        // controller (hp_kernels.cu: peer_exchange_kernel).
        a.finalize = 0;
        n += step(a);
        n += s->K->peer_exchange(rb, s->peer_args[alt ? 1 : 0], a, 0, st);
    } else if (s->comm == nullptr) {
        a.finalize = 1;
        n += step(a);
    } else if (s->small_strip && !(s->cfg.options & HP_OPT_SPLIT_STRIPS)) {
        // Small row strips: there is too little interior work to hide the exchange behind, and the two edge launches
        // cost more than they save.  One launch over all owned rows, then the halo send/recv and the all-reduce of the
        // wave speed in ONE NCCL group, then the time controller.
        a.finalize = 0;
        mark(0); mark(1);
        n += step(a);
        mark(2); mark(3);
        const char* err = hp::comm_exchange_and_allreduce(s->comm, a.dst, s->grid, mh ? 2 : 1, s->rb, s->max_bits, st);
        if (err) return fail(HP_ERR_NCCL, "halo exchange + allreduce: %s", err);
        mark(4);
        n += s->K->advance(rb, a, st);
        mark(5);
        collect();
    } else {
        // Large row strips: edge rows first, their halo exchange overlaps the interior rows, then the
        // wave-speed maximum is all-reduced on the device and one thread runs the time controller.
        a.finalize = 0;
        const int halo = mh ? 2 : 1;
        const int y0 = s->grid.own_y0, y1 = s->grid.own_y1;
        const int e0 = y0 + halo < y1 ? y0 + halo : y1, e1 = y1 - halo > e0 ? y1 - halo : e0;
        hp::StepArgs lo = a, hi = a, mid = a;
        lo.y1 = e0; hi.y0 = e1; mid.y0 = e0; mid.y1 = e1;
        mark(0);
        n += step(lo);
        n += step(hi);
        mark(1);
        if (cudaEventRecord(s->ev_edges, st) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaEventRecord failed");
        if (cudaStreamWaitEvent(s->comm_stream, s->ev_edges, 0) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaStreamWaitEvent failed");
        const char* err = hp::comm_exchange_halos(s->comm, a.dst, s->grid, halo, s->rb, s->comm_stream);
        if (err) return fail(HP_ERR_NCCL, "halo exchange: %s", err);
        if (cudaEventRecord(s->ev_halo, s->comm_stream) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaEventRecord failed");
        n += step(mid, kCommSpareSMs);
        mark(2);
        if (cudaStreamWaitEvent(st, s->ev_halo, 0) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaStreamWaitEvent failed");
        mark(3);
        err = hp::comm_allreduce_max(s->comm, s->max_bits, st);
        if (err) return fail(HP_ERR_NCCL, "allreduce: %s", err);
        mark(4);
        n += s->K->advance(rb, a, st);
        mark(5);
        collect();
    }
    if (launched) *launched += n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(HP_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return HP_OK;
}

int build_graph(hp_scheme* s, int slot, int pairs) {
    cudaStream_t st = s->ex->stream;
    cudaGraph_t graph = nullptr;
    HP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int launched = 0, rc = HP_OK;
    for (int p = 0; p < pairs && rc == HP_OK; ++p) {
        rc = enqueue_iteration(s, false, &launched);
        if (rc == HP_OK) rc = enqueue_iteration(s, true, &launched);
    }
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != HP_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(HP_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->graph_exec[slot], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(HP_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    s->graph_launches[slot] = launched;
    return HP_OK;
}

// zero-filled device allocation; the fill is ordered on the scheme's own stream (a legacy
// default-stream cudaMemset is NOT ordered against a non-blocking stream)
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
int encode_plane_map(void* out128, void* plane, const hp::Grid& g, size_t rb, int box_w, int box_h, int planes = 0,
                     size_t plane_bytes = 0, int box_planes = 0) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        HP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(HP_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(g.cols), static_cast<cuuint64_t>(g.rows), static_cast<cuuint64_t>(planes)};
    const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(g.pitch) * rb, static_cast<cuuint64_t>(plane_bytes)};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), static_cast<cuuint32_t>(box_planes)};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUtensorMap tm;
    // L2 promotion: how much of a line a TMA miss brings into L2.  HIPIMS_TMA_L2_PROMOTION (0 none, 1 64 B, 2 128 B, 3 256 B)
    // is a measurement knob; the default is what measured best.
    static const CUtensorMapL2promotion promotion = []() {
        const char* e = getenv("HIPIMS_TMA_L2_PROMOTION");
        const int v = e ? atoi(e) : 2;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
             : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    const CUresult r = encode(&tm, rb == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, planes > 0 ? 3 : 2, plane, gdim, gstride,
                              box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promotion,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    memcpy(out128, &tm, 128);
    return HP_OK;
}

int build_tma_maps(hp_scheme* s) {
    const int halo = s->cfg.scheme == HP_SCHEME_MUSCL_HANCOCK ? 2 : 1;
    int rc;
    hp::Planes* bufs[2] = {&s->A, &s->B};
    hp::TmaMapsPOD* maps[2] = {&s->maps_a, &s->maps_b};
    const int rbi = static_cast<int>(s->rb);
    if (s->use_march) {
        // marching kernels: one 3-D descriptor over the ten-plane block; a box is one row of six planes
        const int box_w = s->K->march_box_w(static_cast<int>(s->cfg.scheme), rbi, march_width_mode(s));
        return encode_plane_map(s->march_map.bytes[0], s->block, s->grid, s->rb, box_w, 1, 10, s->plane_bytes, 6);
    }
    const int w = hp::tma_box_w(rbi, halo), h = hp::tma_box_h(halo);
    for (int b = 0; b < 2; ++b) {
        if ((rc = encode_plane_map(maps[b]->bytes[0], bufs[b]->eta, s->grid, s->rb, w, h))) return rc;
        if ((rc = encode_plane_map(maps[b]->bytes[1], bufs[b]->qx, s->grid, s->rb, w, h))) return rc;
        if ((rc = encode_plane_map(maps[b]->bytes[2], bufs[b]->qy, s->grid, s->rb, w, h))) return rc;
        if ((rc = encode_plane_map(maps[b]->bytes[3], s->bed, s->grid, s->rb, w, h))) return rc;
    }
    return HP_OK;
}

template <class T> int dev_alloc(hp_scheme* s, T** p, size_t bytes) {
    HP_CUDA(cudaMalloc(reinterpret_cast<void**>(p), bytes));
    HP_CUDA(cudaMemsetAsync(*p, 0, bytes, s->ex->stream));
    return HP_OK;
}

int alloc_block(hp_scheme* s) {
    int rc;
    if ((rc = dev_alloc(s, &s->block, 10 * s->plane_bytes))) return rc;
    auto at = [&](int i) { return static_cast<void*>(s->block + static_cast<size_t>(i) * s->plane_bytes); };
    s->A = hp::Planes{at(0), at(3), at(1), at(2)};           // Planes is {eta, emax, qx, qy}
    s->bed = at(4); s->manning = at(5);
    s->B = hp::Planes{at(6), at(9), at(7), at(8)};
    return HP_OK;
}

int write_clock(hp_scheme* s, double t, double dt, double th, double target) {
    if (s->rb == 8) { auto c = make_clock<double>(t, dt, th, target); HP_CUDA(cudaMemcpyAsync(s->clock, &c, sizeof(c), cudaMemcpyHostToDevice, s->ex->stream)); }
    else { auto c = make_clock<float>(t, dt, th, target); HP_CUDA(cudaMemcpyAsync(s->clock, &c, sizeof(c), cudaMemcpyHostToDevice, s->ex->stream)); }
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    return HP_OK;
}

// writes one real-typed field of the device clock
int write_clock_field(hp_scheme* s, int index, double value) {
    double d = value; float f = static_cast<float>(value);
    char* base = static_cast<char*>(s->clock) + static_cast<size_t>(index) * s->rb;
    HP_CUDA(cudaMemcpyAsync(base, s->rb == 8 ? static_cast<void*>(&d) : static_cast<void*>(&f), s->rb, cudaMemcpyHostToDevice,
                            s->ex->stream));
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    return HP_OK;
}

int rows_to_device(hp_scheme* s, const void* host_aos, int row0, int nrows, bool both) {
    const size_t row_bytes = static_cast<size_t>(s->grid.cols) * 4 * s->rb;
    const char* h = static_cast<const char*>(host_aos);
    for (int r = 0; r < nrows; r += s->staging_rows) {
        const int n = nrows - r < s->staging_rows ? nrows - r : s->staging_rows;
        HP_CUDA(cudaMemcpyAsync(s->staging, h + static_cast<size_t>(r) * row_bytes, static_cast<size_t>(n) * row_bytes,
                                cudaMemcpyHostToDevice, s->ex->stream));
        if (both) {
            s->launches += s->K->aos_to_soa(static_cast<int>(s->rb), s->staging, s->A, s->grid, row0 + r, n, s->ex->stream);
            s->launches += s->K->aos_to_soa(static_cast<int>(s->rb), s->staging, s->B, s->grid, row0 + r, n, s->ex->stream);
        } else {
            s->launches += s->K->aos_to_soa(static_cast<int>(s->rb), s->staging, src_planes(s, s->use_alt), s->grid, row0 + r, n,
                                            s->ex->stream);
        }
    }
    HP_CUDA(cudaGetLastError());
    return HP_OK;
}

int rows_to_host(hp_scheme* s, const hp::Planes& p, void* host_aos, int row0, int nrows) {
    const size_t row_bytes = static_cast<size_t>(s->grid.cols) * 4 * s->rb;
    char* h = static_cast<char*>(host_aos);
    for (int r = 0; r < nrows; r += s->staging_rows) {
        const int n = nrows - r < s->staging_rows ? nrows - r : s->staging_rows;
        s->launches += s->K->soa_to_aos(static_cast<int>(s->rb), p, s->staging, s->grid, row0 + r, n, s->ex->stream);
        HP_CUDA(cudaMemcpyAsync(h + static_cast<size_t>(r) * row_bytes, s->staging, static_cast<size_t>(n) * row_bytes,
                                cudaMemcpyDeviceToHost, s->ex->stream));
    }
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    return HP_OK;
}

int upload_series(hp_scheme* s, const double* host, size_t count, size_t padded, void** out) {
    // converted to the working precision like the reference's host does before upload
    std::vector<char> tmp(padded * s->rb, 0);
    if (s->rb == 8) { double* d = reinterpret_cast<double*>(tmp.data()); for (size_t i = 0; i < count; ++i) d[i] = host[i]; }
    else { float* f = reinterpret_cast<float*>(tmp.data()); for (size_t i = 0; i < count; ++i) f[i] = static_cast<float>(host[i]); }
    HP_CUDA(cudaMalloc(out, tmp.size()));
    cudaError_t e = cudaMemcpyAsync(*out, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, s->ex->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->ex->stream);
    if (e != cudaSuccess) { cudaFree(*out); *out = nullptr; return fail(HP_ERR_CUDA, "timeseries upload failed: %s", cudaGetErrorString(e)); }
    return HP_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int hp_abi_version(void) { return HP_ABI_VERSION; }
const char* hp_last_error(void) { return g_last_error.c_str(); }

int hp_device_count(int* count) {
    if (!count) return fail(HP_ERR_INVALID, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        *count = 0;
        cudaGetLastError();
        return fail(HP_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    *count = n;
    return HP_OK;
}

int hp_executor_create(int device_ordinal, void* external_stream, hp_executor** out) {
    if (!out) return fail(HP_ERR_INVALID, "out is null");
    int n = 0, rc = hp_device_count(&n);
    if (rc != HP_OK) return rc;
    if (device_ordinal < 0 || device_ordinal >= n) return fail(HP_ERR_INVALID, "device ordinal %d out of range [0,%d)", device_ordinal, n);
    HP_CUDA(cudaSetDevice(device_ordinal));
    hp_executor* ex = new hp_executor();
    ex->device = device_ordinal;
    // on any failure below the half-built executor is released again (hp_executor_destroy copes with missing parts)
    auto bail = [&](int code) { const std::string keep = g_last_error; hp_executor_destroy(ex); g_last_error = keep; return code; };
    cudaError_t e = cudaGetDeviceProperties(&ex->prop, device_ordinal);
    if (e != cudaSuccess) return bail(fail(HP_ERR_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e)));
    if (ex->prop.major < 10)
        return bail(fail(HP_ERR_NO_DEVICE, "device is sm_%d%d; this library is built for sm_100a (B200) only", ex->prop.major, ex->prop.minor));
    if (external_stream) { ex->stream = static_cast<cudaStream_t>(external_stream); ex->own_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { ex->stream = nullptr; return bail(fail(HP_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e))); }
        ex->own_stream = true;
    }
    if ((e = cudaEventCreate(&ex->ev0)) != cudaSuccess || (e = cudaEventCreate(&ex->ev1)) != cudaSuccess)
        return bail(fail(HP_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)));
    *out = ex;
    return HP_OK;
}

void hp_executor_destroy(hp_executor* ex) {
    if (!ex) return;
    {
        std::lock_guard<std::recursive_mutex> lock__(ex->mu);
        cudaSetDevice(ex->device);
        if (ex->ev0) cudaEventDestroy(ex->ev0);
        if (ex->ev1) cudaEventDestroy(ex->ev1);
        if (ex->own_stream && ex->stream) cudaStreamDestroy(ex->stream);
    }
    delete ex;
}

int hp_executor_describe(hp_executor* ex, char* name, size_t name_len, int* sm_count, size_t* total_mem) {
    if (!ex) return fail(HP_ERR_INVALID, "executor is null");
    std::lock_guard<std::recursive_mutex> lock__(ex->mu);
    if (name && name_len) { strncpy(name, ex->prop.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = ex->prop.multiProcessorCount;
    if (total_mem) *total_mem = ex->prop.totalGlobalMem;
    return HP_OK;
}

int hp_executor_finish(hp_executor* ex) {
    if (!ex) return fail(HP_ERR_INVALID, "executor is null");
    // the stream handle never changes: block WITHOUT the lock so that another thread can keep enqueueing
    HP_CUDA(cudaStreamSynchronize(ex->stream));
    return HP_OK;
}
int hp_executor_timer_start(hp_executor* ex) {
    if (!ex) return fail(HP_ERR_INVALID, "executor is null");
    std::lock_guard<std::recursive_mutex> lock__(ex->mu);
    HP_CUDA(cudaSetDevice(ex->device));
    HP_CUDA(cudaEventRecord(ex->ev0, ex->stream));
    return HP_OK;
}
int hp_executor_timer_stop(hp_executor* ex, float* ms) {
    if (!ex || !ms) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(ex->mu);
    HP_CUDA(cudaSetDevice(ex->device));
    HP_CUDA(cudaEventRecord(ex->ev1, ex->stream));
    HP_CUDA(cudaEventSynchronize(ex->ev1));
    HP_CUDA(cudaEventElapsedTime(ms, ex->ev0, ex->ev1));
    return HP_OK;
}

int hp_scheme_create(hp_executor* ex, const hp_scheme_config* cfg, hp_scheme** out) {
    if (!ex || !cfg || !out) return fail(HP_ERR_INVALID, "null argument");
    if (cfg->struct_size != sizeof(hp_scheme_config)) return fail(HP_ERR_INVALID, "hp_scheme_config size mismatch (ABI %d)", HP_ABI_VERSION);
    if (cfg->scheme > HP_SCHEME_INERTIAL) return fail(HP_ERR_INVALID, "unknown scheme %u", cfg->scheme);
    if (cfg->real_bytes != 4 && cfg->real_bytes != 8) return fail(HP_ERR_INVALID, "real_bytes must be 4 or 8");
    if (cfg->cols < 3 || cfg->rows < 3 || cfg->cols > 0x3fffffff || cfg->rows > 0x3fffffff) return fail(HP_ERR_INVALID, "bad domain size");
    if (!(cfg->delta > 0.0)) return fail(HP_ERR_INVALID, "delta must be positive");
    const uint64_t own = cfg->rows - cfg->halo_south - cfg->halo_north;
    if (cfg->halo_south + cfg->halo_north >= cfg->rows || cfg->row_offset < cfg->halo_south ||
        cfg->row_offset + own + cfg->halo_north > cfg->global_rows)
        return fail(HP_ERR_INVALID, "strip geometry inconsistent (rows=%llu offset=%llu global=%llu)", (unsigned long long)cfg->rows,
                    (unsigned long long)cfg->row_offset, (unsigned long long)cfg->global_rows);
    HP_CUDA(cudaSetDevice(ex->device));
    hp_scheme* s = new hp_scheme();
    s->ex = ex; s->cfg = *cfg; s->rb = cfg->real_bytes;
    s->K = (cfg->options & HP_OPT_STRICT_FP) ? &hp::strict_kernels() : &hp::fast_kernels();
    hp::Grid& g = s->grid;
    g.cols = static_cast<int>(cfg->cols); g.rows = static_cast<int>(cfg->rows);
    g.pitch = (g.cols + 31) / 32 * 32;
    g.grows = static_cast<int>(cfg->global_rows);
    g.gy0 = static_cast<int>(cfg->row_offset) - static_cast<int>(cfg->halo_south);
    g.own_y0 = static_cast<int>(cfg->halo_south); g.own_y1 = g.rows - static_cast<int>(cfg->halo_north);
    s->params = hp::ParamsD{cfg->dry_threshold, cfg->dry_threshold * 10, cfg->delta, cfg->courant, cfg->end_time,
                            cfg->fixed_timestep, static_cast<int>(cfg->dynamic_timestep), static_cast<int>(cfg->friction),
                            cfg->scheme == HP_SCHEME_INERTIAL ? 1 : 0,
                            (cfg->scheme == HP_SCHEME_GODUNOV && (cfg->quirks & HP_QUIRK_GODUNOV_DT0_KEEP)) ? 1 : 0};
    s->plane_bytes = static_cast<size_t>(g.rows) * g.pitch * s->rb;
    int rc = HP_OK;
    do {
        if ((rc = alloc_block(s))) break;
        if ((rc = dev_alloc(s, reinterpret_cast<char**>(&s->clock), 64))) break;
        if ((rc = dev_alloc(s, &s->max_bits, sizeof(unsigned long long)))) break;
        if ((rc = dev_alloc(s, &s->ticket, sizeof(unsigned int)))) break;
        const size_t row_bytes = static_cast<size_t>(g.cols) * 4 * s->rb;
        size_t rows_fit = (static_cast<size_t>(256) << 20) / row_bytes;
        if (rows_fit < 1) rows_fit = 1;
        if (rows_fit > static_cast<size_t>(g.rows)) rows_fit = g.rows;
        s->staging_rows = static_cast<int>(rows_fit);
        s->staging_bytes = rows_fit * row_bytes;
        if ((rc = dev_alloc(s, reinterpret_cast<char**>(&s->staging), s->staging_bytes))) break;
        if ((rc = write_clock(s, 0.0, cfg->initial_timestep, 0.0, 0.0))) break;
        // TMA-staged kernels: the fast flavour's Godunov step (others use the plain-load kernels)
        s->use_tma = s->K->step_tma != nullptr && !(cfg->options & HP_OPT_NO_TMA);
        s->use_march = s->use_tma && s->K->step_march != nullptr &&
                       (cfg->scheme == HP_SCHEME_INERTIAL ||
                        (cfg->scheme == HP_SCHEME_MUSCL_HANCOCK && !(cfg->options & HP_OPT_TILE_KERNELS)) ||
                        (cfg->scheme == HP_SCHEME_GODUNOV && (cfg->options & HP_OPT_MARCH_GODUNOV)));
        if (s->use_tma && (rc = build_tma_maps(s))) break;
    } while (0);
    if (rc != HP_OK) { hp_scheme_destroy(s); return rc; }
    *out = s;
    return HP_OK;
}

void hp_scheme_destroy(hp_scheme* s) {
    if (!s) return;
    {
        std::lock_guard<std::recursive_mutex> lock__(s->mu);
        cudaSetDevice(s->ex->device);
        cudaStreamSynchronize(s->ex->stream);
        drop_graphs(s);
        if (s->comm) hp::comm_destroy(s->comm);
        if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
        for (auto& e : s->ev_t) if (e) cudaEventDestroy(e);
        if (s->ev_edges) cudaEventDestroy(s->ev_edges);
        if (s->ev_halo) cudaEventDestroy(s->ev_halo);
        for (void* m : s->peer_mapped) cudaIpcCloseMemHandle(m);
        cudaFree(s->peer_box);
        cudaFree(s->block); cudaFree(s->clock); cudaFree(s->max_bits); cudaFree(s->ticket); cudaFree(s->staging);
        for (auto& b : s->bdys) { cudaFree(b.series); cudaFree(b.relations); }
    }
    delete s;
}

int hp_boundary_add_uniform(hp_scheme* s, const hp_bdy_uniform* conf, const double* series) {
    if (!s || !conf || !series) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (conf->entries < 2 || !(conf->interval > 0.0)) return fail(HP_ERR_INVALID, "a boundary timeseries is too short");
    Boundary b; b.kind = 0; b.uniform = *conf;
    int rc = upload_series(s, series, 2 * static_cast<size_t>(conf->entries), 2 * static_cast<size_t>(conf->entries) + 2, &b.series);
    if (rc) return rc;
    s->bdys.push_back(b);
    drop_graphs(s);
    return static_cast<int>(s->bdys.size()) - 1;
}

int hp_boundary_add_gridded(hp_scheme* s, const hp_bdy_gridded* conf, const double* series) {
    if (!s || !conf || !series) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (conf->entries < 1 || conf->rows < 1 || conf->cols < 1 || !(conf->interval > 0.0) || !(conf->resolution > 0.0))
        return fail(HP_ERR_INVALID, "bad gridded boundary configuration");
    Boundary b; b.kind = 1; b.gridded = *conf;
    const size_t frame = static_cast<size_t>(conf->rows * conf->cols);
    // the kernel may index frame `entries` (CLBoundaries.clc:229): keep one zero frame of padding
    int rc = upload_series(s, series, frame * conf->entries, frame * (conf->entries + 1), &b.series);
    if (rc) return rc;
    s->bdys.push_back(b);
    drop_graphs(s);
    return static_cast<int>(s->bdys.size()) - 1;
}

int hp_boundary_add_cell(hp_scheme* s, const hp_bdy_cell* conf, const uint64_t* relations, const double* series) {
    if (!s || !conf || !relations || !series) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (conf->entries < 2 || !(conf->interval > 0.0)) return fail(HP_ERR_INVALID, "a boundary timeseries is too short");
    Boundary b; b.kind = 2; b.cell = *conf; b.count = conf->relations;
    // the kernel reads entry base+1 (CLBoundaries.clc:44,49): one padding entry
    int rc = upload_series(s, series, 4 * static_cast<size_t>(conf->entries), 4 * (static_cast<size_t>(conf->entries) + 1), &b.series);
    if (rc) return rc;
    std::vector<long long> local(conf->relations > 0 ? conf->relations : 1, -1);
    const hp::Grid& g = s->grid;
    for (uint64_t i = 0; i < conf->relations; ++i) {
        const uint64_t gx = relations[i] % static_cast<uint64_t>(g.cols), gy = relations[i] / static_cast<uint64_t>(g.cols);
        if (gy >= static_cast<uint64_t>(g.grows)) { cudaFree(b.series); return fail(HP_ERR_INVALID, "boundary cell %llu outside the domain", (unsigned long long)relations[i]); }
        const long long y = static_cast<long long>(gy) - g.gy0;
        local[i] = (y >= 0 && y < g.rows) ? y * g.pitch + static_cast<long long>(gx) : -1;
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b.relations), local.size() * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.relations, local.data(), local.size() * sizeof(long long), cudaMemcpyHostToDevice, s->ex->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->ex->stream);
    if (e != cudaSuccess) {
        cudaFree(b.series); cudaFree(b.relations);
        return fail(e == cudaErrorMemoryAllocation ? HP_ERR_OOM : HP_ERR_CUDA, "boundary relations upload failed: %s", cudaGetErrorString(e));
    }
    s->bdys.push_back(b);
    drop_graphs(s);
    return static_cast<int>(s->bdys.size()) - 1;
}

int hp_scheme_upload_cells(hp_scheme* s, const void* states, const void* bed, const void* manning) {
    if (!s || !states || !bed || !manning) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    const hp::Grid& g = s->grid;
    const size_t w = static_cast<size_t>(g.cols) * s->rb, dp = static_cast<size_t>(g.pitch) * s->rb;
    HP_CUDA(cudaMemcpy2DAsync(s->bed, dp, bed, w, w, g.rows, cudaMemcpyHostToDevice, s->ex->stream));
    HP_CUDA(cudaMemcpy2DAsync(s->manning, dp, manning, w, w, g.rows, cudaMemcpyHostToDevice, s->ex->stream));
    s->use_alt = false;
    const int rc = rows_to_device(s, states, 0, g.rows, true);
    // strips over peer memory: nobody's first cell update (whose edge rows go straight into the neighbours' halo rows)
    // may start before every strip's upload has landed -- a barrier in stream order, no host synchronisation
    if (rc == HP_OK && s->peers) s->launches += s->K->peer_barrier(s->peer_args[0], s->ex->stream);
    return rc;
}

int hp_scheme_download_cells(hp_scheme* s, void* states) {
    if (!s || !states) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    return rows_to_host(s, src_planes(s, s->use_alt), states, 0, s->grid.rows);
}

int hp_scheme_download_both(hp_scheme* s, void* a, void* b) {
    if (!s || !a || !b) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    int rc = rows_to_host(s, s->A, a, 0, s->grid.rows);
    if (rc) return rc;
    return rows_to_host(s, s->B, b, 0, s->grid.rows);
}

int hp_scheme_read_rows(hp_scheme* s, uint64_t first_row, uint64_t row_count, void* states) {
    if (!s || !states) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    if (first_row + row_count > static_cast<uint64_t>(s->grid.rows)) return fail(HP_ERR_INVALID, "row range outside the scheme");
    HP_CUDA(cudaSetDevice(s->ex->device));
    return rows_to_host(s, src_planes(s, s->use_alt), states, static_cast<int>(first_row), static_cast<int>(row_count));
}

int hp_scheme_write_rows(hp_scheme* s, uint64_t first_row, uint64_t row_count, const void* states) {
    if (!s || !states) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    if (first_row + row_count > static_cast<uint64_t>(s->grid.rows)) return fail(HP_ERR_INVALID, "row range outside the scheme");
    HP_CUDA(cudaSetDevice(s->ex->device));
    return rows_to_device(s, states, static_cast<int>(first_row), static_cast<int>(row_count), false);
}

int hp_scheme_derive_raster(hp_scheme* s, uint32_t value, double nodata, double* out) {
    if (!s || !out) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    if (value > 11u) return fail(HP_ERR_INVALID, "unknown raster value code %u", value);
    HP_CUDA(cudaSetDevice(s->ex->device));
    const hp::Grid& g = s->grid;
    const hp::KernelTable& K = hp::strict_kernels();              // rounded as written, like the reference's host loop
    const size_t row_bytes = static_cast<size_t>(g.cols) * sizeof(double);
    const int chunk = static_cast<int>(s->staging_bytes / row_bytes) > 0 ? static_cast<int>(s->staging_bytes / row_bytes) : 1;
    const int own = g.own_y1 - g.own_y0;
    for (int done = 0; done < own; done += chunk) {               // from the northern edge of the owned rows downwards
        const int n = own - done < chunk ? own - done : chunk;
        s->launches += K.derive_raster(static_cast<int>(s->rb), src_planes(s, s->use_alt), s->bed, static_cast<double*>(s->staging), g,
                                       g.own_y1 - done, n, static_cast<int>(value), s->cfg.delta, nodata, s->ex->stream);
        HP_CUDA(cudaMemcpyAsync(out + static_cast<size_t>(done) * g.cols, s->staging, static_cast<size_t>(n) * row_bytes,
                                cudaMemcpyDeviceToHost, s->ex->stream));
    }
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    HP_CUDA(cudaGetLastError());
    return HP_OK;
}

int hp_scheme_set_target_time(hp_scheme* s, double target) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    return write_clock_field(s, 3, target);
}
int hp_scheme_force_timestep(hp_scheme* s, double timestep) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    return write_clock_field(s, 1, timestep);
}
int hp_scheme_set_clock(hp_scheme* s, double time, double timestep, double th) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    int rc = write_clock_field(s, 0, time);
    if (!rc) rc = write_clock_field(s, 1, timestep);
    if (!rc) rc = write_clock_field(s, 2, th);
    return rc;
}

int hp_scheme_update_timestep(hp_scheme* s) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    // the reference's reduction kernel stays bound to "Cell states" (Q1); for MUSCL-Hancock that
    // buffer is the only one.  Here MUSCL-Hancock ping-pongs, so read the current buffer.
    hp::StepArgs a = base_args(s, false);
    if (s->cfg.scheme == HP_SCHEME_MUSCL_HANCOCK || !(s->cfg.quirks & HP_QUIRK_REDUCE_BUFFER_A)) a.src = src_planes(s, s->use_alt);
    else a.src = s->A;
    s->launches += s->K->reduce_only(static_cast<int>(s->rb), a, s->ex->stream);
    if (s->peers) {
        hp::PeerArgs p = s->peer_args[0];
        p.bytes[0] = p.bytes[1] = 0;                // no rows to move: only the maximum and the clock
        s->launches += s->K->peer_exchange(static_cast<int>(s->rb), p, a, 1, s->ex->stream);
        HP_CUDA(cudaGetLastError());
        return HP_OK;
    }
    if (s->comm) { const char* err = hp::comm_allreduce_max(s->comm, s->max_bits, s->ex->stream); if (err) return fail(HP_ERR_NCCL, "allreduce: %s", err); }
    s->launches += s->K->update_timestep(static_cast<int>(s->rb), a, s->ex->stream);
    HP_CUDA(cudaGetLastError());
    return HP_OK;
}

int hp_scheme_reset_counters(hp_scheme* s) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    int rc = write_clock_field(s, 4, 0.0);
    if (rc) return rc;
    HP_CUDA(cudaMemsetAsync(static_cast<char*>(s->clock) + 5 * s->rb, 0, 2 * sizeof(unsigned int), s->ex->stream));
    return HP_OK;
}

// With a communicator the halo send/recv, the all-reduce and the fork/join onto the communication stream are
// captured too (NCCL records its kernels into the graph), so a strip costs one graph launch per 16 iterations.
// Measured on 4 GPUs: that pays on small strips (4096 x 4096: +1.5 %), but a captured exchange no longer overlaps
// the interior kernel of a large strip (32768 x 4096: 2.74 ms per step against 2.38 ms launched directly), so
// large strips are launched directly -- their launch latency is hidden anyway.  `small_strip` is the same on every
// rank (hp_scheme_attach_comm), so all ranks issue the same NCCL sequence.
static bool uses_graphs(const hp_scheme* s) {
    return !(s->cfg.options & HP_OPT_NO_GRAPH) && !s->strip_timing && (s->peers || s->comm == nullptr || s->small_strip);
}

int hp_scheme_prepare_graphs(hp_scheme* s) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (!uses_graphs(s)) return HP_OK;
    for (int slot = 0; slot < 2; ++slot) {
        if (s->graph_exec[slot]) continue;
        const int rc = build_graph(s, slot, slot == 1 ? kGraphPairs : 1);
        if (rc != HP_OK) return rc;
        // upload the executable graph now, so that its first launch does not pay for it either
        HP_CUDA(cudaGraphUpload(s->graph_exec[slot], s->ex->stream));
    }
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    return HP_OK;
}

int hp_scheme_iterate(hp_scheme* s, uint32_t iterations) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    uint32_t left = iterations;
    const bool use_graph = uses_graphs(s);
    int rc = HP_OK;
    auto direct = [&]() {
        int launched = 0;
        rc = enqueue_iteration(s, s->use_alt, &launched);
        s->launches += launched; s->use_alt = !s->use_alt; ++s->iterations; --left;
        return rc;
    };
    if (use_graph && left >= 2) {
        if (s->use_alt && direct() != HP_OK) return rc;
        for (int slot = 1; slot >= 0; --slot) {
            const uint32_t per = 2u * (slot == 1 ? kGraphPairs : 1);
            if (left < per) continue;
            // built on first use unless hp_scheme_prepare_graphs did it (boundaries and the communicator must be attached
            // before: adding either drops the graphs)
            if (!s->graph_exec[slot] && (rc = build_graph(s, slot, static_cast<int>(per / 2))) != HP_OK) return rc;
            while (left >= per) {
                HP_CUDA(cudaGraphLaunch(s->graph_exec[slot], s->ex->stream));
                s->launches += s->graph_launches[slot]; s->iterations += per; left -= per;
            }
        }
    }
    while (left > 0) if (direct() != HP_OK) return rc;
    return HP_OK;
}

int hp_scheme_sync(hp_scheme* s) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    // blocks WITHOUT the handle's lock (the stream handle never changes): the reference's main thread waits here
    // while its worker thread goes on enqueueing (src/Schemes/CSchemeGodunov.cpp:1116-1141)
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    HP_CUDA(cudaGetLastError());
    return HP_OK;
}

int hp_scheme_read_stats(hp_scheme* s, hp_scheme_stats* out) {
    if (!s || !out) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    char buf[64];
    HP_CUDA(cudaMemcpyAsync(buf, s->clock, 64, cudaMemcpyDeviceToHost, s->ex->stream));
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    if (s->rb == 8) {
        hp::Clock<double> c; memcpy(&c, buf, sizeof(c));
        out->time = c.time; out->timestep = c.timestep; out->time_hydrological = c.time_hydro; out->time_target = c.time_target;
        out->batch_timesteps = c.batch_timesteps; out->batch_successful = c.batch_successful; out->batch_skipped = c.batch_skipped;
    } else {
        hp::Clock<float> c; memcpy(&c, buf, sizeof(c));
        out->time = c.time; out->timestep = c.timestep; out->time_hydrological = c.time_hydro; out->time_target = c.time_target;
        out->batch_timesteps = c.batch_timesteps; out->batch_successful = c.batch_successful; out->batch_skipped = c.batch_skipped;
    }
    out->iterations = s->iterations; out->kernel_launches = s->launches; out->use_alternate = s->use_alt ? 1u : 0u; out->reserved0 = 0;
    if (s->peers) {
        unsigned int err = 0;
        HP_CUDA(cudaMemcpy(&err, &s->peer_box->error, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) return fail(HP_ERR_PEER, "a peer strip did not arrive within the time limit of the exchange kernel");
    }
    return HP_OK;
}

int hp_comm_unique_id(void* id_out) {
    if (!id_out) return fail(HP_ERR_INVALID, "id_out is null");
    const char* err = hp::comm_unique_id(id_out);
    if (err) return fail(HP_ERR_NCCL, "%s", err);
    return HP_OK;
}

int hp_scheme_attach_comm(hp_scheme* s, const void* id, int rank, int world_size) {
    if (!s || !id) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(HP_ERR_INVALID, "bad rank/world size");
    if (s->comm) return fail(HP_ERR_INVALID, "a communicator is already attached");
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (world_size == 1) return HP_OK;
    // Which iteration shape a strip runs (one launch + one NCCL group, or edges / exchange / interior) decides the ORDER
    // of its NCCL calls, so every rank must decide alike: from the largest strip of an even split of the global rows,
    // never from the local row count (edge strips hold one halo, inner strips two, the first strips one row more).
    const long long halo = s->cfg.scheme == HP_SCHEME_MUSCL_HANCOCK ? 2 : 1;
    const long long most_rows = (static_cast<long long>(s->cfg.global_rows) + world_size - 1) / world_size + 2 * halo;
    long long small_cells = kSmallStripCells;
    if (const char* t = getenv("HIPIMS_SMALL_STRIP_CELLS")) small_cells = atoll(t);    // test hook (tests/test_multigpu.py)
    s->small_strip = most_rows * s->grid.cols <= small_cells;
    s->world = world_size;
    const char* err = hp::comm_create(&s->comm, id, rank, world_size);
    if (err) return fail(HP_ERR_NCCL, "%s", err);
    cudaError_t e = cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_edges, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming);
    for (auto& ev : s->ev_t) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e != cudaSuccess) {
        // leave no half-attached communicator behind
        hp::comm_destroy(s->comm); s->comm = nullptr; s->world = 1;
        if (s->comm_stream) { cudaStreamDestroy(s->comm_stream); s->comm_stream = nullptr; }
        if (s->ev_edges) { cudaEventDestroy(s->ev_edges); s->ev_edges = nullptr; }
        if (s->ev_halo) { cudaEventDestroy(s->ev_halo); s->ev_halo = nullptr; }
        for (auto& ev : s->ev_t) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
        return fail(HP_ERR_CUDA, "attaching the communicator failed: %s", cudaGetErrorString(e));
    }
    drop_graphs(s);
    return HP_OK;
}

// What a strip tells its peers (HP_PEER_BLOB_BYTES, opaque to the caller)
struct PeerBlob {
    uint64_t magic;
    int64_t pid;
    int32_t device, rows, own_y0, own_y1, pitch, cols, real_bytes, scheme;
    uint64_t plane_bytes, global_rows;
    void* block; void* box;                     // valid inside the exporting process
    cudaIpcMemHandle_t h_block, h_box;          // ... and for every other process
};
static_assert(sizeof(PeerBlob) <= HP_PEER_BLOB_BYTES, "HP_PEER_BLOB_BYTES");
static_assert(sizeof(hp::PeerBox) <= 1024, "mailbox allocation");
constexpr uint64_t kPeerMagic = 0x4850504545523031ull;   // "HPPEER01"

static int ensure_peer_box(hp_scheme* s) {
    if (s->peer_box) return HP_OK;
    void* p = nullptr;
    HP_CUDA(cudaMalloc(&p, 1024));
    HP_CUDA(cudaMemset(p, 0, 1024));
    s->peer_box = static_cast<hp::PeerBox*>(p);
    return HP_OK;
}

int hp_scheme_peer_export(hp_scheme* s, void* blob_out) {
    if (!s || !blob_out) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    int rc = ensure_peer_box(s);
    if (rc) return rc;
    PeerBlob b{};
    b.magic = kPeerMagic; b.pid = static_cast<int64_t>(getpid()); b.device = s->ex->device;
    b.rows = s->grid.rows; b.own_y0 = s->grid.own_y0; b.own_y1 = s->grid.own_y1; b.pitch = s->grid.pitch; b.cols = s->grid.cols;
    b.real_bytes = static_cast<int32_t>(s->rb); b.scheme = static_cast<int32_t>(s->cfg.scheme);
    b.plane_bytes = s->plane_bytes; b.global_rows = s->cfg.global_rows;
    b.block = s->block; b.box = s->peer_box;
    HP_CUDA(cudaIpcGetMemHandle(&b.h_block, s->block));
    HP_CUDA(cudaIpcGetMemHandle(&b.h_box, s->peer_box));
    memset(blob_out, 0, HP_PEER_BLOB_BYTES);
    memcpy(blob_out, &b, sizeof(b));
    return HP_OK;
}

int hp_scheme_attach_peers(hp_scheme* s, int rank, int world_size, const void* blobs) {
    if (!s || !blobs) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    if (world_size < 1 || world_size > hp::kMaxPeers || rank < 0 || rank >= world_size) return fail(HP_ERR_INVALID, "bad rank/world size (at most %d strips)", hp::kMaxPeers);
    if (s->peers) return fail(HP_ERR_INVALID, "peers are already attached");
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (world_size == 1) return HP_OK;
    int rc = ensure_peer_box(s);
    if (rc) return rc;
    const int halo = s->cfg.scheme == HP_SCHEME_MUSCL_HANCOCK ? 2 : 1;
    std::vector<PeerBlob> all(world_size);
    for (int r = 0; r < world_size; ++r) {
        memcpy(&all[r], static_cast<const char*>(blobs) + static_cast<size_t>(r) * HP_PEER_BLOB_BYTES, sizeof(PeerBlob));
        const PeerBlob& b = all[r];
        if (b.magic != kPeerMagic) return fail(HP_ERR_INVALID, "blob %d is not a peer description", r);
        if (b.pitch != s->grid.pitch || b.cols != s->grid.cols || b.real_bytes != static_cast<int32_t>(s->rb) ||
            b.scheme != static_cast<int32_t>(s->cfg.scheme) || b.global_rows != s->cfg.global_rows)
            return fail(HP_ERR_INVALID, "strip %d was created with a different geometry, precision or scheme", r);
    }
    if (all[rank].box != s->peer_box || all[rank].pid != static_cast<int64_t>(getpid())) return fail(HP_ERR_INVALID, "blob %d is not this strip's own", rank);
    // neighbours must hold the halo rows my edge rows go into
    if (rank > 0 && all[rank - 1].rows - all[rank - 1].own_y1 != halo) return fail(HP_ERR_INVALID, "the southern neighbour has no northern halo of %d rows", halo);
    if (rank + 1 < world_size && all[rank + 1].own_y0 != halo) return fail(HP_ERR_INVALID, "the northern neighbour has no southern halo of %d rows", halo);
    if ((rank > 0 && s->grid.own_y0 != halo) || (rank + 1 < world_size && s->grid.rows - s->grid.own_y1 != halo) || s->grid.own_y1 - s->grid.own_y0 < halo)
        return fail(HP_ERR_INVALID, "this strip's halo rows do not match its position among %d strips", world_size);
    // map every peer's mailbox, the neighbours' plane blocks too
    std::vector<char*> blocks(world_size, nullptr);
    auto map = [&](const PeerBlob& b, bool block, void** out) -> int {
        if (b.pid == static_cast<int64_t>(getpid())) {                         // same process: the pointer itself, peer access enabled
            if (b.device != s->ex->device) {
                int can = 0;
                HP_CUDA(cudaDeviceCanAccessPeer(&can, s->ex->device, b.device));
                if (!can) return fail(HP_ERR_PEER, "device %d cannot access the memory of device %d", s->ex->device, b.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(HP_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
            *out = block ? b.block : b.box;
            return HP_OK;
        }
        const cudaError_t e = cudaIpcOpenMemHandle(out, block ? b.h_block : b.h_box, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(HP_ERR_PEER, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
        s->peer_mapped.push_back(*out);
        return HP_OK;
    };
    hp::PeerArgs base{};
    base.rank = rank; base.world = world_size;
    for (int r = 0; r < world_size; ++r) {
        void* p = nullptr;
        if (r == rank) p = s->peer_box; else if ((rc = map(all[r], false, &p))) return rc;
        base.box[r] = static_cast<hp::PeerBox*>(p);
        if (r == rank - 1 || r == rank + 1) { void* b = nullptr; if ((rc = map(all[r], true, &b))) return rc; blocks[r] = static_cast<char*>(b); }
    }
    // plane q of buffer A / B inside a ten-plane block (alloc_block): A.eta A.qx A.qy A.emax | zb n | B.eta B.qx B.qy B.emax;
    // PeerArgs order is eta emax qx qy
    static const int kPlaneA[4] = {0, 3, 1, 2}, kPlaneB[4] = {6, 9, 7, 8};
    const size_t row = static_cast<size_t>(s->grid.pitch) * s->rb;
    for (int alt = 0; alt < 2; ++alt) {
        hp::PeerArgs a = base;
        const int* mine_planes = alt ? kPlaneA : kPlaneB;         // the step's DESTINATION: B when it read A (alt == 0)
        for (int n = 0; n < 2; ++n) {
            const int r = n == 0 ? rank - 1 : rank + 1;
            if (r < 0 || r >= world_size) { a.bytes[n] = 0; continue; }
            a.bytes[n] = row * halo;
            // my edge rows: the lowest owned rows go south, the highest north; they land in the neighbour's halo rows on
            // the side facing me
            const size_t my_row = n == 0 ? static_cast<size_t>(s->grid.own_y0) : static_cast<size_t>(s->grid.own_y1 - halo);
            const size_t their_row = n == 0 ? static_cast<size_t>(all[r].own_y1) : static_cast<size_t>(all[r].own_y0 - halo);
            for (int q = 0; q < 4; ++q) {
                a.src[n][q] = s->block + static_cast<size_t>(mine_planes[q]) * s->plane_bytes + my_row * row;
                a.dst[n][q] = blocks[r] + static_cast<size_t>(mine_planes[q]) * all[r].plane_bytes + their_row * row;
            }
        }
        s->peer_args[alt] = a;
    }
    // rendezvous: tell every peer that its mailbox is mapped here, wait until all of them have told me
    s->K->peer_hello(base, s->ex->stream);
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        unsigned long long hello[hp::kMaxPeers] = {};
        HP_CUDA(cudaMemcpy(hello, s->peer_box->hello, sizeof(hello), cudaMemcpyDeviceToHost));
        int seen = 0;
        for (int r = 0; r < world_size; ++r) seen += hello[r] != 0ull;
        if (seen == world_size) break;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0)
            return fail(HP_ERR_PEER, "only %d of %d strips attached within 60 s", seen, world_size);
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    s->peers = true; s->peer_rank = rank; s->world = world_size;
    drop_graphs(s);
    return HP_OK;
}

int hp_scheme_strip_timing(hp_scheme* s, int enable) {
    if (!s) return fail(HP_ERR_INVALID, "scheme is null");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    HP_CUDA(cudaSetDevice(s->ex->device));
    if (!s->comm) return fail(HP_ERR_INVALID, "strip timing needs an attached communicator");
    HP_CUDA(cudaStreamSynchronize(s->ex->stream));
    s->strip_timing = enable != 0;
    for (double& t : s->t_phase) t = 0.0;
    s->t_count = 0;
    return HP_OK;
}

int hp_scheme_read_strip_phases(hp_scheme* s, double* ms_per_iteration, uint64_t* iterations) {
    if (!s || !ms_per_iteration || !iterations) return fail(HP_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(s->mu);
    *iterations = s->t_count;
    for (int i = 0; i < HP_STRIP_PHASES; ++i) ms_per_iteration[i] = s->t_count ? s->t_phase[i] / static_cast<double>(s->t_count) : 0.0;
    return HP_OK;
}

}  // extern "C"
