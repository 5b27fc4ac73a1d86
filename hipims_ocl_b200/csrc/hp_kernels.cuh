// hp_kernels.cuh -- kernel-side shared declarations (device structs + the launch interface the
// executor calls).  hp_kernels.cu is compiled twice, with HP_NS=hp_strict (-fmad=false) and
// HP_NS=hp_fast (FMA contraction on); this header is what hp_executor.cu sees of both.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace hp {

// Device-resident clock: the reference's "Time", "Timestep", "Time (hydrological)", "Target time
// (sync)" and batch counter buffers (src/Schemes/CSchemeGodunov.cpp:803-873) in one record, kept
// in the working precision like the reference's buffers.
template <class R> struct Clock {
    R time, timestep, time_hydro, time_target, batch_timesteps;
    unsigned int batch_successful, batch_skipped;
};

// Structure-of-arrays planes of one state buffer ("Cell states" / "Cell states (alternate)").
struct Planes { void *eta, *emax, *qx, *qy; };

// Geometry of the rows one scheme holds (a whole domain or a row strip with halo rows).
struct Grid {
    int cols;        // DOMAIN_COLS
    int rows;        // local rows (owned + halo)
    int pitch;       // elements between consecutive rows of a plane
    int grows;       // DOMAIN_ROWS of the whole domain
    int gy0;         // global row index of local row 0
    int own_y0;      // first owned local row
    int own_y1;      // one past the last owned local row
};

// Scalars every kernel needs, in double; converted to the working precision on the device side.
struct ParamsD {
    double eps, eps10, delta, courant, end_time, fixed_dt;
    int dynamic, friction, simplified_speed;
    int dt0_keep;             // Godunov, dt <= 0: nothing is written (HP_QUIRK_GODUNOV_DT0_KEEP)
};

enum ReduceMode { kReduceNone = 0, kReduceSrc = 1, kReduceDst = 2 };

// One row range [y0, y1) of local rows to update in a launch (edge rows first, interior after).
struct StepArgs {
    Planes src, dst;
    const void *bed, *manning;
    void* clock;              // Clock<R>*
    unsigned long long* max_bits;  // running maximum of the wave speed, as ordered bits
    unsigned int* ticket;     // CTA arrival counter for the in-kernel finaliser
    Grid grid;
    ParamsD params;
    int y0, y1;
    int reduce_mode;          // ReduceMode
    int finalize;             // 1: last CTA runs the time controller (single device)
    unsigned int total_ctas;  // CTAs that will arrive on `ticket` before the finaliser runs
    int march_runs;           // marching kernels: runs of (strip group, row) units per CTA, interleaved across the grid
};

// ---------------------------------------------------------------------------------------------
// Row strips over PEER MEMORY (NVLink 5 / NVSwitch): what hp_comm.cpp does with ncclSend/ncclRecv and ncclAllReduce --
// halo rows to the neighbouring strips, the maximum of the wave speed over all strips -- done by ONE kernel with plain
// stores into the peers' memory, followed in the same kernel by the time controller.
//
// Every strip owns a mailbox that its peers write into:
//   vmax[it & 1][r]  wave-speed bits of strip r for iteration `it` (double-buffered: strip r can be one iteration ahead)
//   sig[r]           the last iteration strip r has published -- written AFTER its halo rows and its vmax, behind a
//                    system-scope fence, so "sig[r] >= it" means both have landed
// and keeps its own iteration count next to it.  A strip never runs more than one iteration ahead of a peer (it waits for
// every sig of iteration `it` before it advances its clock), which is what makes two vmax slots and the ping-pong halo
// rows enough.  A peer that does not show up within kPeerSpinCycles sets `error` instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
constexpr long long kPeerSpinCycles = 8000000000ll;          // ~4 s of SM clock
struct PeerBox {
    unsigned long long sig[kMaxPeers];
    unsigned long long vmax[2][kMaxPeers];
    unsigned long long hello[kMaxPeers];   // attach rendezvous: strip r has mapped this mailbox
    unsigned long long bar[kMaxPeers];     // stream-ordered barrier (after an upload): the last epoch strip r has reached
    unsigned long long bar_epoch;     // local: barriers this strip has passed
    unsigned long long iter;          // local: iterations this strip has completed
    unsigned int ticket;              // local: CTA arrival counter of the exchange kernel
    unsigned int error;               // local: a peer did not arrive in time
};
struct PeerArgs {
    const void* src[2][4];            // [south, north][eta emax qx qy]: first of my `halo` owned edge rows in this iteration's dst
    void* dst[2][4];                  // ... and the first halo row they land in, in the neighbour's dst buffer (peer memory)
    unsigned long long bytes[2];      // bytes per plane and neighbour (halo rows x pitch x sizeof(real)); 0 = no neighbour
    PeerBox* box[kMaxPeers];          // every strip's mailbox by rank (mine included)
    int rank, world;
};

struct BdyUniformArgs {
    Planes state; const void* bed; const void* clock; const void* series;  // series: entries x {t, value}
    Grid grid; unsigned int entries, definition; double interval, length; int cover_x, cover_y;
};
struct BdyGriddedArgs {
    Planes state; const void* clock; const void* series;
    Grid grid; double interval, resolution, offset_x, offset_y, delta;
    unsigned long long entries, definition, grows, gcols; int cover_x, cover_y;
};
struct BdyCellArgs {
    Planes state; const void* bed; const void* clock; const void* series; const long long* relations;  // LOCAL ids
    Grid grid; ParamsD params; unsigned long long entries, count; double interval, length;
    unsigned int def_depth, def_discharge;
};

// Four TMA descriptors (eta, qx, qy of the source buffer and zb), opaque 128-byte CUtensorMaps.
struct alignas(64) TmaMapsPOD { unsigned char bytes[4][128]; };
// haloed tile box of the TMA-staged kernels: 64 x 8 cells plus halo; the box must start on a
// 16-byte boundary of the inner dimension, so it carries 16 / real_bytes halo columns per side
constexpr int kTmaTileX = 32, kTmaTileY = 8;        // Godunov tile
constexpr int kTmaTileXMH = 64;                     // MUSCL-Hancock tile (halo of two)
constexpr int tma_tile_x(int halo) { return halo == 2 ? kTmaTileXMH : kTmaTileX; }
constexpr int tma_box_w(int real_bytes, int halo) { return tma_tile_x(halo) + 2 * (16 / real_bytes); }
constexpr int tma_box_h(int halo) { return kTmaTileY + 2 * halo; }

// Marching kernels (hp_march_kernels.cuh): one warp per 32-column strip, rows streamed through a
// per-warp TMA ring of single rows x six planes; bytes[0] is a 3-D descriptor {cols, rows, 10 planes} over the
// scheme's plane block (A.eta A.qx A.qy A.emax | zb n | B.eta B.qx B.qy B.emax), box = {march_box_w, 1, 6}:
// plane coordinate 0 reads buffer A + zb + n, plane coordinate 4 reads zb + n + buffer B.
struct alignas(64) TmaMaps6POD { unsigned char bytes[1][128]; };
constexpr int kMarchWarps = 4;                      // warps (= adjacent strips) per CTA
constexpr int march_use(int real_bytes, int halo) { return (halo == 2 || real_bytes == 4) ? 28 : 30; }   // cells updated per warp row
constexpr int march_box_w(int real_bytes, int halo) {
    const int a16 = 16 / real_bytes, padl = ((-halo) % a16 + a16) % a16;
    return (32 + padl + a16 - 1) / a16 * a16;
}
// "Wide" marching kernels (hp_march_pair.cuh): two adjacent columns per lane, one warp per 64-column strip of which
// 60 are updated; box = {wide_box_w, 1, 6} of the same descriptor layout.
constexpr int kWideUse = 60;
constexpr int wide_box_w(int real_bytes) { return real_bytes == 8 ? 64 : 68; }

// The launch interface of one compiled flavour (strict / fast).
struct KernelTable {
    // TMA-staged step (fast flavour only, NULL otherwise); returns -1 if the scheme has no such kernel
    int (*step_tma)(int scheme, int real_bytes, const StepArgs& a, const TmaMapsPOD* maps, int sm_count, cudaStream_t st);
    // marching step (fast flavour only, NULL otherwise); returns -1 if the scheme has no such kernel
    // alt: bit 0 = the step reads buffer B, bits 1-2 = kernel width mode (0 default, 1 one column per lane, 2 two columns)
    int (*step_march)(int scheme, int real_bytes, const StepArgs& a, const TmaMaps6POD* maps, int alt, int sm_count, cudaStream_t st);
    // columns of the TMA box the marching kernel of (scheme, precision, width mode) expects
    int (*march_box_w)(int scheme, int real_bytes, int mode);
    // returns the number of kernels launched
    int (*step)(int scheme, int real_bytes, const StepArgs& a, cudaStream_t st);
    int (*reduce_only)(int real_bytes, const StepArgs& a, cudaStream_t st);       // tst_Reduce
    int (*advance)(int real_bytes, const StepArgs& a, cudaStream_t st);           // tst_Advance_Normal
    int (*update_timestep)(int real_bytes, const StepArgs& a, cudaStream_t st);   // tst_UpdateTimestep
    // halo rows -> peers, wave-speed maximum over all strips, then tst_Advance_Normal (or tst_UpdateTimestep), one launch
    int (*peer_exchange)(int real_bytes, const PeerArgs& p, const StepArgs& a, int update_only, cudaStream_t st);
    int (*peer_hello)(const PeerArgs& p, cudaStream_t st);                        // attach rendezvous
    int (*peer_barrier)(const PeerArgs& p, cudaStream_t st);                      // all strips have reached this point of their streams
    int (*bdy_uniform)(int real_bytes, const BdyUniformArgs& a, cudaStream_t st);
    int (*bdy_gridded)(int real_bytes, const BdyGriddedArgs& a, cudaStream_t st);
    int (*bdy_cell)(int real_bytes, const BdyCellArgs& a, cudaStream_t st);
    // AoS (host layout, staged on the device) <-> SoA planes, whole rows
    int (*aos_to_soa)(int real_bytes, const void* aos, Planes dst, Grid g, int row0, int nrows, cudaStream_t st);
    int (*soa_to_aos)(int real_bytes, Planes src, void* aos, Grid g, int row0, int nrows, cudaStream_t st);
    int (*copy_plane_rows)(int real_bytes, const void* dense, void* plane, Grid g, int row0, int nrows, cudaStream_t st);
    // output rasters (CRasterDataset::domainToRaster): `nrows` local rows ending below row `row_end`, written
    // NORTH row first into `out` (nrows x cols doubles)
    int (*derive_raster)(int real_bytes, Planes src, const void* bed, double* out, Grid g, int row_end, int nrows, int value,
                         double resolution, double nodata, cudaStream_t st);
};

const KernelTable& strict_kernels();
const KernelTable& fast_kernels();

}  // namespace hp
