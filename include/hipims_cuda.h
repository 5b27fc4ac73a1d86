/*
 * hipims_cuda.h -- C ABI of the B200 executor for the HiPIMS explicit cell-update path.
 *
 * This library replaces the reference's OpenCL executor layer (src/OpenCL/Executors/ --
 * CExecutorControlOpenCL, COCLDevice, COCLProgram, COCLKernel, COCLBuffer and the
 * runtime-compiled .clc sources) underneath the scheme / domain / boundary classes.  The entry
 * points are what CSchemeGodunov / CSchemeMUSCLHancock / CSchemeInertial, CBoundaryCell /
 * CBoundaryUniform / CBoundaryGridded and CDomainLink call the OpenCL wrappers for today; each
 * declaration cites the reference call site it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the binding a maintainer would add to the reference's C++ host.
 *
 * Conventions
 *   - plain C: opaque handles, POD structs, pointers and sizes; no C++ or torch types.
 *   - every function returns HP_OK (0) or a negative hp_status; hp_last_error() returns the
 *     message for the calling thread.  The reference forwards such failures to
 *     model::doError (src/main.cpp:631-652).
 *   - host arrays keep the reference's host layout (src/Domain/CDomain.h:28-33,
 *     src/Schemes/CSchemeGodunov.cpp:832-845): cell states are `cells` 4-vectors
 *     {eta (free-surface level), eta_max, qx, qy} of `real`, bed elevations and Manning
 *     coefficients are `cells` reals; row-major, row 0 is the SOUTHERN edge
 *     (src/Domain/Cartesian/CLDomainCartesian.clc:26-30).  `real` is double or float according
 *     to hp_scheme_config.real_bytes (the reference's floatingPointPrecision switch,
 *     src/OpenCL/Executors/COCLProgram.cpp:381-399).  On the device the library keeps
 *     structure-of-arrays planes; the conversion happens inside upload/download.
 *   - all work of one scheme is issued on one CUDA stream; calls are asynchronous unless stated.
 *   - handles are thread safe: every entry point takes a per-handle lock, so the reference's pattern -- the main
 *     thread polls hp_scheme_read_stats / reads cells back while the scheme's worker thread enqueues batches
 *     (src/Schemes/CSchemeGodunov.cpp:1116-1141, src/Schemes/CScheme.h:137-139) -- is safe; calls on one handle are
 *     serialised in arrival order.  hp_scheme_sync and hp_executor_finish block WITHOUT the lock.  Destroying a
 *     handle while another thread still uses it is the caller's error, as with any C handle.
 *   - there is NO CPU fallback: without a CUDA device every call fails with HP_ERR_NO_DEVICE.
 */
#ifndef HIPIMS_CUDA_H
#define HIPIMS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HP_ABI_VERSION 2

typedef enum hp_status {
    HP_OK = 0,
    HP_ERR_NO_DEVICE = -1,
    HP_ERR_INVALID = -2,
    HP_ERR_CUDA = -3,
    HP_ERR_NCCL = -4,
    HP_ERR_OOM = -5,
    HP_ERR_PEER = -6      /* a peer strip did not arrive in time (row strips over peer memory) */
} hp_status;

typedef struct hp_executor hp_executor; /* one CUDA device + stream: CExecutorControlOpenCL + COCLDevice */
typedef struct hp_scheme hp_scheme;     /* program + kernels + buffers one CScheme* owns                */

/* <scheme name="..."> (src/Schemes/CScheme.cpp:148-165) */
enum { HP_SCHEME_GODUNOV = 0, HP_SCHEME_MUSCL_HANCOCK = 1, HP_SCHEME_INERTIAL = 2 };

/* Reference behaviours that change results; set to reproduce them. */
enum {
    HP_QUIRK_REDUCE_BUFFER_A  = 1u << 0, /* tst_Reduce always reads buffer "Cell states"
                                            (src/Schemes/CSchemeGodunov.cpp:1629,1634)              */
    HP_QUIRK_BDY_COVERAGE     = 1u << 1, /* bdy_Uniform / bdy_Gridded cover floor(n/8)*8 cells per
                                            axis (src/Boundaries/CBoundaryUniform.cpp:294-295)      */
    HP_QUIRK_MH_NO_BOUNDARIES = 1u << 2, /* MUSCL-Hancock iteration applies no boundaries
                                            (src/Schemes/CSchemeMUSCLHancock.cpp:646-680)           */
    HP_QUIRK_GODUNOV_DT0_KEEP = 1u << 3  /* Godunov, timestep <= 0: leave the destination untouched like
                                            gts_cacheEnabled (src/Schemes/CLSchemeGodunov.clc:477-478)
                                            instead of copying the source through like the default
                                            gts_cacheDisabled (:201-206); not in the reference default  */
};

/* Execution options (not part of the numerical contract). */
enum {
    HP_OPT_STRICT_FP   = 1u << 0, /* kernels built without FMA contraction: bit-comparable with the
                                     reference arithmetic evaluated in IEEE order                   */
    HP_OPT_NO_GRAPH    = 1u << 1, /* launch kernels directly instead of replaying CUDA graphs      */
    HP_OPT_NO_TMA      = 1u << 2, /* use the plain-load kernels instead of the TMA-staged ones     */
    HP_OPT_TILE_KERNELS = 1u << 3, /* MUSCL-Hancock: use the TMA tile kernel (CTA-wide 2-D tiles) instead of
                                      the default marching kernel (one warp per column strip)      */
    HP_OPT_MARCH_GODUNOV = 1u << 4, /* Godunov: use a marching kernel instead of the default TMA tile kernel
                                      (3-7 % faster on fully wet domains, slower on mostly dry ones; one
                                      column per lane in fp64, two in fp32 unless pinned below)    */
    HP_OPT_SPLIT_STRIPS = 1u << 5, /* row strips: always split a step into edge rows + interior rows with
                                      the halo exchange overlapped (default only for large strips) */
    HP_OPT_NARROW_MARCH = 1u << 6, /* marching kernels: one column per lane even where the two-column
                                      ("wide") kernel is the default (inertial, fp32 MUSCL-Hancock,
                                      fp32 marching Godunov)                                       */
    HP_OPT_WIDE_MARCH  = 1u << 7   /* marching kernels: two columns per lane even where the one-column
                                      kernel is the default (fp64 MUSCL-Hancock and fp64 marching
                                      Godunov: fewer instructions but too few resident warps,
                                      DESIGN.md 4.4)                                               */
};

/*
 * Everything the reference bakes into the OpenCL program as "#define"s
 * (src/Schemes/CSchemeGodunov.cpp:667-783) plus the domain geometry.
 */
typedef struct hp_scheme_config {
    uint32_t struct_size;     /* sizeof(hp_scheme_config)                                           */
    uint32_t scheme;          /* HP_SCHEME_*                                                        */
    uint32_t real_bytes;      /* 8 = double, 4 = float                                              */
    uint32_t quirks;          /* HP_QUIRK_*                                                         */
    uint32_t options;         /* HP_OPT_*                                                           */
    uint32_t dynamic_timestep;/* TIMESTEP_DYNAMIC (1) or TIMESTEP_FIXED (0)                         */
    uint32_t friction;        /* FRICTION_ENABLED && FRICTION_IN_FLUX_KERNEL                        */
    uint32_t reserved0;
    uint64_t cols;            /* DOMAIN_COLS                                                        */
    uint64_t rows;            /* rows held by THIS scheme (strip rows incl. halo rows)              */
    double   delta;           /* DOMAIN_DELTAX == DOMAIN_DELTAY                                     */
    double   courant;         /* COURANT_NUMBER                                                     */
    double   dry_threshold;   /* VERY_SMALL; QUITE_SMALL is 10x                                     */
    double   end_time;        /* SCHEME_ENDTIME                                                     */
    double   fixed_timestep;  /* TIMESTEP_FIXED                                                     */
    double   initial_timestep;/* first timestep (src/Schemes/CScheme.cpp:49)                        */
    /* Row-strip decomposition (replaces the overlap zones of src/Domain/Links/CDomainLink.cpp):
     * this scheme holds global rows [row_offset - halo_south, row_offset + own_rows + halo_north).
     * A single-device run has global_rows == rows, row_offset == 0 and no halos.                  */
    uint64_t global_rows;     /* DOMAIN_ROWS of the whole domain                                    */
    uint64_t row_offset;      /* global index of the first OWNED row                                */
    uint32_t halo_south;      /* halo rows below the owned rows (0 on the southern strip)           */
    uint32_t halo_north;      /* halo rows above the owned rows (0 on the northern strip)           */
} hp_scheme_config;

/* What CSchemeGodunov::readKeyStatistics pulls back (src/Schemes/CSchemeGodunov.cpp:1817-1850). */
typedef struct hp_scheme_stats {
    double   time;              /* "Time"                                                           */
    double   timestep;          /* "Timestep" (negative = suspended at the sync time)               */
    double   time_hydrological; /* "Time (hydrological)"                                            */
    double   time_target;       /* "Target time (sync)"                                             */
    double   batch_timesteps;   /* "Batch timesteps cumulative"                                     */
    uint32_t batch_successful;  /* "Batch successful iterations"                                    */
    uint32_t batch_skipped;     /* "Batch skipped iterations"                                       */
    uint64_t iterations;        /* iterations scheduled since creation                              */
    uint64_t kernel_launches;   /* CUDA kernels launched (directly or inside graphs) since creation */
    uint32_t use_alternate;     /* next source buffer is "Cell states (alternate)"                  */
    uint32_t reserved0;
} hp_scheme_stats;

/* Boundary configuration records, as src/Boundaries/CLBoundaries.clh:54-82 with doubles. */
typedef struct hp_bdy_uniform {
    uint32_t entries;           /* TimeseriesEntries                                                */
    uint32_t definition;        /* 0 rain-intensity [mm/h], 1 loss-rate [mm/h]                      */
    double   interval;          /* TimeseriesInterval                                               */
    double   length;            /* TimeseriesLength                                                 */
} hp_bdy_uniform;

typedef struct hp_bdy_gridded {
    double   interval, resolution, offset_x, offset_y;
    uint64_t entries;           /* frames                                                           */
    uint64_t definition;        /* 0 rain-intensity [mm/h], 2 mass-flux [m3/s per cell]             */
    uint64_t rows, cols;        /* coarse grid shape                                                */
} hp_bdy_gridded;

typedef struct hp_bdy_cell {
    uint64_t entries;
    double   interval, length;
    uint64_t relations;         /* RelationCount                                                    */
    uint32_t def_depth;         /* 0 ignore, 1 fsl, 2 depth                                         */
    uint32_t def_discharge;     /* 0 ignore, 1 discharge, 2 velocity, 3 volume                      */
} hp_bdy_cell;

/* ---- library ------------------------------------------------------------------------------ */
int         hp_abi_version(void);
const char* hp_last_error(void);
/* CExecutorControlOpenCL::getDeviceCount (src/OpenCL/Executors/CExecutorControlOpenCL.h:55-63) */
int         hp_device_count(int* count);

/* ---- executor: device + queue --------------------------------------------------------------
 * replaces CExecutorControlOpenCL::createDevices / COCLDevice (context + command queue,
 * src/OpenCL/Executors/COCLDevice.cpp:250-300).  `external_stream` is a cudaStream_t to issue
 * work on (0 = create a private non-blocking stream). */
int  hp_executor_create(int device_ordinal, void* external_stream, hp_executor** out);
void hp_executor_destroy(hp_executor* ex);
/* COCLDevice::getDeviceShortName / capability fields (src/OpenCL/Executors/COCLDevice.h:87-111) */
int  hp_executor_describe(hp_executor* ex, char* name, size_t name_len, int* sm_count, size_t* total_mem);
/* COCLDevice::blockUntilFinished (src/OpenCL/Executors/COCLDevice.cpp:374-401) */
int  hp_executor_finish(hp_executor* ex);
/* device-side timing on the executor's stream (the reference only has wall clock,
 * src/General/CBenchmark.cpp:77-79): records events around whatever is enqueued between. */
int  hp_executor_timer_start(hp_executor* ex);
int  hp_executor_timer_stop(hp_executor* ex, float* milliseconds);

/* ---- scheme: prepareAll ----------------------------------------------------------------------
 * replaces CSchemeGodunov::prepareAll -> prepareCode/compileProgram, prepare1OMemory,
 * prepareGeneralKernels, prepare1OKernels (src/Schemes/CSchemeGodunov.cpp:386-478, 789-988) and
 * the MUSCL-Hancock / inertial equivalents (src/Schemes/CSchemeMUSCLHancock.cpp:230-320,
 * src/Schemes/CSchemeInertial.cpp:120-300). */
int  hp_scheme_create(hp_executor* ex, const hp_scheme_config* cfg, hp_scheme** out);
void hp_scheme_destroy(hp_scheme* s);

/* ---- boundaries: CBoundary*::prepareBoundary -----------------------------------------------
 * Boundaries are applied in the order they are added (the reference's order is hash order on an
 * out-of-order queue).  `series` layouts as the reference uploads them:
 *   uniform: entries x {t, value}             (src/Boundaries/CBoundaryUniform.cpp:255-262)
 *   gridded: entries x rows x cols values     (src/Boundaries/CBoundaryGridded.cpp:254-271)
 *   cell:    entries x {t, depth|fsl, Qx, Qy} (src/Boundaries/CBoundaryCell.cpp:383-400), already
 *            divided by the relation count for dischargeValue=total, and `relations` cell IDs
 *            y * cols + x in GLOBAL coordinates (src/Boundaries/CBoundaryCell.cpp:417-421).
 * All series are doubles on the host and converted to `real`. */
int  hp_boundary_add_uniform(hp_scheme* s, const hp_bdy_uniform* conf, const double* series);
int  hp_boundary_add_gridded(hp_scheme* s, const hp_bdy_gridded* conf, const double* series);
int  hp_boundary_add_cell(hp_scheme* s, const hp_bdy_cell* conf, const uint64_t* relations, const double* series);

/* ---- data movement ---------------------------------------------------------------------------
 * hp_scheme_upload_cells: CSchemeGodunov::prepareSimulation's queueWriteAll of "Cell states",
 * "Cell states (alternate)", "Bed elevations", "Manning coefficients" and the clock buffers
 * (src/Schemes/CSchemeGodunov.cpp:1053-1071).  Pointers are HOST memory (pinned or pageable)
 * covering this scheme's `rows` x `cols` cells. */
int  hp_scheme_upload_cells(hp_scheme* s, const void* states, const void* bed, const void* manning);
/* CSchemeGodunov::readDomainAll / saveCurrentState: reads the next source buffer
 * (src/Schemes/CSchemeGodunov.cpp:1671-1679, 1720-1736).  Synchronous. */
int  hp_scheme_download_cells(hp_scheme* s, void* states);
/* both ping-pong buffers, for parity checks of the "leave dst untouched" rule */
int  hp_scheme_download_both(hp_scheme* s, void* states_a, void* states_b);
/* COCLBuffer::queueReadPartial / queueWritePartial on whole rows of the next source buffer, as
 * CDomainLink::pullFromBuffer / pushToBuffer use them (src/Domain/Links/CDomainLink.cpp:168-197,
 * 252-270).  `first_row` is local to this scheme. */
/* Output rasters.  Replaces the full-state read-back + host loop of CRasterDataset::domainToRaster
 * (src/Datasets/CRasterDataset.cpp:180-280; called from CDomainCartesian::writeOutputs,
 * src/Domain/Cartesian/CDomainCartesian.cpp:738-767): derives one output value per cell on the device, in
 * double and with the reference's no-data rules, and writes the OWNED rows in raster order -- NORTH row first,
 * (rows - halos) x cols doubles -- to `out`.  `value` is a model::rasterDatasets::dataValues code
 * (src/Datasets/CRasterDataset.h:33-46): HP_RASTER_*.  Synchronous. */
enum { HP_RASTER_DEPTH = 1, HP_RASTER_FSL = 2, HP_RASTER_VELOCITY_X = 3, HP_RASTER_VELOCITY_Y = 4, HP_RASTER_DISCHARGE_X = 5,
       HP_RASTER_DISCHARGE_Y = 6, HP_RASTER_MAX_DEPTH = 9, HP_RASTER_MAX_FSL = 10, HP_RASTER_FROUDE = 11 };
int  hp_scheme_derive_raster(hp_scheme* s, uint32_t value, double nodata, double* out);

int  hp_scheme_read_rows(hp_scheme* s, uint64_t first_row, uint64_t row_count, void* states);
int  hp_scheme_write_rows(hp_scheme* s, uint64_t first_row, uint64_t row_count, const void* states);

/* ---- clock -----------------------------------------------------------------------------------
 * write of "Target time (sync)" (src/Schemes/CSchemeGodunov.cpp:1164-1180) */
int  hp_scheme_set_target_time(hp_scheme* s, double target);
/* timestep override at the start of a batch (src/Schemes/CSchemeGodunov.cpp:1213-1232) */
int  hp_scheme_force_timestep(hp_scheme* s, double timestep);
/* sets time, timestep and hydrological accumulator (prepareSimulation / tests) */
int  hp_scheme_set_clock(hp_scheme* s, double time, double timestep, double time_hydrological);
/* tst_Reduce + tst_UpdateTimestep after a sync point (src/Schemes/CSchemeGodunov.cpp:1191-1196) */
int  hp_scheme_update_timestep(hp_scheme* s);
/* tst_ResetCounters (src/Schemes/CLDynamicTimestep.clc:151-161) */
int  hp_scheme_reset_counters(hp_scheme* s);

/* ---- the hot loop ---------------------------------------------------------------------------
 * hp_scheme_iterate(n) enqueues n iterations of
 *   boundaries -> cell update -> CFL reduction -> time advance
 * i.e. n x CSchemeGodunov::scheduleIteration (src/Schemes/CSchemeGodunov.cpp:1617-1666) or
 * CSchemeMUSCLHancock::scheduleIteration (src/Schemes/CSchemeMUSCLHancock.cpp:646-680), including
 * the ping-pong toggle of Threaded_runBatch (src/Schemes/CSchemeGodunov.cpp:1287-1301).
 * Asynchronous; the timestep never leaves the device. */
int  hp_scheme_iterate(hp_scheme* s, uint32_t iterations);
/* Builds (and uploads) the CUDA graphs hp_scheme_iterate replays -- 16 iterations and 2 iterations, with the NCCL
 * halo exchange and all-reduce captured inside when a communicator is attached -- so that the first batch does not
 * pay for capture and instantiation.  The analogue of the reference preparing its kernels and their arguments once
 * in prepare1OKernels / prepareGeneralKernels (src/Schemes/CSchemeGodunov.cpp:895-988) instead of on the first
 * iteration.  Call after the boundaries and the communicator are attached (adding either drops the graphs);
 * collective when a communicator is attached (every rank must call it).  Synchronous. */
int  hp_scheme_prepare_graphs(hp_scheme* s);
/* clFlush + clFinish of the batch (src/Schemes/CSchemeGodunov.cpp:1337-1341) */
int  hp_scheme_sync(hp_scheme* s);
/* CSchemeGodunov::readKeyStatistics (src/Schemes/CSchemeGodunov.cpp:1817-1850). Synchronous. */
int  hp_scheme_read_stats(hp_scheme* s, hp_scheme_stats* out);

/* ---- row-strip multi-GPU (replaces src/MPI/CMPIManager.cpp:555-717, 837-889) -------------------
 * One process per GPU.  Rank r owns a contiguous strip of rows; after each cell update the
 * edge rows are sent to the neighbouring strips' halo rows (ncclSend/ncclRecv) while the
 * interior rows are still being computed, and the CFL wave-speed maximum is all-reduced on the
 * device (ncclAllReduce, ncclMax) before the time controller runs.
 * hp_comm_unique_id fills a 128-byte NCCL id on rank 0; the caller broadcasts it (e.g. with
 * torch.distributed) and every rank passes it to hp_scheme_attach_comm. */
#define HP_COMM_ID_BYTES 128
int  hp_comm_unique_id(void* id_out);
int  hp_scheme_attach_comm(hp_scheme* s, const void* id, int rank, int world_size);
/* ---- row strips over peer memory (NVLink 5 / NVSwitch) ------------------------------------------
 * The same exchange without NCCL: after each cell update ONE kernel stores the strip's edge rows straight into the
 * neighbouring strips' halo rows, publishes the strip's wave-speed maximum to every strip, waits for theirs and runs
 * the time controller -- the reference's CDomainLink::pullFromBuffer -> sendOverMPI -> pushToBuffer
 * (src/Domain/Links/CDomainLink.cpp:168-270) plus CMPIManager::reduceTimeData (src/MPI/CMPIManager.cpp:837-889) as
 * stores over NVLink.  An iteration is then two launches (cell update, exchange + clock), replayed from CUDA graphs
 * whatever the strip's size, from one process per GPU (CUDA IPC) or from one process driving several GPUs.
 *   hp_scheme_peer_export   fills an opaque HP_PEER_BLOB_BYTES description of this strip's buffers and mailbox;
 *   the caller gathers the blobs of all strips (torch.distributed all_gather, MPI_Allgather, or by hand in one process);
 *   hp_scheme_attach_peers  maps them and returns once every strip has done the same (a rendezvous: call it after the
 *                           cells have been uploaded, from every strip, concurrently).
 * With peers attached hp_scheme_iterate / hp_scheme_update_timestep use this path; a communicator need not be
 * attached.  Collective like the NCCL path: every strip must enqueue the same iterations, and hp_scheme_upload_cells
 * becomes collective too (it ends in a stream-ordered barrier over all strips, so that no neighbour's edge rows land in
 * halo rows an upload is still going to overwrite).  A peer that does not
 * arrive within a few seconds fails the iteration (HP_ERR_PEER from hp_scheme_read_stats) instead of hanging. */
#define HP_PEER_BLOB_BYTES 256
int  hp_scheme_peer_export(hp_scheme* s, void* blob_out /* HP_PEER_BLOB_BYTES */);
int  hp_scheme_attach_peers(hp_scheme* s, int rank, int world_size, const void* blobs /* world_size x HP_PEER_BLOB_BYTES, by rank */);
/* Per-phase device times of the strip iteration, the counterpart of the reference's per-domain wall-clock log lines
 * around its exchange (src/CModel.cpp:843-958).  While enabled, iterations are launched directly (no graph replay)
 * with CUDA events between the phases and one host synchronisation per iteration: a diagnostic mode, not a fast one.
 * Phases: 0 edge rows, 1 interior rows (the whole step on small strips), 2 wait for the halo exchange that ran beside
 * the interior rows, 3 all-reduce of the wave speed (on small strips: halo exchange + all-reduce in one NCCL group),
 * 4 time controller.  hp_scheme_strip_timing(s, 1) also clears the accumulated times; collective like iterate. */
#define HP_STRIP_PHASES 5
int  hp_scheme_strip_timing(hp_scheme* s, int enable);
int  hp_scheme_read_strip_phases(hp_scheme* s, double* ms_per_iteration /* [HP_STRIP_PHASES] */, uint64_t* iterations);

#ifdef __cplusplus
}
#endif
#endif /* HIPIMS_CUDA_H */
