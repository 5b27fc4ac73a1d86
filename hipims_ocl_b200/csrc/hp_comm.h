// hp_comm.h -- row-strip communication over NCCL (NVLink 5 / NVSwitch).
//
// Replaces the reference's host-staged MPI path: CDomainLink::pullFromBuffer -> sendOverMPI ->
// pushToBuffer (src/Domain/Links/CDomainLink.cpp:168-270, src/MPI/CMPIManager.cpp:555-717) becomes
// device-to-device ncclSend/ncclRecv of whole halo rows, and the MPI_Allreduce(MIN) of the
// timestep in a helper thread (src/MPI/CMPIManager.cpp:837-889) becomes an in-stream
// ncclAllReduce(ncclMax) on the wave-speed bits.  NCCL is loaded with dlopen so that a
// single-GPU run has no NCCL dependency.  Functions return NULL or an error message.
#pragma once

#include <cstddef>
#include <cuda_runtime.h>

#include "hp_kernels.cuh"

namespace hp {

struct Comm;

const char* comm_unique_id(void* id_out_128);
const char* comm_create(Comm** out, const void* id_128, int rank, int world_size);
void comm_destroy(Comm* c);
// sends the `halo` owned edge rows of every plane of `p` to the neighbouring strips and receives
// their edge rows into this strip's halo rows (rank r-1 is the southern neighbour)
const char* comm_exchange_halos(Comm* c, const Planes& p, const Grid& g, int halo, size_t real_bytes, cudaStream_t st);
// both of the above in ONE NCCL group (small strips: a single aggregated launch instead of two)
const char* comm_exchange_and_allreduce(Comm* c, const Planes& p, const Grid& g, int halo, size_t real_bytes, unsigned long long* value,
                                        cudaStream_t st);
// max over ranks of one unsigned 64-bit value (ordered bits of the wave speed), in place
const char* comm_allreduce_max(Comm* c, unsigned long long* value, cudaStream_t st);

}  // namespace hp
