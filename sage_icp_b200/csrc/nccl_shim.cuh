// Minimal NCCL binding resolved with dlopen at sage_map_comm_init() time, so the library loads (and the
// single-GPU path runs) on hosts without libnccl.  Only what the sharded ICP needs: one communicator and a
// sum all-reduce of 17 doubles per Gauss-Newton iteration.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace sage {
struct NcclComm;
void nccl_unique_id(uint8_t out[128]);
NcclComm *nccl_comm_create(int rank, int world, const uint8_t id[128]);
void nccl_comm_destroy(NcclComm *c);
void nccl_allreduce_sum_f64(NcclComm *c, double *dev_buf, int count, cudaStream_t stream);
}  // namespace sage
