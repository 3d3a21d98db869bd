// nccl_hook.cpp — libglio_nccl.so: the NCCL implementation of glio_allreduce_fn (include/glio_b200.h).
// Kept out of libglio_b200.so so the core library has no NCCL link dependency; a C++ host (the reference's
// Estimator is one process per robot; a multi-GPU batch solve is one process per GPU) creates the communicator
// here, or passes its own ncclComm_t as `user` to glio_nccl_allreduce.
// Payload per evaluation at BASELINE cfg 4: K x 28 + P x 36 doubles = 2000*28*8 + 24000*36*8 B = 7.4 MB ->
// latency/launch bound on NVLink 5 / NVSwitch, one ncclAllReduce(sum, double) per buffer.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <string.h>

extern "C" {

int glio_nccl_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

int glio_nccl_get_unique_id(void* out) {
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return -5;
  memcpy(out, &id, sizeof(id));
  return 0;
}

int glio_nccl_comm_create(int nranks, int rank, const void* id_bytes, void** comm_out) {
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm;
  if (ncclCommInitRank(&comm, nranks, id, rank) != ncclSuccess) return -5;
  *comm_out = (void*)comm;
  return 0;
}

void glio_nccl_comm_destroy(void* comm) { if (comm) ncclCommDestroy((ncclComm_t)comm); }

// glio_allreduce_fn: user = ncclComm_t
int glio_nccl_allreduce(void* user, double* d_buf, int64_t count, void* cuda_stream) {
  const ncclResult_t r = ncclAllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)user, (cudaStream_t)cuda_stream);
  return r == ncclSuccess ? 0 : -5;
}

}  // extern "C"
