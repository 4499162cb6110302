// a14: cross-replica gradient SUM over NCCL (train.py:75,110; keras_nerf/model/nerf/nerf.py:455-458).
// Interface used by pipeline.cu; the C ABI (knerf_comm_*) is in comm.cu.
#pragma once
#include "common.cuh"

struct knerf_comm;   // opaque in knerf.h

namespace knerf {

// enqueue on `comm_stream`: wait until everything enqueued on `after` so far has finished, then all-reduce
// grads[n] in place.  slot (0 / 1) selects one of the communicator's two hand-over events.
int comm_allreduce_after(knerf_comm* comm, float* grads, int64_t n, cudaStream_t after, cudaStream_t comm_stream,
                         int slot);
// `waiter` continues only after everything enqueued on `comm_stream` so far has finished
int comm_join(knerf_comm* comm, cudaStream_t comm_stream, cudaStream_t waiter);

}  // namespace knerf
