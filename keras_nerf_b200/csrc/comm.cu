// a14: the gradient all-reduce that tf.distribute.MirroredStrategy performs inside apply_gradients
// (train.py:75,110; keras_nerf/model/nerf/nerf.py:455-458; SUM, not mean: train.py:130-136), as one NCCL
// all-reduce per network over the flat fp32 gradient buffer (595,844 floats = 2.4 MB each).
//
// NCCL is bound at run time: dlopen("libnccl.so.2") picks up the copy the host process has already loaded (the
// one bundled with PyTorch / TensorFlow) or the system library; libknerf.so itself has no link-time dependency on
// it, so single-GPU hosts need no NCCL.  Only five entry points with a stable ABI are used, declared here.
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>
#include <new>

#include <mutex>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[KNERF_COMM_ID_BYTES]; } ncclUniqueId;   // NCCL_UNIQUE_ID_BYTES = 128
typedef int ncclResult_t;                                              // ncclSuccess = 0
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

const NcclApi* nccl_api() {
  static NcclApi api{};
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);   // the host's copy, if any
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

}  // namespace

extern "C" int knerf_comm_destroy(knerf_comm* comm);

struct knerf_comm {
  ncclComm_t nccl;
  int rank, nranks;
  bool owned;             // created by knerf_comm_create (destroyed with the object) vs adopted from the host
  cudaEvent_t ready[2];   // "this network's backward has finished" (stream -> comm_stream)
  cudaEvent_t done;       // "the reductions have finished" (comm_stream -> stream)
};

namespace knerf {

#define KN_NCCL(api, expr)                                                                                     \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != 0) return fail(KNERF_ERR_NCCL, "%s failed: %s", #expr, (api)->GetErrorString(_r));               \
  } while (0)

static int make_comm(ncclComm_t nccl, int rank, int nranks, bool owned, knerf_comm** out) {
  knerf_comm* c = new (std::nothrow) knerf_comm{nccl, rank, nranks, owned, {nullptr, nullptr}, nullptr};
  if (c == nullptr) return fail(KNERF_ERR_INVALID, "out of host memory");
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ready[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
  if (e != cudaSuccess) {   // nothing half-built escapes: the caller gets no handle
    c->owned = owned;
    knerf_comm_destroy(c);
    return fail(KNERF_ERR_CUDA, "cudaEventCreateWithFlags failed: %s", cudaGetErrorString(e));
  }
  *out = c;
  return KNERF_OK;
}

int comm_allreduce_after(knerf_comm* comm, float* grads, int64_t n, cudaStream_t after, cudaStream_t comm_stream,
                         int slot) {
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(KNERF_ERR_NCCL, "libnccl.so.2 not found");
  if (comm->nranks == 1) return KNERF_OK;
  KN_CUDA(cudaEventRecord(comm->ready[slot & 1], after));
  KN_CUDA(cudaStreamWaitEvent(comm_stream, comm->ready[slot & 1], 0));
  KN_NCCL(api, api->AllReduce(grads, grads, (size_t)n, kNcclFloat32, kNcclSum, comm->nccl, comm_stream));
  return KNERF_OK;
}

int comm_join(knerf_comm* comm, cudaStream_t comm_stream, cudaStream_t waiter) {
  if (comm->nranks == 1) return KNERF_OK;
  KN_CUDA(cudaEventRecord(comm->done, comm_stream));
  KN_CUDA(cudaStreamWaitEvent(waiter, comm->done, 0));
  return KNERF_OK;
}

}  // namespace knerf

using namespace knerf;

extern "C" int knerf_comm_unique_id(void* id_host) {
  KN_CHECK_ARG(id_host != nullptr, "knerf_comm_unique_id: null pointer");
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(KNERF_ERR_NCCL, "libnccl.so.2 not found");
  ncclUniqueId id;
  KN_NCCL(api, api->GetUniqueId(&id));
  memcpy(id_host, id.internal, KNERF_COMM_ID_BYTES);
  return KNERF_OK;
}

extern "C" int knerf_comm_create(const void* id_host, int rank, int nranks, knerf_comm** comm) {
  KN_CHECK_ARG(id_host != nullptr && comm != nullptr && nranks >= 1 && rank >= 0 && rank < nranks,
               "knerf_comm_create: bad arguments (rank %d of %d)", rank, nranks);
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(KNERF_ERR_NCCL, "libnccl.so.2 not found");
  ncclUniqueId id;
  memcpy(id.internal, id_host, KNERF_COMM_ID_BYTES);
  ncclComm_t nccl = nullptr;
  KN_NCCL(api, api->CommInitRank(&nccl, nranks, id, rank));
  return make_comm(nccl, rank, nranks, true, comm);
}

extern "C" int knerf_comm_adopt(void* nccl_comm, int rank, int nranks, knerf_comm** comm) {
  KN_CHECK_ARG(nccl_comm != nullptr && comm != nullptr && nranks >= 1 && rank >= 0 && rank < nranks,
               "knerf_comm_adopt: bad arguments (rank %d of %d)", rank, nranks);
  if (nccl_api() == nullptr) return fail(KNERF_ERR_NCCL, "libnccl.so.2 not found");
  return make_comm(reinterpret_cast<ncclComm_t>(nccl_comm), rank, nranks, false, comm);
}

extern "C" int knerf_comm_destroy(knerf_comm* comm) {
  if (comm == nullptr) return KNERF_OK;
  for (int i = 0; i < 2; ++i)
    if (comm->ready[i]) cudaEventDestroy(comm->ready[i]);
  if (comm->done) cudaEventDestroy(comm->done);
  int rc = KNERF_OK;
  if (comm->owned && comm->nccl != nullptr) {
    const NcclApi* api = nccl_api();
    if (api != nullptr && api->CommDestroy(comm->nccl) != 0) rc = fail(KNERF_ERR_NCCL, "ncclCommDestroy failed");
  }
  delete comm;
  return rc;
}

extern "C" int knerf_comm_rank(const knerf_comm* comm, int* rank, int* nranks) {
  KN_CHECK_ARG(comm != nullptr, "knerf_comm_rank: null communicator");
  if (rank) *rank = comm->rank;
  if (nranks) *nranks = comm->nranks;
  return KNERF_OK;
}

extern "C" int knerf_allreduce_grads(knerf_comm* comm, float* grads, int64_t n, void* stream) {
  KN_CHECK_ARG(comm != nullptr && grads != nullptr && n >= 0, "knerf_allreduce_grads: bad arguments");
  const NcclApi* api = nccl_api();
  if (api == nullptr) return fail(KNERF_ERR_NCCL, "libnccl.so.2 not found");
  if (comm->nranks == 1 || n == 0) return KNERF_OK;
  KN_NCCL(api, api->AllReduce(grads, grads, (size_t)n, kNcclFloat32, kNcclSum, comm->nccl, (cudaStream_t)stream));
  return KNERF_OK;
}
