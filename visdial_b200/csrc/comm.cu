// Data-parallel gradient all-reduce: one ncclAllReduce(SUM) of the flat fp32 gradient per
// training step.  The reference has no counterpart (train.lua is single-GPU).  NCCL is bound with dlopen so
// that the library loads (and every single-GPU path works) on a host without libnccl, and so that a Python
// process that already carries torch's libnccl.so.2 reuses that copy instead of loading a second one.
#include "engine.h"
#include <dlfcn.h>
#include <string.h>

namespace vd {
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api() {
  static NcclApi a;
  if (a.handle) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  VD_REQUIRE(a.handle != nullptr, VD_E_COMM, "cannot dlopen libnccl.so.2");
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
  VD_REQUIRE(a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy, VD_E_COMM, "libnccl lacks required symbols");
  return a;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r == 0) return;
  char buf[256];
  snprintf(buf, sizeof(buf), "NCCL %s failed: %s", what, api().GetErrorString ? api().GetErrorString(r) : "?");
  throw CudaError(VD_E_COMM, buf);
}
}  // namespace

void comm_unique_id(void* out) {
  ncclUniqueId id;
  nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
  static_assert(sizeof(ncclUniqueId) == VD_COMM_ID_BYTES, "unique id size");
  memcpy(out, &id, sizeof(id));
}

void comm_init(Engine* e, const void* idbytes, int rank, int world) {
  VD_REQUIRE(world >= 1 && rank >= 0 && rank < world, VD_E_BADARG, "bad rank/world");
  VD_CUDA_CHECK(cudaSetDevice(e->cfg.gpuid));
  if (e->nccl_comm) { api().CommDestroy((ncclComm_t)e->nccl_comm); e->nccl_comm = nullptr; }
  e->rank = rank; e->world = world;
  if (world == 1) return;
  ncclUniqueId id;
  memcpy(&id, idbytes, sizeof(id));
  ncclComm_t comm;
  nccl_check(api().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  e->nccl_comm = comm;
}

void comm_destroy(Engine* e) {
  if (e->nccl_comm) { api().CommDestroy((ncclComm_t)e->nccl_comm); e->nccl_comm = nullptr; }
}

void Engine::allreduce_grads() {
  if (world <= 1) return;
  VD_REQUIRE(nccl_comm != nullptr, VD_E_STATE, "communicator not initialised");
  join_options_backward();
  LaunchCtx::Scope sc(&cx, "allreduce", 0.0, 2.0 * 4.0 * (double)nparams);
  nccl_check(api().AllReduce(dW, dW, (size_t)nparams, ncclFloat32, ncclSum, (ncclComm_t)nccl_comm, cx.stream), "ncclAllReduce");
}

}  // namespace vd
