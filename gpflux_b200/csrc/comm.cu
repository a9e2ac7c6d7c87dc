// Data-parallel gradient exchange for the SVGP path (SURVEY §8e / §8b `gpfx_allreduce_grads`): the
// minibatch rows are sharded over GPUs, parameters and the M x M factorisations are replicated, and
// the flat FP64 gradient buffer is summed (or averaged) over ranks with ONE NCCL all-reduce per
// bucket, stream-ordered on the caller's stream (CUDA-graph capturable).
//
// The reference has no distributed code: this is an addition.  NCCL is bound at run time with
// dlopen("libnccl.so.2") so that the library has no link-time dependency on it (a single-GPU host
// never touches NCCL) and shares the NCCL that the host framework already loaded, if any.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <mutex>

#include "../../include/gpfx.h"
#include "common.cuh"

namespace gpfx {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
  char err[256] = "missing symbols";
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) {
      const char* e = dlerror();
      snprintf(a.err, sizeof a.err, "%s", e ? e : "dlopen failed");
      return;
    }
    auto sym = [&](const char* s) { return dlsym(a.handle, s); };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
  });
  return a;
}

int need_api() {
  if (!api().ok) {
    set_last_error("NCCL is not available: dlopen(libnccl.so.2) failed (%s)", api().err);
    return GPFX_E_NCCL;
  }
  return GPFX_OK;
}

#define GPFX_NCCL_CHECK(expr)                                                                 \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, api().GetErrorString(_r)); \
      return GPFX_E_NCCL;                                                                     \
    }                                                                                         \
  } while (0)

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
};

}  // namespace
}  // namespace gpfx

using namespace gpfx;

extern "C" {

int gpfx_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

int gpfx_comm_nccl_version(void) {
  int v = 0;
  if (!api().ok || !api().GetVersion || api().GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int gpfx_comm_unique_id(void* id_out) {
  GPFX_TRY(need_api());
  if (!id_out) {
    set_last_error("gpfx_comm_unique_id: null pointer");
    return GPFX_E_SHAPE;
  }
  GPFX_NCCL_CHECK(api().GetUniqueId(reinterpret_cast<ncclUniqueId*>(id_out)));
  return GPFX_OK;
}

int gpfx_comm_init(const void* id, int rank, int world, gpfx_comm_t* comm_out) {
  if (!id || !comm_out || world < 1 || rank < 0 || rank >= world) {
    set_last_error("gpfx_comm_init: bad arguments (rank %d of %d)", rank, world);
    return GPFX_E_SHAPE;
  }
  GPFX_TRY(need_api());
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId uid = *reinterpret_cast<const ncclUniqueId*>(id);
  ncclResult_t r = api().CommInitRank(&c->nccl, world, uid, rank);
  if (r != ncclSuccess) {
    set_last_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, api().GetErrorString(r));
    delete c;
    return GPFX_E_NCCL;
  }
  *comm_out = c;
  return GPFX_OK;
}

int gpfx_comm_destroy(gpfx_comm_t comm) {
  if (!comm) return GPFX_OK;
  Comm* c = reinterpret_cast<Comm*>(comm);
  int st = GPFX_OK;
  if (c->nccl && api().ok && api().CommDestroy(c->nccl) != ncclSuccess) st = GPFX_E_NCCL;
  delete c;
  return st;
}

int gpfx_allreduce_grads(gpfx_comm_t comm, double* flat_buf, long long count, int average,
                         void* stream) {
  if (!comm || (!flat_buf && count > 0) || count < 0) {
    set_last_error("gpfx_allreduce_grads: bad arguments");
    return GPFX_E_SHAPE;
  }
  if (count == 0) return GPFX_OK;
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c->world == 1) return GPFX_OK;
  GPFX_TRY(need_api());
  GPFX_NCCL_CHECK(api().AllReduce(flat_buf, flat_buf, (size_t)count, ncclFloat64,
                                  average ? ncclAvg : ncclSum, c->nccl,
                                  reinterpret_cast<cudaStream_t>(stream)));
  return GPFX_OK;
}

}  // extern "C"
