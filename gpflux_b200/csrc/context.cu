// gpfx_create / gpfx_destroy: the library context (see context.h).  SURVEY §8b: "stream-ordered and
// re-entrant per handle, no global mutable state".
#include "context.h"

#include <cstdlib>

namespace gpfx {

namespace {
Context& default_context() {
  static Context c;
  return c;
}
thread_local Context* t_current = nullptr;

bool make_pool(DevPool& pool) {
  bool ok = true;
  // highest priority: the side strands are short latency-bound chains; their blocks should take the SM
  // slots the large GEMMs' retiring CTAs free, ahead of those GEMMs' own queued CTAs
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) lo = hi = 0;
  for (int i = 0; ok && i < kAuxPool; ++i) {
    ok = cudaStreamCreateWithPriority(&pool.ctx[i].aux, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int e = 0; ok && e < 4; ++e)
      ok = cudaEventCreateWithFlags(&pool.ctx[i].ev[e], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    pool.used = kAuxPool;  // nothing usable: every caller runs in line
  }
  pool.ready = true;
  return ok;
}

void free_pool(DevPool& pool) {
  for (int i = 0; i < kAuxPool; ++i) {
    for (int e = 0; e < 4; ++e)
      if (pool.ctx[i].ev[e]) cudaEventDestroy(pool.ctx[i].ev[e]);
    if (pool.ctx[i].aux) cudaStreamDestroy(pool.ctx[i].aux);
    pool.ctx[i] = AuxCtx();
  }
  pool.bound.clear();
  pool.used = 0;
  pool.ready = false;
}
}  // namespace

Context* current_context() { return t_current ? t_current : &default_context(); }

void context_set_current(Context* c) { t_current = c; }

int context_create(int device, Context** out) {
  int ndev = 0;
  GPFX_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (out == nullptr || device < 0 || device >= ndev) {
    set_last_error("gpfx_create: bad device ordinal %d (%d devices)", device, ndev);
    return GPFX_E_SHAPE;
  }
  int prev = 0;
  GPFX_CUDA_CHECK(cudaGetDevice(&prev));
  GPFX_CUDA_CHECK(cudaSetDevice(device));
  Context* c = new Context();
  c->device = device;
  make_pool(c->pools[device]);  // streams / events exist before any (possibly captured) call
  cudaSetDevice(prev);
  *out = c;
  return GPFX_OK;
}

int context_destroy(Context* c) {
  if (c == nullptr || c == &default_context()) return GPFX_OK;
  if (t_current == c) t_current = nullptr;
  int prev = 0;
  cudaGetDevice(&prev);
  {
    std::lock_guard<std::mutex> lock(c->mu);
    for (auto& kv : c->pools) {
      cudaSetDevice(kv.first);
      free_pool(kv.second);
    }
  }
  cudaSetDevice(prev);
  cudaGetLastError();
  delete c;
  return GPFX_OK;
}

AuxCtx* context_aux(cudaStream_t main) {
  static const bool enabled = [] {
    const char* e = getenv("GPFX_AUX_STREAM");
    return !(e && e[0] == '0');
  }();
  if (!enabled) return nullptr;
  // The pool of a device is created at gpfx_create, or (default context) on the first call that is
  // NOT inside a stream capture - creating streams / events is not capture-safe; afterwards caller
  // streams, including the capture stream of a CUDA graph, are bound to pool entries without any
  // CUDA call.
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  Context* c = current_context();
  std::lock_guard<std::mutex> lock(c->mu);
  DevPool& pool = c->pools[dev];
  if (!pool.ready) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return nullptr;
    make_pool(pool);
  }
  auto it = pool.bound.find(main);
  if (it != pool.bound.end()) return it->second;
  AuxCtx* a = (pool.used < kAuxPool) ? &pool.ctx[pool.used++] : nullptr;  // pool exhausted: run in line
  pool.bound[main] = a;
  return a;
}

}  // namespace gpfx
