// Library context ("handle" of include/gpfx.h): owns what used to be process-global - the pool of
// auxiliary streams / events the layer entry points fork side strands onto, the GEMM tuning override
// and the last error text.  Every entry point works on the calling thread's CURRENT context
// (gpfx_set_current); threads that never set one share the process default context.
#pragma once
#include <map>
#include <mutex>

#include "common.cuh"

namespace gpfx {

// One side strand: a non-blocking stream forked from / joined into the caller's stream with events.
struct AuxCtx {
  cudaStream_t aux = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

constexpr int kAuxPool = 8;
struct DevPool {
  AuxCtx ctx[kAuxPool];
  int used = 0;
  bool ready = false;
  std::map<cudaStream_t, AuxCtx*> bound;
};

struct Context {
  int device = -1;  // >= 0: bound to that device at creation (its pool exists from the start); -1: default
  std::mutex mu;
  std::map<int, DevPool> pools;
  int forced_cfg = -1;  // gpfx_tune_gemm_config
  // contraction precision of the layer entry points (gpfx_set_contraction): 0 = FP64 DMMA (default),
  // 1 = 3xTF32 on tcgen05 for the large plain products (float32 opt-in), with its caller-owned workspace
  int contraction = 0;
  float* tf32_ws = nullptr;
  size_t tf32_ws_bytes = 0;
  char err[512] = "";
};

Context* current_context();  // never null
int context_create(int device, Context** out);
int context_destroy(Context* c);
void context_set_current(Context* c);  // nullptr = process default

// side-strand context bound to caller stream `main` (nullptr: run in line), see layer.cu
AuxCtx* context_aux(cudaStream_t main);

inline int forced_gemm_cfg() { return current_context()->forced_cfg; }

}  // namespace gpfx
