// Shared helpers for the gpflux_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#define GPFX_OK 0
#define GPFX_E_SHAPE (-1)
#define GPFX_E_DTYPE (-2)
#define GPFX_E_LAYOUT (-3)
#define GPFX_E_NOT_PD (-4)
#define GPFX_E_CUDA (-5)
#define GPFX_E_NCCL (-6)
#define GPFX_E_UNSUPPORTED (-7)
#define GPFX_E_WORKSPACE (-8)

namespace gpfx {

void set_last_error(const char* fmt, ...);
void count_launch();

#define GPFX_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::gpfx::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                             cudaGetErrorString(_e));                                      \
      return GPFX_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

// every kernel launch site goes through this: it bumps the library's launch counter
// (gpfx_kernel_launch_count) and surfaces launch-configuration errors
#define GPFX_LAUNCH_CHECK()             \
  do {                                  \
    ::gpfx::count_launch();             \
    GPFX_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

#define GPFX_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != GPFX_OK) return _s; \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember per device (bit =
// device ordinal) which kernels already had it raised - one process may drive several GPUs.
struct SmemAttrOnce {
  std::atomic<unsigned long long> done{0};
};
template <typename F>
static inline int ensure_dyn_smem(F func, size_t bytes, SmemAttrOnce& once) {
  int dev = 0;
  GPFX_CUDA_CHECK(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (once.done.load(std::memory_order_acquire) & bit) return GPFX_OK;
  GPFX_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  once.done.fetch_or(bit, std::memory_order_release);
  return GPFX_OK;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- warp / block reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in thread 0 (and broadcast to all via smem).  `red` must hold
// >= 32 doubles.  Deterministic (fixed tree).
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// ---- cp.async (LDGSTS) -------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- FP64 tensor-core MMA: D[8x8] += A[8x4] * B[4x8]  (SASS: DMMA.8x8x4) -----------------
// fragment ownership (lane = 0..31): a = A[lane>>2][lane&3], b = B[lane&3][lane>>2],
// c0,c1 = C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace gpfx
