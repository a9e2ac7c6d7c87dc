// DMMA.8x8x4 ceilings on one B200: (1) register-only issue rate for MI x NI warp tiles at 4..16 warps
// per SM, (2) the same loops with their fragments re-read from shared memory every k4 step (the
// GEMM mainloop without its global-memory feed).  148 CTAs, one per SM.  Prints TFLOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MI, int NI, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_reg(double* out, int iters, double a0, double b0) {
  double acc[MI][NI][2];
  double a[MI], b[NI];
#pragma unroll
  for (int i = 0; i < MI; ++i) a[i] = a0 + threadIdx.x + i;
#pragma unroll
  for (int j = 0; j < NI; ++j) b[j] = b0 + j;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4)
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fragments from smem: K-major 128-byte rows with the TMA 128B swizzle (16-byte chunk ^= row & 7)
template <int MI, int NI, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_smem(double* out, int iters, int wn_count) {
  extern __shared__ __align__(1024) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  for (int i = tid; i < 3 * 512 * 16; i += blockDim.x) sm[i] = 1.0 + (i & 7);
  __syncthreads();
  const int wm = warp / wn_count, wn = warp % wn_count;
  // A tile rows: (wm*MI + i)*8 + lr ; B tile rows: 256 + (wn*NI + j)*8 + lr  (both k-major, 16 doubles per row)
  const char* base = reinterpret_cast<const char*>(sm);
  unsigned offk[4];
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) offk[k4] = (((2 * k4) ^ ((lc >> 1) ^ lr)) << 4) + (lc & 1) * 8 + lr * 128;
  const char* abase = base + (wm * MI * 8) * 128;
  const char* bbase = base + (256 + wn * NI * 8) * 128;
  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const char* ab = abase + (it % 3) * (512 * 128);
    const char* bb = bbase + (it % 3) * (512 * 128);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      double a[MI], b[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double*>(ab + offk[k4] + i * 1024);
#pragma unroll
      for (int j = 0; j < NI; ++j) b[j] = *reinterpret_cast<const double*>(bb + offk[k4] + j * 1024);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MI, int NI, int MAXT = 512>
void run(int warps, int wn_count, double* out) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double flop = 148.0 * warps * iters * 4.0 * MI * NI * 512.0;
  float ms;
  k_reg<MI, NI, MAXT><<<148, warps * 32>>>(out, 64, 1.0, 2.0);
  cudaEventRecord(e0);
  k_reg<MI, NI, MAXT><<<148, warps * 32>>>(out, iters, 1.0, 2.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double tf_reg = flop / ms * 1e-9;
  const size_t smem = 3 * 512 * 128;
  cudaFuncSetAttribute(k_smem<MI, NI, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_smem<MI, NI, MAXT><<<148, warps * 32, smem>>>(out, 64, wn_count);
  cudaEventRecord(e0);
  k_smem<MI, NI, MAXT><<<148, warps * 32, smem>>>(out, iters, wn_count);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double tf_sm = flop / ms * 1e-9;
  cudaError_t err = cudaGetLastError();
  printf("MI=%d NI=%d warps/SM=%2d : registers %.2f TF   smem fragments %.2f TF  %s\n", MI, NI, warps, tf_reg, tf_sm,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 512);
  for (int w : {4, 8, 12, 16}) run<4, 4>(w, 4, out);
  for (int w : {4, 8, 12, 16}) run<2, 4>(w, 4, out);
  for (int w : {4, 8}) run<8, 4, 256>(w, 4, out);
  for (int w : {4, 8}) run<4, 8, 256>(w, 2, out);
  for (int w : {4, 8, 16}) run<2, 2>(w, 4, out);
  return 0;
}
