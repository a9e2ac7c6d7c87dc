// Micro-benchmarks that size the latency-bound small-matrix kernels (Cholesky diagonal block):
// dependent-chain latency of DFMA / DMUL / MUFU.RSQ64H, __syncthreads round trip with 8 warps,
// shared-memory publish->barrier->read round trip.  One CTA of 256 threads; clock64 deltas.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
    x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 512;
}
__global__ void k_dfma_tp(double* out, long long* cyc, double a, double b) {
  double x[8];
  for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x + j;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, a);
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);  // 512 DFMA per warp, 8 warps
}
__global__ void k_rsq(double* out, long long* cyc, double a) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    x = y + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;  // MUFU + DADD
}
__global__ void k_bar(double* out, long long* cyc) {
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 64;
  out[threadIdx.x] = 0;
}
__global__ void k_pub(double* out, long long* cyc) {
  __shared__ double s[2][64];
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
    if (threadIdx.x == (i & 63)) s[i & 1][i & 63] = x + 1.0;
    __syncthreads();
    x = s[i & 1][i & 63] * 0.5 + x;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 64;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64);
  for (int r = 0; r < 2; ++r) {
    k_dfma<<<1, 256>>>(out, cyc, 1.0, 0.999);
    k_dfma_tp<<<1, 256>>>(out, cyc, 1.0, 0.999);
    k_rsq<<<1, 256>>>(out, cyc, 2.0);
    k_bar<<<1, 256>>>(out, cyc);
    k_pub<<<1, 256>>>(out, cyc);
    cudaDeviceSynchronize();
  }
  printf("DFMA dependent latency (8 warps resident): %lld cyc\n", cyc[0]);
  printf("DFMA throughput: 512 per warp x 8 warps in %lld cyc -> %.2f cyc per warp-instr per SMSP\n", cyc[1], cyc[1] / (512.0 * 2));
  printf("MUFU.RSQ64H + DADD dependent: %lld cyc\n", cyc[2]);
  printf("__syncthreads (8 warps): %lld cyc\n", cyc[3]);
  printf("publish -> barrier -> read -> DFMA: %lld cyc\n", cyc[4]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
