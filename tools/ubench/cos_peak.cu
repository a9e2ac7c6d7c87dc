// FP64 cos ceiling: every thread runs 4 independent chains of cos on register operands (arguments
// spread over [0, 2 pi L) like the random-feature projections), 148 x 8 CTAs of 256 threads.
// Prints cos/s - the denominator of the Wilson-sampling kernel's roofline (c5 is bound by
// the FP64 transcendental issue rate).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cos_peak cos_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_cos(double* out, int iters, double step) {
  double x0 = 0.1 + threadIdx.x * 0.37 + blockIdx.x * 1.3, x1 = x0 + 11.0, x2 = x0 + 23.0, x3 = x0 + 37.0;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    s0 += cos(x0); s1 += cos(x1); s2 += cos(x2); s3 += cos(x3);
    x0 += step; x1 += step; x2 += step; x3 += step;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

int main() {
  const int grid = 148 * 8, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_cos<<<grid, 256>>>(out, 100, 0.731);
  cudaEventRecord(e0);
  k_cos<<<grid, 256>>>(out, iters, 0.731);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("fp64 cos: %.3e cos/s (%d x 256 threads x 4 chains x %d, %.2f ms)\n", 4.0 * grid * 256 * (double)iters / ms * 1e3, grid,
         iters, ms);
  return 0;
}
