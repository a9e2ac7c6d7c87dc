/*
 * A host with no Python and no torch in it: one GPLayer forward + backward through the C ABI of
 * libgpflux_b200.so (include/gpfx.h), written in C99 against the CUDA runtime API.
 *
 *   svgp_layer_step <in.bin> <out.bin>
 *
 * in.bin : int32 header {N, M, D, P, K, kernel, whiten, 0}, then float64:
 *          jitter, X[N,D], Z[K,M,D], lengthscales[K,D], variance[K], q_mu[M,P], q_sqrt[P,M,M],
 *          eps[N,P], g_fsample[N,P], g_fmean[N,P], g_fvar[N,P], g_kl
 * out.bin: float64 fmean[N,P], fvar[N,P], fsample[N,P], kl, dX[N,D], dZ[K,M,D], dlengthscales[K,D],
 *          dvariance[K], dq_mu[M,P], dq_sqrt[P,M,M], dmean[N,P]
 *
 * What it replaces on the reference side: GPLayer.call (gpflux/layers/gp_layer.py:226-363) and the
 * GradientTape around it; tests/test_gpu_c_host.py builds this file, runs it and checks out.bin
 * against the CPU oracle at 1e-9.
 *
 * Build: gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_host/svgp_layer_step.c \
 *            -Lgpflux_b200/lib -lgpflux_b200 -L/usr/local/cuda/lib64 -lcudart -o svgp_layer_step
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "gpfx.h"

#define CUDA_OK(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));      \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)
#define GPFX_CALL(call)                                                                \
  do {                                                                                 \
    int s_ = (call);                                                                   \
    if (s_ != GPFX_OK) {                                                               \
      fprintf(stderr, "%s:%d: gpfx error %d: %s\n", __FILE__, __LINE__, s_, gpfx_last_error()); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

/* host array -> new device array */
static double* to_device(const double* h, size_t n) {
  double* d = NULL;
  if (cudaMalloc((void**)&d, (n ? n : 1) * sizeof(double)) != cudaSuccess) return NULL;
  if (n && cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return d;
}
static double* device_zeros(size_t n) {
  double* d = NULL;
  if (cudaMalloc((void**)&d, (n ? n : 1) * sizeof(double)) != cudaSuccess) return NULL;
  if (cudaMemset(d, 0, (n ? n : 1) * sizeof(double)) != cudaSuccess) return NULL;
  return d;
}
/* device array -> appended to the output file */
static int dump(FILE* f, const double* d, size_t n) {
  double* h = (double*)malloc((n ? n : 1) * sizeof(double));
  if (!h) return 1;
  if (cudaMemcpy(h, d, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  const size_t w = fwrite(h, sizeof(double), n, f);
  free(h);
  return w == n ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc != 3) {
    fprintf(stderr, "usage: %s <in.bin> <out.bin>\n", argv[0]);
    return 2;
  }
  FILE* fin = fopen(argv[1], "rb");
  if (!fin) {
    perror(argv[1]);
    return 2;
  }
  int32_t hdr[8];
  if (fread(hdr, sizeof(int32_t), 8, fin) != 8) return 2;
  const size_t N = (size_t)hdr[0], M = (size_t)hdr[1], D = (size_t)hdr[2], P = (size_t)hdr[3], K = (size_t)hdr[4];
  const size_t counts[12] = {1, N * D, K * M * D, K * D, K, M * P, P * M * M, N * P, N * P, N * P, N * P, 1};
  double* host[12];
  for (int i = 0; i < 12; ++i) {
    host[i] = (double*)malloc((counts[i] ? counts[i] : 1) * sizeof(double));
    if (!host[i] || fread(host[i], sizeof(double), counts[i], fin) != counts[i]) {
      fprintf(stderr, "short read of input array %d\n", i);
      return 2;
    }
  }
  fclose(fin);

  gpfx_layer_desc desc;
  desc.N = hdr[0]; desc.M = hdr[1]; desc.D = hdr[2]; desc.P = hdr[3]; desc.K = hdr[4];
  desc.kernel = hdr[5];
  desc.whiten = hdr[6];
  desc.solver = GPFX_SOLVER_INVERSE_GEMM;
  desc.jitter = host[0][0];

  CUDA_OK(cudaSetDevice(0));
  gpfx_handle handle = NULL;
  GPFX_CALL(gpfx_create(0, &handle)); /* owns the side-stream pool; created outside of any capture */
  GPFX_CALL(gpfx_set_current(handle));
  cudaStream_t stream;
  CUDA_OK(cudaStreamCreate(&stream));

  double* dev[12];
  for (int i = 1; i < 12; ++i) {
    dev[i] = to_device(host[i], counts[i]);
    if (!dev[i]) {
      fprintf(stderr, "device allocation failed\n");
      return 1;
    }
  }
  const double *X = dev[1], *Z = dev[2], *ls = dev[3], *var = dev[4], *q_mu = dev[5], *q_sqrt = dev[6], *eps = dev[7];
  const double *g_f = dev[8], *g_m = dev[9], *g_v = dev[10], *g_kl = dev[11];

  /* the caller owns every buffer: saved-for-backward context and scratch come from the size queries */
  void *ctx = NULL, *ws = NULL;
  CUDA_OK(cudaMalloc(&ctx, gpfx_layer_ctx_bytes(&desc)));
  CUDA_OK(cudaMalloc(&ws, gpfx_layer_workspace_bytes(&desc)));
  double *fmean = device_zeros(N * P), *fvar = device_zeros(N * P), *fsample = device_zeros(N * P), *kl = device_zeros(1);
  double *dX = device_zeros(N * D), *dZ = device_zeros(K * M * D), *dls = device_zeros(K * D), *dvar = device_zeros(K);
  double *dq_mu = device_zeros(M * P), *dq_sqrt = device_zeros(P * M * M), *dmean = device_zeros(N * P);
  if (!fmean || !fvar || !fsample || !kl || !dX || !dZ || !dls || !dvar || !dq_mu || !dq_sqrt || !dmean) return 1;

  GPFX_CALL(gpfx_layer_forward(&desc, X, Z, ls, var, q_mu, q_sqrt, /*mean_add=*/NULL, eps, fmean, fvar, fsample, kl,
                               ctx, ws, stream));
  int32_t* info = (int32_t*)malloc(K * sizeof(int32_t));
  if (gpfx_layer_status(&desc, ctx, info, stream) != GPFX_OK) { /* what TF reports as InvalidArgumentError */
    fprintf(stderr, "Kuu is not positive definite: %s\n", gpfx_last_error());
    return 3;
  }
  GPFX_CALL(gpfx_layer_backward(&desc, X, Z, ls, var, q_mu, eps, fvar, g_f, g_m, g_v, g_kl, dX, dZ, dls, dvar, dq_mu,
                                dq_sqrt, dmean, ctx, ws, stream));
  CUDA_OK(cudaStreamSynchronize(stream));

  FILE* fout = fopen(argv[2], "wb");
  if (!fout) {
    perror(argv[2]);
    return 2;
  }
  int bad = 0;
  bad |= dump(fout, fmean, N * P);
  bad |= dump(fout, fvar, N * P);
  bad |= dump(fout, fsample, N * P);
  bad |= dump(fout, kl, 1);
  bad |= dump(fout, dX, N * D);
  bad |= dump(fout, dZ, K * M * D);
  bad |= dump(fout, dls, K * D);
  bad |= dump(fout, dvar, K);
  bad |= dump(fout, dq_mu, M * P);
  bad |= dump(fout, dq_sqrt, P * M * M);
  bad |= dump(fout, dmean, N * P);
  fclose(fout);
  if (bad) {
    fprintf(stderr, "writing %s failed\n", argv[2]);
    return 1;
  }
  printf("gpfx %d: N=%zu M=%zu D=%zu P=%zu K=%zu, %lld kernels launched\n", gpfx_version(), N, M, D, P, K,
         gpfx_kernel_launch_count());

  CUDA_OK(cudaStreamDestroy(stream));
  GPFX_CALL(gpfx_set_current(NULL));
  GPFX_CALL(gpfx_destroy(handle));
  return 0;
}
