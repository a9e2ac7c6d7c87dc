// Batched FP64 Cholesky + triangular inverse of the (padded) Kuu stack and the Cholesky backward.
//
// Replaces tf.linalg.cholesky / tf.linalg.triangular_solve in GPflow's base_conditional as reached
// from GPLayer.predict (reference gpflux/layers/gp_layer.py:257-266) and
// gpflux/math.py:25-62.  Design: 64x64 diagonal blocks are factorised and inverted inside one CTA's
// shared memory; everything off the diagonal is DMMA GEMM work (left-looking panel updates,
// recursive-doubling assembly of L^-1), batched over the K kernels.  Matrices are padded to
// Mp = 64 * 2^q with an identity block so every level is uniform.
#include "chol.h"

#include "gemm_f64.h"

namespace gpfx {
namespace {

constexpr int NB = 64;

// One CTA per matrix k.  In place: L[j0:j0+NB, j0:j0+NB] <- chol(lower part); zero to the right of
// the diagonal inside these rows; Linv diagonal block <- inverse of the factor.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ L,
                                                         double* __restrict__ Linv, int Mp,
                                                         long long sL, int j0, int* info) {
  extern __shared__ double diag_smem[];
  double(*S)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(diag_smem);
  double(*X)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(diag_smem + NB * (NB + 1));
  const int tid = threadIdx.x;
  double* Lk = L + blockIdx.x * sL;
  double* Xk = Linv + blockIdx.x * sL;

  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    S[r][c] = (c <= r) ? Lk[(long long)(j0 + r) * Mp + j0 + c] : 0.0;
    X[r][c] = 0.0;
  }
  __syncthreads();

  for (int c = 0; c < NB; ++c) {
    const double d = S[c][c];
    if (!(d > 0.0) && tid == 0) atomicCAS(info + blockIdx.x, 0, j0 + c + 1);
    const double l = sqrt(d);
    __syncthreads();
    if (tid == 0) S[c][c] = l;
    for (int r = c + 1 + tid; r < NB; r += 256) S[r][c] /= l;
    __syncthreads();
    const int n = NB - c - 1;
    for (int e = tid; e < n * n; e += 256) {
      const int r = c + 1 + e / n, cc = c + 1 + e % n;
      if (cc <= r) S[r][cc] -= S[r][c] * S[cc][c];
    }
    __syncthreads();
  }

  // inverse by forward substitution: a quad of lanes per column j
  {
    const int j = tid >> 2, part = tid & 3;
    for (int i = 0; i < NB; ++i) {
      double s = 0.0;
      if (i > j)
        for (int k = j + part; k < i; k += 4) s += S[i][k] * X[k][j];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0 && i >= j) X[i][j] = ((i == j ? 1.0 : 0.0) - s) / S[i][i];
      __syncwarp();
    }
  }
  __syncthreads();

  for (int e = tid; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    Lk[(long long)(j0 + r) * Mp + j0 + c] = S[r][c];
    Xk[(long long)(j0 + r) * Mp + j0 + c] = X[r][c];
  }
  const int right = Mp - (j0 + NB);
  for (long long e = tid; e < (long long)NB * right; e += 256) {
    const int r = (int)(e / right), c = (int)(e % right);
    Lk[(long long)(j0 + r) * Mp + j0 + NB + c] = 0.0;
  }
}

}  // namespace

int chol_padded_dim(int M) {
  int Mp = NB;
  while (Mp < M) Mp *= 2;
  return Mp;
}

size_t chol_scratch_doubles(int K, int Mp) { return (size_t)K * Mp * Mp / 4; }

int potrf_trtri_batched(double* L, double* Linv, double* T, int K, int Mp, int* info,
                        cudaStream_t stream) {
  const long long sL = (long long)Mp * Mp;
  GPFX_CUDA_CHECK(cudaMemsetAsync(Linv, 0, sizeof(double) * sL * K, stream));
  GPFX_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int) * K, stream));
  const int nblk = Mp / NB;
  const size_t diag_smem_bytes = 2 * NB * (NB + 1) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    GPFX_CUDA_CHECK(cudaFuncSetAttribute(potrf_diag_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)diag_smem_bytes));
    attr_set = true;
  }
  for (int j = 0; j < nblk; ++j) {
    const int j0 = j * NB;
    if (j > 0) {
      GemmArgs g;
      g.A = L + (long long)j0 * Mp;
      g.B = L + (long long)j0 * Mp;
      g.C = L + (long long)j0 * Mp + j0;
      g.M = Mp - j0; g.N = NB; g.K = j0;
      g.lda = g.ldb = g.ldc = Mp;
      g.transB = 1;
      g.batch1 = K; g.sA1 = g.sB1 = g.sC1 = sL;
      g.alpha = -1.0; g.beta = 1.0;
      GPFX_TRY(launch_gemm(g, stream));
    }
    potrf_diag_kernel<<<K, 256, diag_smem_bytes, stream>>>(L, Linv, Mp, sL, j0, info);
    GPFX_LAUNCH_CHECK();
    if (j < nblk - 1) {
      GemmArgs g;
      g.A = L + (long long)(j0 + NB) * Mp + j0;
      g.B = Linv + (long long)j0 * Mp + j0;
      g.C = L + (long long)(j0 + NB) * Mp + j0;  // in place: every CTA tile spans all NB columns
      g.M = Mp - j0 - NB; g.N = NB; g.K = NB;
      g.lda = g.ldb = g.ldc = Mp;
      g.transB = 1;
      g.batch1 = K; g.sA1 = g.sB1 = g.sC1 = sL;
      GPFX_TRY(launch_gemm(g, stream));
    }
  }
  // L^-1 by recursive doubling: Linv21 = -Linv22 * (L21 * Linv11)
  for (int s = NB; s < Mp; s *= 2) {
    const int nq = Mp / (2 * s);
    const long long pair = 2LL * s * (Mp + 1);
    GemmArgs g1;
    g1.A = L + (long long)s * Mp; g1.lda = Mp;
    g1.B = Linv; g1.ldb = Mp;
    g1.C = T; g1.ldc = s;
    g1.M = g1.N = g1.K = s;
    g1.batch1 = K; g1.batch2 = nq;
    g1.sA1 = sL; g1.sA2 = pair;
    g1.sB1 = sL; g1.sB2 = pair;
    g1.sC1 = (long long)nq * s * s; g1.sC2 = (long long)s * s;
    g1.tri = TRI_LOWER_B;
    GPFX_TRY(launch_gemm(g1, stream));
    GemmArgs g2;
    g2.A = Linv + (long long)s * (Mp + 1); g2.lda = Mp;
    g2.B = T; g2.ldb = s;
    g2.C = Linv + (long long)s * Mp; g2.ldc = Mp;
    g2.M = g2.N = g2.K = s;
    g2.batch1 = K; g2.batch2 = nq;
    g2.sA1 = sL; g2.sA2 = pair;
    g2.sB1 = g1.sC1; g2.sB2 = g1.sC2;
    g2.sC1 = sL; g2.sC2 = pair;
    g2.tri = TRI_LOWER_A;
    g2.alpha = -1.0;
    GPFX_TRY(launch_gemm(g2, stream));
  }
  return GPFX_OK;
}

// dK = L^-T Phi(L^T dL) L^-1  (un-symmetrised; consumers contract it with a symmetric dK/dtheta)
int chol_backward_batched(const double* L, const double* Linv, const double* dL, double* T1,
                          double* T2, double* dK, int K, int Mp, cudaStream_t stream) {
  const long long sL = (long long)Mp * Mp;
  GemmArgs g;
  g.M = g.N = g.K = Mp;
  g.lda = g.ldb = g.ldc = Mp;
  g.batch1 = K; g.sA1 = g.sB1 = g.sC1 = sL;
  // T1 = Phi(L^T dL)
  g.A = L; g.transA = 1; g.B = dL; g.C = T1;
  g.tri = TRI_UPPER_A | TRI_LOWER_B; g.tril_out = 1; g.diag_scale = 0.5;
  GPFX_TRY(launch_gemm(g, stream));
  // T2 = T1 Linv
  g.A = T1; g.transA = 0; g.B = Linv; g.C = T2;
  g.tri = TRI_LOWER_A | TRI_LOWER_B; g.tril_out = 0; g.diag_scale = 1.0;
  GPFX_TRY(launch_gemm(g, stream));
  // dK = Linv^T T2
  g.A = Linv; g.transA = 1; g.B = T2; g.C = dK;
  g.tri = TRI_UPPER_A | TRI_LOWER_B;
  GPFX_TRY(launch_gemm(g, stream));
  return GPFX_OK;
}

}  // namespace gpfx
