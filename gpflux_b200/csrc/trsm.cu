// True triangular solves for the SVGP conditional: A = L^-1 Kuf by blocked
// forward substitution, dKuf = L^-T dA by blocked back substitution - the reference-faithful form of
// gpflow's base_conditional `tf.linalg.triangular_solve(Lm, Kmn)` (reached from
// gpflux/layers/gp_layer.py:257-266) and of its autodiff transpose.  The default path multiplies by an
// explicit L^-1 on the DMMA GEMM (layer.cu); this one never forms or uses an inverse:
//
//   for each 64-row block j (ascending; descending for the transposed solve):
//     X_j  = B_j - L[j, 0:j] X[0:j]          one DMMA GEMM launch (gemm_f64.cu), alpha = -1, beta = 1
//     X_j <- L_jj^-1 X_j by SUBSTITUTION      trsm_diag_kernel: one thread per right-hand-side column,
//                                             the 64 unknowns of the block in registers, L_jj broadcast
//                                             from shared memory, divisions by the diagonal
//
// 2 nb - 1 launches per solve (nb = ceil(M / 64)); the off-diagonal work (all but 1/nb of the flops) stays
// on the tensor pipe.  Plus the column reductions that the inverse-GEMM path gets from its fused epilogue.
#include "trsm.h"

#include "gemm_f64.h"

namespace gpfx {
namespace {

constexpr int TB = 64;        // diagonal block
constexpr int TCOLS = 128;    // right-hand-side columns per CTA (one per thread)

// X[b][j0 + i][n] (i < r) <- solution of L_jj x = X (TRANS: L_jj^T x = X) for the columns of this CTA.
// L: [K][ldl][ldl] lower triangular; X: [K][M][N] (row stride N).
template <bool TRANS>
__global__ void __launch_bounds__(TCOLS) trsm_diag_kernel(const double* __restrict__ L, long long sL,
                                                          int ldl, double* __restrict__ X,
                                                          long long sX, int N, int j0, int r) {
  __shared__ double Ls[TB][TB + 1];
  const int b = blockIdx.y;
  const double* Lb = L + b * sL + (long long)j0 * ldl + j0;
  for (int e = threadIdx.x; e < TB * TB; e += TCOLS) {
    const int i = e / TB, k = e % TB;
    Ls[i][k] = (i < r && k <= i) ? Lb[(long long)i * ldl + k] : (i == k ? 1.0 : 0.0);
  }
  __syncthreads();
  const int n = blockIdx.x * TCOLS + threadIdx.x;
  if (n >= N) return;
  double* Xb = X + b * sX + (long long)j0 * N + n;
  double x[TB];
#pragma unroll
  for (int i = 0; i < TB; ++i) x[i] = (i < r) ? Xb[(long long)i * N] : 0.0;
  if (!TRANS) {
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      double s = x[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s = fma(-Ls[i][k], x[k], s);
      x[i] = s / Ls[i][i];
    }
  } else {
#pragma unroll
    for (int i = TB - 1; i >= 0; --i) {
      double s = x[i];
#pragma unroll
      for (int k = TB - 1; k > i; --k) s = fma(-Ls[k][i], x[k], s);
      x[i] = s / Ls[i][i];
    }
  }
#pragma unroll
  for (int i = 0; i < TB; ++i)
    if (i < r) Xb[(long long)i * N] = x[i];
}

// sq[b][n] = sum_m A[b][m][n]^2;  ws[b][q][n] = sum_m A[b][m][n] * w[b * sW + m * ldw + q]  (q < nw)
// one thread per column, rows streamed (coalesced across the CTA); fixed summation order
template <int QC>
__global__ void __launch_bounds__(128) col_reduce_kernel(const double* __restrict__ A, long long sA,
                                                         int M, int N, const double* __restrict__ w,
                                                         long long sW, int ldw, int nw, int q0,
                                                         double* __restrict__ sq, double* __restrict__ ws) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  const double* Ab = A + b * sA + n;
  const double* wb = w + b * sW;
  double s = 0.0, acc[QC];
#pragma unroll
  for (int q = 0; q < QC; ++q) acc[q] = 0.0;
  for (int m = 0; m < M; ++m) {
    const double a = Ab[(long long)m * N];
    s = fma(a, a, s);
#pragma unroll
    for (int q = 0; q < QC; ++q)
      if (q0 + q < nw) acc[q] = fma(a, wb[(long long)m * ldw + q0 + q], acc[q]);
  }
  if (q0 == 0) sq[(long long)b * N + n] = s;
#pragma unroll
  for (int q = 0; q < QC; ++q)
    if (q0 + q < nw) ws[((long long)b * nw + q0 + q) * N + n] = acc[q];
}

}  // namespace

int trsm_lower_batched(const double* L, long long sL, int ldl, double* X, long long sX, int K, int M,
                       int N, bool transposed, cudaStream_t stream) {
  if (K <= 0 || M <= 0 || N <= 0) return GPFX_OK;
  const int nb = ceil_div(M, TB);
  dim3 dgrid(ceil_div(N, TCOLS), K);
  for (int step = 0; step < nb; ++step) {
    const int j = transposed ? nb - 1 - step : step;
    const int j0 = j * TB;
    const int r = (M - j0 < TB) ? M - j0 : TB;
    if (step > 0) {
      GemmArgs g;
      g.M = r; g.N = N;
      g.B = nullptr;
      g.ldb = N; g.sB1 = sX;
      g.C = X + (long long)j0 * N; g.ldc = N; g.sC1 = sX;
      g.batch1 = K;
      g.alpha = -1.0; g.beta = 1.0;
      g.lda = ldl; g.sA1 = sL;
      if (!transposed) {
        // X_j -= L[j0 : j0 + r, 0 : j0] X[0 : j0]
        g.A = L + (long long)j0 * ldl;
        g.B = X;
        g.K = j0;
      } else {
        // X_j -= L[j0 + r : M, j0 : j0 + r]^T X[j0 + r : M]
        g.A = L + (long long)(j0 + r) * ldl + j0;
        g.transA = 1;
        g.B = X + (long long)(j0 + r) * N;
        g.K = M - (j0 + r);
      }
      GPFX_TRY(launch_gemm(g, stream));
    }
    if (transposed)
      trsm_diag_kernel<true><<<dgrid, TCOLS, 0, stream>>>(L, sL, ldl, X, sX, N, j0, r);
    else
      trsm_diag_kernel<false><<<dgrid, TCOLS, 0, stream>>>(L, sL, ldl, X, sX, N, j0, r);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

int launch_col_reduce(const double* A, long long sA, int K, int M, int N, const double* w,
                      long long sW, int ldw, int nw, double* sq, double* ws, cudaStream_t stream) {
  if (K <= 0 || M <= 0 || N <= 0) return GPFX_OK;
  dim3 grid(ceil_div(N, 128), K);
  for (int q0 = 0; q0 < (nw > 0 ? nw : 1); q0 += 8) {
    col_reduce_kernel<8><<<grid, 128, 0, stream>>>(A, sA, M, N, w, sW, ldw, nw, q0, sq, ws);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace gpfx
