// FP64 DMMA GEMM with triangular k-range skipping, k-segment accumulation, split-K and fused
// column-reduction epilogues.  This is the workhorse of the SVGP conditional (SURVEY §8a rows
// a4-a7, a14): A = Linv*Kuf, B_p = Lq_p^T*A, and every backward contraction.
#pragma once
#include "common.cuh"

struct gpfx_gemm_sample;

namespace gpfx {

enum : int {
  TRI_NONE = 0,
  TRI_LOWER_A = 1,  // op(A)[m,k] == 0 for k > m  -> k <  m0 + BM
  TRI_UPPER_A = 2,  // op(A)[m,k] == 0 for k < m  -> k >= m0
  TRI_LOWER_B = 4,  // op(B)[k,n] == 0 for k < n  -> k >= n0
  TRI_UPPER_B = 8,  // op(B)[k,n] == 0 for k > n  -> k <  n0 + BN
};

struct GemmArgs {
  // C[b] = alpha * (sum_seg op(A[b,seg]) * op(B[b,seg])) (*colscale[b][n]) + beta * C[b]
  const double* A = nullptr;
  const double* B = nullptr;
  double* C = nullptr;
  int M = 0, N = 0, K = 0;  // op(A) [M,K], op(B) [K,N], C [M,N]
  int lda = 0, ldb = 0, ldc = 0;
  // two-level batch: b = b1 * batch2 + b2; offsets b1*s?1 + b2*s?2 (elements)
  int batch1 = 1, batch2 = 1;
  long long sA1 = 0, sB1 = 0, sC1 = 0;
  long long sA2 = 0, sB2 = 0, sC2 = 0;
  int nseg = 1;
  long long gA = 0, gB = 0;  // k-segment strides
  int transA = 0;            // 0: A stored [M][K] (k contiguous); 1: stored [K][M]
  int transB = 0;            // 0: B stored [K][N] (n contiguous); 1: stored [N][K]
  int tri = TRI_NONE;
  int tril_out = 0;  // keep only the lower triangle of C (rest written as beta*C with acc = 0)
  double alpha = 1.0, beta = 0.0;
  double diag_scale = 1.0;  // multiplies acc on the diagonal (r == c); used for Phi() in chol-bwd
  const double* colscale = nullptr;  // [batch][N], multiplies acc columns (after alpha)
  long long sColscale = 0;
  // per-k scale of the contraction (TMA kernel, NT products - ask gemm_kscale_supported first):
  //   C[b] = alpha * sum_k op(A)[m,k] * kscale[b][k] * op(B)[k,n]
  // (the layer backward: dLq_p = tril(A diag(2 gv_p) B_p^T) without a scaled copy of B_p)
  const double* kscale = nullptr;
  long long sKscale = 0;
  // optional fused column reductions over the CTA's BM rows of the *stored* value v:
  double* sumsq = nullptr;  // [batch][gridM][N]      sum_m v^2
  long long sSumsq = 0;     // batch stride (= gridM*N); gridM = ceil(M/BM)
  const double* wvec = nullptr;  // [batch][M][ldw] weights, nw columns used
  int ldw = 0, nw = 0;
  long long sW = 0;
  double* wsum = nullptr;  // [batch][gridM][nw][N]  sum_m v * w[m][j]
  long long sWsum = 0;
  // optional add-in terms of the stored value, fused into the epilogue (TMA kernel, interior tiles of a
  // plain product only - ask gemm_addin_supported first):
  //   C[b][m][n] = alpha * acc + addscale * addcol[b][n] * addmat[b][m][n] + r1u[b][m] * r1v[b][n]
  // and, when rowdot is set, rowdot[tile_x][b][m] = sum over the tile's 128 columns of
  // addmat[b][m][n] * r1v[b][n]   (the layer backward: dA = q_mu gm^T - 2 A diag(gvsum) + Lq Bs and
  // dq_mu = A gm in the one pass that reads A)
  const double* addmat = nullptr;
  int ldadd = 0;
  long long sAdd = 0;
  const double* addcol = nullptr;
  long long sAddcol = 0;
  double addscale = 1.0;
  const double* r1u = nullptr;  // element m at r1u[b*sR1u + m*ldr1u]
  int ldr1u = 1;
  long long sR1u = 0;
  const double* r1v = nullptr;  // [batch][N]
  long long sR1v = 0;
  double* rowdot = nullptr;     // [ceil(N/128)][batch][M]
  // split-K: partial products go to splitws [splits][batch][M][N] (ld = N), then reduced
  int splits = 1;
  double* splitws = nullptr;
  int small_tile = 0;  // 1: force the 64x64 CTA tile
  int cfg = -1;        // >= 0: force a tile configuration (see pick_cfg in gemm_f64.cu)
};

// Rows per CTA tile for the configuration the launcher will pick (needed to size sumsq/wsum).
int gemm_block_m(const GemmArgs& a);
int gemm_grid_m(const GemmArgs& a);
int launch_gemm(const GemmArgs& a, cudaStream_t stream);
// TMA-fed kernel (gemm_f64_tma.cu): CTA tile height when the problem shape is eligible (0 otherwise),
// pointer-alignment check, launcher
int gemm_tma_block_m(const GemmArgs& a);
bool gemm_tma_pointers_ok(const GemmArgs& a);
int launch_gemm_tma(const GemmArgs& a, cudaStream_t stream);
// true when launch_gemm will honour the add-in terms (addmat / addcol / r1u / r1v / rowdot) of `a`
bool gemm_addin_supported(const GemmArgs& a);
// true when launch_gemm will honour `kscale` of `a`
bool gemm_kscale_supported(const GemmArgs& a);
// tuning hook: force the tile configuration of large problems (-1 = automatic)
void gemm_force_cfg(int cfg);
// per-launch device timing (see gpfx_gemm_profile_* in include/gpfx.h)
void gemm_profile_enable(int on);
int gemm_profile_read(::gpfx_gemm_sample* out, int max);

}  // namespace gpfx
