// Blocked triangular solves without an explicit inverse (trsm.cu): the reference-faithful form of
// base_conditional's tf.linalg.triangular_solve.
#pragma once
#include "common.cuh"

namespace gpfx {

// In place: X[b] <- L[b]^-1 X[b] (transposed: L[b]^-T X[b]) for b < K.  L [K][ldl][ldl] lower triangular
// (batch stride sL), X [K][M][N] row-major (batch stride sX).  Forward / back substitution in 64-row
// blocks: substitution kernels on the diagonal blocks, DMMA GEMM updates off the diagonal.
int trsm_lower_batched(const double* L, long long sL, int ldl, double* X, long long sX, int K, int M,
                       int N, bool transposed, cudaStream_t stream);

// Column reductions the inverse-GEMM path gets from its fused epilogue:
// sq[b][n] = sum_m A[b][m][n]^2,  ws[b][q][n] = sum_m A[b][m][n] w[b * sW + m * ldw + q]  (q < nw).
int launch_col_reduce(const double* A, long long sA, int K, int M, int N, const double* w,
                      long long sW, int ldw, int nw, double* sq, double* ws, cudaStream_t stream);

}  // namespace gpfx
