// Batched Cholesky / triangular inverse / Cholesky backward on identity-padded stacks (chol.cu).
#pragma once
#include "common.cuh"

namespace gpfx {

// Mp = 64 * 2^q >= M
int chol_padded_dim(int M);
// scratch doubles needed by potrf_trtri_batched (T)
size_t chol_scratch_doubles(int K, int Mp);
// L [K][Mp][Mp]: in = padded Kuu (lower triangle read), out = Cholesky factor (upper zeroed).
// Linv [K][Mp][Mp]: out = L^-1.  info [K]: 0, or 1 + index of the first non-positive pivot.
int potrf_trtri_batched(double* L, double* Linv, double* T, int K, int Mp, int* info,
                        cudaStream_t stream);
// dK [K][Mp][Mp] = L^-T Phi(L^T dL) L^-1 (un-symmetrised).  T1, T2: [K][Mp][Mp] scratch.
int chol_backward_batched(const double* L, const double* Linv, const double* dL, double* T1,
                          double* T2, double* dK, int K, int Mp, cudaStream_t stream);

}  // namespace gpfx
