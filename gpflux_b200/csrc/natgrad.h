// Natural-gradient step for q(u) (natgrad.cu).
#pragma once
#include "common.cuh"

namespace gpfx {

size_t natgrad_workspace_bytes(int M, int P);
// q_mu [M][P], q_sqrt [P][M][M] (lower triangle read), ordinary gradients dq_mu / dq_sqrt of the
// same shapes; writes the updated parameters (may alias the inputs) and info [3][P] (Cholesky
// status of S, of the new precision, of the new covariance).
int natgrad_step(int M, int P, double gamma, const double* q_mu, const double* q_sqrt,
                 const double* dq_mu, const double* dq_sqrt, double* q_mu_new, double* q_sqrt_new,
                 int* info, void* workspace, cudaStream_t stream);

}  // namespace gpfx
