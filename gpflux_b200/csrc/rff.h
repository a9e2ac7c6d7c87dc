// Fused Wilson-sample evaluation (rff.cu).
#pragma once
#include "common.cuh"

namespace gpfx {

struct WilsonArgs {
  int kind = 0;
  int S = 1, N = 0, D = 0, P = 0, L = 0, M = 0;
  const double* X = nullptr;  // [S][N][D] (sX = N*D) or shared [N][D] (sX = 0)
  long long sX = 0;
  const double* W = nullptr;    // [L][D]  feature weights
  const double* b = nullptr;    // [L]     feature phases
  const double* ls = nullptr;   // [D]
  const double* var = nullptr;  // [1]
  const double* Z = nullptr;    // [M][D]
  const double* w = nullptr;    // [S][L][P] prior weights
  const double* v = nullptr;    // [S][M][P] function-space update weights
  const double* mean_add = nullptr;  // [S][N][P] or null
  double* out = nullptr;             // [S][N][P]
};

int launch_wilson_eval(const WilsonArgs& a, cudaStream_t stream);

// One-off set-up of S Matheron-rule draws (reference sample.py:158-181 + math.py:42-62):
//   w_s = sqrt(lambda) * eps_w[s];  u_s = q_mu + (q_sqrt eps_u[s])^T;  u_s <- L_uu u_s (whitened);
//   v_s = Kuu^-1 (u_s - Phi(Z) w_s)
struct WilsonSetupArgs {
  int kind = 0, whiten = 1;
  int S = 1, M = 0, D = 0, P = 0, L = 0;
  double jitter = 1e-6;
  const double* Z = nullptr;       // [M][D]
  const double* ls = nullptr;      // [D]
  const double* var = nullptr;     // [1]
  const double* q_mu = nullptr;    // [M][P]
  const double* q_sqrt = nullptr;  // [P][M][M]
  const double* W = nullptr;       // [L][D]
  const double* b = nullptr;       // [L]
  const double* coeffs = nullptr;  // [L]  feature coefficients lambda
  const double* eps_w = nullptr;   // [S][L][P]
  const double* eps_u = nullptr;   // [S][P][M]
  double* w = nullptr;             // [S][L][P]
  double* v = nullptr;             // [S][M][P]
  int* info = nullptr;             // [1] Cholesky status of Kuu
  double* scratch = nullptr;       // wilson_setup_scratch_doubles
};
size_t wilson_setup_scratch_doubles(int S, int M, int D, int P, int L);
int launch_wilson_setup(const WilsonSetupArgs& a, cudaStream_t stream);

}  // namespace gpfx
