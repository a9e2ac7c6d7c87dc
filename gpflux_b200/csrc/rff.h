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

}  // namespace gpfx
