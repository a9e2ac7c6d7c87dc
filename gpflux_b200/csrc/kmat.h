// Kernel-matrix assembly / backward launchers (see kmat.cu).
#pragma once
#include "common.cuh"

#define GPFX_KERNEL_SE 0
#define GPFX_KERNEL_MATERN12 1
#define GPFX_KERNEL_MATERN32 2
#define GPFX_KERNEL_MATERN52 3

namespace gpfx {

struct KmatArgs {
  // out[k][i][j] = k_k(U[k][i], V[k][j]) (+ jitter on the diagonal); identity padding for
  // i >= nu or j >= nv up to rows_out x cols_out.
  const double* U = nullptr;  // [K][nu][D]   (sU = 0 -> shared)
  const double* V = nullptr;  // [K][nv][D]   (sV = 0 -> shared)
  const double* ls = nullptr;   // [K][D]
  const double* var = nullptr;  // [K]
  double* out = nullptr;
  int K = 1, nu = 0, nv = 0, D = 0;
  long long sU = 0, sV = 0, sOut = 0;
  int ldo = 0, rows_out = 0, cols_out = 0;
  int kind = GPFX_KERNEL_SE;
  int add_jitter = 0;
  double jitter = 0.0;
  // mode 0: stationary kernel k(r2).  Random Fourier features (fourier_features/base.py:59-76):
  // mode 1: out[i][j] = c cos((u_i/l).v_j + bias[j]),            c = sqrt(2 var / nv)
  // mode 2: out[i][j] = c sin((u_i/l).v_j), out[i][nv+j] = c cos, c = sqrt(2 var / (2 nv))
  // (V holds the feature weights W and is NOT divided by the lengthscales)
  int mode = 0;
  const double* bias = nullptr;  // [K][nv] (mode 1)
  long long sBias = 0;
  // mode 3 (backward): out[i][j] = dK_in[i][j] * dk/dr2(r2_ij)  (same layout as out; in place is
  // fine) and part_out[k][tile] = sum over the tile of dK_in * k / var
  const double* dK_in = nullptr;
  double* part_out = nullptr;
  // modes 4 / 5 (backward of modes 1 / 2): dK_in = dPhi [K][nu][ld_dk] (batch stride sdK),
  //   mode 4: out[i][j] = -dPhi[i][j] c sin(z_ij + bias[j])
  //   mode 5: out[i][j] =  dPhi[i][j] c cos(z_ij) - dPhi[i][nv+j] c sin(z_ij)
  // = dLoss/dz, and part_out[k][tile] = sum over the tile of dPhi * Phi / (2 var)  (d variance)
  long long sdK = 0;
  int ld_dk = 0;
};

// Backward of the random-Fourier-feature map (FourierFeaturesBase.call): gradients w.r.t. X (shared by
// the K feature sets: summed), the lengthscales and the variance.  W and the bias are not trainable in
// the reference (fourier_features/random/base.py).
struct RffBwdArgs {
  int K = 1, N = 0, D = 0, L = 0, concat = 0;
  const double* X = nullptr;      // [N][D]
  const double* W = nullptr;      // [K][L][D]
  const double* bias = nullptr;   // [K][L] (cosine variant)
  const double* ls = nullptr;     // [K][D]
  const double* var = nullptr;    // [K]
  const double* dPhi = nullptr;   // [K][N][L] or [K][N][2L]
  double* dX = nullptr;           // [N][D]
  double* dls = nullptr;          // [K][D]
  double* dvar = nullptr;         // [K]
  double* scratch = nullptr;      // rff_bwd_scratch_doubles
};
size_t rff_bwd_scratch_doubles(int K, int N, int D, int L);
int launch_rff_bwd(const RffBwdArgs& a, cudaStream_t stream);

struct KmatBwdArgs {
  const double* U = nullptr;
  const double* V = nullptr;
  const double* ls = nullptr;
  const double* var = nullptr;
  double* dK = nullptr;  // [K][nu][ldk] in: dL/dK ; out: R = dK * dk/dr2 (scratch)
  int K = 1, nu = 0, nv = 0, D = 0;
  long long sU = 0, sV = 0, sK = 0;
  int ldk = 0;
  int kind = GPFX_KERNEL_SE;
  double* dU = nullptr;  // [K][nu][D]
  long long sdU = 0;
  int accumulate_dU = 0;
  double* dV = nullptr;  // [nv][D] when V is shared (sV == 0 && sdV == 0: summed over k) else [K][nv][D]
  long long sdV = 0;
  int accumulate_dV = 0;
  double* dls = nullptr;   // [K][D]
  double* dvar = nullptr;  // [K]
  int accumulate_hyp = 0;
  double* scratch = nullptr;  // kmat_bwd_scratch_doubles(K, nu, nv, D)
  // optional forward kernel matrix (same layout as dK): lets the SE backward skip the exp
  const double* Kmat = nullptr;
  long long sKmat = 0;
  int ldkmat = 0;
  // 1: GEMM form (R pass on the tensor pipe + two DMMA GEMMs against [V,1] / [U,1]); used for the
  // large Kuf matrices.  0: direct row/column passes (better conditioned; used for Kuu).
  // 2: fused single pass (D <= 15): R tile in shared memory, both contractions on DMMA, R never
  // written; falls back to the GEMM form when the shape does not qualify.
  int gemm_form = 0;
  // carved from scratch by the launcher
  double* partU = nullptr;
  double* part = nullptr;
  double* partV = nullptr;
};

int launch_kmat_fwd(const KmatArgs& a, cudaStream_t stream);
size_t kmat_bwd_scratch_doubles(int K, int nu, int nv, int D);
int launch_kmat_bwd(const KmatBwdArgs& a, cudaStream_t stream);

}  // namespace gpfx
