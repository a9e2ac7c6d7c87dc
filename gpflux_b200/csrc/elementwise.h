// HBM-bound streaming kernels of the SVGP layer (elementwise.cu): whitened KL + tril copy,
// finalize (mean/variance assembly + reparameterised sample), backward preparation, Gaussian
// likelihood.
#pragma once
#include "common.cuh"

namespace gpfx {

// ---- a9: whitened KL (gp_layer.py:303-311 -> gauss_kl) + clean lower-triangular copy of q_sqrt
size_t kl_scratch_doubles(int M, int P);
// Lq (nullable) [P][M][M] <- tril(q_sqrt);  kl[0] <- KL[q(v) || N(0, I)]
int launch_kl_white(const double* q_mu, const double* q_sqrt, double* Lq, double* kl,
                    double* scratch, int M, int P, cudaStream_t stream);
// dq_mu[m][p] (+)= g * q_mu ; dq_sqrt[p][i][j] (+)= g * (Lq_ij - [i==j]/Lq_ii)  (lower part)
int launch_kl_white_bwd(const double* q_mu, const double* Lq, const double* g_kl, double* dq_mu,
                        double* dq_sqrt, int M, int P, int accumulate, cudaStream_t stream);

// dst[p] = tril(src[p]) for P matrices of size M x M
int launch_tril_copy(const double* src, double* dst, int P, int M, cudaStream_t stream);

// ---- a5-a8: assemble mean / variance from the GEMM column partials and draw the sample
struct FinalizeArgs {
  int N = 0, P = 0, K = 1;
  int gridM_A = 0, gridM_B = 0;
  const double* wsum = nullptr;   // [K][gridM_A][nw][N], nw = P (shared) or 1 (separate)
  const double* sqA = nullptr;    // [K][gridM_A][N]
  const double* sqB = nullptr;    // [P][gridM_B][N]
  const double* variance = nullptr;  // [K]
  const double* mean_add = nullptr;  // [N][P] or null
  const double* eps = nullptr;       // [N][P] or null
  double* fmean = nullptr;           // [N][P]
  double* fvar = nullptr;            // [N][P]
  double* fsample = nullptr;         // [N][P] or null
};
int launch_finalize(const FinalizeArgs& a, cudaStream_t stream);

// standalone a8: f = mean + sqrt(var) * eps  (n elements)
int launch_reparam_sample(const double* mean, const double* var, const double* eps, double* f,
                          long long n, cudaStream_t stream);

// a8 with the noise generated in the kernel (Philox4x32-10; counter = (pair index, offset), key = seed).
// raw: out [n_pairs][4] uint32, the generator's words (what the bit-exact test reads).
// normal: mean == nullptr -> out = eps; else out = mean + sqrt(var) eps, eps also stored if eps_out.
int launch_philox_raw(unsigned long long seed, unsigned long long offset, uint32_t* out,
                      long long n_pairs, cudaStream_t stream);
int launch_philox_normal(unsigned long long seed, unsigned long long offset, const double* mean,
                         const double* var, double* out, double* eps_out, long long n,
                         cudaStream_t stream);

// ---- a14 helpers
struct BwdPrepArgs {
  int N = 0, P = 0, K = 1;
  const double* g_f = nullptr;     // [N][P] or null : dLoss/d fsample
  const double* g_mean = nullptr;  // [N][P] or null : direct dLoss/d fmean
  const double* g_var = nullptr;   // [N][P] or null : direct dLoss/d fvar
  const double* eps = nullptr;     // [N][P] or null
  const double* fvar = nullptr;    // [N][P]
  double* dmean = nullptr;         // [N][P] total dLoss/d fmean (for the host's mean function)
  double* gmT = nullptr;           // [P][N]
  double* gv2T = nullptr;          // [P][N]  2 * dLoss/d fvar
  double* gvsum = nullptr;         // [K][N]  sum_{p in k} dLoss/d fvar
};
int launch_bwd_prep(const BwdPrepArgs& a, cudaStream_t stream);

// B[p][m][n] *= s[p][n]
int launch_scale_cols(double* B, const double* s, int P, int M, int N, cudaStream_t stream);

// dA[k][m][n] = sum_{p in k} q_mu[m][p] * gmT[p][n] - 2 * A[k][m][n] * gvsum[k][n]
int launch_dA_init(const double* q_mu, const double* gmT, const double* A, const double* gvsum,
                   double* dA, int M, int N, int P, int K, cudaStream_t stream);

// dq_mu[m][p] = sum_n A[k(p)][m][n] * gmT[p][n]
// dq_mu = A gm (row-wise long contraction): scratch of dqmu_scratch_doubles; needs even N and
// 16-byte aligned A / gmT (dqmu_supported), else the caller falls back to the split-K GEMM
size_t dqmu_scratch_doubles(int M, int N, int P, int K);
bool dqmu_supported(const double* A, const double* gmT, int N);
int launch_dqmu(const double* A, const double* gmT, double* dq_mu, double* scratch, int M, int N,
                int P, int K, cudaStream_t stream);

// out[r] (+)= sum_n x[r][n]
int launch_dqmu_finish(const double* part, double* dq_mu, int M, int P, int K, int nseg, cudaStream_t stream);
int launch_row_sums(const double* x, int rows, int N, double* out, int accumulate,
                    cudaStream_t stream);

// ---- Gaussian likelihood (likelihood_layer.py:84-93): sum of variational expectations + grads
// ---- LatentVariableLayer (gpflux/layers/latent_variable_layer.py:140-228), diagonal Gaussians:
// out[n] = concat(X[n] (D), mean[n] + std[n] * eps[n] (W));  kl_sum = sum_{n,w} KL[N(m,s^2) || N(m0,s0^2)].
// post_stride = W (per-row posterior) or 0 (mean/std are [W] vectors, e.g. sampling the prior).
size_t latent_scratch_doubles(long long N, int W);
int launch_latent_forward(const double* X, const double* mean, const double* std, int post_stride,
                          const double* prior_mean, const double* prior_std, const double* eps,
                          double* out, double* kl_sum, double* scratch, long long N, int D, int W,
                          cudaStream_t stream);
// dX = g_out[:, :D]; dmean = g_out[:, D:] + g_kl (m - m0)/s0^2; dstd = g_out[:, D:] eps + g_kl (s/s0^2 - 1/s)
int launch_latent_backward(const double* mean, const double* std, const double* prior_mean,
                           const double* prior_std, const double* eps, const double* g_out,
                           const double* g_kl, double* dX, double* dmean, double* dstd, long long N,
                           int D, int W, cudaStream_t stream);

size_t gaussian_ve_scratch_doubles(long long n);
// ve_sum[0] = sum_{n,q} VE ; if scale != null: dFmu, dFvar = scale * dVE/d., dnoise[0] = scale * sum dVE/dnoise
int launch_gaussian_ve(const double* Fmu, const double* Fvar, const double* Y,
                       const double* noise_var, long long n, double* ve_sum, const double* scale,
                       double* dFmu, double* dFvar, double* dnoise, double* scratch,
                       cudaStream_t stream);

}  // namespace gpfx
