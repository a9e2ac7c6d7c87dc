// Natural-gradient update of q(u) = N(q_mu, q_sqrt q_sqrt^T) on the device.
//
// Replaces gpflow.optimizers.NaturalGradient._natgrad_apply_gradients as called by
// NatGradModel._apply_backwards_pass (reference gpflux/optimization/keras_natgrad.py:162-193) with
// the default XiNat parameterisation.  GPflow evaluates the step by autodiff through
// expectation -> (mean, chol(var)); in closed form, per latent GP p with m = q_mu[:,p], L = q_sqrt_p,
// S = L L^T and the ordinary gradients gm = dLoss/dm, G = dLoss/dL (lower triangular):
//
//   Shat   = sym( L^-T Phi(L^T G) L^-1 )                 = dLoss/dS  (Cholesky backward, symmetrised)
//   dL/deta1 = gm - 2 Shat m ,  dL/deta2 = Shat           (eta1 = m, eta2 = S + m m^T)
//   nat1' = S^-1 m - gamma dL/deta1 = Lam' m - gamma gm   with  Lam' = -2 nat2' = S^-1 + 2 gamma Shat
//   C = chol(Lam'),  S' = C^-T C^-1,  m' = S' nat1',  L' = chol(S')     (natural_to_meanvarsqrt)
//
// Everything M^3-shaped runs on the batched DMMA GEMM / Cholesky kernels of this library, batched
// over the P latent GPs; matrices are identity-padded to Mp = 64 * 2^q like the Kuu stack.
#include "chol.h"
#include "common.cuh"
#include "gemm_f64.h"
#include "natgrad.h"

namespace gpfx {
namespace {

// dst [P][Mp][Mp] <- tril(src [P][M][M]); zero above the diagonal, `pad_diag` on the padded diagonal
__global__ void ng_pad_tril_kernel(const double* __restrict__ src, double* __restrict__ dst, int M,
                                   int Mp, double pad_diag) {
  const long long p = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)Mp * Mp;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / Mp), j = (int)(e % Mp);
    double v = 0.0;
    if (i < M && j <= i) v = src[p * M * M + (long long)i * M + j];
    else if (i >= M && i == j) v = pad_diag;
    dst[p * Mp * Mp + e] = v;
  }
}

// Lam[i][j] = Sinv[i][j] + gamma (dS[i][j] + dS[j][i])     (in place over Sinv; 2 gamma * sym(dS))
__global__ void ng_lambda_kernel(double* __restrict__ Sinv, const double* __restrict__ dS, int Mp,
                                 double gamma) {
  __shared__ double tile[32][33];
  const long long p = blockIdx.z;
  const double* d = dS + p * Mp * Mp;
  double* s = Sinv + p * Mp * Mp;
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) tile[r][tx] = d[(long long)(bj + r) * Mp + bi + tx];  // dS^T tile
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long o = (long long)(bi + r) * Mp + bj + tx;
    s[o] = s[o] + gamma * (d[o] + tile[tx][r]);
  }
}

// y(p,i) = sum_j A[p][i][j] x(p,j) - gamma sub(p,i)   for i < M ; one warp per row
__global__ void __launch_bounds__(256) ng_matvec_kernel(const double* __restrict__ A, int Mp, int M,
                                                        const double* __restrict__ x, long long sxp,
                                                        long long sxj, const double* __restrict__ sub,
                                                        long long ssp, long long ssi, double gamma,
                                                        double* __restrict__ y, long long syp,
                                                        long long syi) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= M) return;
  const double* row = A + ((long long)p * Mp + i) * Mp;
  double s = 0.0;
  for (int j = lane; j < M; j += 32) s += row[j] * x[p * sxp + j * sxj];
  s = warp_sum(s);
  if (lane == 0) {
    if (sub) s -= gamma * sub[p * ssp + i * ssi];
    y[p * syp + i * syi] = s;
  }
}

// q_sqrt_new[p] = tril(Lp[p][:M,:M])
__global__ void ng_unpad_tril_kernel(const double* __restrict__ Lp, double* __restrict__ out, int M,
                                     int Mp) {
  const long long p = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * M;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    out[p * M * M + e] = (j <= i) ? Lp[p * Mp * Mp + (long long)i * Mp + j] : 0.0;
  }
}

struct NgPlan {
  int Mp;
  size_t mat;      // doubles of one [P][Mp][Mp] stack
  size_t o_buf[7]; // seven stacks
  size_t o_T, o_vec, total;
};

NgPlan ng_plan(int M, int P) {
  NgPlan pl;
  pl.Mp = chol_padded_dim(M);
  pl.mat = (size_t)P * pl.Mp * pl.Mp;
  size_t off = 0;
  for (int i = 0; i < 7; ++i) {
    pl.o_buf[i] = off;
    off += pl.mat;
  }
  pl.o_T = off;
  off += align_up(chol_scratch_doubles(P, pl.Mp), 32);
  pl.o_vec = off;
  off += align_up((size_t)P * pl.Mp, 32);
  pl.total = off;
  return pl;
}

}  // namespace

size_t natgrad_workspace_bytes(int M, int P) { return ng_plan(M, P).total * sizeof(double); }

int natgrad_step(int M, int P, double gamma, const double* q_mu, const double* q_sqrt,
                 const double* dq_mu, const double* dq_sqrt, double* q_mu_new, double* q_sqrt_new,
                 int* info, void* workspace, cudaStream_t stream) {
  if (M <= 0 || P <= 0) {
    set_last_error("gpfx_natgrad_step: bad shape M=%d P=%d", M, P);
    return GPFX_E_SHAPE;
  }
  const NgPlan pl = ng_plan(M, P);
  const int Mp = pl.Mp;
  const long long sM = (long long)Mp * Mp;
  double* ws = reinterpret_cast<double*>(workspace);
  double* B0 = ws + pl.o_buf[0];  // Lq (padded)        -> later Sinv / Lam' / C
  double* B1 = ws + pl.o_buf[1];  // S -> L = chol(S)
  double* B2 = ws + pl.o_buf[2];  // L^-1
  double* B3 = ws + pl.o_buf[3];  // G (padded)         -> later C^-1
  double* B4 = ws + pl.o_buf[4];  // T1                 -> later S' / L'
  double* B5 = ws + pl.o_buf[5];  // T2                 -> later L'^-1 (by-product)
  double* B6 = ws + pl.o_buf[6];  // dS (un-symmetrised)
  double* T = ws + pl.o_T;
  double* nat1 = ws + pl.o_vec;   // [P][Mp]

  const int nb = ceil_div(Mp * Mp, 256) < 1024 ? ceil_div(Mp * Mp, 256) : 1024;
  dim3 pgrid(nb, P);
  ng_pad_tril_kernel<<<pgrid, 256, 0, stream>>>(q_sqrt, B0, M, Mp, 1.0);
  GPFX_LAUNCH_CHECK();
  ng_pad_tril_kernel<<<pgrid, 256, 0, stream>>>(dq_sqrt, B3, M, Mp, 0.0);
  GPFX_LAUNCH_CHECK();

  auto mm = [&](const double* A, int tA, const double* B, int tB, double* C, int tri, int tril_out) {
    GemmArgs g;
    g.A = A; g.B = B; g.C = C;
    g.M = g.N = g.K = Mp;
    g.lda = g.ldb = g.ldc = Mp;
    g.transA = tA; g.transB = tB;
    g.batch1 = P; g.sA1 = g.sB1 = g.sC1 = sM;
    g.tri = tri; g.tril_out = tril_out;
    return launch_gemm(g, stream);
  };
  // S = Lq Lq^T (lower triangle is all the factorisation reads), L = chol(S), L^-1
  GPFX_TRY(mm(B0, 0, B0, 1, B1, TRI_LOWER_A | TRI_UPPER_B, 1));
  GPFX_TRY(potrf_trtri_batched(B1, B2, T, P, Mp, info, stream));
  // dS = L^-T Phi(L^T G) L^-1
  GPFX_TRY(chol_backward_batched(B1, B2, B3, B4, B5, B6, P, Mp, stream));
  // Sinv = L^-T L^-1 ; Lam' = Sinv + 2 gamma sym(dS)
  GPFX_TRY(mm(B2, 1, B2, 0, B0, TRI_UPPER_A | TRI_LOWER_B, 0));
  {
    dim3 grid(Mp / 32, Mp / 32, P);
    ng_lambda_kernel<<<grid, 256, 0, stream>>>(B0, B6, Mp, gamma);
    GPFX_LAUNCH_CHECK();
  }
  // nat1' = Lam' m - gamma gm        (q_mu, dq_mu are [M][P])
  {
    dim3 grid(ceil_div(M, 8), P);
    ng_matvec_kernel<<<grid, 256, 0, stream>>>(B0, Mp, M, q_mu, 1, P, dq_mu, 1, P, gamma, nat1, Mp, 1);
    GPFX_LAUNCH_CHECK();
  }
  // C = chol(Lam'), C^-1 ; S' = C^-T C^-1 ; m' = S' nat1' ; L' = chol(S')
  GPFX_TRY(potrf_trtri_batched(B0, B3, T, P, Mp, info + P, stream));
  GPFX_TRY(mm(B3, 1, B3, 0, B4, TRI_UPPER_A | TRI_LOWER_B, 0));
  {
    dim3 grid(ceil_div(M, 8), P);
    ng_matvec_kernel<<<grid, 256, 0, stream>>>(B4, Mp, M, nat1, Mp, 1, nullptr, 0, 0, 0.0, q_mu_new, 1, P);
    GPFX_LAUNCH_CHECK();
  }
  GPFX_TRY(potrf_trtri_batched(B4, B5, T, P, Mp, info + 2 * P, stream));
  {
    dim3 grid(ceil_div(M * M, 256) < 1024 ? ceil_div(M * M, 256) : 1024, P);
    ng_unpad_tril_kernel<<<grid, 256, 0, stream>>>(B4, q_sqrt_new, M, Mp);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace gpfx
