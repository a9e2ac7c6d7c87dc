// Efficient posterior sampling (Wilson et al. 2020; reference gpflux/sampling/sample.py:184-198):
//   f(X) = Phi(X) w + K_XZ v,   Phi(X) = sqrt(2 var / L) cos((X/l) W^T + b)
// fused so that neither Phi [N,L] (5.1 TB per layer at cfg-5 sizes) nor K_XZ [N,M] ever reaches HBM:
// both contractions run on the FP64 tensor pipe with the cosine / kernel function applied to the
// accumulator fragments in between (wilson_dmma_kernel).  The kernel is bound by FP64 cos throughput
// (SURVEY §7 hard part 6), not by DMMA or HBM.
#include "rff.h"

#include <math.h>

#include <algorithm>
#include <type_traits>

#include "chol.h"
#include "gemm_f64.h"
#include "kmat.h"

namespace gpfx {
namespace {

__device__ __forceinline__ double k_of_r2_dev(int kind, double r2, double var) {
  if (kind == GPFX_KERNEL_SE) return var * exp(-0.5 * r2);
  const double r = sqrt(fmax(r2, 1e-36));
  if (kind == GPFX_KERNEL_MATERN12) return var * exp(-r);
  if (kind == GPFX_KERNEL_MATERN32) {
    const double s = 1.7320508075688772 * r;
    return var * (1.0 + s) * exp(-s);
  }
  const double s = 2.23606797749979 * r;
  return var * (1.0 + s + (5.0 / 3.0) * (r * r)) * exp(-s);
}

constexpr int WT = 128;  // threads per block (4 warps)
constexpr int FC = 64;   // features / inducing points staged per chunk
constexpr int RB = 4;    // 8-row blocks per warp -> 32 rows per warp, 128 rows per CTA

// smallest row stride (in doubles) >= n with stride % 16 == 4: the 8 rows x 4 doubles of a DMMA
// B fragment then fall into distinct banks
__host__ __device__ constexpr int frag_stride(int n) { return ((n + 11) / 16) * 16 + 4; }

// Flash-style evaluation on the FP64 tensor pipe.  Per warp: 32 query rows as DMMA A fragments (kept
// in registers for the whole kernel); per group of 8 features
//   proj  = X~ W_g^T            DMMA, k = D (padded to 4 DK)
//   phi   = cos(proj + b_g)     on the accumulator fragment, in registers
//   out  += phi w_g             DMMA, k = 8: the C fragment of the first product IS the A operand of the
//                               second - lane (lr, lc) holds phi[lr][2 lc], phi[lr][2 lc + 1], which serve
//                               as k-slot lc of two k4 steps whose B rows are the features 2 lc / 2 lc + 1
// and the same for the function-space update with k(x, z) in place of the cosine.  Neither Phi(X) nor
// K_XZ exists outside registers.   DK = ceil(D / 4), NP = ceil(P / 8).
template <int DK, int NP>
__global__ void __launch_bounds__(WT) wilson_dmma_kernel(const WilsonArgs a) {
  constexpr int Dp = 4 * DK, Pp = 8 * NP;
  constexpr int SW = frag_stride(Dp);  // row stride of the staged W / Z chunk
  constexpr int SP = Pp + 4;           // row stride of the staged w / v chunk (rows 2 lc: stride % 8 == 4)
  __shared__ double Ws[FC * SW];
  __shared__ double ws[FC * SP];
  __shared__ double bs[FC];  // feature phases, later |z|^2
  const int s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const double* X = a.X + (long long)s * a.sX;
  const double* wS = a.w + (long long)s * a.L * a.P;
  const double* vS = a.v + (long long)s * a.M * a.P;
  const int row0 = blockIdx.x * (WT / 32 * RB * 8) + warp * (RB * 8);

  // A fragments of the scaled inputs and the row norms
  double xa[RB][DK], xn[RB];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    const int r = row0 + rb * 8 + lr;
    double n2 = 0.0;
#pragma unroll
    for (int k4 = 0; k4 < DK; ++k4) {
      const int d = k4 * 4 + lc;
      const double v = (d < a.D && r < a.N) ? X[(long long)r * a.D + d] / a.ls[d] : 0.0;
      xa[rb][k4] = v;
      n2 += v * v;
    }
    n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
    n2 += __shfl_xor_sync(0xffffffffu, n2, 2);
    xn[rb] = n2;
  }
  double acc[RB][NP][2];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
#pragma unroll
    for (int np = 0; np < NP; ++np) acc[rb][np][0] = acc[rb][np][1] = 0.0;

  const double var = a.var[0];
  // one pass over nl staged rows: PRIOR = cosine features, else kernel values
  auto contract = [&](int nl, auto prior_c) {
    constexpr bool PRIOR = decltype(prior_c)::value;
    for (int g = 0; g * 8 < nl; ++g) {
      double b1[DK], wb0[NP], wb1[NP];
#pragma unroll
      for (int k4 = 0; k4 < DK; ++k4) b1[k4] = Ws[(g * 8 + lr) * SW + k4 * 4 + lc];
#pragma unroll
      for (int np = 0; np < NP; ++np) {
        wb0[np] = ws[(g * 8 + 2 * lc) * SP + np * 8 + lr];
        wb1[np] = ws[(g * 8 + 2 * lc + 1) * SP + np * 8 + lr];
      }
      const double e0 = bs[g * 8 + 2 * lc], e1 = bs[g * 8 + 2 * lc + 1];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < DK; ++k4) dmma884(c0, c1, xa[rb][k4], b1[k4]);
        double f0, f1;
        if (PRIOR) {
          f0 = cos(c0 + e0);
          f1 = cos(c1 + e1);
        } else {
          f0 = k_of_r2_dev(a.kind, fmax(xn[rb] + e0 - 2.0 * c0, 0.0), var);
          f1 = k_of_r2_dev(a.kind, fmax(xn[rb] + e1 - 2.0 * c1, 0.0), var);
        }
#pragma unroll
        for (int np = 0; np < NP; ++np) {
          dmma884(acc[rb][np][0], acc[rb][np][1], f0, wb0[np]);
          dmma884(acc[rb][np][0], acc[rb][np][1], f1, wb1[np]);
        }
      }
    }
  };
  // stage rows [r0, r0 + nl) of src [rows][D] (optionally scaled by 1/l) and wsrc [rows][P]; rows beyond nl
  // are zero-filled (their weights are zero, so they contribute nothing)
  auto stage = [&](const double* src, const double* wsrc, int r0, int nl, bool scale) {
    __syncthreads();
    for (int e = tid; e < FC * Dp; e += WT) {
      const int l = e / Dp, d = e - l * Dp;
      double v = 0.0;
      if (l < nl && d < a.D) {
        v = src[(long long)(r0 + l) * a.D + d];
        if (scale) v /= a.ls[d];
      }
      Ws[l * SW + d] = v;
    }
    for (int e = tid; e < FC * Pp; e += WT) {
      const int l = e / Pp, q = e - l * Pp;
      ws[l * SP + q] = (l < nl && q < a.P) ? wsrc[(long long)(r0 + l) * a.P + q] : 0.0;
    }
  };

  // ---- weight-space prior: sum_l cos(x.W_l + b_l) w_l
  for (int l0 = 0; l0 < a.L; l0 += FC) {
    const int nl = min(FC, a.L - l0);
    stage(a.W, wS, l0, nl, false);
    for (int e = tid; e < FC; e += WT) bs[e] = (e < nl) ? a.b[l0 + e] : 0.0;
    __syncthreads();
    contract(nl, std::true_type{});
  }
  const double c = sqrt(2.0 * var / (double)a.L);
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
#pragma unroll
    for (int np = 0; np < NP; ++np) {
      acc[rb][np][0] *= c;
      acc[rb][np][1] *= c;
    }

  // ---- function-space update: sum_m k(x, z_m) v_m
  for (int m0 = 0; m0 < a.M; m0 += FC) {
    const int nm = min(FC, a.M - m0);
    stage(a.Z, vS, m0, nm, true);
    __syncthreads();
    for (int e = tid; e < FC; e += WT) {  // |z|^2 of the staged (scaled) rows
      double n2 = 0.0;
      for (int d = 0; d < Dp; ++d) n2 += Ws[e * SW + d] * Ws[e * SW + d];
      bs[e] = n2;
    }
    __syncthreads();
    contract(nm, std::false_type{});
  }

#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    const int r = row0 + rb * 8 + lr;
    if (r >= a.N) continue;
    const long long o = ((long long)s * a.N + r) * a.P;
#pragma unroll
    for (int np = 0; np < NP; ++np)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = np * 8 + 2 * lc + e;
        if (q < a.P) a.out[o + q] = acc[rb][np][e] + (a.mean_add ? a.mean_add[o + q] : 0.0);
      }
  }
}

// Scalar form for tiny input / output dims (D, P <= 2): two FMAs per cosine, where the DMMA form would
// spend three mostly-padding tensor instructions per 64 cosines.  Each thread owns R query rows.
constexpr int SWT = 128;
constexpr int SFC = 128;  // features / inducing points staged per chunk (scalar kernel)

template <int DMAX, int PMAX, int R>
__global__ void __launch_bounds__(SWT) wilson_scalar_kernel(const WilsonArgs a) {
  extern __shared__ double sm[];
  double* Ws = sm;                 // [SFC][D]
  double* bs = Ws + SFC * a.D;      // [SFC]
  double* ws = bs + SFC;            // [SFC][P]
  const int s = blockIdx.y;
  const int tid = threadIdx.x;
  const double* X = a.X + (long long)s * a.sX;
  const double* wS = a.w + (long long)s * a.L * a.P;
  const double* vS = a.v + (long long)s * a.M * a.P;

  double x[R][DMAX];
  double acc[R][PMAX];
  int rows[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rows[r] = (blockIdx.x * R + r) * SWT + tid;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      x[r][d] = (d < a.D && rows[r] < a.N) ? X[(long long)rows[r] * a.D + d] / a.ls[d] : 0.0;
#pragma unroll
    for (int p = 0; p < PMAX; ++p) acc[r][p] = 0.0;
  }

  // ---- weight-space prior: sum_l cos(x.W_l + b_l) w_l
  for (int l0 = 0; l0 < a.L; l0 += SFC) {
    const int nl = min(SFC, a.L - l0);
    __syncthreads();
    for (int e = tid; e < nl * a.D; e += SWT) Ws[e] = a.W[(long long)l0 * a.D + e];
    for (int e = tid; e < nl; e += SWT) bs[e] = a.b[l0 + e];
    for (int e = tid; e < nl * a.P; e += SWT) ws[e] = wS[(long long)l0 * a.P + e];
    __syncthreads();
    for (int l = 0; l < nl; ++l) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double proj = bs[l];
#pragma unroll
        for (int d = 0; d < DMAX; ++d)
          if (d < a.D) proj += x[r][d] * Ws[l * a.D + d];
        const double phi = cos(proj);
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
          if (p < a.P) acc[r][p] += phi * ws[l * a.P + p];
      }
    }
  }
  const double var = a.var[0];
  const double c = sqrt(2.0 * var / (double)a.L);
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int p = 0; p < PMAX; ++p) acc[r][p] *= c;

  // ---- function-space update: sum_m k(x, z_m) v_m   (Z scaled by 1/l while staging)
  for (int m0 = 0; m0 < a.M; m0 += SFC) {
    const int nm = min(SFC, a.M - m0);
    __syncthreads();
    for (int e = tid; e < nm * a.D; e += SWT) Ws[e] = a.Z[(long long)m0 * a.D + e] / a.ls[e % a.D];
    for (int e = tid; e < nm * a.P; e += SWT) ws[e] = vS[(long long)m0 * a.P + e];
    __syncthreads();
    for (int m = 0; m < nm; ++m) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < DMAX; ++d)
          if (d < a.D) {
            const double t = x[r][d] - Ws[m * a.D + d];
            r2 += t * t;
          }
        const double kv = k_of_r2_dev(a.kind, r2, var);
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
          if (p < a.P) acc[r][p] += kv * ws[m * a.P + p];
      }
    }
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (rows[r] >= a.N) continue;
    const long long o = ((long long)s * a.N + rows[r]) * a.P;
#pragma unroll
    for (int p = 0; p < PMAX; ++p)
      if (p < a.P) a.out[o + p] = acc[r][p] + (a.mean_add ? a.mean_add[o + p] : 0.0);
  }
}

template <int DMAX, int PMAX, int R>
int launch_scalar(const WilsonArgs& a, cudaStream_t stream) {
  dim3 grid(ceil_div(a.N, SWT * R), a.S);
  const size_t smem = ((size_t)SFC * a.D + SFC + (size_t)SFC * a.P) * sizeof(double);
  wilson_scalar_kernel<DMAX, PMAX, R><<<grid, SWT, smem, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

template <int DK, int NP>
int launch_t(const WilsonArgs& a, cudaStream_t stream) {
  dim3 grid(ceil_div(a.N, WT / 32 * RB * 8), a.S);
  wilson_dmma_kernel<DK, NP><<<grid, WT, 0, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

template <int DK>
int launch_p(const WilsonArgs& a, cudaStream_t stream) {
  if (a.P <= 8) return launch_t<DK, 1>(a, stream);
  if (a.P <= 16) return launch_t<DK, 2>(a, stream);
  return launch_t<DK, 4>(a, stream);
}

}  // namespace

int launch_wilson_eval(const WilsonArgs& a, cudaStream_t stream) {
  if (a.S <= 0 || a.N <= 0) return GPFX_OK;
  if (a.D <= 0 || a.P <= 0 || a.L <= 0 || a.M <= 0 || a.D > 32 || a.P > 32) {
    set_last_error("gpfx_wilson_eval: unsupported shape D=%d P=%d L=%d M=%d (D, P <= 32)", a.D, a.P,
                   a.L, a.M);
    return GPFX_E_UNSUPPORTED;
  }
  if (a.D == 1 && a.P == 1) return launch_scalar<1, 1, 4>(a, stream);
  if (a.D <= 2 && a.P <= 2) return launch_scalar<2, 2, 4>(a, stream);
  if (a.D <= 4) return launch_p<1>(a, stream);
  if (a.D <= 8) return launch_p<2>(a, stream);
  if (a.D <= 16) return launch_p<4>(a, stream);
  return launch_p<8>(a, stream);
}


// ---- Matheron-rule set-up (a13) ---------------------------------------------------------------
namespace {
// w[s][l][p] = sqrt(lambda_l) eps_w[s][l][p]
__global__ void wilson_prior_weights_kernel(const double* __restrict__ coeffs, const double* __restrict__ eps,
                                            double* __restrict__ w, long long total, int L, int P) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((e / P) % L);
    w[e] = sqrt(coeffs[l]) * eps[e];
  }
}
// u[s][m][p] = q_mu[m][p] + T[p][m][s]
__global__ void wilson_u_kernel(const double* __restrict__ q_mu, const double* __restrict__ T, double* __restrict__ u,
                                int S, int M, int P) {
  const long long total = (long long)S * M * P;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(e % P);
    const int m = (int)((e / P) % M);
    const int s = (int)(e / ((long long)P * M));
    u[e] = q_mu[(long long)m * P + p] + T[((long long)p * M + m) * S + s];
  }
}
inline int grid_for(long long n) { return (int)std::min<long long>(ceil_div_ll(n, 256), 2048); }
}  // namespace

size_t wilson_setup_scratch_doubles(int S, int M, int D, int P, int L) {
  const size_t Mp = (size_t)chol_padded_dim(M);
  return 2 * Mp * Mp + chol_scratch_doubles(1, (int)Mp) + (size_t)M * L + 3 * (size_t)S * M * P + 64;
}

int launch_wilson_setup(const WilsonSetupArgs& a, cudaStream_t stream) {
  const int S = a.S, M = a.M, D = a.D, P = a.P, L = a.L;
  if (S <= 0 || M <= 0 || D <= 0 || P <= 0 || L <= 0) {
    set_last_error("gpfx_wilson_setup: bad shape");
    return GPFX_E_SHAPE;
  }
  const int Mp = chol_padded_dim(M);
  double* Lu = a.scratch;                        // [Mp][Mp] Kuu + jitter I, then its Cholesky factor
  double* Linv = Lu + (size_t)Mp * Mp;           // [Mp][Mp]
  double* T = Linv + (size_t)Mp * Mp;            // Cholesky scratch
  double* phiZ = T + chol_scratch_doubles(1, Mp);  // [M][L]
  double* tmp = phiZ + (size_t)M * L;            // [P][M][S], later t_s [S][M][P]
  double* u = tmp + (size_t)S * M * P;           // [S][M][P]
  double* u2 = u + (size_t)S * M * P;            // [S][M][P]
  const long long MP = (long long)M * P;

  wilson_prior_weights_kernel<<<grid_for((long long)S * L * P), 256, 0, stream>>>(a.coeffs, a.eps_w, a.w,
                                                                               (long long)S * L * P, L, P);
  GPFX_LAUNCH_CHECK();
  {
    KmatArgs k;  // Kuu + jitter I, identity padded to Mp
    k.U = a.Z; k.V = a.Z; k.ls = a.ls; k.var = a.var; k.out = Lu;
    k.K = 1; k.nu = M; k.nv = M; k.D = D;
    k.sOut = (long long)Mp * Mp; k.ldo = Mp; k.rows_out = Mp; k.cols_out = Mp;
    k.kind = a.kind; k.add_jitter = 1; k.jitter = a.jitter;
    GPFX_TRY(launch_kmat_fwd(k, stream));
  }
  GPFX_TRY(potrf_trtri_batched(Lu, Linv, T, 1, Mp, a.info, stream));
  {
    KmatArgs k;  // Phi(Z) [M][L]
    k.U = a.Z; k.V = a.W; k.ls = a.ls; k.var = a.var; k.out = phiZ;
    k.K = 1; k.nu = M; k.nv = L; k.D = D;
    k.sOut = (long long)M * L; k.ldo = L; k.rows_out = M; k.cols_out = L;
    k.mode = 1; k.bias = a.b; k.sBias = L;
    GPFX_TRY(launch_kmat_fwd(k, stream));
  }
  {
    GemmArgs g;  // tmp[p] [M][S] = q_sqrt[p] eps_u[:, p, :]^T
    g.A = a.q_sqrt; g.lda = M; g.sA1 = (long long)M * M;
    g.B = a.eps_u; g.ldb = P * M; g.sB1 = M; g.transB = 1;
    g.C = tmp; g.ldc = S; g.sC1 = (long long)M * S;
    g.M = M; g.N = S; g.K = M;
    g.batch1 = P;
    GPFX_TRY(launch_gemm(g, stream));
  }
  wilson_u_kernel<<<grid_for((long long)S * MP), 256, 0, stream>>>(a.q_mu, tmp, u, S, M, P);
  GPFX_LAUNCH_CHECK();
  double* cur = u;
  if (a.whiten) {
    GemmArgs g;  // u_s <- L_uu u_s
    g.A = Lu; g.lda = Mp; g.sA1 = 0;
    g.B = u; g.ldb = P; g.sB1 = MP;
    g.C = u2; g.ldc = P; g.sC1 = MP;
    g.M = M; g.N = P; g.K = M;
    g.batch1 = S;
    g.tri = TRI_LOWER_A;
    GPFX_TRY(launch_gemm(g, stream));
    cur = u2;
  }
  {
    GemmArgs g;  // diff_s = u_s - Phi(Z) w_s   (in place)
    g.A = phiZ; g.lda = L; g.sA1 = 0;
    g.B = a.w; g.ldb = P; g.sB1 = (long long)L * P;
    g.C = cur; g.ldc = P; g.sC1 = MP;
    g.M = M; g.N = P; g.K = L;
    g.batch1 = S;
    g.alpha = -1.0; g.beta = 1.0;
    GPFX_TRY(launch_gemm(g, stream));
  }
  {
    GemmArgs g;  // t_s = L^-1 diff_s
    g.A = Linv; g.lda = Mp; g.sA1 = 0;
    g.B = cur; g.ldb = P; g.sB1 = MP;
    g.C = tmp; g.ldc = P; g.sC1 = MP;
    g.M = M; g.N = P; g.K = M;
    g.batch1 = S;
    g.tri = TRI_LOWER_A;
    GPFX_TRY(launch_gemm(g, stream));
  }
  {
    GemmArgs g;  // v_s = L^-T t_s = Kuu^-1 diff_s   (compute_A_inv_b, reference math.py:42-62)
    g.A = Linv; g.lda = Mp; g.sA1 = 0; g.transA = 1;
    g.B = tmp; g.ldb = P; g.sB1 = MP;
    g.C = a.v; g.ldc = P; g.sC1 = MP;
    g.M = M; g.N = P; g.K = M;
    g.batch1 = S;
    g.tri = TRI_UPPER_A;
    GPFX_TRY(launch_gemm(g, stream));
  }
  return GPFX_OK;
}

}  // namespace gpfx
