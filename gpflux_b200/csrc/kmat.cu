// Kernel-matrix assembly (Kuu, Kuf) and its backward for stationary kernels.
//
// Forward (reference call site gpflux/layers/gp_layer.py:257-266 ->
// gpflow.covariances.Kuu/Kuf -> gpflow.kernels.Stationary.K): r2 = |u/l|^2 + |v/l|^2 - 2 (u/l).(v/l)
// (GPflow's square_distance expansion) with the cross term on the FP64 tensor pipe (DMMA.8x8x4),
// and the exp / Matern polynomial / variance scale / jitter epilogue applied to the accumulator
// fragments in registers.
//
// Backward (row a14): given dK, R = dK * dk/dr2 is formed on the fly from recomputed distances;
// a row pass (one warp per u-row) reduces du, d lengthscale, d variance and stores R in place, a
// column pass (one thread per v-column) reduces dv.
#include "kmat.h"

#include <math.h>

namespace gpfx {
namespace {

__device__ __forceinline__ double k_of_r2(int kind, double r2, double var) {
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

// dk/d(r2) and k/var from the exact squared distance
__device__ __forceinline__ void dk_of_r2(int kind, double r2, double var, double& k_over_var,
                                         double& dk_dr2) {
  if (kind == GPFX_KERNEL_SE) {
    const double e = exp(-0.5 * r2);
    k_over_var = e;
    dk_dr2 = -0.5 * var * e;
    return;
  }
  const double r2c = fmax(r2, 1e-36);
  const double r = sqrt(r2c);
  // gradient of max(r2, 1e-36) is 0 below the clamp (TF's maximum gradient)
  const double live = (r2 > 1e-36) ? 1.0 : 0.0;
  if (kind == GPFX_KERNEL_MATERN12) {
    const double e = exp(-r);
    k_over_var = e;
    dk_dr2 = live * var * (-e) / (2.0 * r);
    return;
  }
  if (kind == GPFX_KERNEL_MATERN32) {
    const double s3 = 1.7320508075688772;
    const double e = exp(-s3 * r);
    k_over_var = (1.0 + s3 * r) * e;
    // dk/dr = var * (-3 r) e ; dr/dr2 = 1/(2r)
    dk_dr2 = live * var * (-1.5) * e;
    return;
  }
  const double s5 = 2.23606797749979;
  const double e = exp(-s5 * r);
  // GPflow Matern52.K_r: (1 + s5 r + 5/3 r^2) e  with r^2 = square(r) (clamped r)
  k_over_var = (1.0 + s5 * r + (5.0 / 3.0) * r2c) * e;
  // dk/dr = var * (-(5/3) r (1 + s5 r)) e ; * 1/(2r)
  dk_dr2 = live * var * (-(5.0 / 6.0)) * (1.0 + s5 * r) * e;
}

// ---------------------------------------------------------------------------- forward
constexpr int FT = 64;   // tile edge
constexpr int FDC = 32;  // contraction chunk (D is processed FDC dims at a time)

__global__ void __launch_bounds__(128) kmat_fwd_kernel(const KmatArgs a) {
  __shared__ double Us[FT][FDC + 4];
  __shared__ double Vs[FT][FDC + 4];
  __shared__ double un[FT], vn[FT];
  const int k = blockIdx.z;
  const int i0 = blockIdx.y * FT, j0 = blockIdx.x * FT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lc = lane & 3;
  double* out = a.out + (long long)k * a.sOut;

  // tiles entirely inside the identity padding (Kuu only)
  if (i0 >= a.nu || j0 >= a.nv) {
    for (int e = tid; e < FT * FT; e += 128) {
      const int i = i0 + e / FT, j = j0 + e % FT;
      if (i < a.rows_out && j < a.cols_out) out[(long long)i * a.ldo + j] = (i == j) ? 1.0 : 0.0;
    }
    return;
  }

  const double* U = a.U + (long long)k * a.sU;
  const double* V = a.V + (long long)k * a.sV;
  const double* ls = a.ls + (long long)k * a.D;
  const double var = a.var[k];

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double my_norm = 0.0;  // threads 0..63: |u_i|^2 ; 64..127: |v_j|^2

  for (int d0 = 0; d0 < a.D; d0 += FDC) {
    __syncthreads();
    for (int e = tid; e < FT * FDC; e += 128) {
      const int r = e / FDC, d = e % FDC;
      const int gd = d0 + d;
      double u = 0.0, v = 0.0;
      if (gd < a.D) {
        const double l = ls[gd];
        if (i0 + r < a.nu) u = U[(long long)(i0 + r) * a.D + gd] / l;
        if (j0 + r < a.nv) v = V[(long long)(j0 + r) * a.D + gd] / l;
      }
      Us[r][d] = u;
      Vs[r][d] = v;
    }
    __syncthreads();
    {
      const int r = tid & 63;
      const double(*S)[FDC + 4] = (tid < 64) ? Us : Vs;
      double s = 0.0;
      const int dmax = min(FDC, a.D - d0);
      for (int d = 0; d < dmax; ++d) s += S[r][d] * S[r][d];
      my_norm += s;
    }
    const int nk4 = (min(FDC, a.D - d0) + 3) / 4;
    for (int k4 = 0; k4 < nk4; ++k4) {
      const int kk = k4 * 4 + lc;
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Us[wm * 32 + i * 8 + lr][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Vs[wn * 32 + j * 8 + lr][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  if (tid < 64) un[tid] = my_norm;
  else vn[tid - 64] = my_norm;
  __syncthreads();

  const bool vec = ((a.ldo & 1) == 0) && ((((uintptr_t)out) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int li = wm * 32 + i * 8 + lr;
    const int gi = i0 + li;
    if (gi >= a.rows_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lj = wn * 32 + j * 8 + 2 * lc;
      const int gj = j0 + lj;
      double v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gje = gj + e;
        if (gi < a.nu && gje < a.nv) {
          const double r2 = un[li] + vn[lj + e] - 2.0 * acc[i][j][e];
          double kv = k_of_r2(a.kind, r2, var);
          if (a.add_jitter && gi == gje) kv += a.jitter;
          v[e] = kv;
        } else {
          v[e] = (gi == gje) ? 1.0 : 0.0;  // identity padding
        }
      }
      double* dst = out + (long long)gi * a.ldo + gj;
      if (vec && gj + 1 < a.cols_out) {
        *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      } else {
        if (gj < a.cols_out) dst[0] = v[0];
        if (gj + 1 < a.cols_out) dst[1] = v[1];
      }
    }
  }
}

// ---------------------------------------------------------------------------- backward
// Row pass: one warp per (k, i).  Lanes stride over j.  Produces
//   dU[k][i][d]  (+)= sum_j R_ij * 2 (u_id - v_jd) / l_d^2
//   part[k][blk][0..D)  = sum_{i in blk} sum_j -R_ij * 2 (u_id - v_jd)^2 / l_d^3     (d lengthscale)
//   part[k][blk][D]     = sum_{i in blk} sum_j dK_ij * k_ij / var                      (d variance)
// and overwrites dK with R.
template <int DC>
__global__ void __launch_bounds__(256) kmat_bwd_rows_kernel(const KmatBwdArgs a) {
  extern __shared__ double sm[];
  constexpr int TJ = 128;
  const int Dp = a.D + 1;
  double* Vs = sm;                  // [TJ][Dp]
  double* Ui = Vs + TJ * Dp;        // [8][Dp]   this block's 8 rows
  double* red = Ui + 8 * Dp;        // [8][DC + 1]
  const int k = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x * 8 + warp;
  const bool row_ok = i < a.nu;
  const double* U = a.U + (long long)k * a.sU;
  const double* V = a.V + (long long)k * a.sV;
  const double* ls = a.ls + (long long)k * a.D;
  const double var = a.var[k];
  double* G = a.dK + (long long)k * a.sK;

  for (int e = tid; e < 8 * a.D; e += 256) {
    const int r = e / a.D, d = e % a.D;
    const int gi = blockIdx.x * 8 + r;
    Ui[r * Dp + d] = (gi < a.nu) ? U[(long long)gi * a.D + d] : 0.0;
  }

  double dvar_acc = 0.0;
  for (int d0 = 0; d0 < a.D; d0 += DC) {
    double du[DC], dl[DC];
#pragma unroll
    for (int d = 0; d < DC; ++d) du[d] = dl[d] = 0.0;
    for (int j0 = 0; j0 < a.nv; j0 += TJ) {
      __syncthreads();
      for (int e = tid; e < TJ * a.D; e += 256) {
        const int r = e / a.D, d = e % a.D;
        Vs[r * Dp + d] = (j0 + r < a.nv) ? V[(long long)(j0 + r) * a.D + d] : 0.0;
      }
      __syncthreads();
      if (row_ok) {
        for (int jj = lane; jj < TJ && j0 + jj < a.nv; jj += 32) {
          const int j = j0 + jj;
          double R;
          double* gp = G + (long long)i * a.ldk + j;
          if (d0 == 0) {
            double r2 = 0.0;
            for (int d = 0; d < a.D; ++d) {
              const double t = (Ui[warp * Dp + d] - Vs[jj * Dp + d]) / ls[d];
              r2 += t * t;
            }
            double kov, dk;
            dk_of_r2(a.kind, r2, var, kov, dk);
            const double g = *gp;
            dvar_acc += g * kov;
            R = g * dk;
            *gp = R;
          } else {
            R = *gp;  // already converted by this same lane in the first chunk
          }
#pragma unroll
          for (int d = 0; d < DC; ++d) {
            if (d0 + d < a.D) {
              const double l = ls[d0 + d];
              const double t = Ui[warp * Dp + d0 + d] - Vs[jj * Dp + d0 + d];
              const double w = 2.0 * R * t / (l * l);
              du[d] += w;
              dl[d] -= w * t / l;
            }
          }
        }
      }
    }
    // warp reduce, write du, stash dl for the block reduce
#pragma unroll
    for (int d = 0; d < DC; ++d) {
      du[d] = warp_sum(du[d]);
      dl[d] = warp_sum(dl[d]);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int d = 0; d < DC; ++d) {
        if (d0 + d < a.D) {
          if (row_ok) {
            double* dst = a.dU + (long long)k * a.sdU + (long long)i * a.D + d0 + d;
            *dst = a.accumulate_dU ? (*dst + du[d]) : du[d];
          }
          red[warp * (DC + 1) + d] = row_ok ? dl[d] : 0.0;
        }
      }
    }
    __syncthreads();
    if (tid < DC && d0 + tid < a.D) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w * (DC + 1) + tid];
      a.part[((long long)k * gridDim.x + blockIdx.x) * (a.D + 1) + d0 + tid] = s;
    }
  }
  dvar_acc = warp_sum(dvar_acc);
  __syncthreads();
  if (lane == 0) red[warp] = row_ok ? dvar_acc : 0.0;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    a.part[((long long)k * gridDim.x + blockIdx.x) * (a.D + 1) + a.D] = s;
  }
}

// part[k][nblk][D+1] -> dls[k][d] (+)=, dvar[k] (+)= ; one block per k
__global__ void kmat_bwd_finish_kernel(const double* part, int nblk, int D, double* dls,
                                       double* dvar, int accumulate) {
  const int k = blockIdx.x;
  for (int c = threadIdx.x; c <= D; c += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[((long long)k * nblk + b) * (D + 1) + c];
    double* dst = (c < D) ? (dls + (long long)k * D + c) : (dvar + k);
    *dst = accumulate ? (*dst + s) : s;
  }
}

// Column pass: one thread per column j; loops over rows i (and over k when V is shared).
//   dV[j][d] (+)= sum_k sum_i -R_ij * 2 (u_id - v_jd) / l_d^2
template <int DC>
__global__ void __launch_bounds__(128) kmat_bwd_cols_kernel(const KmatBwdArgs a) {
  extern __shared__ double sm[];
  constexpr int TI = 64;
  double* Us = sm;  // [TI][D]
  double* il2 = Us + TI * a.D;  // [D]  2 / l^2
  const int tid = threadIdx.x;
  const int j = blockIdx.x * 128 + tid;
  const bool ok = j < a.nv;
  const bool shared_v = (a.sV == 0 && a.sdV == 0);
  const int k_begin = shared_v ? 0 : blockIdx.y;
  const int k_end = shared_v ? a.K : blockIdx.y + 1;

  for (int d0 = 0; d0 < a.D; d0 += DC) {
    double dv[DC], vj[DC];
#pragma unroll
    for (int d = 0; d < DC; ++d) dv[d] = 0.0;
    for (int k = k_begin; k < k_end; ++k) {
      const double* U = a.U + (long long)k * a.sU;
      const double* V = a.V + (long long)k * a.sV;
      const double* ls = a.ls + (long long)k * a.D;
      const double* G = a.dK + (long long)k * a.sK;
      __syncthreads();
      for (int d = tid; d < a.D; d += 128) il2[d] = 2.0 / (ls[d] * ls[d]);
#pragma unroll
      for (int d = 0; d < DC; ++d)
        vj[d] = (ok && d0 + d < a.D) ? V[(long long)j * a.D + d0 + d] : 0.0;
      double dk[DC];
#pragma unroll
      for (int d = 0; d < DC; ++d) dk[d] = 0.0;
      for (int i0 = 0; i0 < a.nu; i0 += TI) {
        __syncthreads();
        for (int e = tid; e < TI * a.D; e += 128) {
          const int r = e / a.D, d = e % a.D;
          Us[e] = (i0 + r < a.nu) ? U[(long long)(i0 + r) * a.D + d] : 0.0;
        }
        __syncthreads();
        if (ok) {
          const int imax = min(TI, a.nu - i0);
          for (int ii = 0; ii < imax; ++ii) {
            const double R = G[(long long)(i0 + ii) * a.ldk + j];
#pragma unroll
            for (int d = 0; d < DC; ++d)
              if (d0 + d < a.D) dk[d] -= R * (Us[ii * a.D + d0 + d] - vj[d]);
          }
        }
      }
#pragma unroll
      for (int d = 0; d < DC; ++d)
        if (d0 + d < a.D) dv[d] += dk[d] * il2[d0 + d];
    }
    if (ok) {
      const long long kk = shared_v ? 0 : blockIdx.y;
#pragma unroll
      for (int d = 0; d < DC; ++d) {
        if (d0 + d < a.D) {
          double* dst = a.dV + kk * a.sdV + (long long)j * a.D + d0 + d;
          *dst = a.accumulate_dV ? (*dst + dv[d]) : dv[d];
        }
      }
    }
  }
}

}  // namespace

int launch_kmat_fwd(const KmatArgs& a, cudaStream_t stream) {
  if (a.K <= 0 || a.rows_out <= 0 || a.cols_out <= 0) return GPFX_OK;
  dim3 grid(ceil_div(a.cols_out, FT), ceil_div(a.rows_out, FT), a.K);
  kmat_fwd_kernel<<<grid, 128, 0, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

size_t kmat_bwd_part_doubles(int K, int nu, int D) {
  return (size_t)K * ceil_div(nu, 8) * (D + 1);
}

int launch_kmat_bwd(const KmatBwdArgs& a, cudaStream_t stream) {
  if (a.K <= 0 || a.nu <= 0 || a.nv <= 0) return GPFX_OK;
  const int nblk = ceil_div(a.nu, 8);
  {
    dim3 grid(nblk, a.K);
    if (a.D <= 8) {
      const size_t smem = ((size_t)(128 + 8) * (a.D + 1) + 8 * 9) * sizeof(double);
      kmat_bwd_rows_kernel<8><<<grid, 256, smem, stream>>>(a);
    } else {
      const size_t smem = ((size_t)(128 + 8) * (a.D + 1) + 8 * 33) * sizeof(double);
      if (smem > 48 * 1024) {
        static bool set = false;
        if (!set) {
          GPFX_CUDA_CHECK(cudaFuncSetAttribute(kmat_bwd_rows_kernel<32>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
          set = true;
        }
        if (smem > 200 * 1024) {
          set_last_error("kmat_bwd: input dim %d too large", a.D);
          return GPFX_E_UNSUPPORTED;
        }
      }
      kmat_bwd_rows_kernel<32><<<grid, 256, smem, stream>>>(a);
    }
    GPFX_LAUNCH_CHECK();
    kmat_bwd_finish_kernel<<<a.K, 64, 0, stream>>>(a.part, nblk, a.D, a.dls, a.dvar,
                                                    a.accumulate_hyp);
    GPFX_LAUNCH_CHECK();
  }
  if (a.dV != nullptr) {
    const bool shared_v = (a.sV == 0 && a.sdV == 0);
    dim3 grid(ceil_div(a.nv, 128), shared_v ? 1 : a.K);
    const size_t smem = ((size_t)64 * a.D + a.D) * sizeof(double);
    if (smem > 48 * 1024) {
      set_last_error("kmat_bwd: input dim %d too large", a.D);
      return GPFX_E_UNSUPPORTED;
    }
    if (a.D <= 8) kmat_bwd_cols_kernel<8><<<grid, 128, smem, stream>>>(a);
    else kmat_bwd_cols_kernel<32><<<grid, 128, smem, stream>>>(a);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace gpfx
