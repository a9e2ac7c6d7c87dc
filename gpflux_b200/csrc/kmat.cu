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

#include <algorithm>

#include "gemm_f64.h"

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

__global__ void __launch_bounds__(128, 4) kmat_fwd_kernel(const KmatArgs a) {
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

  __shared__ double invl[FDC], lval[FDC];
  const bool scale_v = (a.mode == 0 || a.mode == 3);
  // x / l without a division per staged element: q = x * (1/l), then one Markstein correction
  // q' = q + (x - q l) (1/l) with the residual exact in an FMA - the correctly rounded quotient (the
  // reciprocal is itself correctly rounded), i.e. bit-identical to x / l at 3 FP64 operations
  auto div_l = [&](double x, int d) {
    const double r = invl[d];
    const double q = x * r;
    return fma(fma(-q, lval[d], x), r, q);
  };
  for (int d0 = 0; d0 < a.D; d0 += FDC) {
    __syncthreads();
    const int Dc4 = ((min(FDC, a.D - d0) + 3) / 4) * 4;  // staged columns (zero-padded to a k4 step)
    // one reciprocal per input dimension and CTA (a division per staged element costs as much FP64
    // issue as the exponential of the epilogue)
    if (tid < FDC) {
      const double l = (d0 + tid < a.D) ? ls[d0 + tid] : 1.0;
      lval[tid] = l;
      invl[tid] = 1.0 / l;
    }
    __syncthreads();
#pragma unroll 4
    for (int e = tid; e < FT * Dc4; e += 128) {
      const int r = e / Dc4, d = e - r * Dc4;
      const int gd = d0 + d;
      double u = 0.0, v = 0.0;
      if (gd < a.D) {
        if (i0 + r < a.nu) u = div_l(U[(long long)(i0 + r) * a.D + gd], d);
        if (j0 + r < a.nv) {
          v = V[(long long)(j0 + r) * a.D + gd];
          if (scale_v) v = div_l(v, d);
        }
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

  if (a.mode == 4 || a.mode == 5) {
    // backward of the random-Fourier-feature epilogue: dLoss/dz on the accumulator fragments
    const double c = sqrt(2.0 * var / (double)(a.mode == 5 ? 2 * a.nv : a.nv));
    const double* bias = (a.mode == 4) ? a.bias + (long long)k * a.sBias : nullptr;
    const double* dPhi = a.dK_in + (long long)k * a.sdK;
    double dv = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = i0 + wm * 32 + i * 8 + lr;
      if (gi >= a.nu) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gj = j0 + wn * 32 + j * 8 + 2 * lc + e;
          if (gj >= a.nv) continue;
          const double* grow = dPhi + (long long)gi * a.ld_dk;
          double sn, cs;
          sincos(acc[i][j][e] + (bias ? bias[gj] : 0.0), &sn, &cs);
          if (a.mode == 4) {
            const double g = grow[gj];
            out[(long long)gi * a.ldo + gj] = -g * c * sn;
            dv += g * c * cs;
          } else {
            const double gs = grow[gj], gc = grow[a.nv + gj];
            out[(long long)gi * a.ldo + gj] = c * (gs * cs - gc * sn);
            dv += c * (gs * sn + gc * cs);
          }
        }
      }
    }
    __syncthreads();
    const double t = block_sum(dv, un);
    if (tid == 0) a.part_out[((long long)k * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t / (2.0 * var);
    return;
  }
  if (a.mode == 1 || a.mode == 2) {
    // random-Fourier-feature epilogue on the accumulator fragments
    const double c = sqrt(2.0 * var / (double)(a.mode == 2 ? 2 * a.nv : a.nv));
    const double* bias = (a.mode == 1) ? a.bias + (long long)k * a.sBias : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = i0 + wm * 32 + i * 8 + lr;
      if (gi >= a.nu) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gj = j0 + wn * 32 + j * 8 + 2 * lc + e;
          if (gj >= a.nv) continue;
          double* row = out + (long long)gi * a.ldo;
          if (a.mode == 1) {
            row[gj] = c * cos(acc[i][j][e] + bias[gj]);
          } else {
            double sn, cs;
            sincos(acc[i][j][e], &sn, &cs);
            row[gj] = c * sn;
            row[a.nv + gj] = c * cs;
          }
        }
      }
    }
    return;
  }

  const bool vec = ((a.ldo & 1) == 0) && ((((uintptr_t)out) & 15) == 0);
  const double* dKin = (a.mode == 3) ? a.dK_in + (long long)k * a.sOut : nullptr;
  double dvar_local = 0.0;
  // SquaredExponential fast path: the thread's 32 exponentials are evaluated in one branch-free
  // block (independent dependency chains the compiler can interleave) instead of one by one
  // behind the kernel-family switch; acc is reused to hold exp(-r2/2).
  const bool se_fast = (a.kind == GPFX_KERNEL_SE);
  if (se_fast) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double uni = un[wm * 32 + i * 8 + lr];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int lj = wn * 32 + j * 8 + 2 * lc;
#pragma unroll
        for (int e = 0; e < 2; ++e)
          acc[i][j][e] = exp(-0.5 * (uni + vn[lj + e] - 2.0 * acc[i][j][e]));
      }
    }
  }
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
          if (se_fast) {
            const double ee = acc[i][j][e];
            if (a.mode == 3) {
              const double g = dKin[(long long)gi * a.ldo + gje];
              dvar_local += g * ee;
              v[e] = g * (-0.5 * var * ee);
            } else {
              double kv = var * ee;
              if (a.add_jitter && gi == gje) kv += a.jitter;
              v[e] = kv;
            }
            continue;
          }
          const double r2 = un[li] + vn[lj + e] - 2.0 * acc[i][j][e];
          if (a.mode == 3) {
            // backward: R = dK * dk/dr2 (written over dK), and sum dK * k / var for d variance
            double kov, dk;
            dk_of_r2(a.kind, r2, var, kov, dk);
            const double g = dKin[(long long)gi * a.ldo + gje];
            dvar_local += g * kov;
            v[e] = g * dk;
          } else {
            double kv = k_of_r2(a.kind, r2, var);
            if (a.add_jitter && gi == gje) kv += a.jitter;
            v[e] = kv;
          }
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
  if (a.mode == 3) {
    __syncthreads();
    const double t = block_sum(dvar_local, un);  // un[64] is free now
    if (tid == 0)
      a.part_out[((long long)k * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
  }
}

// out[j][0..D) = src[j][:], out[j][D] = 1, zero padding up to Dp columns
__global__ void augment_ones_kernel(const double* __restrict__ src, double* __restrict__ out,
                                    long long rows, int D, int Dp) {
  const long long total = rows * Dp;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / Dp;
    const int d = (int)(e - j * Dp);
    out[e] = (d < D) ? src[j * D + d] : (d == D ? 1.0 : 0.0);
  }
}

// GEMM-form backward, final assembly.  RV[k][i][0..D) = sum_j R_ij v_j, RV[k][i][D] = sum_j R_ij;
// RtU likewise with the roles swapped.
//   dU[k][i][d] (+)= 2/l^2 (u_id rs_i - RV_id)
__global__ void kmat_bwdg_dU_kernel(const KmatBwdArgs a, const double* __restrict__ RV, int Dp) {
  const int k = blockIdx.y;
  const long long nUD = (long long)a.nu * a.D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nUD) return;
  const long long i = t / a.D;
  const int d = (int)(t - i * a.D);
  const double l = a.ls[(long long)k * a.D + d];
  const double* rv = RV + ((long long)k * a.nu + i) * Dp;
  const double u = a.U[(long long)k * a.sU + t];
  const double g = (2.0 / (l * l)) * (u * rv[a.D] - rv[d]);
  double* dst = a.dU + (long long)k * a.sdU + t;
  *dst = a.accumulate_dU ? (*dst + g) : g;
}

//   dV[j][d] (+)= sum_k -2/l_k^2 (RtU_k[j][d] - v_jd cs_kj)      (shared V: sum over k)
__global__ void kmat_bwdg_dV_kernel(const KmatBwdArgs a, const double* __restrict__ RtU, int Dp,
                                    int shared_v, int centered = 0) {
  const long long nVD = (long long)a.nv * a.D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nVD) return;
  const long long j = t / a.D;
  const int d = (int)(t - j * a.D);
  const int k0 = shared_v ? 0 : blockIdx.y, k1 = shared_v ? a.K : blockIdx.y + 1;
  double s = 0.0;
  for (int k = k0; k < k1; ++k) {
    const double l = a.ls[(long long)k * a.D + d];
    const double* r = RtU + ((long long)k * a.nv + j) * Dp;
    // centered: RtU was formed against U - U_k[0] (fused pass), so v is taken relative to it too
    const double v = a.V[(long long)k * a.sV + t] - (centered ? a.U[(long long)k * a.sU + d] : 0.0);
    s -= (2.0 / (l * l)) * (r[d] - v * r[a.D]);
  }
  double* dst = a.dV + (shared_v ? 0 : (long long)blockIdx.y * a.sdV) + t;
  *dst = a.accumulate_dV ? (*dst + s) : s;
}

//   dls[k][d] (+)= -2/l^3 [ sum_i u_id^2 rs_i + sum_j v_jd^2 cs_j - 2 sum_i u_id RV_id ]
//   dvar[k]   (+)= sum of the R-kernel's block partials             (one block per (d, k); d == D -> dvar)
__global__ void __launch_bounds__(256) kmat_bwdg_hyp_kernel(const KmatBwdArgs a,
                                                            const double* __restrict__ RV,
                                                            const double* __restrict__ RtU, int Dp,
                                                            const double* __restrict__ part,
                                                            int nparts) {
  __shared__ double red[32];
  const int d = blockIdx.x, k = blockIdx.y;
  double s = 0.0;
  if (d == a.D) {
    for (int e = threadIdx.x; e < nparts; e += 256) s += part[(long long)k * nparts + e];
    s = block_sum(s, red);
    if (threadIdx.x == 0) a.dvar[k] = a.accumulate_hyp ? a.dvar[k] + s : s;
    return;
  }
  const double* U = a.U + (long long)k * a.sU;
  const double* V = a.V + (long long)k * a.sV;
  for (int i = threadIdx.x; i < a.nu; i += 256) {
    const double u = U[(long long)i * a.D + d];
    const double* rv = RV + ((long long)k * a.nu + i) * Dp;
    s += u * (u * rv[a.D] - 2.0 * rv[d]);
  }
  for (int j = threadIdx.x; j < a.nv; j += 256) {
    const double v = V[(long long)j * a.D + d];
    s += v * v * RtU[((long long)k * a.nv + j) * Dp + a.D];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double l = a.ls[(long long)k * a.D + d];
    const double g = -2.0 / (l * l * l) * s;
    double* dst = a.dls + (long long)k * a.D + d;
    *dst = a.accumulate_hyp ? (*dst + g) : g;
  }
}

// Row pass: grid (nchunks, ceil(nu/8), K); one warp per row i, lanes stride over the columns of this
// block's column chunk.  Produces
//   partU[k][chunk][i][d]     = sum_{j in chunk} R_ij * 2 (u_id - v_jd) / l_d^2
//   part[k][rb][chunk][0..D)  = sum_{i in rb} sum_{j in chunk} -R_ij * 2 (u_id - v_jd)^2 / l_d^3
//   part[k][rb][chunk][D]     = sum dK_ij * k_ij / var
// and overwrites dK with R = dK * dk/dr2.  When the kernel is SE and the forward kernel matrix is
// available (Kmat), R = -dK*K/2 needs neither the distance nor the exp.
template <int DC>
__global__ void __launch_bounds__(256) kmat_bwd_rows_kernel(const KmatBwdArgs a, int nchunks,
                                                            int chunk_cols, int stile) {
  extern __shared__ double sm[];
  constexpr int TJ = 128;
  const int Dp = a.D + 1;
  double* Vs = sm;                  // [stile][Dp]
  double* Ui = Vs + (size_t)stile * Dp;  // [8][Dp]   this block's 8 rows
  double* il = Ui + 8 * Dp;         // [D]  1 / l
  double* red = il + a.D;           // [8][DC + 1]
  const int k = blockIdx.z, rb = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = rb * 8 + warp;
  const bool row_ok = i < a.nu;
  const int c0 = chunk * chunk_cols, c1 = min(a.nv, c0 + chunk_cols);
  const double* U = a.U + (long long)k * a.sU;
  const double* V = a.V + (long long)k * a.sV;
  const double* ls = a.ls + (long long)k * a.D;
  const double var = a.var[k];
  double* G = a.dK + (long long)k * a.sK;
  const double* Kf = a.Kmat ? a.Kmat + (long long)k * a.sKmat : nullptr;
  const bool use_k = (Kf != nullptr) && (a.kind == GPFX_KERNEL_SE);

  for (int e = tid; e < 8 * a.D; e += 256) {
    const int r = e / a.D, d = e % a.D;
    const int gi = rb * 8 + r;
    Ui[r * Dp + d] = (gi < a.nu) ? U[(long long)gi * a.D + d] : 0.0;
  }
  for (int d = tid; d < a.D; d += 256) il[d] = 1.0 / ls[d];

  double dvar_acc = 0.0;
  double* part = a.part + (((long long)k * gridDim.y + rb) * nchunks + chunk) * (a.D + 1);
  for (int d0 = 0; d0 < a.D; d0 += DC) {
    double du[DC], dl[DC];
#pragma unroll
    for (int d = 0; d < DC; ++d) du[d] = dl[d] = 0.0;
    // V is staged in super-tiles of `stile` columns (one barrier pair each); inside a super-tile
    // the 128-column sub-tiles run barrier-free with the next sub-tile's dK / K loads already in
    // flight (software pipelining: the stores of R to dK would otherwise fence the loads).
    for (int s0 = c0; s0 < c1; s0 += stile) {
      const int s1 = min(c1, s0 + stile);
      __syncthreads();
      for (int e = tid; e < (s1 - s0) * a.D; e += 256) {
        const int r = e / a.D, d = e % a.D;
        Vs[r * Dp + d] = V[(long long)(s0 + r) * a.D + d];
      }
      double gq[TJ / 32], kq[TJ / 32];
#pragma unroll
      for (int q = 0; q < TJ / 32; ++q) {
        const int j = s0 + lane + 32 * q;
        const bool ok = row_ok && j < s1;
        gq[q] = ok ? G[(long long)i * a.ldk + j] : 0.0;
        kq[q] = (ok && use_k && d0 == 0) ? Kf[(long long)i * a.ldkmat + j] : 0.0;
      }
      __syncthreads();
      for (int j0 = s0; j0 < s1; j0 += TJ) {
        double gn[TJ / 32], kn[TJ / 32];
#pragma unroll
        for (int q = 0; q < TJ / 32; ++q) {
          const int j = j0 + TJ + lane + 32 * q;
          const bool ok = row_ok && j < s1;
          gn[q] = ok ? G[(long long)i * a.ldk + j] : 0.0;
          kn[q] = (ok && use_k && d0 == 0) ? Kf[(long long)i * a.ldkmat + j] : 0.0;
        }
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < TJ / 32; ++q) {
            const int j = j0 + lane + 32 * q;
            if (j >= s1) continue;
            const int jj = j - s0;
            double R;
            if (d0 == 0) {
              const double g = gq[q];
              if (use_k) {
                const double gk = g * kq[q];
                dvar_acc += gk;
                R = -0.5 * gk;
              } else {
                double r2 = 0.0;
                for (int d = 0; d < a.D; ++d) {
                  const double t = (Ui[warp * Dp + d] - Vs[jj * Dp + d]) * il[d];
                  r2 += t * t;
                }
                double kov, dk;
                dk_of_r2(a.kind, r2, var, kov, dk);
                dvar_acc += g * kov;
                R = g * dk;
              }
              G[(long long)i * a.ldk + j] = R;
            } else {
              R = gq[q];  // already converted by this same lane in the first chunk
            }
            const double R2 = 2.0 * R;
#pragma unroll
            for (int d = 0; d < DC; ++d) {
              if (d0 + d < a.D) {
                const double t = Ui[warp * Dp + d0 + d] - Vs[jj * Dp + d0 + d];
                const double w = R2 * t;
                du[d] += w;
                dl[d] += w * t;
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < TJ / 32; ++q) {
          gq[q] = gn[q];
          kq[q] = kn[q];
        }
      }
    }
    // warp reduce; apply the per-dimension lengthscale factors once per row
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
          const double inv = il[d0 + d];
          if (row_ok)
            a.partU[(((long long)k * nchunks + chunk) * a.nu + i) * a.D + d0 + d] = du[d] * inv * inv;
          red[warp * (DC + 1) + d] = row_ok ? -dl[d] * inv * inv * inv : 0.0;
        }
      }
    }
    __syncthreads();
    if (tid < DC && d0 + tid < a.D) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w * (DC + 1) + tid];
      part[d0 + tid] = s;
    }
  }
  dvar_acc = warp_sum(dvar_acc);
  __syncthreads();
  if (lane == 0) red[warp] = row_ok ? dvar_acc : 0.0;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    part[a.D] = use_k ? s / var : s;
  }
}

// dU[k][i][d] (+)= sum_chunk partU ; dls[k][d] (+)= sum part ; dvar[k] (+)= sum part
__global__ void kmat_bwd_rows_finish_kernel(const KmatBwdArgs a, int nchunks, int nrb) {
  const int k = blockIdx.y;
  const long long nUD = (long long)a.nu * a.D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < nUD) {
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += a.partU[((long long)k * nchunks + c) * nUD + t];
    double* dst = a.dU + (long long)k * a.sdU + t;
    *dst = a.accumulate_dU ? (*dst + s) : s;
  } else if (t < nUD + a.D + 1) {
    const int c = (int)(t - nUD);
    double s = 0.0;
    const long long np = (long long)nrb * nchunks;
    for (long long b = 0; b < np; ++b) s += a.part[((long long)k * np + b) * (a.D + 1) + c];
    double* dst = (c < a.D) ? (a.dls + (long long)k * a.D + c) : (a.dvar + k);
    *dst = a.accumulate_hyp ? (*dst + s) : s;
  }
}

// Column pass: grid (ceil(nv/128), RC, K); one thread per column j, looping over the rows of row
// chunk rc.   partV[k][rc][j][d] = sum_{i in rc} -R_ij * 2 (u_id - v_jd) / l_d^2
template <int DC>
__global__ void __launch_bounds__(128) kmat_bwd_cols_kernel(const KmatBwdArgs a, int RC,
                                                            int chunk_rows) {
  extern __shared__ double sm[];
  constexpr int TI = 64;
  double* Us = sm;  // [TI][D]
  const int tid = threadIdx.x;
  const int j = blockIdx.x * 128 + tid;
  const bool ok = j < a.nv;
  const int rc = blockIdx.y, k = blockIdx.z;
  const int r0 = rc * chunk_rows, r1 = min(a.nu, r0 + chunk_rows);
  const double* U = a.U + (long long)k * a.sU;
  const double* V = a.V + (long long)k * a.sV;
  const double* ls = a.ls + (long long)k * a.D;
  const double* G = a.dK + (long long)k * a.sK;
  double* out = a.partV + (((long long)k * RC + rc) * a.nv) * a.D;

  for (int d0 = 0; d0 < a.D; d0 += DC) {
    double dk[DC], vj[DC];
#pragma unroll
    for (int d = 0; d < DC; ++d) {
      dk[d] = 0.0;
      vj[d] = (ok && d0 + d < a.D) ? V[(long long)j * a.D + d0 + d] : 0.0;
    }
    for (int i0 = r0; i0 < r1; i0 += TI) {
      __syncthreads();
      for (int e = tid; e < TI * a.D; e += 128) {
        const int r = e / a.D, d = e % a.D;
        Us[e] = (i0 + r < r1) ? U[(long long)(i0 + r) * a.D + d] : 0.0;
      }
      __syncthreads();
      if (ok) {
        const int imax = min(TI, r1 - i0);
        for (int ii = 0; ii < imax; ii += 8) {
          double Rv[8];  // eight independent loads in flight per thread
#pragma unroll
          for (int u = 0; u < 8; ++u)
            Rv[u] = (ii + u < imax) ? G[(long long)(i0 + ii + u) * a.ldk + j] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int d = 0; d < DC; ++d)
              if (d0 + d < a.D) dk[d] -= Rv[u] * (Us[(ii + u) * a.D + d0 + d] - vj[d]);
          }
        }
      }
    }
    if (ok) {
#pragma unroll
      for (int d = 0; d < DC; ++d)
        if (d0 + d < a.D) {
          const double l = ls[d0 + d];
          out[(long long)j * a.D + d0 + d] = dk[d] * (2.0 / (l * l));
        }
    }
  }
}

// dV (+)= sum over partials: shared V: sum over (k, rc) -> [nv][D]; else per k: sum over rc
__global__ void kmat_bwd_cols_finish_kernel(const KmatBwdArgs a, int RC, int shared_v) {
  const long long nVD = (long long)a.nv * a.D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nVD) return;
  if (shared_v) {
    double s = 0.0;
    for (int p = 0; p < a.K * RC; ++p) s += a.partV[(long long)p * nVD + t];
    a.dV[t] = a.accumulate_dV ? (a.dV[t] + s) : s;
  } else {
    const int k = blockIdx.y;
    double s = 0.0;
    for (int p = 0; p < RC; ++p) s += a.partV[((long long)k * RC + p) * nVD + t];
    double* dst = a.dV + (long long)k * a.sdV + t;
    *dst = a.accumulate_dV ? (*dst + s) : s;
  }
}

struct BwdSplit {
  int nrb, nchunks, chunk_cols, RC, chunk_rows;
};

BwdSplit bwd_split(int K, int nu, int nv) {
  BwdSplit s;
  s.nrb = ceil_div(nu, 8);
  const int max_chunks = ceil_div(nv, 128);
  long long want = ceil_div_ll(160, (long long)s.nrb * K);
  s.nchunks = (int)(want < 1 ? 1 : (want > max_chunks ? max_chunks : want));
  s.chunk_cols = ceil_div(ceil_div(nv, s.nchunks), 128) * 128;
  s.nchunks = ceil_div(nv, s.chunk_cols);
  const int cb = ceil_div(nv, 128);
  const int max_rc = ceil_div(nu, 64);
  want = ceil_div_ll(300, (long long)cb * K);
  s.RC = (int)(want < 1 ? 1 : (want > max_rc ? max_rc : want));
  s.chunk_rows = ceil_div(ceil_div(nu, s.RC), 64) * 64;
  s.RC = ceil_div(nu, s.chunk_rows);
  return s;
}

}  // namespace

int launch_kmat_fwd(const KmatArgs& a, cudaStream_t stream) {
  if (a.K <= 0 || a.rows_out <= 0 || a.cols_out <= 0) return GPFX_OK;
  dim3 grid(ceil_div(a.mode == 0 ? a.cols_out : a.nv, FT), ceil_div(a.rows_out, FT), a.K);
  kmat_fwd_kernel<<<grid, 128, 0, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

namespace {

// GEMM-form plan: column width of the augmented operands and the split-K factors
struct BwdGemmPlan {
  int Dp, tiles, splits1, splits2;
  size_t o_Uaug, o_Vaug, o_RV, o_RtU, o_part, o_split, total;
};

BwdGemmPlan bwd_gemm_plan(int K, int nu, int nv, int D, bool shared_v) {
  BwdGemmPlan p;
  p.Dp = ((D + 1) + 1) / 2 * 2;  // even leading dimension -> 16-byte cp.async path
  p.tiles = ceil_div(nu, FT) * ceil_div(nv, FT);
  auto splits = [](long long tiles, int kdim) {
    long long s = ceil_div_ll(296, tiles > 0 ? tiles : 1);
    const long long smax = kdim / 512 > 1 ? kdim / 512 : 1;
    return (int)(s < 1 ? 1 : (s > smax ? smax : s));
  };
  p.splits1 = splits((long long)ceil_div(nu, 64) * K, nv);
  p.splits2 = splits((long long)ceil_div(nv, 64) * K, nu);
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += (n + 31) / 32 * 32;
    return o;
  };
  p.o_Uaug = take((size_t)K * nu * p.Dp);
  p.o_Vaug = take((size_t)(shared_v ? 1 : K) * nv * p.Dp);
  p.o_RV = take((size_t)K * nu * p.Dp);
  p.o_RtU = take((size_t)K * nv * p.Dp);
  p.o_part = take((size_t)K * p.tiles);
  const size_t s1 = p.splits1 > 1 ? (size_t)p.splits1 * K * nu * p.Dp : 0;
  const size_t s2 = p.splits2 > 1 ? (size_t)p.splits2 * K * nv * p.Dp : 0;
  p.o_split = take(s1 > s2 ? s1 : s2);
  p.total = off;
  return p;
}


// ---------------------------------------------------------------------------- fused backward
// Single pass over dK (and the saved forward matrix) for small input dims (D + 1 <= 8 NF, NF <= 2):
// one CTA forms a 64 x 128 tile of R = dK * dk/dr2 in shared memory and contracts it on the FP64
// tensor pipe against BOTH augmented operands while it is there:
//   RV  += R   [V, 1]    (rows: sum_j R_ij v_j, sum_j R_ij)       m = 64,  n = 8 NF, k = 128
//   RtU += R^T [U, 1]    (columns: sum_i R_ij u_i, sum_i R_ij)    m = 128, n = 8 NF, k = 64
// R never reaches HBM; dK and K are read exactly once.  Partials per column chunk / row block are
// summed in a fixed order by kmat_bwdf_assemble_kernel (bit-reproducible), which also forms dU, dV
// and the per-block lengthscale sums; kmat_bwdf_hyp_kernel finishes dls / dvar.
template <int NF>
__global__ void __launch_bounds__(256, (NF <= 2) ? 2 : 1) kmat_bwd_fused_kernel(const KmatBwdArgs a,
                                                                double* __restrict__ RVp,
                                                                double* __restrict__ RtUp,
                                                                double* __restrict__ part) {
  constexpr int TI = 64, TJ = 128, W = 8 * NF, LDA = W + 4, LDR = TJ + 4;
  extern __shared__ double sm[];
  double* Rs = sm;              // [TI][LDR]
  double* Ua = Rs + TI * LDR;   // [TI][LDA]  rows of [U, 1, 0...]
  double* Va = Ua + TI * LDA;   // [TJ][LDA]  rows of [V, 1, 0...]
  double* il = Va + TJ * LDA;   // [W]  1 / lengthscale
  double* red = il + W;         // [32]
  double* Us = red + 32;        // [TI][W]  rows of (U - c) / l  (distance path only)
  const int chunk = blockIdx.x, rb = blockIdx.y, k = blockIdx.z;
  const int nchunks = gridDim.x, nrb = gridDim.y;
  const int i0 = rb * TI, j0 = chunk * TJ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const double* U = a.U + (long long)k * a.sU;
  const double* V = a.V + (long long)k * a.sV;
  const double* ls = a.ls + (long long)k * a.D;
  const double var = a.var[k];
  const double* G = a.dK + (long long)k * a.sK;
  const double* Kf = a.Kmat ? a.Kmat + (long long)k * a.sKmat : nullptr;
  const bool use_k = (Kf != nullptr) && (a.kind == GPFX_KERNEL_SE);

  // Both operands are shifted by the same point c = U[0] (stationary kernels only see differences):
  // the contractions R [V - c, 1] and R^T [U - c, 1] then cancel like sum_j R_ij (u_i - v_j) itself,
  // whatever the offset of the inputs from the origin - which lets this form serve Kuu (u = v) too.
  for (int e = tid; e < TI * W; e += 256) {
    const int r = e / W, d = e - r * W;
    const int gi = i0 + r;
    double v = 0.0;
    if (gi < a.nu) v = (d < a.D) ? U[(long long)gi * a.D + d] - U[d] : (d == a.D ? 1.0 : 0.0);
    Ua[r * LDA + d] = v;
  }
  for (int e = tid; e < TJ * W; e += 256) {
    const int r = e / W, d = e - r * W;
    const int gj = j0 + r;
    double v = 0.0;
    if (gj < a.nv) v = (d < a.D) ? V[(long long)gj * a.D + d] - U[d] : (d == a.D ? 1.0 : 0.0);
    Va[r * LDA + d] = v;
  }
  if (tid < W) il[tid] = (tid < a.D) ? 1.0 / ls[tid] : 0.0;
  __syncthreads();
  if (!use_k) {
    // distance path: rows of U scaled once per CTA (one broadcast load per term of r2 below)
    for (int e = tid; e < TI * W; e += 256) {
      const int r = e / W, d = e - r * W;
      Us[e] = Ua[r * LDA + d] * il[d];  // il = 0 from the ones column on
    }
    __syncthreads();
  }

  // R tile: thread t owns column (t & 127) of rows (t >> 7) + 2 q  -> coalesced 1 KB row segments
  double dvar_acc = 0.0;
  {
    const int c = tid & (TJ - 1), rbase = tid >> 7;
    const int gj = j0 + c;
    const bool col_ok = gj < a.nv;
    // distance path: this thread's column of V, scaled, stays in registers for its 32 rows
    double vs[W];
    if (!use_k) {
#pragma unroll
      for (int d = 0; d < W; ++d) vs[d] = Va[c * LDA + d] * il[d];
    }
#pragma unroll 1
    for (int q0 = 0; q0 < TI / 2; q0 += 8) {
      double g[8], kv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int gi = i0 + rbase + 2 * (q0 + q);
        const bool ok = col_ok && gi < a.nu;
        g[q] = ok ? G[(long long)gi * a.ldk + gj] : 0.0;
        kv[q] = (ok && use_k) ? Kf[(long long)gi * a.ldkmat + gj] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = rbase + 2 * (q0 + q);
        double R;
        if (use_k) {
          const double gk = g[q] * kv[q];
          dvar_acc += gk;
          R = -0.5 * gk;
        } else {
          double r2 = 0.0;
          const double2* us2 = reinterpret_cast<const double2*>(Us + r * W);
#pragma unroll
          for (int d = 0; d < W; d += 2) {  // (entries from the ones column on are zero on both sides)
            const double2 u2 = us2[d >> 1];
            const double t0 = u2.x - vs[d], t1 = u2.y - vs[d + 1];
            r2 = fma(t0, t0, r2);
            r2 = fma(t1, t1, r2);
          }
          double kov, dk;
          dk_of_r2(a.kind, r2, var, kov, dk);
          dvar_acc += g[q] * kov;
          R = g[q] * dk;  // g = 0 outside the matrix
        }
        Rs[r * LDR + c] = R;
      }
    }
  }
  __syncthreads();

  // RV: warp w owns rows 8w .. 8w+7, contracts over the 128 columns
  {
    double acc[NF][2];
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) acc[nf][0] = acc[nf][1] = 0.0;
    const double* ar = Rs + (8 * warp + lr) * LDR + lc;
    const double* br = Va + lc * LDA + lr;
#pragma unroll 8
    for (int k4 = 0; k4 < TJ / 4; ++k4) {
      const double af = ar[4 * k4];
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) dmma884(acc[nf][0], acc[nf][1], af, br[4 * k4 * LDA + 8 * nf]);
    }
    const int gi = i0 + 8 * warp + lr;
    if (gi < a.nu) {
      double* dst = RVp + (((long long)k * nchunks + chunk) * a.nu + gi) * W + 2 * lc;
#pragma unroll
      for (int nf = 0; nf < NF; ++nf)
        *reinterpret_cast<double2*>(dst + 8 * nf) = make_double2(acc[nf][0], acc[nf][1]);
    }
  }
  // RtU: warp w owns columns 16w .. 16w+15, contracts over the 64 rows
  {
    double acc[2][NF][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    const double* ar = Rs + lc * LDR + 16 * warp + lr;
    const double* br = Ua + lc * LDA + lr;
#pragma unroll 4
    for (int k4 = 0; k4 < TI / 4; ++k4) {
      double bf[NF];
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) bf[nf] = br[4 * k4 * LDA + 8 * nf];
#pragma unroll
      for (int mf = 0; mf < 2; ++mf) {
        const double af = ar[4 * k4 * LDR + 8 * mf];
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], af, bf[nf]);
      }
    }
#pragma unroll
    for (int mf = 0; mf < 2; ++mf) {
      const int gj = j0 + 16 * warp + 8 * mf + lr;
      if (gj < a.nv) {
        double* dst = RtUp + (((long long)k * nrb + rb) * a.nv + gj) * W + 2 * lc;
#pragma unroll
        for (int nf = 0; nf < NF; ++nf)
          *reinterpret_cast<double2*>(dst + 8 * nf) = make_double2(acc[mf][nf][0], acc[mf][nf][1]);
      }
    }
  }
  const double t = block_sum(dvar_acc, red);
  if (tid == 0) part[((long long)k * nrb + rb) * nchunks + chunk] = use_k ? t / var : t;
}

// Assembly after the fused pass, one warp per row of U (blocks [0, nbU)) or column of V (blocks
// [nbU, nbU + nbV)): sums the per-chunk / per-row-block partials in a fixed order (lanes c < W hold
// the W augmented columns, the 32 / W lane groups split the partial index), then forms
//   dU[i][d] = 2/l^2 (u_id rs_i - RV_id)         dV[j][d] = -2/l^2 (RtU_jd - v_jd cs_j)
// and this block's contribution to the lengthscale sums  sum u (u rs - 2 RV)  /  sum v^2 cs.
template <int W>
__global__ void __launch_bounds__(256) kmat_bwdf_assemble_kernel(
    const KmatBwdArgs a, const double* __restrict__ RVp, const double* __restrict__ RtUp,
    double* __restrict__ RtU, double* __restrict__ part2, int nchunks, int nrb, int nbU,
    int direct_dV, int blk0, int nblk_total) {
  __shared__ double hs[8][W];
  constexpr int G = 32 / W;
  const int k = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = lane % W, g = lane / W;
  const int blk = blk0 + (int)blockIdx.x;  // U blocks are [0, nbU), V blocks [nbU, nblk_total)
  const bool isU = blk < nbU;
  const int row = (isU ? blk : blk - nbU) * 8 + warp;
  const int nrows = isU ? a.nu : a.nv;
  const int nparts = isU ? nchunks : nrb;
  double s = 0.0;
  if (row < nrows) {
    const double* src = (isU ? RVp : RtUp) + ((long long)k * nparts * nrows + row) * W + c;
    for (int q = g; q < nparts; q += G) s += src[(long long)q * nrows * W];
  }
#pragma unroll
  for (int o = W; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const double tot = __shfl_sync(0xffffffffu, s, a.D);  // row sum rs_i / column sum cs_j
  double term = 0.0;
  if (row < nrows && g == 0 && c < a.D) {
    const double l = a.ls[(long long)k * a.D + c];
    const double cen = a.U[(long long)k * a.sU + c];  // the fused pass worked on U - U[0], V - U[0]
    if (isU) {
      const long long t = (long long)row * a.D + c;
      const double u = a.U[(long long)k * a.sU + t] - cen;
      const double gr = (2.0 / (l * l)) * (u * tot - s);
      double* dst = a.dU + (long long)k * a.sdU + t;
      *dst = a.accumulate_dU ? (*dst + gr) : gr;
      term = u * (u * tot - 2.0 * s);
    } else {
      const long long t = (long long)row * a.D + c;
      const double v = a.V[(long long)k * a.sV + t] - cen;
      if (direct_dV) {
        const double gr = -(2.0 / (l * l)) * (s - v * tot);
        double* dst = a.dV + (long long)k * a.sdV + t;
        *dst = a.accumulate_dV ? (*dst + gr) : gr;
      }
      term = v * v * tot;
    }
  }
  if (!isU && !direct_dV && row < nrows && g == 0) RtU[((long long)k * a.nv + row) * W + c] = s;
  if (g == 0) hs[warp][c] = term;
  __syncthreads();
  if (threadIdx.x < W) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += hs[w][threadIdx.x];
    part2[((long long)k * nblk_total + blk) * W + threadIdx.x] = t;
  }
}

// The same assembly for augmented widths that do not divide the warp (W = 24, 32, 40): a lane owns
// columns lane and lane + 32, partial indices are summed in order by every lane.
__global__ void __launch_bounds__(256) kmat_bwdf_assemble_generic_kernel(
    const KmatBwdArgs a, const double* __restrict__ RVp, const double* __restrict__ RtUp,
    double* __restrict__ RtU, double* __restrict__ part2, int W, int nchunks, int nrb, int nbU,
    int direct_dV, int blk0, int nblk_total) {
  __shared__ double hs[8][64];
  const int k = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk = blk0 + (int)blockIdx.x;
  const bool isU = blk < nbU;
  const int row = (isU ? blk : blk - nbU) * 8 + warp;
  const int nrows = isU ? a.nu : a.nv;
  const int nparts = isU ? nchunks : nrb;
  double sv[2] = {0.0, 0.0};
  if (row < nrows) {
    const double* src = (isU ? RVp : RtUp) + ((long long)k * nparts * nrows + row) * W;
    for (int q = 0; q < nparts; ++q) {
      const double* sq = src + (long long)q * nrows * W;
      if (lane < W) sv[0] += sq[lane];
      if (lane + 32 < W) sv[1] += sq[lane + 32];
    }
  }
  const double tot = (a.D < 32) ? __shfl_sync(0xffffffffu, sv[0], a.D) : __shfl_sync(0xffffffffu, sv[1], a.D - 32);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = lane + 32 * h;
    double term = 0.0;
    if (row < nrows && c < a.D) {
      const double s = sv[h];
      const double l = a.ls[(long long)k * a.D + c];
      const double cen = a.U[(long long)k * a.sU + c];
      const long long t = (long long)row * a.D + c;
      if (isU) {
        const double u = a.U[(long long)k * a.sU + t] - cen;
        const double gr = (2.0 / (l * l)) * (u * tot - s);
        double* dst = a.dU + (long long)k * a.sdU + t;
        *dst = a.accumulate_dU ? (*dst + gr) : gr;
        term = u * (u * tot - 2.0 * s);
      } else {
        const double v = a.V[(long long)k * a.sV + t] - cen;
        if (direct_dV) {
          const double gr = -(2.0 / (l * l)) * (s - v * tot);
          double* dst = a.dV + (long long)k * a.sdV + t;
          *dst = a.accumulate_dV ? (*dst + gr) : gr;
        }
        term = v * v * tot;
      }
    }
    if (!isU && !direct_dV && row < nrows && c < W) RtU[((long long)k * a.nv + row) * W + c] = sv[h];
    if (c < 64) hs[warp][c] = term;
  }
  __syncthreads();
  if ((int)threadIdx.x < W) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += hs[w][threadIdx.x];
    part2[((long long)k * nblk_total + blk) * W + threadIdx.x] = t;
  }
}

// dls[k][d] (+)= -2/l^3 sum_blocks part2[k][b][d] ;  dvar[k] (+)= sum part[k][:]     grid (D + 1, K)
__global__ void __launch_bounds__(256) kmat_bwdf_hyp_kernel(const KmatBwdArgs a,
                                                            const double* __restrict__ part2, int W,
                                                            int nblocks,
                                                            const double* __restrict__ part,
                                                            int nparts) {
  __shared__ double red[32];
  const int d = blockIdx.x, k = blockIdx.y;
  double s = 0.0;
  if (d == a.D) {
    for (int e = threadIdx.x; e < nparts; e += 256) s += part[(long long)k * nparts + e];
    s = block_sum(s, red);
    if (threadIdx.x == 0) a.dvar[k] = a.accumulate_hyp ? a.dvar[k] + s : s;
    return;
  }
  for (int b = threadIdx.x; b < nblocks; b += 256) s += part2[((long long)k * nblocks + b) * W + d];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double l = a.ls[(long long)k * a.D + d];
    const double gr = -2.0 / (l * l * l) * s;
    double* dst = a.dls + (long long)k * a.D + d;
    *dst = a.accumulate_hyp ? (*dst + gr) : gr;
  }
}

struct BwdFusedPlan {
  int NF, W, nchunks, nrb, nbU, nbV;
  size_t o_RVp, o_RtUp, o_RV, o_RtU, o_part, o_part2, total;
};

// up to 40 augmented columns; beyond 16 only for Kuu-sized column counts (the partials grow with
// K * (nu / 64) * nv * W)
bool bwd_fused_ok(int nv, int D) { return nv >= 128 && (D + 1 <= 16 || (D + 1 <= 40 && nv <= 2048)); }

BwdFusedPlan bwd_fused_plan(int K, int nu, int nv, int D) {
  BwdFusedPlan p;
  p.NF = ceil_div(D + 1, 8);
  p.W = 8 * p.NF;
  p.nchunks = ceil_div(nv, 128);
  p.nrb = ceil_div(nu, 64);
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += (n + 31) / 32 * 32;
    return o;
  };
  p.o_RVp = take((size_t)K * p.nchunks * nu * p.W);
  p.o_RtUp = take((size_t)K * p.nrb * nv * p.W);
  p.o_RV = take((size_t)K * nu * p.W);
  p.o_RtU = take((size_t)K * nv * p.W);
  p.o_part = take((size_t)K * p.nrb * p.nchunks);
  p.nbU = ceil_div(nu, 8);
  p.nbV = ceil_div(nv, 8);
  p.o_part2 = take((size_t)K * (p.nbU + p.nbV) * p.W);
  p.total = off;
  return p;
}

int launch_kmat_bwd_fused(const KmatBwdArgs& a, cudaStream_t stream) {
  const bool shared_v = (a.sV == 0 && a.sdV == 0);
  const BwdFusedPlan p = bwd_fused_plan(a.K, a.nu, a.nv, a.D);
  double* RVp = a.scratch + p.o_RVp;
  double* RtUp = a.scratch + p.o_RtUp;
  double* RtU = a.scratch + p.o_RtU;
  double* part = a.scratch + p.o_part;
  {
    dim3 grid(p.nchunks, p.nrb, a.K);
    const int LDA = p.W + 4;
    const size_t smem = ((size_t)64 * 132 + (size_t)(64 + 128) * LDA + p.W + 32 + (size_t)64 * p.W) * sizeof(double);
    static SmemAttrOnce once[5];
#define GPFX_FUSED_CASE(NFV)                                                                   \
  case NFV:                                                                                    \
    GPFX_TRY(ensure_dyn_smem(kmat_bwd_fused_kernel<NFV>, 160 * 1024, once[NFV - 1]));          \
    kmat_bwd_fused_kernel<NFV><<<grid, 256, smem, stream>>>(a, RVp, RtUp, part);               \
    break;
    switch (p.NF) {
      GPFX_FUSED_CASE(1)
      GPFX_FUSED_CASE(2)
      GPFX_FUSED_CASE(3)
      GPFX_FUSED_CASE(4)
      GPFX_FUSED_CASE(5)
      default:
        set_last_error("kmat_bwd (fused): input dim %d too large", a.D);
        return GPFX_E_UNSUPPORTED;
    }
#undef GPFX_FUSED_CASE
    GPFX_LAUNCH_CHECK();
  }
  double* part2 = a.scratch + p.o_part2;
  const int direct_dV = (!shared_v || a.K == 1) ? 1 : 0;
  const int nblk = p.nbU + p.nbV;
  // Kuu: dU and dV are the same array (u = v): the row side must land before the column side adds to it
  const bool aliased = (a.dU == a.dV);
  for (int pass = 0; pass < (aliased ? 2 : 1); ++pass) {
    const int blk0 = aliased ? (pass == 0 ? 0 : p.nbU) : 0;
    const int nb = aliased ? (pass == 0 ? p.nbU : p.nbV) : nblk;
    dim3 grid(nb, a.K);
    if (p.NF == 1)
      kmat_bwdf_assemble_kernel<8><<<grid, 256, 0, stream>>>(a, RVp, RtUp, RtU, part2, p.nchunks, p.nrb,
                                                              p.nbU, direct_dV, blk0, nblk);
    else if (p.NF == 2)
      kmat_bwdf_assemble_kernel<16><<<grid, 256, 0, stream>>>(a, RVp, RtUp, RtU, part2, p.nchunks, p.nrb,
                                                               p.nbU, direct_dV, blk0, nblk);
    else
      kmat_bwdf_assemble_generic_kernel<<<grid, 256, 0, stream>>>(a, RVp, RtUp, RtU, part2, p.W, p.nchunks,
                                                                  p.nrb, p.nbU, direct_dV, blk0, nblk);
    GPFX_LAUNCH_CHECK();
  }
  if (!direct_dV) {  // V shared by K > 1 kernels: dV sums over k
    dim3 g2((unsigned)ceil_div_ll((long long)a.nv * a.D, 128), 1);
    kmat_bwdg_dV_kernel<<<g2, 128, 0, stream>>>(a, RtU, p.W, 1, 1);
    GPFX_LAUNCH_CHECK();
  }
  {
    dim3 g3(a.D + 1, a.K);
    kmat_bwdf_hyp_kernel<<<g3, 256, 0, stream>>>(a, part2, p.W, nblk, part, p.nrb * p.nchunks);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

// SquaredExponential with the forward kernel matrix saved: R = dK * dk/dr2 = -dK K / 2 needs neither
// the distances nor the exp - one streaming pass (read dK and K, write R in place), with the
// sum dK K / var for d variance reduced per block.  grid (nparts, K): block b owns rows b, b + nparts, ...
__global__ void __launch_bounds__(256) kmat_R_from_K_kernel(const KmatBwdArgs a,
                                                            double* __restrict__ part, int vec) {
  __shared__ double red[32];
  const int k = blockIdx.y;
  double* G = a.dK + (long long)k * a.sK;
  const double* Kf = a.Kmat + (long long)k * a.sKmat;
  double acc = 0.0;
  for (int i = blockIdx.x; i < a.nu; i += gridDim.x) {
    double* g = G + (long long)i * a.ldk;
    const double* kf = Kf + (long long)i * a.ldkmat;
    if (vec) {
      for (int j = 2 * threadIdx.x; j < a.nv; j += 512) {
        double2 gv = *reinterpret_cast<const double2*>(g + j);
        const double2 kv = *reinterpret_cast<const double2*>(kf + j);
        const double p0 = gv.x * kv.x, p1 = gv.y * kv.y;
        acc += p0 + p1;
        gv.x = -0.5 * p0;
        gv.y = -0.5 * p1;
        *reinterpret_cast<double2*>(g + j) = gv;
      }
    } else {
      for (int j = threadIdx.x; j < a.nv; j += 256) {
        const double p0 = g[j] * kf[j];
        acc += p0;
        g[j] = -0.5 * p0;
      }
    }
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) part[(long long)k * gridDim.x + blockIdx.x] = t / a.var[k];
}

// GEMM form of the backward (used for the large Kuf matrices):
//   R = dK * dk/dr2            kmat_fwd_kernel mode 3 (cross term on DMMA, written over dK)
//   RV  = R  [V, 1]            DMMA GEMM  -> sum_j R_ij v_j and the row sums
//   RtU = R^T [U, 1]           DMMA GEMM  -> sum_i R_ij u_i and the column sums
//   dU = 2/l^2 (u rs - RV),  dV = -2/l^2 (RtU - v cs),  dl = -2/l^3 (sum u^2 rs + sum v^2 cs - 2 sum u RV)
int launch_kmat_bwd_gemm(const KmatBwdArgs& a, cudaStream_t stream) {
  const bool shared_v = (a.sV == 0 && a.sdV == 0);
  const BwdGemmPlan p = bwd_gemm_plan(a.K, a.nu, a.nv, a.D, shared_v);
  double* Uaug = a.scratch + p.o_Uaug;
  double* Vaug = a.scratch + p.o_Vaug;
  double* RV = a.scratch + p.o_RV;
  double* RtU = a.scratch + p.o_RtU;
  double* part = a.scratch + p.o_part;
  double* split = a.scratch + p.o_split;
  if (a.Kmat != nullptr && a.kind == GPFX_KERNEL_SE) {
    const int vec = ((a.nv & 1) == 0) && ((a.ldk & 1) == 0) && ((a.ldkmat & 1) == 0) && ((a.sK & 1) == 0) &&
                    ((a.sKmat & 1) == 0) && ((((uintptr_t)a.dK | (uintptr_t)a.Kmat) & 15) == 0);
    dim3 grid(p.tiles, a.K);
    kmat_R_from_K_kernel<<<grid, 256, 0, stream>>>(a, part, vec);
    GPFX_LAUNCH_CHECK();
  } else {
    KmatArgs k;
    k.U = a.U; k.V = a.V; k.ls = a.ls; k.var = a.var; k.out = a.dK;
    k.K = a.K; k.nu = a.nu; k.nv = a.nv; k.D = a.D;
    k.sU = a.sU; k.sV = a.sV; k.sOut = a.sK;
    k.ldo = a.ldk; k.rows_out = a.nu; k.cols_out = a.nv;
    k.kind = a.kind; k.mode = 3; k.dK_in = a.dK; k.part_out = part;
    GPFX_TRY(launch_kmat_fwd(k, stream));
  }
  {
    const long long nU = (long long)a.K * a.nu, nV = (long long)(shared_v ? 1 : a.K) * a.nv;
    augment_ones_kernel<<<(int)std::min<long long>(ceil_div_ll(nU * p.Dp, 256), 2048), 256, 0, stream>>>(
        a.U, Uaug, nU, a.D, p.Dp);
    GPFX_LAUNCH_CHECK();
    augment_ones_kernel<<<(int)std::min<long long>(ceil_div_ll(nV * p.Dp, 256), 2048), 256, 0, stream>>>(
        a.V, Vaug, nV, a.D, p.Dp);
    GPFX_LAUNCH_CHECK();
  }
  {
    GemmArgs g;
    g.A = a.dK; g.lda = a.ldk; g.sA1 = a.sK;
    g.B = Vaug; g.ldb = p.Dp; g.sB1 = shared_v ? 0 : (long long)a.nv * p.Dp;
    g.C = RV; g.ldc = p.Dp; g.sC1 = (long long)a.nu * p.Dp;
    g.M = a.nu; g.N = p.Dp; g.K = a.nv;
    g.batch1 = a.K;
    g.splits = p.splits1; g.splitws = split;
    GPFX_TRY(launch_gemm(g, stream));
  }
  if (a.dV != nullptr) {
    GemmArgs g;
    g.A = a.dK; g.lda = a.ldk; g.sA1 = a.sK; g.transA = 1;
    g.B = Uaug; g.ldb = p.Dp; g.sB1 = (long long)a.nu * p.Dp;
    g.C = RtU; g.ldc = p.Dp; g.sC1 = (long long)a.nv * p.Dp;
    g.M = a.nv; g.N = p.Dp; g.K = a.nu;
    g.batch1 = a.K;
    g.splits = p.splits2; g.splitws = split;
    GPFX_TRY(launch_gemm(g, stream));
  } else {
    // the lengthscale gradient needs the column sums even when dV is not requested
    set_last_error("kmat_bwd (GEMM form): dV must be requested");
    return GPFX_E_UNSUPPORTED;
  }
  {
    dim3 g1((unsigned)ceil_div_ll((long long)a.nu * a.D, 128), a.K);
    kmat_bwdg_dU_kernel<<<g1, 128, 0, stream>>>(a, RV, p.Dp);
    GPFX_LAUNCH_CHECK();
    dim3 g2((unsigned)ceil_div_ll((long long)a.nv * a.D, 128), shared_v ? 1 : a.K);
    kmat_bwdg_dV_kernel<<<g2, 128, 0, stream>>>(a, RtU, p.Dp, shared_v ? 1 : 0);
    GPFX_LAUNCH_CHECK();
    dim3 g3(a.D + 1, a.K);
    kmat_bwdg_hyp_kernel<<<g3, 256, 0, stream>>>(a, RV, RtU, p.Dp, part, p.tiles);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace

// ---- random-Fourier-feature backward ---------------------------------------------------------
namespace {
// dX[n][d] = sum_k dXs[k][n][d] / l_kd  (X is shared by the K feature sets)
__global__ void rff_bwd_dx_kernel(const double* __restrict__ dXs, const double* __restrict__ ls,
                                  double* __restrict__ dX, int K, long long N, int D) {
  const long long total = N * D;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += dXs[(long long)k * total + e] / ls[k * D + d];
    dX[e] = s;
  }
}
// grid (D + 1, K): dls[k][d] = -sum_n dXs[k][n][d] x_nd / l_kd^2 ;  block D: dvar[k] = sum of the tile partials
__global__ void __launch_bounds__(256) rff_bwd_hyp_kernel(const double* __restrict__ dXs, const double* __restrict__ X,
                                                          const double* __restrict__ ls, const double* __restrict__ part,
                                                          int ntiles, double* __restrict__ dls,
                                                          double* __restrict__ dvar, long long N, int D) {
  __shared__ double red[32];
  const int d = blockIdx.x, k = blockIdx.y;
  double acc = 0.0;
  if (d < D) {
    for (long long n = threadIdx.x; n < N; n += 256) acc += dXs[((long long)k * N + n) * D + d] * X[n * D + d];
  } else {
    for (int t = threadIdx.x; t < ntiles; t += 256) acc += part[(long long)k * ntiles + t];
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    if (d < D) {
      const double l = ls[k * D + d];
      dls[k * D + d] = -tot / (l * l);
    } else {
      dvar[k] = tot;
    }
  }
}
}  // namespace

size_t rff_bwd_scratch_doubles(int K, int N, int D, int L) {
  const size_t tiles = (size_t)ceil_div(N, FT) * ceil_div(L, FT);
  return (size_t)K * N * L + (size_t)K * N * D + (size_t)K * tiles + 16;
}

int launch_rff_bwd(const RffBwdArgs& a, cudaStream_t stream) {
  if (a.K <= 0 || a.N <= 0 || a.L <= 0 || a.D <= 0) return GPFX_OK;
  const int tiles = ceil_div(a.N, FT) * ceil_div(a.L, FT);
  double* G = a.scratch;                                   // [K][N][L]  dLoss/dz
  double* dXs = G + (size_t)a.K * a.N * a.L;               // [K][N][D]  dLoss/d(X / l_k)
  double* part = dXs + (size_t)a.K * a.N * a.D;            // [K][tiles]
  {
    KmatArgs k;
    k.U = a.X; k.V = a.W; k.ls = a.ls; k.var = a.var; k.out = G;
    k.K = a.K; k.nu = a.N; k.nv = a.L; k.D = a.D;
    k.sU = 0; k.sV = (long long)a.L * a.D;
    k.sOut = (long long)a.N * a.L; k.ldo = a.L; k.rows_out = a.N; k.cols_out = a.L;
    k.mode = a.concat ? 5 : 4;
    k.bias = a.bias; k.sBias = a.L;
    const int width = a.concat ? 2 * a.L : a.L;
    k.dK_in = a.dPhi; k.sdK = (long long)a.N * width; k.ld_dk = width;
    k.part_out = part;
    GPFX_TRY(launch_kmat_fwd(k, stream));
  }
  {
    GemmArgs g;  // dXs_k = G_k W_k
    g.A = G; g.lda = a.L; g.sA1 = (long long)a.N * a.L;
    g.B = a.W; g.ldb = a.D; g.sB1 = (long long)a.L * a.D;
    g.C = dXs; g.ldc = a.D; g.sC1 = (long long)a.N * a.D;
    g.M = a.N; g.N = a.D; g.K = a.L;
    g.batch1 = a.K;
    GPFX_TRY(launch_gemm(g, stream));
  }
  {
    const long long total = (long long)a.N * a.D;
    rff_bwd_dx_kernel<<<(int)std::min<long long>(ceil_div_ll(total, 256), 2048), 256, 0, stream>>>(dXs, a.ls, a.dX,
                                                                                                  a.K, a.N, a.D);
    GPFX_LAUNCH_CHECK();
    dim3 grid(a.D + 1, a.K);
    rff_bwd_hyp_kernel<<<grid, 256, 0, stream>>>(dXs, a.X, a.ls, part, tiles, a.dls, a.dvar, a.N, a.D);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

size_t kmat_bwd_scratch_doubles(int K, int nu, int nv, int D) {
  const BwdSplit s = bwd_split(K, nu, nv);
  const size_t direct = (size_t)K * s.nchunks * nu * D + (size_t)K * s.nrb * s.nchunks * (D + 1) +
                        (size_t)K * s.RC * nv * D + 64;
  const size_t gemm = std::max(bwd_gemm_plan(K, nu, nv, D, true).total,
                               bwd_gemm_plan(K, nu, nv, D, false).total);
  const size_t fused = bwd_fused_ok(nv, D) ? bwd_fused_plan(K, nu, nv, D).total : 0;
  return std::max(std::max(direct, gemm), fused);
}

int launch_kmat_bwd(const KmatBwdArgs& a_in, cudaStream_t stream) {
  KmatBwdArgs a = a_in;
  if (a.K <= 0 || a.nu <= 0 || a.nv <= 0) return GPFX_OK;
  if (a.gemm_form == 2 && a.dV != nullptr && bwd_fused_ok(a.nv, a.D)) return launch_kmat_bwd_fused(a, stream);
  if (a.gemm_form && a.dV != nullptr) return launch_kmat_bwd_gemm(a, stream);
  const BwdSplit s = bwd_split(a.K, a.nu, a.nv);
  a.partU = a.scratch;
  a.part = a.partU + (size_t)a.K * s.nchunks * a.nu * a.D;
  a.partV = a.part + (size_t)a.K * s.nrb * s.nchunks * (a.D + 1);
  {
    dim3 grid(s.nchunks, s.nrb, a.K);
    const int DC = (a.D <= 8) ? 8 : 32;
    // super-tile: as many 128-column sub-tiles of V as fit in ~150 KB of shared memory
    const size_t fixed = ((size_t)8 * (a.D + 1) + a.D + 8 * (DC + 1)) * sizeof(double);
    int stile = (int)(((150 * 1024 - fixed) / ((a.D + 1) * sizeof(double))) / 128) * 128;
    stile = std::max(128, std::min(stile, s.chunk_cols));
    const size_t smem = (size_t)stile * (a.D + 1) * sizeof(double) + fixed;
    if (smem > 200 * 1024) {
      set_last_error("kmat_bwd: input dim %d too large", a.D);
      return GPFX_E_UNSUPPORTED;
    }
    static SmemAttrOnce once8, once32;
    if (a.D <= 8) {
      GPFX_TRY(ensure_dyn_smem(kmat_bwd_rows_kernel<8>, 200 * 1024, once8));
      kmat_bwd_rows_kernel<8><<<grid, 256, smem, stream>>>(a, s.nchunks, s.chunk_cols, stile);
    } else {
      GPFX_TRY(ensure_dyn_smem(kmat_bwd_rows_kernel<32>, 200 * 1024, once32));
      kmat_bwd_rows_kernel<32><<<grid, 256, smem, stream>>>(a, s.nchunks, s.chunk_cols, stile);
    }
    GPFX_LAUNCH_CHECK();
    const long long work = (long long)a.nu * a.D + a.D + 1;
    dim3 fgrid((unsigned)ceil_div_ll(work, 128), a.K);
    kmat_bwd_rows_finish_kernel<<<fgrid, 128, 0, stream>>>(a, s.nchunks, s.nrb);
    GPFX_LAUNCH_CHECK();
  }
  if (a.dV != nullptr) {
    const bool shared_v = (a.sV == 0 && a.sdV == 0);
    dim3 grid(ceil_div(a.nv, 128), s.RC, a.K);
    const size_t smem = ((size_t)64 * a.D) * sizeof(double);
    if (smem > 48 * 1024) {
      set_last_error("kmat_bwd: input dim %d too large", a.D);
      return GPFX_E_UNSUPPORTED;
    }
    if (a.D <= 8) kmat_bwd_cols_kernel<8><<<grid, 128, smem, stream>>>(a, s.RC, s.chunk_rows);
    else kmat_bwd_cols_kernel<32><<<grid, 128, smem, stream>>>(a, s.RC, s.chunk_rows);
    GPFX_LAUNCH_CHECK();
    dim3 fgrid((unsigned)ceil_div_ll((long long)a.nv * a.D, 128), shared_v ? 1 : a.K);
    kmat_bwd_cols_finish_kernel<<<fgrid, 128, 0, stream>>>(a, s.RC, shared_v ? 1 : 0);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace gpfx
