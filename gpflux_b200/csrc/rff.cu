// Efficient posterior sampling (Wilson et al. 2020; reference gpflux/sampling/sample.py:184-198):
//   f(X) = Phi(X) w + K_XZ v,   Phi(X) = sqrt(2 var / L) cos((X/l) W^T + b)
// fused so that neither Phi [N,L] (5.1 TB per layer at cfg-5 sizes) nor K_XZ [N,M] ever reaches HBM:
// each thread owns R query rows, streams the feature / inducing parameters through shared memory
// (broadcast reads) and keeps its P running sums in registers.  The kernel is bound by FP64 cos
// throughput (SURVEY §7 hard part 6), not by DMMA or HBM.
#include "rff.h"

#include <math.h>

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

constexpr int WT = 128;  // threads per block
constexpr int FC = 128;  // features / inducing points staged per chunk

template <int DMAX, int PMAX, int R>
__global__ void __launch_bounds__(WT) wilson_eval_kernel(const WilsonArgs a) {
  extern __shared__ double sm[];
  double* Ws = sm;                 // [FC][D]
  double* bs = Ws + FC * a.D;      // [FC]
  double* ws = bs + FC;            // [FC][P]
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
    rows[r] = (blockIdx.x * R + r) * WT + tid;
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      x[r][d] = (d < a.D && rows[r] < a.N) ? X[(long long)rows[r] * a.D + d] / a.ls[d] : 0.0;
#pragma unroll
    for (int p = 0; p < PMAX; ++p) acc[r][p] = 0.0;
  }

  // ---- weight-space prior: sum_l cos(x.W_l + b_l) w_l
  for (int l0 = 0; l0 < a.L; l0 += FC) {
    const int nl = min(FC, a.L - l0);
    __syncthreads();
    for (int e = tid; e < nl * a.D; e += WT) Ws[e] = a.W[(long long)l0 * a.D + e];
    for (int e = tid; e < nl; e += WT) bs[e] = a.b[l0 + e];
    for (int e = tid; e < nl * a.P; e += WT) ws[e] = wS[(long long)l0 * a.P + e];
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
  for (int m0 = 0; m0 < a.M; m0 += FC) {
    const int nm = min(FC, a.M - m0);
    __syncthreads();
    for (int e = tid; e < nm * a.D; e += WT) Ws[e] = a.Z[(long long)m0 * a.D + e] / a.ls[e % a.D];
    for (int e = tid; e < nm * a.P; e += WT) ws[e] = vS[(long long)m0 * a.P + e];
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
int launch_t(const WilsonArgs& a, cudaStream_t stream) {
  dim3 grid(ceil_div(a.N, WT * R), a.S);
  const size_t smem = ((size_t)FC * a.D + FC + (size_t)FC * a.P) * sizeof(double);
  wilson_eval_kernel<DMAX, PMAX, R><<<grid, WT, smem, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

}  // namespace

int launch_wilson_eval(const WilsonArgs& a, cudaStream_t stream) {
  if (a.S <= 0 || a.N <= 0) return GPFX_OK;
  if (a.D <= 0 || a.P <= 0 || a.L <= 0 || a.M <= 0 || a.D > 32 || a.P > 32) {
    set_last_error("gpfx_wilson_eval: unsupported shape D=%d P=%d L=%d M=%d (D, P <= 32)", a.D, a.P,
                   a.L, a.M);
    return GPFX_E_UNSUPPORTED;
  }
  if (a.D == 1 && a.P == 1) return launch_t<1, 1, 4>(a, stream);
  if (a.D <= 8 && a.P <= 8) return launch_t<8, 8, 2>(a, stream);
  return launch_t<32, 32, 1>(a, stream);
}

}  // namespace gpfx
