// HBM-bound streaming kernels of the SVGP layer.
// All reductions are two-stage with fixed trees (no floating-point atomics) so results are
// bit-reproducible run to run.
#include "elementwise.h"

#include <math.h>

namespace gpfx {
namespace {

constexpr int KL_ROWS = 8;  // rows of q_sqrt per block (one warp per row)

constexpr int KL_FWD_ROWS = 32;  // rows of q_sqrt per block of the forward kernel (four per warp)

// grid (ceil(M/32), P); each warp owns rows base + warp + 8 r (interleaved, so the triangle's row
// lengths balance across warps).  A row is read with 128-bit loads, four per lane in flight, and
// only up to the diagonal (the zeros above it are written, not read).
__global__ void __launch_bounds__(256) kl_tril_kernel(const double* __restrict__ q_mu,
                                                      const double* __restrict__ q_sqrt,
                                                      double* __restrict__ Lq,
                                                      double* __restrict__ scratch, int M, int P,
                                                      int vec) {
  __shared__ double red[32];
  const int p = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sq = 0.0, ld = 0.0, mah = 0.0;
  if (threadIdx.x < KL_FWD_ROWS) {  // this block's slice of q_mu[:, p]^2 (the Mahalanobis term)
    const int m = blockIdx.x * KL_FWD_ROWS + threadIdx.x;  // grid.x * 32 >= M (see launch_kl_white)
    if (m < M) {
      const double v = q_mu[(long long)m * P + p];
      mah = v * v;
    }
  }
  // rows are paired (t, M-1-t) so that every block streams the same number of triangle elements
  const int half = (M + 1) / 2;
  for (int r = 0; r < KL_FWD_ROWS / 8; ++r) {
    const int idx = warp + 8 * r;
    const int t = blockIdx.x * (KL_FWD_ROWS / 2) + (idx & (KL_FWD_ROWS / 2 - 1));
    if (t >= half) continue;
    const int i = (idx < KL_FWD_ROWS / 2) ? t : M - 1 - t;
    if (idx >= KL_FWD_ROWS / 2 && i == t) continue;  // middle row of an odd M: counted once
    const double* src = q_sqrt + ((long long)p * M + i) * M;
    double* dst = Lq ? Lq + ((long long)p * M + i) * M : nullptr;
    if (vec) {
      const int jread = i + 1;             // elements j <= i are read
      const int jend = dst ? M : jread;    // ... and the whole row is written when Lq is kept
      for (int j0 = 0; j0 < jend; j0 += 256) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 64 * u + 2 * lane;
          v[u] = (j < jread) ? *reinterpret_cast<const double2*>(src + j) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 64 * u + 2 * lane;
          if (j + 1 > i) v[u].y = 0.0;  // j + 1 above the diagonal (j <= i holds or v is already 0)
          if (j == i) ld += log(v[u].x * v[u].x);
          if (j + 1 == i) ld += log(v[u].y * v[u].y);
          sq += v[u].x * v[u].x + v[u].y * v[u].y;
          if (dst && j < M) *reinterpret_cast<double2*>(dst + j) = v[u];
        }
      }
    } else {
      for (int j = lane; j < M; j += 32) {
        const double v = (j <= i) ? src[j] : 0.0;
        if (dst) dst[j] = v;
        sq += v * v;
        if (j == i) ld += log(v * v);
      }
    }
  }
  const double tsq = block_sum(sq, red);
  const double tld = block_sum(ld, red);
  const double tmah = block_sum(mah, red);
  if (threadIdx.x == 0) {
    double* o = scratch + ((long long)p * gridDim.x + blockIdx.x) * 3;
    o[0] = tsq;
    o[1] = tld;
    o[2] = tmah;
  }
}

__global__ void __launch_bounds__(256) kl_finish_kernel(const double* __restrict__ scratch,
                                                        int nparts, int M, int P,
                                                        double* __restrict__ kl) {
  __shared__ double red[32];
  double mah = 0.0, sq = 0.0, ld = 0.0;
  for (int e = threadIdx.x; e < nparts; e += 256) {
    sq += scratch[3 * e];
    ld += scratch[3 * e + 1];
    mah += scratch[3 * e + 2];
  }
  mah = block_sum(mah, red);
  sq = block_sum(sq, red);
  ld = block_sum(ld, red);
  if (threadIdx.x == 0) kl[0] = 0.5 * (mah - (double)M * (double)P - ld + sq);
}

// grid (ceil(M/8) + 1, P); one warp per row of dq_sqrt_p (128-bit accesses, Lq read only up to the
// diagonal); the extra block row handles dq_mu[:, p].
__global__ void __launch_bounds__(256) kl_white_bwd_kernel(const double* __restrict__ q_mu,
                                                           const double* __restrict__ Lq,
                                                           const double* __restrict__ g_kl,
                                                           double* __restrict__ dq_mu,
                                                           double* __restrict__ dq_sqrt, int M, int P,
                                                           int accumulate, int vec) {
  const double g = g_kl[0];
  const int p = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x == gridDim.x - 1) {
    for (int m = threadIdx.x; m < M; m += 256) {
      const long long t = (long long)m * P + p;
      const double v = g * q_mu[t];
      dq_mu[t] = accumulate ? dq_mu[t] + v : v;
    }
    return;
  }
  const int i = blockIdx.x * KL_ROWS + warp;
  if (i >= M) return;
  const double* src = Lq + ((long long)p * M + i) * M;
  double* dst = dq_sqrt + ((long long)p * M + i) * M;
  if (vec) {
    for (int j0 = 0; j0 < M; j0 += 256) {
      double2 v[4], o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 64 * u + 2 * lane;
        v[u] = (j <= i) ? *reinterpret_cast<const double2*>(src + j) : make_double2(0.0, 0.0);
        o[u] = (accumulate && j < M) ? *reinterpret_cast<const double2*>(dst + j) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 64 * u + 2 * lane;
        if (j >= M) continue;
        double a = 0.0, b = 0.0;
        if (j <= i) a = g * ((j == i) ? (v[u].x - 1.0 / v[u].x) : v[u].x);
        if (j + 1 <= i) b = g * ((j + 1 == i) ? (v[u].y - 1.0 / v[u].y) : v[u].y);
        *reinterpret_cast<double2*>(dst + j) = make_double2(o[u].x + a, o[u].y + b);
      }
    }
  } else {
    for (int j = lane; j < M; j += 32) {
      double v = 0.0;
      if (j <= i) {
        const double l = src[j];
        v = g * ((i == j) ? (l - 1.0 / l) : l);
      }
      dst[j] = accumulate ? dst[j] + v : v;
    }
  }
}

__global__ void finalize_kernel(const FinalizeArgs a) {
  const long long total = (long long)a.N * a.P;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e / a.P), p = (int)(e % a.P);
    const int k = (a.K > 1) ? p : 0;
    const int j = (a.K > 1) ? 0 : p;
    const int nw = (a.K > 1) ? 1 : a.P;
    double mean = 0.0, s0 = 0.0, s1 = 0.0;
    for (int b = 0; b < a.gridM_A; ++b) {
      mean += a.wsum[(((long long)k * a.gridM_A + b) * nw + j) * a.N + n];
      s0 += a.sqA[((long long)k * a.gridM_A + b) * a.N + n];
    }
    for (int b = 0; b < a.gridM_B; ++b) s1 += a.sqB[((long long)p * a.gridM_B + b) * a.N + n];
    const double var = (a.variance[k] - s0) + s1;
    if (a.mean_add) mean += a.mean_add[e];
    a.fmean[e] = mean;
    a.fvar[e] = var;
    if (a.fsample) a.fsample[e] = a.eps ? mean + sqrt(var) * a.eps[e] : mean;
  }
}

__global__ void reparam_sample_kernel(const double* __restrict__ mean,
                                      const double* __restrict__ var,
                                      const double* __restrict__ eps, double* __restrict__ f,
                                      long long n, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec) {
    const long long n2 = n >> 1;
    const double2* m2 = reinterpret_cast<const double2*>(mean);
    const double2* v2 = reinterpret_cast<const double2*>(var);
    const double2* e2 = reinterpret_cast<const double2*>(eps);
    double2* f2 = reinterpret_cast<double2*>(f);
    for (long long i = t; i < n2; i += stride) {
      const double2 m = m2[i], v = v2[i], e = e2[i];
      f2[i] = make_double2(m.x + sqrt(v.x) * e.x, m.y + sqrt(v.y) * e.y);
    }
    if (t == 0 && (n & 1)) f[n - 1] = mean[n - 1] + sqrt(var[n - 1]) * eps[n - 1];
  } else {
    for (long long i = t; i < n; i += stride) f[i] = mean[i] + sqrt(var[i]) * eps[i];
  }
}


// ---- a8, large inputs: the same map streamed through TMA-staged shared memory -----------------
// One elected thread per CTA drives 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of the three input
// tiles into a 3-stage shared-memory ring and arms an mbarrier with the expected byte count; all
// threads wait on the barrier's phase, compute f = mean + sqrt(var) eps on 128-bit shared-memory
// accesses and the elected thread streams the result tile back with a bulk shared->global store.
// Persistent CTAs (one per SM), so the ring stays full for the whole array.
namespace tma {
constexpr int TILE = 2048;  // doubles per array per stage (16 KB)
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr size_t SMEM = (size_t)STAGES * 4 * TILE * sizeof(double) + STAGES * sizeof(unsigned long long);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, unsigned bytes,
                                          unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
}  // namespace tma

__global__ void __launch_bounds__(tma::THREADS, 1)
    reparam_sample_tma_kernel(const double* __restrict__ mean, const double* __restrict__ var,
                              const double* __restrict__ eps, double* __restrict__ f, long long n) {
  using namespace tma;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* buf = reinterpret_cast<double*>(smem_raw);  // [STAGES][4][TILE]: mean, var, eps, out
  unsigned long long* full = reinterpret_cast<unsigned long long*>(buf + (size_t)STAGES * 4 * TILE);
  const int tid = threadIdx.x;
  const long long n_even = n & ~1LL;  // bulk copies move multiples of 16 bytes
  const long long ntiles = (n_even + TILE - 1) / TILE;
  const long long first = blockIdx.x, step = gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  auto tile_elems = [&](long long t) -> unsigned {
    const long long left = n_even - t * TILE;
    return (unsigned)(left < TILE ? left : TILE);
  };
  auto issue_load = [&](long long t, int s) {
    const unsigned bytes = tile_elems(t) * (unsigned)sizeof(double);
    double* b = buf + (size_t)s * 4 * TILE;
    mbar_expect_tx(&full[s], 3 * bytes);
    bulk_load(b, mean + t * TILE, bytes, &full[s]);
    bulk_load(b + TILE, var + t * TILE, bytes, &full[s]);
    bulk_load(b + 2 * TILE, eps + t * TILE, bytes, &full[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      const long long t = first + s * step;
      if (t < ntiles) issue_load(t, s);
    }
  }
  int s = 0;
  unsigned parity = 0;
  for (long long t = first; t < ntiles; t += step) {
    // the bulk store issued STAGES tiles ago has finished reading this stage's output buffer
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(STAGES - 1) : "memory");
    __syncthreads();
    mbar_wait(&full[s], parity);
    double* b = buf + (size_t)s * 4 * TILE;
    const unsigned ne = tile_elems(t);
    const double2* m2 = reinterpret_cast<const double2*>(b);
    const double2* v2 = reinterpret_cast<const double2*>(b + TILE);
    const double2* e2 = reinterpret_cast<const double2*>(b + 2 * TILE);
    double2* o2 = reinterpret_cast<double2*>(b + 3 * TILE);
#pragma unroll
    for (int i = 0; i < TILE / 2 / THREADS; ++i) {
      const unsigned j = tid + i * THREADS;
      if (2 * j < ne) {
        const double2 m = m2[j], v = v2[j], e = e2[j];
        o2[j] = make_double2(m.x + sqrt(v.x) * e.x, m.y + sqrt(v.y) * e.y);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes -> async proxy
    __syncthreads();
    if (tid == 0) {
      bulk_store(f + t * TILE, b + 3 * TILE, ne * (unsigned)sizeof(double));
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      const long long tn = t + (long long)STAGES * step;
      if (tn < ntiles) issue_load(tn, s);
    }
    if (++s == STAGES) {
      s = 0;
      parity ^= 1u;
    }
  }
  if (tid == 0) {
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    if (blockIdx.x == 0 && (n & 1)) f[n - 1] = mean[n - 1] + sqrt(var[n - 1]) * eps[n - 1];
  }
}

// thread per row n
__global__ void bwd_prep_kernel(const BwdPrepArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  double acc = 0.0;
  for (int p = 0; p < a.P; ++p) {
    const long long e = (long long)n * a.P + p;
    double gm = a.g_mean ? a.g_mean[e] : 0.0;
    double gv = a.g_var ? a.g_var[e] : 0.0;
    if (a.g_f) {
      const double gf = a.g_f[e];
      gm += gf;
      if (a.eps) gv += gf * a.eps[e] / (2.0 * sqrt(a.fvar[e]));
    }
    if (a.dmean) a.dmean[e] = gm;
    a.gmT[(long long)p * a.N + n] = gm;
    a.gv2T[(long long)p * a.N + n] = 2.0 * gv;
    if (a.K > 1) a.gvsum[(long long)p * a.N + n] = gv;
    else acc += gv;
  }
  if (a.K == 1) a.gvsum[n] = acc;
}

constexpr int EW_ROWS = 16;  // rows per block of the column-oriented streaming kernels below

// B_p[m][n] *= s[p][n].  grid (ceil(N/512), ceil(M/16), P): a thread owns two adjacent columns
// (128-bit accesses, its scale pair in registers) and walks the block's rows.
__global__ void __launch_bounds__(256) scale_cols_kernel(double* __restrict__ B,
                                                         const double* __restrict__ s, int M, int N,
                                                         int vec) {
  const int p = blockIdx.z;
  const long long total = (long long)M * N;
  double* Bp = B + p * total;
  const double* sp = s + (long long)p * N;
  const int m0 = blockIdx.y * EW_ROWS, m1 = min(M, m0 + EW_ROWS);
  const int n = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (n >= N) return;
  if (vec) {
    const double2 sv = *reinterpret_cast<const double2*>(sp + n);
#pragma unroll 4
    for (int m = m0; m < m1; ++m) {
      double2* q = reinterpret_cast<double2*>(Bp + (long long)m * N + n);
      double2 v = *q;
      v.x *= sv.x;
      v.y *= sv.y;
      *q = v;
    }
  } else {
    const double s0 = sp[n], s1 = (n + 1 < N) ? sp[n + 1] : 0.0;
    for (int m = m0; m < m1; ++m) {
      double* q = Bp + (long long)m * N + n;
      q[0] *= s0;
      if (n + 1 < N) q[1] *= s1;
    }
  }
}

// dA_k[m][n] = -2 A_k[m][n] gvsum[k][n] + sum_p q_mu[m][p] gmT[p][n]   (p = k when K > 1).
// Same thread layout as scale_cols_kernel; the gm pairs of up to eight latent GPs stay in registers.
__global__ void __launch_bounds__(256) dA_init_kernel(const double* __restrict__ q_mu,
                                                      const double* __restrict__ gmT,
                                                      const double* __restrict__ A,
                                                      const double* __restrict__ gvsum,
                                                      double* __restrict__ dA, int M, int N, int P,
                                                      int K, int vec) {
  const int k = blockIdx.z;
  const long long total = (long long)M * N;
  const int p0 = (K > 1) ? k : 0, p1 = (K > 1) ? k + 1 : P;
  const int m0 = blockIdx.y * EW_ROWS, m1 = min(M, m0 + EW_ROWS);
  const int n = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (n >= N) return;
  const bool two = vec || (n + 1 < N);
  const double* Ak = A + k * total;
  double* dAk = dA + k * total;
  const double gv0 = gvsum[(long long)k * N + n], gv1 = two ? gvsum[(long long)k * N + n + 1] : 0.0;
  for (int pb = p0; pb < p1; pb += 8) {
    double g0[8], g1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const bool ok = pb + q < p1;
      g0[q] = ok ? gmT[(long long)(pb + q) * N + n] : 0.0;
      g1[q] = (ok && two) ? gmT[(long long)(pb + q) * N + n + 1] : 0.0;
    }
    for (int m = m0; m < m1; ++m) {
      const long long e = (long long)m * N + n;
      double v0, v1;
      if (pb == p0) {
        if (vec) {
          const double2 a = *reinterpret_cast<const double2*>(Ak + e);
          v0 = -2.0 * a.x * gv0;
          v1 = -2.0 * a.y * gv1;
        } else {
          v0 = -2.0 * Ak[e] * gv0;
          v1 = two ? -2.0 * Ak[e + 1] * gv1 : 0.0;
        }
      } else {
        v0 = dAk[e];
        v1 = two ? dAk[e + 1] : 0.0;
      }
      const double* qm = q_mu + (long long)m * P + pb;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (pb + q < p1) {
          const double w = qm[q];
          v0 += w * g0[q];
          v1 += w * g1[q];
        }
      }
      if (vec) {
        *reinterpret_cast<double2*>(dAk + e) = make_double2(v0, v1);
      } else {
        dAk[e] = v0;
        if (two) dAk[e + 1] = v1;
      }
    }
  }
}

// dq_mu[m][p] = sum_n A_k[m][n] gmT[p][n]   (k = p for SeparateIndependent, else 0): the long-N,
// few-output contraction a GEMM tile is wasted on.  grid (ceil(M/4), ceil(N/1024), K): a block owns
// four rows of A_k over one 1024-column segment (A is read exactly once), a thread four columns
// (two 128-bit loads per row) and the gm values of up to eight latent GPs in registers; partial sums
// go to part[segment][k][m][Pw] and dqmu_finish_kernel adds the segments in a fixed order.
constexpr int DQ_SEG = 1024;

// NP = latent GPs handled together (8 for a shared kernel, 1 for SeparateIndependent where k = p);
// ROWS = rows of A_k per block (4 x 8 or 16 x 1 running sums per thread)
template <int NP, int ROWS>
__global__ void __launch_bounds__(256) dqmu_partial_kernel(const double* __restrict__ A,
                                                           const double* __restrict__ gmT,
                                                           double* __restrict__ part, int M, int N,
                                                           int P, int K) {
  __shared__ double red[8][ROWS * NP];
  const int k = blockIdx.z, seg = blockIdx.y, m0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Pw = (K > 1) ? 1 : P;
  const int p0 = (K > 1) ? k : 0, p1 = (K > 1) ? k + 1 : P;
  const int c0 = seg * DQ_SEG + 2 * tid;  // this thread's columns: c0, c0+1, c0+512, c0+513
  const bool ok0 = c0 < N, ok1 = c0 + 512 < N;  // N is even on this path
  const double* Ak = A + (long long)k * M * N;
  for (int pb = p0; pb < p1; pb += NP) {
    double2 g0[NP], g1[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const bool pok = pb + q < p1;
      const double* gp = gmT + (long long)(pb + q) * N;
      g0[q] = (pok && ok0) ? *reinterpret_cast<const double2*>(gp + c0) : make_double2(0.0, 0.0);
      g1[q] = (pok && ok1) ? *reinterpret_cast<const double2*>(gp + c0 + 512) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int r4 = 0; r4 < ROWS; r4 += 4) {
      double2 a0[4], a1[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool rok = m0 + r4 + r < M;
        const double* ar = Ak + (long long)(m0 + r4 + r) * N;
        a0[r] = (rok && ok0) ? *reinterpret_cast<const double2*>(ar + c0) : make_double2(0.0, 0.0);
        a1[r] = (rok && ok1) ? *reinterpret_cast<const double2*>(ar + c0 + 512) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const double v = warp_sum(a0[r].x * g0[q].x + a0[r].y * g0[q].y + a1[r].x * g1[q].x + a1[r].y * g1[q].y);
          if (lane == 0) red[warp][(r4 + r) * NP + q] = v;
        }
    }
    __syncthreads();
    if (tid < ROWS * NP) {
      const int r = tid / NP, q = tid % NP;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][tid];
      if (m0 + r < M && pb + q < p1)
        part[(((long long)seg * K + k) * M + m0 + r) * Pw + (pb - p0) + q] = t;
    }
    __syncthreads();
  }
}

// dq_mu[m][p] = sum_seg part[seg][k][m][.]
__global__ void dqmu_finish_kernel(const double* __restrict__ part, double* __restrict__ dq_mu, int M,
                                   int P, int K, int nseg) {
  const int Pw = (K > 1) ? 1 : P;
  const long long total = (long long)K * M * Pw;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  double s = 0.0;
  for (int g = 0; g < nseg; ++g) s += part[(long long)g * total + t];
  const int k = (int)(t / ((long long)M * Pw));
  const long long rem = t - (long long)k * M * Pw;
  const int m = (int)(rem / Pw), q = (int)(rem - (long long)m * Pw);
  dq_mu[(long long)m * P + ((K > 1) ? k : q)] = s;
}

// one block per row
__global__ void __launch_bounds__(256) row_sums_kernel(const double* __restrict__ x, int N,
                                                       double* __restrict__ out, int accumulate) {
  __shared__ double red[32];
  const double* r = x + (long long)blockIdx.x * N;
  double s = 0.0;
  for (int n = threadIdx.x; n < N; n += 256) s += r[n];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = accumulate ? out[blockIdx.x] + s : s;
}

constexpr int VE_THREADS = 256;
constexpr int VE_PER_BLOCK = 2048;

__global__ void __launch_bounds__(VE_THREADS) gaussian_ve_kernel(
    const double* __restrict__ Fmu, const double* __restrict__ Fvar, const double* __restrict__ Y,
    const double* __restrict__ noise_var, long long n, const double* __restrict__ scale,
    double* __restrict__ dFmu, double* __restrict__ dFvar, double* __restrict__ scratch) {
  __shared__ double red[32];
  const double s2 = noise_var[0];
  const double inv = 1.0 / s2;
  const double c0 = -0.5 * 1.8378770664093453 - 0.5 * log(s2);  // log(2 pi)
  const double sc = scale ? scale[0] : 0.0;
  double ve = 0.0, dn = 0.0;
  const long long base = (long long)blockIdx.x * VE_PER_BLOCK;
  for (int t = threadIdx.x; t < VE_PER_BLOCK; t += VE_THREADS) {
    const long long i = base + t;
    if (i < n) {
      const double r = Y[i] - Fmu[i];
      const double q = r * r + Fvar[i];
      ve += c0 - 0.5 * q * inv;
      if (scale) {
        dFmu[i] = sc * r * inv;
        dFvar[i] = sc * (-0.5 * inv);
        dn += -0.5 * inv + 0.5 * q * inv * inv;
      }
    }
  }
  ve = block_sum(ve, red);
  dn = block_sum(dn, red);
  if (threadIdx.x == 0) {
    scratch[2 * blockIdx.x] = ve;
    scratch[2 * blockIdx.x + 1] = dn;
  }
}

__global__ void __launch_bounds__(256) gaussian_ve_finish_kernel(const double* __restrict__ scratch,
                                                                 int nblk,
                                                                 const double* __restrict__ scale,
                                                                 double* __restrict__ ve_sum,
                                                                 double* __restrict__ dnoise) {
  __shared__ double red[32];
  double ve = 0.0, dn = 0.0;
  for (int e = threadIdx.x; e < nblk; e += 256) {
    ve += scratch[2 * e];
    dn += scratch[2 * e + 1];
  }
  ve = block_sum(ve, red);
  dn = block_sum(dn, red);
  if (threadIdx.x == 0) {
    ve_sum[0] = ve;
    if (scale && dnoise) dnoise[0] = scale[0] * dn;
  }
}

inline int grid_for(long long total, int threads, int max_blocks = 148 * 16) {
  long long b = ceil_div_ll(total, threads);
  if (b < 1) b = 1;
  return (int)(b < max_blocks ? b : max_blocks);
}

}  // namespace

namespace {
__global__ void tril_copy_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                 long long total, int M) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % M);
    const int i = (int)((e / M) % M);
    dst[e] = (j <= i) ? src[e] : 0.0;
  }
}
}  // namespace

int launch_tril_copy(const double* src, double* dst, int P, int M, cudaStream_t stream) {
  const long long total = (long long)P * M * M;
  tril_copy_kernel<<<grid_for(total, 256), 256, 0, stream>>>(src, dst, total, M);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

size_t kl_scratch_doubles(int M, int P) { return (size_t)P * ceil_div(M, KL_ROWS) * 2; }

int launch_kl_white(const double* q_mu, const double* q_sqrt, double* Lq, double* kl,
                    double* scratch, int M, int P, cudaStream_t stream) {
  dim3 grid(ceil_div(M, KL_FWD_ROWS), P);
  const int vec = ((M & 1) == 0) && ((((uintptr_t)q_sqrt | (uintptr_t)Lq) & 15) == 0);
  kl_tril_kernel<<<grid, 256, 0, stream>>>(q_mu, q_sqrt, Lq, scratch, M, P, vec);
  GPFX_LAUNCH_CHECK();
  kl_finish_kernel<<<1, 256, 0, stream>>>(scratch, (int)(grid.x * grid.y), M, P, kl);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_kl_white_bwd(const double* q_mu, const double* Lq, const double* g_kl, double* dq_mu,
                        double* dq_sqrt, int M, int P, int accumulate, cudaStream_t stream) {
  dim3 grid(ceil_div(M, KL_ROWS) + 1, P);
  const int vec = ((M & 1) == 0) && ((((uintptr_t)Lq | (uintptr_t)dq_sqrt) & 15) == 0);
  kl_white_bwd_kernel<<<grid, 256, 0, stream>>>(q_mu, Lq, g_kl, dq_mu, dq_sqrt, M, P, accumulate, vec);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_finalize(const FinalizeArgs& a, cudaStream_t stream) {
  finalize_kernel<<<grid_for((long long)a.N * a.P, 256), 256, 0, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_reparam_sample(const double* mean, const double* var, const double* eps, double* f,
                          long long n, cudaStream_t stream) {
  const int vec = ((((uintptr_t)mean | (uintptr_t)var | (uintptr_t)eps | (uintptr_t)f) & 15) == 0);
  if (vec && n >= (1LL << 20)) {  // >= 8 MB per array: TMA-staged streaming kernel, one CTA per SM
    static SmemAttrOnce once;
    GPFX_TRY(ensure_dyn_smem(reparam_sample_tma_kernel, tma::SMEM, once));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = ((n & ~1LL) + tma::TILE - 1) / tma::TILE;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    reparam_sample_tma_kernel<<<grid, tma::THREADS, tma::SMEM, stream>>>(mean, var, eps, f, n);
    GPFX_LAUNCH_CHECK();
    return GPFX_OK;
  }
  reparam_sample_kernel<<<grid_for(n / 2 + 1, 256, 148 * 8), 256, 0, stream>>>(mean, var, eps, f, n,
                                                                                vec);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

// ---- a8, noise generated in the kernel (Philox4x32-10, counter based) --------------------------
// Not a parity path: the reference's noise comes from TFP's generator.  One counter per PAIR of
// elements: key = seed, counter = (pair index, offset); the four output words give two 53-bit
// uniforms in (0,1) and Box-Muller turns them into elements 2i and 2i+1.  Because the stream is a
// pure function of (seed, offset, index), the backward pass can regenerate eps instead of reading
// it back: the forward then moves 24 B per element (mean, var in; f out) instead of 32.
namespace philox {
__device__ __forceinline__ uint4 round10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 words(unsigned long long seed, unsigned long long offset,
                                       unsigned long long i) {
  return round10(make_uint4((unsigned)i, (unsigned)(i >> 32), (unsigned)offset, (unsigned)(offset >> 32)),
                 make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}
__device__ __forceinline__ double2 normal_pair(uint4 w) {
  const double u1 = ((double)((((unsigned long long)w.y << 32) | w.x) >> 11) + 0.5) * 0x1p-53;
  const double u2 = ((double)((((unsigned long long)w.w << 32) | w.z) >> 11) + 0.5) * 0x1p-53;
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  return make_double2(r * c, r * s);
}
}  // namespace philox

__global__ void philox_raw_kernel(unsigned long long seed, unsigned long long offset,
                                  uint4* __restrict__ out, long long n_pairs) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pairs; i += stride)
    out[i] = philox::words(seed, offset, (unsigned long long)i);
}

// mean == nullptr: out = eps (plain N(0,1) fill).  Otherwise out = mean + sqrt(var) eps and, if
// eps_out != nullptr, eps is stored as well.
__global__ void philox_normal_kernel(unsigned long long seed, unsigned long long offset,
                                     const double* __restrict__ mean, const double* __restrict__ var,
                                     double* __restrict__ out, double* __restrict__ eps_out,
                                     long long n, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_pairs = (n + 1) >> 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
    const double2 e = philox::normal_pair(philox::words(seed, offset, (unsigned long long)i));
    const long long j = 2 * i;
    if (vec && j + 1 < n) {
      double2 f = e;
      if (mean) {
        const double2 m = *reinterpret_cast<const double2*>(mean + j);
        const double2 v = *reinterpret_cast<const double2*>(var + j);
        f = make_double2(m.x + sqrt(v.x) * e.x, m.y + sqrt(v.y) * e.y);
      }
      *reinterpret_cast<double2*>(out + j) = f;
      if (eps_out) *reinterpret_cast<double2*>(eps_out + j) = e;
    } else {
      out[j] = mean ? mean[j] + sqrt(var[j]) * e.x : e.x;
      if (eps_out) eps_out[j] = e.x;
      if (j + 1 < n) {
        out[j + 1] = mean ? mean[j + 1] + sqrt(var[j + 1]) * e.y : e.y;
        if (eps_out) eps_out[j + 1] = e.y;
      }
    }
  }
}

int launch_philox_raw(unsigned long long seed, unsigned long long offset, uint32_t* out,
                      long long n_pairs, cudaStream_t stream) {
  if (n_pairs <= 0) return GPFX_OK;
  if (((uintptr_t)out & 15) != 0) {
    set_last_error("philox_raw: out must be 16-byte aligned");
    return GPFX_E_LAYOUT;
  }
  philox_raw_kernel<<<grid_for(n_pairs, 256, 148 * 8), 256, 0, stream>>>(seed, offset,
                                                                        reinterpret_cast<uint4*>(out), n_pairs);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_philox_normal(unsigned long long seed, unsigned long long offset, const double* mean,
                         const double* var, double* out, double* eps_out, long long n,
                         cudaStream_t stream) {
  if (n <= 0) return GPFX_OK;
  const int vec = ((((uintptr_t)mean | (uintptr_t)var | (uintptr_t)out | (uintptr_t)eps_out) & 15) == 0);
  philox_normal_kernel<<<grid_for((n + 1) / 2, 256, 148 * 8), 256, 0, stream>>>(seed, offset, mean, var, out,
                                                                               eps_out, n, vec);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_bwd_prep(const BwdPrepArgs& a, cudaStream_t stream) {
  bwd_prep_kernel<<<ceil_div(a.N, 128), 128, 0, stream>>>(a);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_scale_cols(double* B, const double* s, int P, int M, int N, cudaStream_t stream) {
  dim3 grid(ceil_div(N, 512), ceil_div(M, EW_ROWS), P);
  const int vec = ((N & 1) == 0) && ((((uintptr_t)B | (uintptr_t)s) & 15) == 0);
  scale_cols_kernel<<<grid, 256, 0, stream>>>(B, s, M, N, vec);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_dA_init(const double* q_mu, const double* gmT, const double* A, const double* gvsum,
                   double* dA, int M, int N, int P, int K, cudaStream_t stream) {
  dim3 grid(ceil_div(N, 512), ceil_div(M, EW_ROWS), K);
  const int vec = ((N & 1) == 0) && ((((uintptr_t)A | (uintptr_t)dA) & 15) == 0);
  dA_init_kernel<<<grid, 256, 0, stream>>>(q_mu, gmT, A, gvsum, dA, M, N, P, K, vec);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

size_t dqmu_scratch_doubles(int M, int N, int P, int K) {
  return (size_t)ceil_div(N, DQ_SEG) * K * M * ((K > 1) ? 1 : P);
}

bool dqmu_supported(const double* A, const double* gmT, int N) {
  return ((N & 1) == 0) && ((((uintptr_t)A | (uintptr_t)gmT) & 15) == 0);
}

int launch_dqmu(const double* A, const double* gmT, double* dq_mu, double* scratch, int M, int N,
                int P, int K, cudaStream_t stream) {
  const int nseg = ceil_div(N, DQ_SEG);
  if (K > 1) {
    dim3 grid(ceil_div(M, 16), nseg, K);
    dqmu_partial_kernel<1, 16><<<grid, 256, 0, stream>>>(A, gmT, scratch, M, N, P, K);
  } else {
    dim3 grid(ceil_div(M, 4), nseg, K);
    dqmu_partial_kernel<8, 4><<<grid, 256, 0, stream>>>(A, gmT, scratch, M, N, P, K);
  }
  GPFX_LAUNCH_CHECK();
  const long long total = (long long)K * M * ((K > 1) ? 1 : P);
  dqmu_finish_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(scratch, dq_mu, M, P, K, nseg);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

// dq_mu from the row-dot partials [nseg][K][M] a GEMM epilogue left (K > 1: one latent GP per kernel)
int launch_dqmu_finish(const double* part, double* dq_mu, int M, int P, int K, int nseg, cudaStream_t stream) {
  const long long total = (long long)K * M * ((K > 1) ? 1 : P);
  dqmu_finish_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(part, dq_mu, M, P, K, nseg);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_row_sums(const double* x, int rows, int N, double* out, int accumulate,
                    cudaStream_t stream) {
  row_sums_kernel<<<rows, 256, 0, stream>>>(x, N, out, accumulate);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

namespace {

constexpr int LAT_ROWS = 256;  // rows per block

// grid (ceil(N/256)); one thread per row
__global__ void __launch_bounds__(LAT_ROWS) latent_forward_kernel(
    const double* __restrict__ X, const double* __restrict__ mean, const double* __restrict__ sd,
    int post_stride, const double* __restrict__ pm, const double* __restrict__ ps,
    const double* __restrict__ eps, double* __restrict__ out, double* __restrict__ scratch,
    long long N, int D, int W) {
  __shared__ double red[32];
  const long long n = blockIdx.x * (long long)LAT_ROWS + threadIdx.x;
  double kl = 0.0;
  if (n < N) {
    double* o = out + n * (D + W);
    for (int d = 0; d < D; ++d) o[d] = X[n * D + d];
    for (int w = 0; w < W; ++w) {
      const double m = mean[n * post_stride + w], s = sd[n * post_stride + w];
      const double m0 = pm[w], s0 = ps[w];
      o[D + w] = m + s * eps[n * W + w];
      if (post_stride) {
        const double dm = m - m0;
        kl += log(s0 / s) + (s * s + dm * dm) / (2.0 * s0 * s0) - 0.5;
      }
    }
  }
  kl = block_sum(kl, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = kl;
}

__global__ void __launch_bounds__(256) latent_kl_finish_kernel(const double* __restrict__ scratch,
                                                               int nblk, double* __restrict__ kl) {
  __shared__ double red[32];
  double s = 0.0;
  for (int e = threadIdx.x; e < nblk; e += 256) s += scratch[e];
  s = block_sum(s, red);
  if (threadIdx.x == 0) kl[0] = s;
}

__global__ void latent_backward_kernel(const double* __restrict__ mean, const double* __restrict__ sd,
                                       const double* __restrict__ pm, const double* __restrict__ ps,
                                       const double* __restrict__ eps, const double* __restrict__ g_out,
                                       const double* __restrict__ g_kl, double* __restrict__ dX,
                                       double* __restrict__ dmean, double* __restrict__ dstd,
                                       long long N, int D, int W) {
  const double gk = g_kl ? g_kl[0] : 0.0;
  const long long total = N * (D + W);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / (D + W);
    const int c = (int)(e - n * (D + W));
    const double g = g_out[e];
    if (c < D) {
      dX[n * D + c] = g;
    } else {
      const int w = c - D;
      const double m = mean[n * W + w], s = sd[n * W + w], m0 = pm[w], s0 = ps[w];
      const double is2 = 1.0 / (s0 * s0);
      dmean[n * W + w] = g + gk * (m - m0) * is2;
      dstd[n * W + w] = g * eps[n * W + w] + gk * (s * is2 - 1.0 / s);
    }
  }
}

}  // namespace

size_t latent_scratch_doubles(long long N, int W) {
  (void)W;
  return (size_t)ceil_div_ll(N > 0 ? N : 1, LAT_ROWS);
}

int launch_latent_forward(const double* X, const double* mean, const double* std, int post_stride,
                          const double* prior_mean, const double* prior_std, const double* eps,
                          double* out, double* kl_sum, double* scratch, long long N, int D, int W,
                          cudaStream_t stream) {
  const int nblk = (int)ceil_div_ll(N > 0 ? N : 1, LAT_ROWS);
  latent_forward_kernel<<<nblk, LAT_ROWS, 0, stream>>>(X, mean, std, post_stride, prior_mean,
                                                       prior_std, eps, out, scratch, N, D, W);
  GPFX_LAUNCH_CHECK();
  latent_kl_finish_kernel<<<1, 256, 0, stream>>>(scratch, nblk, kl_sum);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int launch_latent_backward(const double* mean, const double* std, const double* prior_mean,
                           const double* prior_std, const double* eps, const double* g_out,
                           const double* g_kl, double* dX, double* dmean, double* dstd, long long N,
                           int D, int W, cudaStream_t stream) {
  latent_backward_kernel<<<grid_for(N * (D + W), 256), 256, 0, stream>>>(
      mean, std, prior_mean, prior_std, eps, g_out, g_kl, dX, dmean, dstd, N, D, W);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

size_t gaussian_ve_scratch_doubles(long long n) {
  return (size_t)(2 * ceil_div_ll(n > 0 ? n : 1, VE_PER_BLOCK));
}

int launch_gaussian_ve(const double* Fmu, const double* Fvar, const double* Y,
                       const double* noise_var, long long n, double* ve_sum, const double* scale,
                       double* dFmu, double* dFvar, double* dnoise, double* scratch,
                       cudaStream_t stream) {
  const int nblk = (int)ceil_div_ll(n > 0 ? n : 1, VE_PER_BLOCK);
  gaussian_ve_kernel<<<nblk, VE_THREADS, 0, stream>>>(Fmu, Fvar, Y, noise_var, n, scale, dFmu,
                                                      dFvar, scratch);
  GPFX_LAUNCH_CHECK();
  gaussian_ve_finish_kernel<<<1, 256, 0, stream>>>(scratch, nblk, scale, ve_sum, dnoise);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

}  // namespace gpfx
