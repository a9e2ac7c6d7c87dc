// FP64 tensor-core (DMMA.8x8x4) GEMM for sm_100a: cp.async multi-stage smem pipeline, padded
// conflict-free fragment layouts, triangular k-range skipping, k-segment accumulation, split-K,
// fused column sum-of-squares / weighted-column-sum epilogues.
//
// Replaces the tf.linalg.triangular_solve / tf.matmul calls of GPflow's base_conditional that
// GPLayer.predict reaches (reference gpflux/layers/gp_layer.py:257-266)
// and their autodiff transposes (row a14).
#include "gemm_f64.h"

#include <algorithm>

namespace gpfx {

namespace {

constexpr int BK = 16;

template <int BM, int BN, int WARPS_M, int WARPS_N, bool TA, bool TB, int STAGES>
struct GemmCfg {
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int MI = WM / 8, NI = WN / 8;
  static constexpr int LDA_S = TA ? (BM + 4) : (BK + 4);
  static constexpr int LDB_S = TB ? (BK + 4) : (BN + 4);
  static constexpr int A_ELEMS = TA ? BK * LDA_S : BM * LDA_S;
  static constexpr int B_ELEMS = TB ? BN * LDB_S : BK * LDB_S;
  static constexpr size_t SMEM = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(double);
};

// Stage one operand tile.  `rows` index the non-contracted dim (m or n), k the contracted dim.
//   TR == false: element (r,k) at g[r*ld + k]  -> s[r*(BK+4) + k]      (k contiguous)
//   TR == true : element (r,k) at g[k*ld + r]  -> s[k*(ROWS+4) + r]    (r contiguous)
// Out-of-range elements are zero-filled (cp.async src-size < cp-size).
template <int ROWS, bool TR, int NT>
__device__ __forceinline__ void stage_tile(double* s, const double* __restrict__ g, int ld, int r0,
                                           int k0, int rmax, int kmax, bool vec16, int tid) {
  if (!TR) {
    constexpr int CPR = BK / 2;  // 16B chunks per row
    constexpr int TOTAL = ROWS * CPR;
#pragma unroll
    for (int c = tid; c < TOTAL; c += NT) {
      const int r = c / CPR, kc = (c % CPR) * 2;
      const int gr = r0 + r, gk = k0 + kc;
      double* dst = s + r * (BK + 4) + kc;
      int nvalid = (gr < rmax) ? min(max(kmax - gk, 0), 2) : 0;
      const double* src = nvalid ? (g + (long long)gr * ld + gk) : g;
      if (vec16) {
        cp_async16(dst, src, nvalid * 8);
      } else {
        cp_async8(dst, src, nvalid >= 1 ? 8 : 0);
        cp_async8(dst + 1, nvalid >= 2 ? src + 1 : g, nvalid >= 2 ? 8 : 0);
      }
    }
  } else {
    constexpr int CPR = ROWS / 2;
    constexpr int TOTAL = BK * CPR;
#pragma unroll
    for (int c = tid; c < TOTAL; c += NT) {
      const int kk = c / CPR, rc = (c % CPR) * 2;
      const int gk = k0 + kk, gr = r0 + rc;
      double* dst = s + kk * (ROWS + 4) + rc;
      int nvalid = (gk < kmax) ? min(max(rmax - gr, 0), 2) : 0;
      const double* src = nvalid ? (g + (long long)gk * ld + gr) : g;
      if (vec16) {
        cp_async16(dst, src, nvalid * 8);
      } else {
        cp_async8(dst, src, nvalid >= 1 ? 8 : 0);
        cp_async8(dst + 1, nvalid >= 2 ? src + 1 : g, nvalid >= 2 ? 8 : 0);
      }
    }
  }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool TA, bool TB, int STAGES, int MINB>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB)
    gemm_f64_kernel(const GemmArgs p, const bool vecA, const bool vecB) {
  using Cfg = GemmCfg<BM, BN, WARPS_M, WARPS_N, TA, TB, STAGES>;
  constexpr int NT = Cfg::NT, WM = Cfg::WM, WN = Cfg::WN, MI = Cfg::MI, NI = Cfg::NI;
  constexpr int LDA_S = Cfg::LDA_S, LDB_S = Cfg::LDB_S;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_ELEMS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int lr = lane >> 2, lc = lane & 3;
  const int bz = blockIdx.z;
  const int split = bz % p.splits, b = bz / p.splits;
  const int b1 = b / p.batch2, b2 = b % p.batch2;
  int by = blockIdx.y;
  if (p.tri & TRI_LOWER_A) by = gridDim.y - 1 - by;  // longest k-ranges first
  const int m0 = by * BM, n0 = blockIdx.x * BN;
  const double* Ab = p.A + b1 * p.sA1 + b2 * p.sA2;
  const double* Bb = p.B + b1 * p.sB1 + b2 * p.sB2;

  int klo = 0, khi = p.K;
  if (p.tri & TRI_UPPER_A) klo = max(klo, m0);
  if (p.tri & TRI_LOWER_B) klo = max(klo, n0);
  if (p.tri & TRI_LOWER_A) khi = min(khi, m0 + BM);
  if (p.tri & TRI_UPPER_B) khi = min(khi, n0 + BN);
  klo = (klo / BK) * BK;
  const bool dead = p.tril_out && (n0 >= m0 + BM);
  const int tps = (khi > klo && !dead) ? (khi - klo + BK - 1) / BK : 0;
  const int T = tps * p.nseg;
  int t_begin = 0, t_end = T;
  if (p.splits > 1) {
    const int cps = (T + p.splits - 1) / p.splits;
    t_begin = min(T, split * cps);
    t_end = min(T, t_begin + cps);
  }

  auto load = [&](int t, int slot) {
    const int seg = t / tps;
    const int k0 = klo + (t - seg * tps) * BK;
    const double* Ag = Ab + seg * p.gA;
    const double* Bg = Bb + seg * p.gB;
    stage_tile<BM, TA, NT>(As + slot * Cfg::A_ELEMS, Ag, p.lda, m0, k0, p.M, p.K, vecA, tid);
    stage_tile<BN, !TB, NT>(Bs + slot * Cfg::B_ELEMS, Bg, p.ldb, n0, k0, p.N, p.K, vecB, tid);
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (t_begin + s < t_end) load(t_begin + s, s);
    cp_async_commit();
  }

  for (int t = t_begin; t < t_end; ++t) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int tn = t + STAGES - 1;
      if (tn < t_end) load(tn, (tn - t_begin) % STAGES);
      cp_async_commit();
    }
    const double* as = As + ((t - t_begin) % STAGES) * Cfg::A_ELEMS;
    const double* bs = Bs + ((t - t_begin) % STAGES) * Cfg::B_ELEMS;
    // Does this k-slice cross a known-zero triangle inside this warp's rows / columns?  Then whole
    // 8x4 / 4x8 fragments are zero and their DMMAs are skipped (warp-uniform predicates), which
    // makes the executed work track the triangular (algorithmic) flop count.
    const int k0t = klo + (t % tps) * BK;
    const int wr0 = m0 + wm * WM, wc0 = n0 + wn * WN;
    const bool edge = ((p.tri & TRI_LOWER_A) && k0t + BK - 1 > wr0) ||
                      ((p.tri & TRI_UPPER_A) && k0t < wr0 + WM - 1) ||
                      ((p.tri & TRI_LOWER_B) && k0t < wc0 + WN - 1) ||
                      ((p.tri & TRI_UPPER_B) && k0t + BK - 1 > wc0) ||
                      (p.tril_out && wc0 + WN - 1 > wr0);  // warp tile reaches above the diagonal
    if (!edge) {
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        const int k = k4 * 4 + lc;
        double af[MI], bf[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int r = wm * WM + i * 8 + lr;
          af[i] = TA ? as[k * LDA_S + r] : as[r * LDA_S + k];
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int c = wn * WN + j * 8 + lr;
          bf[j] = TB ? bs[c * LDB_S + k] : bs[k * LDB_S + c];
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    } else {
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        const int kk = k0t + k4 * 4;
        unsigned amask = 0, bmask = 0;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int r0 = wr0 + i * 8;
          const bool off = ((p.tri & TRI_LOWER_A) && kk > r0 + 7) || ((p.tri & TRI_UPPER_A) && kk + 3 < r0);
          amask |= (off ? 0u : 1u) << i;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int c0 = wc0 + j * 8;
          const bool off = ((p.tri & TRI_LOWER_B) && kk + 3 < c0) || ((p.tri & TRI_UPPER_B) && kk > c0 + 7);
          bmask |= (off ? 0u : 1u) << j;
        }
        if (amask == 0 || bmask == 0) continue;
        const int k = k4 * 4 + lc;
        double af[MI], bf[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int r = wm * WM + i * 8 + lr;
          af[i] = TA ? as[k * LDA_S + r] : as[r * LDA_S + k];
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int c = wn * WN + j * 8 + lr;
          bf[j] = TB ? bs[c * LDB_S + k] : bs[k * LDB_S + c];
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          if (!((amask >> i) & 1u)) continue;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            // tril_out: an 8x8 output fragment strictly above the diagonal is never stored
            const bool above = p.tril_out && (wc0 + j * 8 > wr0 + i * 8 + 7);
            if (((bmask >> j) & 1u) && !above) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  // ------------------------------------------------------------------ epilogue
  if (p.splits > 1) {
    const long long nb = (long long)p.batch1 * p.batch2;
    double* W = p.splitws + ((long long)split * nb + b) * (long long)p.M * p.N;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int r = m0 + wm * WM + i * 8 + lr;
      if (r >= p.M) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int c = n0 + wn * WN + j * 8 + 2 * lc;
        if (c < p.N) W[(long long)r * p.N + c] = acc[i][j][0];
        if (c + 1 < p.N) W[(long long)r * p.N + c + 1] = acc[i][j][1];
      }
    }
    return;
  }

  double* Cb = p.C + b1 * p.sC1 + b2 * p.sC2;
  const double* cs = p.colscale ? p.colscale + b * p.sColscale : nullptr;
  const bool vecC = ((p.ldc & 1) == 0) && ((((uintptr_t)Cb) & 15) == 0);
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int c = n0 + wn * WN + j * 8 + 2 * lc;
    double s0 = 1.0, s1 = 1.0;
    if (cs) {
      s0 = (c < p.N) ? cs[c] : 0.0;
      s1 = (c + 1 < p.N) ? cs[c + 1] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int r = m0 + wm * WM + i * 8 + lr;
      double v0 = p.alpha * acc[i][j][0] * s0;
      double v1 = p.alpha * acc[i][j][1] * s1;
      if (p.tril_out) {
        if (c > r) v0 = 0.0;
        if (c + 1 > r) v1 = 0.0;
      }
      if (c == r) v0 *= p.diag_scale;
      if (c + 1 == r) v1 *= p.diag_scale;
      if (r < p.M) {
        double* dst = Cb + (long long)r * p.ldc + c;
        if (c + 1 < p.N && vecC) {
          if (p.beta != 0.0) {
            const double2 o = *reinterpret_cast<const double2*>(dst);
            v0 += p.beta * o.x;
            v1 += p.beta * o.y;
          }
          *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        } else {
          if (c < p.N) {
            if (p.beta != 0.0) v0 += p.beta * dst[0];
            dst[0] = v0;
          } else {
            v0 = 0.0;
          }
          if (c + 1 < p.N) {
            if (p.beta != 0.0) v1 += p.beta * dst[1];
            dst[1] = v1;
          } else {
            v1 = 0.0;
          }
        }
      } else {
        v0 = v1 = 0.0;
      }
      acc[i][j][0] = v0;
      acc[i][j][1] = v1;
    }
  }

  if (p.sumsq == nullptr && p.wsum == nullptr) return;
  __syncthreads();  // pipeline smem is free now
  double* red = smem;  // [WARPS_M][BN]

  if (p.sumsq) {
#pragma unroll
    for (int j = 0; j < NI; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < MI; ++i) s += acc[i][j][e] * acc[i][j][e];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (lr == 0) red[wm * BN + wn * WN + j * 8 + 2 * lc + e] = s;
      }
    }
    __syncthreads();
    for (int c = tid; c < BN; c += NT) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < WARPS_M; ++w) tot += red[w * BN + c];
      if (n0 + c < p.N) p.sumsq[b * p.sSumsq + (long long)by * p.N + n0 + c] = tot;
    }
    __syncthreads();
  }

  if (p.wsum) {
    const double* wv = p.wvec + b * p.sW;
    for (int q = 0; q < p.nw; ++q) {
      double wgt[MI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = m0 + wm * WM + i * 8 + lr;
        wgt[i] = (r < p.M) ? wv[(long long)r * p.ldw + q] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < MI; ++i) s += acc[i][j][e] * wgt[i];
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          if (lr == 0) red[wm * BN + wn * WN + j * 8 + 2 * lc + e] = s;
        }
      }
      __syncthreads();
      for (int c = tid; c < BN; c += NT) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS_M; ++w) tot += red[w * BN + c];
        if (n0 + c < p.N)
          p.wsum[b * p.sWsum + ((long long)by * p.nw + q) * p.N + n0 + c] = tot;
      }
      __syncthreads();
    }
  }
}

// C = alpha * sum_s ws[s] (*colscale) (tril) + beta * C
__global__ void splitk_reduce_kernel(const GemmArgs p) {
  const long long nb = (long long)p.batch1 * p.batch2;
  const long long MN = (long long)p.M * p.N;
  const long long total = nb * MN;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / MN;
    const long long rem = i - b * MN;
    const int r = (int)(rem / p.N), c = (int)(rem - (long long)r * p.N);
    double s = 0.0;
    if (!(p.tril_out && c > r))
      for (int k = 0; k < p.splits; ++k) s += p.splitws[(k * nb + b) * MN + rem];
    double v = p.alpha * s;
    if (r == c) v *= p.diag_scale;
    if (p.colscale) v *= p.colscale[b * p.sColscale + c];
    const int b1 = (int)(b / p.batch2), b2 = (int)(b % p.batch2);
    double* dst = p.C + b1 * p.sC1 + b2 * p.sC2 + (long long)r * p.ldc + c;
    if (p.beta != 0.0) v += p.beta * *dst;
    *dst = v;
  }
}

bool use_small(const GemmArgs& a) {
  if (a.small_tile) return true;
  if (a.M <= 64 || a.N <= 64) return true;
  // 128x128 tiles halve the L2->smem traffic per flop; fall back to 64x64 only when the big
  // tiles could not occupy most of the 148 SMs
  const long long big_tiles = (long long)ceil_div(a.M, 128) * ceil_div(a.N, 128) * a.batch1 *
                              a.batch2 * (a.splits > 0 ? a.splits : 1);
  return big_tiles < 96;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int MINB>
int launch_cfg(const GemmArgs& a, cudaStream_t stream) {
  auto aligned = [](const void* p, int ld, long long s1, long long s2, long long g) {
    return (((uintptr_t)p) & 15) == 0 && (ld & 1) == 0 && (s1 & 1) == 0 && (s2 & 1) == 0 &&
           (g & 1) == 0;
  };
  const bool vecA = aligned(a.A, a.lda, a.sA1, a.sA2, a.gA);
  const bool vecB = aligned(a.B, a.ldb, a.sB1, a.sB2, a.gB);
  dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM), a.batch1 * a.batch2 * a.splits);
  dim3 block(WARPS_M * WARPS_N * 32);
#define GPFX_GEMM_LAUNCH(TA, TB)                                                              \
  do {                                                                                        \
    using Cfg = GemmCfg<BM, BN, WARPS_M, WARPS_N, TA, TB, STAGES>;                            \
    auto kern = gemm_f64_kernel<BM, BN, WARPS_M, WARPS_N, TA, TB, STAGES, MINB>;              \
    static bool attr_set = false;                                                             \
    if (!attr_set) {                                                                          \
      GPFX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)Cfg::SMEM));                                  \
      attr_set = true;                                                                        \
    }                                                                                         \
    kern<<<grid, block, Cfg::SMEM, stream>>>(a, vecA, vecB);                                  \
  } while (0)
  if (a.transA) {
    if (a.transB) GPFX_GEMM_LAUNCH(true, true);
    else GPFX_GEMM_LAUNCH(true, false);
  } else {
    if (a.transB) GPFX_GEMM_LAUNCH(false, true);
    else GPFX_GEMM_LAUNCH(false, false);
  }
#undef GPFX_GEMM_LAUNCH
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

}  // namespace

int gemm_block_m(const GemmArgs& a) { return use_small(a) ? 64 : 128; }
int gemm_grid_m(const GemmArgs& a) { return ceil_div(a.M, gemm_block_m(a)); }

int launch_gemm(const GemmArgs& a_in, cudaStream_t stream) {
  GemmArgs a = a_in;
  if (a.splits < 1) a.splits = 1;
  if (a.M <= 0 || a.N <= 0 || a.batch1 <= 0 || a.batch2 <= 0) return GPFX_OK;
  if (a.splits > 1 && (a.splitws == nullptr || a.sumsq || a.wsum)) {
    set_last_error("launch_gemm: split-K needs a workspace and no fused column reductions");
    return GPFX_E_WORKSPACE;
  }
  int st;
  if (use_small(a)) st = launch_cfg<64, 64, 2, 2, 4, 2>(a, stream);
  else st = launch_cfg<128, 128, 2, 4, 4, 1>(a, stream);
  if (st != GPFX_OK) return st;
  if (a.splits > 1) {
    const long long total = (long long)a.batch1 * a.batch2 * a.M * a.N;
    const int threads = 256;
    const int blocks = (int)std::min<long long>(ceil_div_ll(total, threads), 148 * 8);
    splitk_reduce_kernel<<<blocks, threads, 0, stream>>>(a);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace gpfx
