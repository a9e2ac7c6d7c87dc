// FP64 tensor-core (DMMA.8x8x4) GEMM for sm_100a: cp.async multi-stage smem pipeline, padded
// conflict-free fragment layouts, triangular k-range skipping, k-segment accumulation, split-K,
// fused column sum-of-squares / weighted-column-sum epilogues.
//
// Replaces the tf.linalg.triangular_solve / tf.matmul calls of GPflow's base_conditional that
// GPLayer.predict reaches (reference gpflux/layers/gp_layer.py:257-266)
// and their autodiff transposes (row a14).
#include "gemm_f64.h"

#include "context.h"

#include "../../include/gpfx.h"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <tuple>
#include <vector>

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

// Per-thread descriptor of the 16-byte chunks it stages for one operand.  `rows` index the
// non-contracted dim (m or n), k the contracted dim.
//   TR == false: element (r,k) at g[r*ld + k]  -> s[r*(BK+4) + k]      (k contiguous)
//   TR == true : element (r,k) at g[k*ld + r]  -> s[k*(ROWS+4) + r]    (r contiguous)
// Everything that does not depend on the k-tile is computed once (offsets, row validity); a tile
// that lies fully inside the matrix takes a predicate-free path.  Out-of-range elements are
// zero-filled (cp.async src-size < cp-size).
template <int ROWS, bool TR, int NT>
struct TileLoader {
  static constexpr int TOTAL = TR ? BK * (ROWS / 2) : ROWS * (BK / 2);
  static constexpr int CH = (TOTAL + NT - 1) / NT;
  long long goff[CH];  // element offset of the chunk for k0 = 0
  int soff[CH];        // smem element offset
  int kofs[CH];        // k index of the chunk inside the tile
  int rvalid[CH];      // TR: valid elements (0..2) along r ; !TR: 2 if the row is inside, else 0
  long long kstride;   // element stride of one k step in global memory
  bool interior;       // all rows of this CTA tile are inside the matrix
  bool vec16;

  __device__ __forceinline__ void init(int ld, int r0, int rmax, bool vec, int tid) {
    vec16 = vec;
    kstride = TR ? ld : 1;
    interior = (r0 + ROWS <= rmax);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = tid + i * NT;
      if (!TR) {
        constexpr int CPR = BK / 2;
        const int r = c / CPR, kc = (c % CPR) * 2;
        goff[i] = (long long)(r0 + r) * ld + kc;
        soff[i] = r * (BK + 4) + kc;
        kofs[i] = kc;
        rvalid[i] = (c < TOTAL && r0 + r < rmax) ? 2 : 0;
      } else {
        constexpr int CPR = ROWS / 2;
        const int kk = c / CPR, rc = (c % CPR) * 2;
        goff[i] = (long long)kk * ld + r0 + rc;
        soff[i] = kk * (ROWS + 4) + rc;
        kofs[i] = kk;
        rvalid[i] = (c < TOTAL) ? min(max(rmax - (r0 + rc), 0), 2) : 0;
      }
    }
  }

  // g: operand base (batch/segment applied); k0: first k of the tile; kmax: K
  __device__ __forceinline__ void load(double* s, const double* __restrict__ g, int k0, int kmax,
                                       int tid) const {
    const double* gk = g + (long long)k0 * kstride;
    if (interior && vec16 && k0 + BK <= kmax && (TOTAL % NT == 0)) {
#pragma unroll
      for (int i = 0; i < CH; ++i) cp_async16(s + soff[i], gk + goff[i], 16);
      return;
    }
    const int kleft = kmax - k0;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (TOTAL % NT != 0 && tid + i * NT >= TOTAL) continue;
      int nvalid;
      if (!TR) nvalid = rvalid[i] ? min(max(kleft - kofs[i], 0), 2) : 0;
      else nvalid = (kofs[i] < kleft) ? rvalid[i] : 0;
      const double* src = nvalid ? gk + goff[i] : g;
      double* dst = s + soff[i];
      if (vec16) {
        cp_async16(dst, src, nvalid * 8);
      } else {
        cp_async8(dst, src, nvalid >= 1 ? 8 : 0);
        cp_async8(dst + 1, nvalid >= 2 ? src + 1 : g, nvalid >= 2 ? 8 : 0);
      }
    }
  }
};

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
  // odd warp rows run right to left: the two warps that share a scheduler (warp id mod 4) then sit at
  // mirrored columns, which balances the live 8x8 fragments of a tile on the diagonal of a tril output
  const int wm = warp / WARPS_N;
  const int wn = (wm & 1) ? WARPS_N - 1 - (warp % WARPS_N) : warp % WARPS_N;
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

  TileLoader<BM, TA, NT> ldA;
  TileLoader<BN, !TB, NT> ldB;
  ldA.init(p.lda, m0, p.M, vecA, tid);
  ldB.init(p.ldb, n0, p.N, vecB, tid);

  // running (segment, tile-in-segment) counters of the load stream and of the compute stream
  int l_seg = 0, l_kt = 0, c_kt = 0;
  if (tps > 0) {
    l_seg = t_begin / tps;
    l_kt = t_begin - l_seg * tps;
    c_kt = l_kt;
  }
  auto load_next = [&](int slot) {
    const int k0 = klo + l_kt * BK;
    ldA.load(As + slot * Cfg::A_ELEMS, Ab + l_seg * p.gA, k0, p.K, tid);
    ldB.load(Bs + slot * Cfg::B_ELEMS, Bb + l_seg * p.gB, k0, p.K, tid);
    if (++l_kt == tps) {
      l_kt = 0;
      ++l_seg;
    }
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (t_begin + s < t_end) load_next(s);
    cp_async_commit();
  }

  // warp-level triangular structure: k-ranges in which every fragment of this warp is known zero
  const int wr0 = m0 + wm * WM, wc0 = n0 + wn * WN;
  const bool warp_dead = p.tril_out && (wc0 > wr0 + WM - 1);  // warp tile strictly above diagonal
  int w_klo = 0, w_khi = 0x7fffffff;  // this warp has work only for w_klo <= kk + 3, kk <= w_khi
  if (p.tri & TRI_LOWER_A) w_khi = min(w_khi, wr0 + WM - 1);
  if (p.tri & TRI_UPPER_B) w_khi = min(w_khi, wc0 + WN - 1);
  if (p.tri & TRI_UPPER_A) w_klo = max(w_klo, wr0);
  if (p.tri & TRI_LOWER_B) w_klo = max(w_klo, wc0);
  if (warp_dead) w_khi = -1;

  // tril output, tile on the diagonal: 8x8 fragments strictly above the diagonal are never stored, so
  // their DMMAs are predicated off (bit j of live[i] = fragment (i, j) touches the lower triangle)
  const bool diag_tile = p.tril_out && !dead && (n0 + BN - 1 > m0);
  unsigned live[MI];
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    live[i] = 0u;
#pragma unroll
    for (int j = 0; j < NI; ++j)
      if (!diag_tile || (wr0 + i * 8 + 7 >= wc0 + j * 8)) live[i] |= 1u << j;
  }
  unsigned any_live = 0u;
#pragma unroll
  for (int i = 0; i < MI; ++i) any_live |= live[i];

  int slot_c = 0, slot_l = STAGES - 1;
  for (int t = t_begin; t < t_end; ++t) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (t + STAGES - 1 < t_end) load_next(slot_l);
    cp_async_commit();
    if (++slot_l == STAGES) slot_l = 0;
    const double* as = As + slot_c * Cfg::A_ELEMS;
    const double* bs = Bs + slot_c * Cfg::B_ELEMS;
    if (++slot_c == STAGES) slot_c = 0;
    const int k0t = klo + c_kt * BK;
    if (++c_kt == tps) c_kt = 0;
    // Warp-level skipping makes the executed DMMAs track the triangular (algorithmic) flop count:
    // a k4 step is dropped when every fragment of this warp is known zero there.
    if (diag_tile) {
      if (any_live) {
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
          const int kk = k0t + k4 * 4;
          if (kk + 3 < w_klo || kk > w_khi) continue;
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
            for (int j = 0; j < NI; ++j)
              if (live[i] & (1u << j)) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
    } else if (k0t + 3 >= w_klo && k0t + BK - 4 <= w_khi) {
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
        if (kk + 3 < w_klo || kk > w_khi) continue;
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

// C = alpha * sum_s ws[s] (*colscale) (tril) + beta * C.   grid (M, ceil(N/256), batch): one
// block row per output row, a thread owns two adjacent columns (128-bit partial loads, four splits
// in flight); fixed summation order -> bit-reproducible.
__global__ void __launch_bounds__(128) splitk_reduce_kernel(const GemmArgs p) {
  const long long nb = (long long)p.batch1 * p.batch2;
  const long long MN = (long long)p.M * p.N;
  const int r = blockIdx.x;
  const long long b = blockIdx.z;
  const int c = (blockIdx.y * 128 + threadIdx.x) * 2;
  if (c >= p.N) return;
  const bool two = c + 1 < p.N;
  const bool vec = ((p.N & 1) == 0) && ((((uintptr_t)p.splitws) & 15) == 0);
  const double* src = p.splitws + b * MN + (long long)r * p.N + c;
  const long long ss = nb * MN;
  double s0 = 0.0, s1 = 0.0;
  if (!(p.tril_out && c > r)) {
    if (vec) {
      int k = 0;
      for (; k + 4 <= p.splits; k += 4) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const double2*>(src + (k + u) * ss);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          s0 += v[u].x;
          s1 += v[u].y;
        }
      }
      for (; k < p.splits; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(src + k * ss);
        s0 += v.x;
        s1 += v.y;
      }
    } else {
      for (int k = 0; k < p.splits; ++k) {
        s0 += src[k * ss];
        if (two) s1 += src[k * ss + 1];
      }
    }
  }
  if (p.tril_out && c + 1 > r) s1 = 0.0;
  double v0 = p.alpha * s0, v1 = p.alpha * s1;
  if (r == c) v0 *= p.diag_scale;
  if (r == c + 1) v1 *= p.diag_scale;
  if (p.colscale) {
    v0 *= p.colscale[b * p.sColscale + c];
    if (two) v1 *= p.colscale[b * p.sColscale + c + 1];
  }
  const int b1 = (int)(b / p.batch2), b2 = (int)(b % p.batch2);
  double* dst = p.C + b1 * p.sC1 + b2 * p.sC2 + (long long)r * p.ldc + c;
  if (p.beta != 0.0) {
    v0 += p.beta * dst[0];
    if (two) v1 += p.beta * dst[1];
  }
  dst[0] = v0;
  if (two) dst[1] = v1;
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
    static SmemAttrOnce once;                                                                 \
    GPFX_TRY(ensure_dyn_smem(kern, Cfg::SMEM, once));                                         \
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


// tile configurations
//   0: 128x128, 8 warps (64x32 warp tiles)   - default for large problems
//   1:  64x64,  4 warps (32x32), 2 CTAs/SM   - small / latency-bound problems
//   2: 128x128, 16 warps (32x32)             - more warps per scheduler (tuning experiments)
//   3:  64x256, 8 warps (32x64)              - triangular-A products: 64-row skipping granularity
//   4:  64x128, 4 warps (32x64), 2 CTAs/SM   - as 3, with one CTA's epilogue under the other's mainloop
//   5:  32x32,  4 warps (16x16), 4 CTAs/SM   - tiny latency-bound products (Cholesky panels, L^-1 assembly)
//   6:  32x128, 4 warps (32x32) side by side, 3 stages, 3 CTAs/SM - short triangular products: every
//       warp of a CTA has the same k-range (no barrier stalls behind warps that skip), 32-row granularity
//   7:  32x256, 4 warps (32x64) side by side, 2 stages, 2 CTAs/SM - as 6 with the 64x128 warp tile
//   8: 128x40,  4 warps (32x40) stacked, 2 CTAs/SM - skinny products (N <= 40: the kernel-matrix backward's
//       R [V,1] and R^T [U,1] with D + 2 <= 40 columns), 40 instead of 64 columns of DMMA work per tile
namespace {

struct CfgInfo {
  int bm, bn;
  double eff;    // relative DMMA efficiency of the tile (smaller tiles pay more L2 traffic)
  int slots;     // CTAs resident per SM
  double fixed;  // prologue + epilogue cost in units of one 128x128x16 k-tile
};
const CfgInfo kCfgInfo[9] = {{128, 128, 1.0, 1, 2.5}, {64, 64, 0.8, 2, 1.0}, {128, 128, 1.0, 1, 2.5},
                             {64, 256, 1.0, 1, 2.5},  {64, 128, 0.92, 2, 0.6}, {32, 32, 0.5, 4, 0.4},
                             {32, 128, 1.0, 3, 0.45}, {32, 256, 0.9, 2, 0.6}, {128, 40, 0.8, 2, 1.0}};

// Predicted makespan (in 128x128x16 k-tile units) of the grid on 148 SMs: the CTAs are handed to
// the first free slot in launch order, exactly like the hardware block scheduler does, with the
// per-CTA k-range the kernel will actually run (triangular skipping, dead tiles, split-K).
double predict_makespan(const GemmArgs& a, int cfg) {
  const CfgInfo& ci = kCfgInfo[cfg];
  const int gx = ceil_div(a.N, ci.bn), gy = ceil_div(a.M, ci.bm);
  const long long gz = (long long)a.batch1 * a.batch2 * a.splits;
  // CTAs that actually share an SM: a grid smaller than one wave leaves every CTA alone on its SM
  const long long nctas_all = (long long)gx * gy * gz;
  const int share = (int)std::max<long long>(1, std::min<long long>(ci.slots, ceil_div_ll(nctas_all, 148)));
  const int servers = 148 * share;
  // measured on B200 (profiles/gemm_sweep_r01.json): the 64x128 two-CTA configuration runs the
  // n-contiguous-B products ~8 % faster than 128x128 (one CTA's barriers and epilogue hide under
  // the other's DMMAs) but the k-contiguous-B (NT) products ~15 % slower
  double eff = ci.eff;
  if (cfg == 4) eff = a.transB ? 0.85 : 1.08;
  // 32x128 (measured, profiles/gemm_sweep2_r01.json): same per-tile efficiency as 64x128 on the
  // n-contiguous-B products - its gain is the finer triangular granularity the k-ranges below already
  // account for, plus no barrier stalls behind skipping warps - but slower on the NT split-K products
  if (cfg == 6) eff = a.transB ? 0.75 : 1.08;
  // 32x32 on the k-contiguous-B (NT) M x M = (M x N)(N x M)^T products: measured
  // (profiles/gemm_sweep3_r01.json) 8 % faster than 64x64 on c2's dLq - the 32-row triangular
  // granularity outweighs the smaller tile
  if (cfg == 5 && a.transB) eff = 0.75;
  const double unit = (double)ci.bm * ci.bn / (128.0 * 128.0) / eff * share;
  std::vector<double> cost((size_t)gx * gy);
  double total = 0.0, cmax = 0.0;
  for (int yy = 0; yy < gy; ++yy) {
    const int by = (a.tri & TRI_LOWER_A) ? gy - 1 - yy : yy;
    const int m0 = by * ci.bm;
    for (int bx = 0; bx < gx; ++bx) {
      const int n0 = bx * ci.bn;
      int klo = 0, khi = a.K;
      if (a.tri & TRI_UPPER_A) klo = std::max(klo, m0);
      if (a.tri & TRI_LOWER_B) klo = std::max(klo, n0);
      if (a.tri & TRI_LOWER_A) khi = std::min(khi, m0 + ci.bm);
      if (a.tri & TRI_UPPER_B) khi = std::min(khi, n0 + ci.bn);
      klo = (klo / BK) * BK;
      const bool dead = a.tril_out && (n0 >= m0 + ci.bm);
      const int tps = (khi > klo && !dead) ? ceil_div(khi - klo, BK) : 0;
      const int T = ceil_div(tps * a.nseg, a.splits);
      const double c = (T * unit + ci.fixed * unit);
      cost[(size_t)yy * gx + bx] = c;
      total += c;
      cmax = std::max(cmax, c);
    }
  }
  total *= (double)gz;
  const long long nctas = (long long)gx * gy * gz;
  if (nctas > 60000) return std::max(total / servers, cmax);
  std::priority_queue<double, std::vector<double>, std::greater<double>> free_at;
  for (int i = 0; i < servers; ++i) free_at.push(0.0);
  double makespan = 0.0;
  for (long long z = 0; z < gz; ++z)
    for (size_t i = 0; i < cost.size(); ++i) {
      const double t = free_at.top() + cost[i];
      free_at.pop();
      free_at.push(t);
      makespan = std::max(makespan, t);
    }
  return makespan;
}

}  // namespace

int pick_cfg(const GemmArgs& a) {
  if (a.cfg >= 0) return a.cfg;
  static const int env_default = [] {
    const char* e = getenv("GPFX_GEMM_CFG");
    return e ? atoi(e) : -1;
  }();
  // values >= 100 address the TMA-fed kernel (gemm_f64_tma.cu): 100 = off, 101.. = its variants
  const int fc = forced_gemm_cfg();
  const int forced = (fc >= 100) ? -1 : ((fc >= 0) ? fc : env_default);
  const bool small = a.small_tile || a.M <= 64 || a.N <= 64;
  if (!small && forced >= 0) return forced;
  if (!a.small_tile && forced < 0 && a.N > 32 && a.N <= 40 && a.M >= 512 && a.tri == TRI_NONE && !a.tril_out) return 8;
  // cached cost-model choice per problem shape
  using Key = std::tuple<int, int, int, long long, int, int, int, int>;
  static std::mutex mu;
  static std::map<Key, int> cache;
  const Key key(a.M, a.N, a.K, (long long)a.batch1 * a.batch2, a.nseg, a.tri + 16 * a.transB + 32 * (small ? 1 : 0),
                a.tril_out, a.splits);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int best = 0;
  double best_t = 1e300;
  // latency-bound problems (a handful of CTAs) may prefer 32x32 tiles: 4x the CTAs, each 4x shorter
  static const int kLarge[] = {0, 3, 4, 6, 1, 5};
  static const int kSmall[] = {1, 5};
  const int* cand = small ? kSmall : kLarge;
  const int ncand = small ? 2 : 6;
  for (int ci = 0; ci < ncand; ++ci) {
    const int cfg = cand[ci];
    const double t = predict_makespan(a, cfg);
    if (t < best_t * 0.97) {  // prefer the earlier (larger-tile) candidate on near ties
      best_t = t;
      best = cfg;
    }
  }
  cache[key] = best;
  return best;
}

void gemm_force_cfg(int cfg) { current_context()->forced_cfg = cfg; }

int gemm_block_m(const GemmArgs& a) {
  if (const int bm = gemm_tma_block_m(a)) return bm;
  const int c = pick_cfg(a);
  return (c == 5 || c == 6 || c == 7) ? 32 : ((c == 1 || c == 3 || c == 4) ? 64 : 128);  // 0, 2, 8: 128
}
int gemm_grid_m(const GemmArgs& a) { return ceil_div(a.M, gemm_block_m(a)); }

// ---- per-launch device timing (gpfx_gemm_profile_*; bench.py's roofline leg) ---------------------
namespace {
struct ProfEntry {
  gpfx_gemm_sample s;
  cudaEvent_t e0, e1;
};
struct Prof {
  std::mutex mu;
  bool on = false;
  std::vector<ProfEntry> log;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
};
Prof& prof() {
  static Prof p;
  return p;
}
constexpr size_t kProfMax = 16384;
}  // namespace

void gemm_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(prof().mu);
  prof().on = on != 0;
}

int gemm_profile_read(gpfx_gemm_sample* out, int max) {
  Prof& p = prof();
  std::lock_guard<std::mutex> lock(p.mu);
  int n = 0;
  for (ProfEntry& e : p.log) {
    float ms = 0.f;
    if (cudaEventSynchronize(e.e1) == cudaSuccess && cudaEventElapsedTime(&ms, e.e0, e.e1) == cudaSuccess &&
        out != nullptr && n < max) {
      out[n] = e.s;
      out[n].ms = ms;
      ++n;
    }
    p.pool.emplace_back(e.e0, e.e1);
  }
  cudaGetLastError();
  p.log.clear();
  return n;
}

static int launch_gemm_impl(const GemmArgs& a, int cfg, cudaStream_t stream);

int launch_gemm(const GemmArgs& a_in, cudaStream_t stream) {
  GemmArgs a = a_in;
  if (a.splits < 1) a.splits = 1;
  if (a.M <= 0 || a.N <= 0 || a.batch1 <= 0 || a.batch2 <= 0) return GPFX_OK;
  if (a.splits > 1 && (a.splitws == nullptr || a.sumsq || a.wsum)) {
    set_last_error("launch_gemm: split-K needs a workspace and no fused column reductions");
    return GPFX_E_WORKSPACE;
  }
  if (a.addmat != nullptr && !gemm_addin_supported(a)) {
    set_last_error("launch_gemm: add-in epilogue terms are not available for this problem (gemm_addin_supported)");
    return GPFX_E_UNSUPPORTED;
  }
  if (a.kscale != nullptr && !gemm_kscale_supported(a)) {
    set_last_error("launch_gemm: a per-k scale is not available for this problem (gemm_kscale_supported)");
    return GPFX_E_UNSUPPORTED;
  }
  // large aligned problems take the TMA-fed kernel (profile cfg ids 101 / 102 = its variants); when only
  // the pointers are misaligned the cp.async kernel runs with the same tile height, so column-partial
  // buffers keep their size
  int cfg;
  if (const int bm = gemm_tma_block_m(a)) cfg = gemm_tma_pointers_ok(a) ? (bm == 64 ? 102 : 101) : (bm == 64 ? 4 : 0);
  else cfg = pick_cfg(a);
  Prof& p = prof();
  if (!p.on) return launch_gemm_impl(a, cfg, stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
    return launch_gemm_impl(a, cfg, stream);
  ProfEntry e;
  {
    std::lock_guard<std::mutex> lock(p.mu);
    if (p.log.size() >= kProfMax) return launch_gemm_impl(a, cfg, stream);
    if (!p.pool.empty()) {
      e.e0 = p.pool.back().first;
      e.e1 = p.pool.back().second;
      p.pool.pop_back();
    } else {
      GPFX_CUDA_CHECK(cudaEventCreate(&e.e0));
      GPFX_CUDA_CHECK(cudaEventCreate(&e.e1));
    }
  }
  e.s = gpfx_gemm_sample{a.M, a.N, a.K, a.batch1 * a.batch2, a.nseg, a.tri, a.transA, a.transB,
                         a.tril_out, a.splits, cfg, 0, 0.0};
  GPFX_CUDA_CHECK(cudaEventRecord(e.e0, stream));
  const int st = launch_gemm_impl(a, cfg, stream);
  GPFX_CUDA_CHECK(cudaEventRecord(e.e1, stream));
  std::lock_guard<std::mutex> lock(p.mu);
  p.log.push_back(e);
  return st;
}

static int launch_gemm_impl(const GemmArgs& a, int cfg, cudaStream_t stream) {
  int st;
  switch (cfg) {
    case 101:
    case 102: st = launch_gemm_tma(a, stream); break;
    case 1: st = launch_cfg<64, 64, 2, 2, 4, 2>(a, stream); break;
    case 2: st = launch_cfg<128, 128, 4, 4, 4, 1>(a, stream); break;
    case 3: st = launch_cfg<64, 256, 2, 4, 4, 1>(a, stream); break;
    case 4: st = launch_cfg<64, 128, 2, 2, 4, 2>(a, stream); break;
    case 5: st = launch_cfg<32, 32, 2, 2, 4, 4>(a, stream); break;
    case 6: st = launch_cfg<32, 128, 1, 4, 3, 3>(a, stream); break;
    case 7: st = launch_cfg<32, 256, 1, 4, 2, 2>(a, stream); break;
    case 8: st = launch_cfg<128, 40, 4, 1, 4, 2>(a, stream); break;
    default: st = launch_cfg<128, 128, 2, 4, 4, 1>(a, stream); break;
  }
  if (st != GPFX_OK) return st;
  if (a.splits > 1) {
    dim3 rgrid(a.M, ceil_div(a.N, 256), a.batch1 * a.batch2);
    splitk_reduce_kernel<<<rgrid, 128, 0, stream>>>(a);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

}  // namespace gpfx
