// Batched FP64 Cholesky + triangular inverse of the (padded) Kuu stack and the Cholesky backward.
//
// Replaces tf.linalg.cholesky / tf.linalg.triangular_solve in GPflow's base_conditional as reached
// from GPLayer.predict (reference gpflux/layers/gp_layer.py:257-266) and
// gpflux/math.py:25-62.  Design: 64x64 diagonal blocks are factorised and inverted inside one CTA's
// shared memory; everything off the diagonal is DMMA GEMM work (left-looking panel updates,
// recursive-doubling assembly of L^-1), batched over the K kernels.  Matrices are padded to
// Mp = 64 * 2^q with an identity block so every level is uniform.
#include "chol.h"

#include "gemm_f64.h"

#include <cstdlib>

namespace gpfx {
namespace {

constexpr int NB = 64;

// 1/sqrt(d) for the pivot of each column step.  The pivot sits on the factorisation's serial
// dependency chain (update -> rsqrt -> scale -> publish -> update), where libdevice's rsqrt()
// (special-case handling, ~20 dependent FP64 ops) costs more than the rest of the step.  Hardware
// seed (MUFU.RSQ64H, ~2^-22 relative) + two Newton-Raphson steps: error ~2^-22 -> 2^-43 -> below one
// ulp of the iteration's own rounding.  Non-positive / non-finite pivots are reported through info
// by the caller (the value returned for them is irrelevant).
__device__ __forceinline__ double pivot_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  double e = fma(-(h * y), y, 0.5);
  y = fma(y, e, y);
  e = fma(-(h * y), y, 0.5);
  y = fma(y, e, y);
  return y;
}

// 16 column steps c = 16*CB + cm of the register-resident 64x64 factorisation (see
// potrf_diag_kernel).  CB is a template parameter so that every comparison between an element's
// 16-block index (a, b) and the pivot's block is resolved at compile time; only five per-thread
// predicates on (tr, tc) vs cm remain at run time.
template <int CB>
__device__ __forceinline__ void diag_steps(double (&S)[4][4], double (&X)[4][4],
                                           double (*colS)[NB], double (*rowX)[NB], int tr, int tc,
                                           int j0, int* info_k) {
  const bool tc_le_tr = tc <= tr;
#pragma unroll 1
  for (int cm = 0; cm < 16; ++cm) {
    const int c = CB * 16 + cm;
    const int buf = cm & 1;
    if (tc == cm) {
#pragma unroll
      for (int a = 0; a < 4; ++a) colS[buf][tr + 16 * a] = S[a][CB];
    }
    if (tr == cm) {
#pragma unroll
      for (int b = 0; b < 4; ++b) rowX[buf][tc + 16 * b] = X[CB][b];
    }
    __syncthreads();
    const double d = colS[buf][c];
    if (!(d > 0.0) && threadIdx.x == 0) atomicCAS(info_k, 0, j0 + c + 1);
    const double inv = pivot_rsqrt(d);
    double lr[4], lc[4], xr[4];
#pragma unroll
    for (int a = CB; a < 4; ++a) lr[a] = colS[buf][tr + 16 * a] * inv;  // L[r][c], my rows
#pragma unroll
    for (int b = CB; b < 4; ++b) lc[b] = colS[buf][tc + 16 * b] * inv;  // L[cc][c], my columns
#pragma unroll
    for (int b = 0; b <= CB; ++b) xr[b] = rowX[buf][tc + 16 * b] * inv;  // X[c][j], final
    const bool tr_gt = tr > cm, tr_eq = tr == cm;
    const bool tc_gt = tc > cm, tc_eq = tc == cm, tc_le = tc <= cm;
#pragma unroll
    for (int a = CB; a < 4; ++a) {
      const bool row_gt = (a > CB) ? true : tr_gt;   // r > c
      const bool row_eq = (a == CB) ? tr_eq : false;  // r == c
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (b >= CB) {  // columns that can lie right of the pivot: S -= l l^T on c < cc <= r
          const bool col_gt = (b > CB) ? true : tc_gt;
          const bool cc_le_r = (b < a) ? true : ((b == a) ? tc_le_tr : false);
          if (row_gt && col_gt && cc_le_r) S[a][b] -= lr[a] * lc[b];
        }
        if (b <= CB) {  // columns j <= c: X -= l x_c^T below the pivot row, X[c][:] final
          const bool col_le = (b < CB) ? true : tc_le;
          if (row_gt && col_le) X[a][b] -= lr[a] * xr[b];
          if (row_eq && col_le) X[a][b] = xr[b];
        }
        if (b == CB) {  // column c itself becomes final (lr = d * rsqrt(d) = l on the diagonal)
          if ((row_gt || row_eq) && tc_eq) S[a][b] = lr[a];
        }
      }
    }
  }
}

// One CTA (256 threads) per matrix k.  In place: L[j0:j0+NB, j0:j0+NB] <- chol(lower part) (zero
// above the diagonal); Linv diagonal block <- inverse of the factor.  Kept as the reference
// implementation of the diagonal block (GPFX_POTRF_OLD=1); the panel kernel below is the default.
//
// The 64x64 block lives in REGISTERS: thread (tr, tc) of a 16x16 grid owns the cyclic 4x4
// sub-block rows tr+16a, cols tc+16b (cyclic keeps every thread busy as the active window
// shrinks).  Right-looking Cholesky and the forward substitution X = L^-1 advance together, one
// column per step, with ONE __syncthreads per step: the owners of column c of S and of row c of X
// publish them through double-buffered shared memory, then everybody applies the two rank-1
// updates  S -= l l^T  and  X -= l x_c^T  to its registers.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ L,
                                                         double* __restrict__ Linv, int Mp,
                                                         long long sL, int j0, int* info) {
  __shared__ double colS[2][NB];
  __shared__ double rowX[2][NB];
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tc = tid & 15;
  double* Lk = L + blockIdx.x * sL + (long long)j0 * Mp + j0;
  double* Xk = Linv + blockIdx.x * sL + (long long)j0 * Mp + j0;

  if (j0 == 0 && tid == 0) info[blockIdx.x] = 0;  // only thread 0 of this CTA touches info[k]
  double S[4][4], X[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = tr + 16 * a, c = tc + 16 * b;
      S[a][b] = (c <= r) ? Lk[(long long)r * Mp + c] : 0.0;
      X[a][b] = (r == c) ? 1.0 : 0.0;
    }

  diag_steps<0>(S, X, colS, rowX, tr, tc, j0, info + blockIdx.x);
  diag_steps<1>(S, X, colS, rowX, tr, tc, j0, info + blockIdx.x);
  diag_steps<2>(S, X, colS, rowX, tr, tc, j0, info + blockIdx.x);
  diag_steps<3>(S, X, colS, rowX, tr, tc, j0, info + blockIdx.x);

#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = tr + 16 * a, c = tc + 16 * b;
      Lk[(long long)r * Mp + c] = (c <= r) ? S[a][b] : 0.0;
      Xk[(long long)r * Mp + c] = (c <= r) ? X[a][b] : 0.0;
    }

}

// ---------------------------------------------------------------------------------------------
// Panel-blocked 64x64 diagonal-block kernel (potrf + inverse), shared-memory resident.
//
// The register-resident kernel above applies BOTH rank-1 updates (S -= l l^T and X -= l x^T) to
// the whole 64x64 block at every column: ~170 instructions per thread per column, ~900 cycles per
// column (measured), i.e. issue-bound.  Here the block is factorised in four 16-column panels:
//   * inside a panel the same publish -> barrier -> scale -> rank-1 scheme runs, but only on the
//     panel's 16 columns (thread (row, quad) owns 4 entries): ~40 instructions per column step, so
//     the step costs little more than its dependency chain (publish/barrier/read ~100 cycles +
//     MUFU.RSQ64H and two Newton steps ~90 + scale + update);
//   * the trailing block takes one rank-16 update per panel on the FP64 tensor pipe
//     (S22 -= L21 L21^T, 8x8 tiles, four DMMA each);
//   * the inverse is formed afterwards: 16x16 diagonal inverses by column-wise forward
//     substitution (64 independent columns, one thread each), then two levels of recursive doubling
//     X21 = -X22 (L21 X11) on DMMA tiles.
// Numerically this is the standard blocked right-looking Cholesky + blocked triangular inverse.
constexpr int PB = 16;         // panel width
constexpr int LDS64 = NB + 4;  // shared-memory leading dimension (== 4 mod 16: conflict-free fragments)

// c += sign * A[r0.., ka..ka+kn) * B on one 8x8 tile, DMMA fragments straight from shared memory.
//   B_IS_T: B[k][n] = Bm[n0 + n][kb + k]   else   B[k][n] = Bm[kb + k][n0 + n]
template <bool B_IS_T>
__device__ __forceinline__ void tile_mma(double& c0, double& c1, const double* Am, int r0, int ka,
                                         const double* Bm, int n0, int kb, int kn, double sign,
                                         int lane) {
  const int lr = lane >> 2, lc = lane & 3;
  for (int k4 = 0; k4 < kn; k4 += 4) {
    const double a = sign * Am[(r0 + lr) * LDS64 + ka + k4 + lc];
    const double b = B_IS_T ? Bm[(n0 + lr) * LDS64 + kb + k4 + lc] : Bm[(kb + k4 + lc) * LDS64 + n0 + lr];
    dmma884(c0, c1, a, b);
  }
}

// Factorises the 64x64 block held in S (lower part; entries above the diagonal are ignored) in
// place and builds its inverse in X (which must be zero on entry).  All 256 threads of the CTA take
// part; S and X are consistent for every thread on return.  info_k may be null (no status report).
__device__ __forceinline__ void factor_and_invert_smem(double* S, double* X, double* T, double* colS,
                                                       double* dinv, int* info_k, int j0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int r = tid >> 2, q = tid & 3;  // thread (row, quad): entries [r][c0 + 4q .. c0 + 4q + 3] of a panel
#pragma unroll 1
  for (int kb = 0; kb < NB / PB; ++kb) {
    const int c0 = kb * PB;
    // ---- panel factorisation: 16 column steps on rows c0..63 ---------------------------------
    double P[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) P[u] = S[r * LDS64 + c0 + 4 * q + u];
#pragma unroll
    for (int j = 0; j < PB; ++j) {
      const int buf = j & 1;
      if (q == (j >> 2)) colS[buf * NB + r] = P[j & 3];  // column c0 + j of every row
      __syncthreads();
      const double d = colS[buf * NB + c0 + j];
      if (!(d > 0.0) && tid == 0 && info_k) atomicCAS(info_k, 0, j0 + c0 + j + 1);
      const double inv = pivot_rsqrt(d);
      const double l_r = colS[buf * NB + r] * inv;  // L[r][c0 + j] for r > c0 + j; sqrt(d) on the diagonal
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = 4 * q + u;  // panel column of P[u]
        if (c == j) {
          P[u] = l_r;
        } else if (c > j) {
          const double l_c = colS[buf * NB + c0 + c] * inv;  // L[c0 + c][c0 + j]
          P[u] -= l_r * l_c;
        }
      }
      if (tid == 0) dinv[c0 + j] = inv;
    }
    // write the panel back (entries above the diagonal are never read; rows above the panel idle)
    if (r >= c0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) S[r * LDS64 + c0 + 4 * q + u] = (c0 + 4 * q + u <= r) ? P[u] : 0.0;
    }
    __syncthreads();
    // ---- trailing update S22 -= L21 L21^T on DMMA: lower 8x8 tiles ---------------------------
    const int nb = NB - c0 - PB;
    if (nb > 0) {
      const int nt = nb / 8;
      const int ntiles = nt * (nt + 1) / 2;
      for (int t = warp; t < ntiles; t += 8) {
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        const int tj = t - ti * (ti + 1) / 2;
        const int r0 = c0 + PB + 8 * ti, n0 = c0 + PB + 8 * tj;
        double* cp = S + (r0 + lr) * LDS64 + n0 + 2 * lc;
        double a0 = cp[0], a1 = cp[1];
        tile_mma<true>(a0, a1, S, r0, c0, S, n0, c0, PB, -1.0, lane);
        cp[0] = a0;
        cp[1] = a1;
      }
      __syncthreads();
    }
  }

  // ---- 16x16 diagonal inverses: thread c < 64 owns column c of its block (forward substitution)
  if (tid < NB) {
    const int b0 = tid & ~(PB - 1), c = tid & (PB - 1);
    double x[PB];
#pragma unroll
    for (int rr = 0; rr < PB; ++rr) {
      double acc = (rr == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < PB; ++k)
        if (k < rr) acc -= S[(b0 + rr) * LDS64 + b0 + k] * ((k >= c) ? x[k] : 0.0);
      x[rr] = (rr >= c) ? acc * dinv[b0 + rr] : 0.0;
    }
#pragma unroll
    for (int rr = 0; rr < PB; ++rr) X[(b0 + rr) * LDS64 + b0 + c] = x[rr];
  }
  __syncthreads();
  // ---- recursive doubling: X21 = -X22 (L21 X11) ----------------------------------------------
#pragma unroll 1
  for (int s = PB; s < NB; s *= 2) {
    const int npair = NB / (2 * s), tps = s / 8;
    const int ntl = npair * tps * tps;
    for (int t = warp; t < ntl; t += 8) {  // T_q = L21_q X11_q
      const int qq = t / (tps * tps), ti = (t / tps) % tps, tj = t % tps;
      const int o = 2 * s * qq;
      double a0 = 0.0, a1 = 0.0;
      tile_mma<false>(a0, a1, S, o + s + 8 * ti, o, X, o + 8 * tj, o, s, 1.0, lane);
      double* tp = T + (8 * ti + lr) * LDS64 + s * qq + 8 * tj + 2 * lc;  // pairs side by side
      tp[0] = a0;
      tp[1] = a1;
    }
    __syncthreads();
    for (int t = warp; t < ntl; t += 8) {  // X21_q = -X22_q T_q
      const int qq = t / (tps * tps), ti = (t / tps) % tps, tj = t % tps;
      const int o = 2 * s * qq;
      double a0 = 0.0, a1 = 0.0;
      tile_mma<false>(a0, a1, X, o + s + 8 * ti, o + s, T, s * qq + 8 * tj, 0, s, -1.0, lane);
      double* xp = X + (o + s + 8 * ti + lr) * LDS64 + o + 8 * tj + 2 * lc;
      xp[0] = a0;
      xp[1] = a1;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) potrf_diag_panel_kernel(double* __restrict__ L,
                                                               double* __restrict__ Linv, int Mp,
                                                               long long sL, int j0, int* info) {
  extern __shared__ double sm64[];
  double* S = sm64;                   // [64][LDS64]  block being factorised (lower part)
  double* X = S + NB * LDS64;         // [64][LDS64]  its inverse
  double* T = X + NB * LDS64;         // [32][LDS64]  temporary of the recursive-doubling products
  double* colS = T + 32 * LDS64;      // [2][64]      published column (double-buffered)
  double* dinv = colS + 2 * NB;       // [64]         1 / L[c][c]
  const int tid = threadIdx.x;
  double* Lk = L + blockIdx.x * sL + (long long)j0 * Mp + j0;
  double* Xk = Linv + blockIdx.x * sL + (long long)j0 * Mp + j0;
  int* info_k = info + blockIdx.x;
  if (j0 == 0 && tid == 0) *info_k = 0;  // only thread 0 of this CTA touches info[k]

  {
    // sixteen independent loads in flight per thread (one L2 round trip), then the shared-memory fill
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = tid + 256 * u;
      const int rr = e >> 6, c = e & 63;
      v[u] = (c <= rr) ? Lk[(long long)rr * Mp + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = tid + 256 * u;
      const int rr = e >> 6, c = e & 63;
      S[rr * LDS64 + c] = v[u];
      X[rr * LDS64 + c] = 0.0;
    }
  }
  __syncthreads();

  factor_and_invert_smem(S, X, T, colS, dinv, info_k, j0);

  {
    // all shared-memory reads first, then the global stores (32 independent stores per thread)
    double lv[16], xv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = tid + 256 * u;
      const int rr = e >> 6, c = e & 63;
      lv[u] = (c <= rr) ? S[rr * LDS64 + c] : 0.0;
      xv[u] = (c <= rr) ? X[rr * LDS64 + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = tid + 256 * u;
      const int rr = e >> 6, c = e & 63;
      Lk[(long long)rr * Mp + c] = lv[u];
      Xk[(long long)rr * Mp + c] = xv[u];
    }
  }

}

// ---- one launch per 64-column block for small matrices (2..4 blocks, i.e. Mp = 128 / 256) ----------
// The blocked chain above costs three launches per block column (panel update, diagonal block,
// panel scale) plus four for the assembly of L^-1, each a latency-bound GEMM on one to six CTAs.
// Here block column j is ONE launch of nblk CTAs per matrix, and L^-1 is built row by row on the way:
//   every CTA   : S = A_jj - sum_{k<j} L_jk L_jk^T, then factorises S itself (the same 16 us on every
//                 CTA, in parallel - redundant work on otherwise idle SMs instead of a dependent launch);
//   CTA c == j  : stores L_jj and Linv_jj;
//   CTA c >  j  : L_cj = (A_cj - sum_{k<j} L_ck L_jk^T) Linv_jj^T          (panel update + scale);
//   CTA c <  j  : Linv_jc = -Linv_jj sum_{k=c}^{j-1} L_jk Linv_kc          (row j of the inverse).
// A_jj is read from a copy of the diagonal blocks (Dg, filled by chol_init_kernel) because CTA j
// overwrites it in place while the other CTAs of the launch may not have read it yet.  Results are
// identical on every CTA (same instruction sequence on the same data), so no value crosses CTAs
// inside a launch.
constexpr size_t FUSED_SMEM = ((size_t)(6 * NB + 32) * LDS64 + 3 * NB) * sizeof(double);

// 64x64 block global -> shared by cp.async (eight 16-byte copies per thread); caller commits / waits
__device__ __forceinline__ void cp_block64(double* dst, const double* __restrict__ src, int ld, int tid) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int idx = tid + 256 * u;
    const int row = idx >> 5, seg = idx & 31;
    cp_async16(dst + row * LDS64 + 2 * seg, src + (long long)row * ld + 2 * seg, 16);
  }
}

// acc[t] += sign * (row tile rt of Am) * op(Bm) for the four column tiles t0 .. t0 + 3 (K = 64);
// independent accumulators keep the DMMA pipe busy.
template <bool B_IS_T>
__device__ __forceinline__ void row_mma4(double (&c0)[4], double (&c1)[4], const double* Am, int rt,
                                         const double* Bm, int t0, double sign, int lane) {
  const int lr = lane >> 2, lc = lane & 3;
#pragma unroll 4
  for (int k4 = 0; k4 < NB; k4 += 4) {
    const double a = sign * Am[(8 * rt + lr) * LDS64 + k4 + lc];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double b = B_IS_T ? Bm[(8 * (t0 + t) + lr) * LDS64 + k4 + lc] : Bm[(k4 + lc) * LDS64 + 8 * (t0 + t) + lr];
      dmma884(c0[t], c1[t], a, b);
    }
  }
}

// Work split of one launch (block column j): blockIdx.y = 0 is the diagonal CTA; every off-diagonal
// block c != j is shared by TWO CTAs - the panel blocks L_cj (c > j) by row halves, the inverse
// blocks Linv_jc (c < j) by column halves (its product chain X * (sum L_jk Linv_kc) only separates
// by columns).  The 36 lower tiles of the diagonal update are dealt round-robin to the eight warps.
__global__ void __launch_bounds__(256) potrf_column_fused_kernel(double* __restrict__ L,
                                                                 double* __restrict__ Linv,
                                                                 const double* __restrict__ Dg, int Mp,
                                                                 long long sL, int j, int* info) {
  extern __shared__ double sm64[];
  double* S = sm64;
  double* X = S + NB * LDS64;
  double* T = X + NB * LDS64;
  double* colS = T + 32 * LDS64;
  double* dinv = colS + 2 * NB;
  double* Pa = dinv + NB;            // [2][64][LDS64]  L_jk, double-buffered (cp.async)
  double* Pb = Pa + 2 * NB * LDS64;  // [2][64][LDS64]  L_ck or Linv_kc, double-buffered
  double* R = Pa;                    // after the products: this CTA's part of its off-diagonal block
  double* O = Pb;                    // ... and the staging of its result
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int k = blockIdx.x, nblk = Mp / NB, j0 = j * NB;
  int c = j, half = 0;
  if (blockIdx.y > 0) {
    const int i = (blockIdx.y - 1) >> 1;
    half = (blockIdx.y - 1) & 1;
    c = (i < j) ? i : i + 1;
  }
  double* Lm = L + k * sL;
  double* Xm = Linv + k * sL;
  const double* Dk = Dg + ((long long)k * nblk + j) * NB * NB;
  const int role = (c == j) ? 0 : (c > j ? 1 : 2);
  if (j == 0 && blockIdx.y == 0 && tid == 0) info[k] = 0;  // thread 0 of the diagonal CTA alone touches info[k]
  // this warp's four tiles of the off-diagonal block: row tile rt, column tiles t0 .. t0 + 3
  const int rt = (role == 1) ? 4 * half + (warp >> 1) : warp;
  const int t0 = (role == 1) ? 4 * (warp & 1) : 4 * half;

  auto issue = [&](int kb) {  // blocks of step kb into buffer kb & 1
    const int b = kb & 1;
    cp_block64(Pa + b * NB * LDS64, Lm + (long long)j0 * Mp + kb * NB, Mp, tid);
    if (role == 1) cp_block64(Pb + b * NB * LDS64, Lm + (long long)c * NB * Mp + kb * NB, Mp, tid);
    else if (role == 2 && kb >= c) cp_block64(Pb + b * NB * LDS64, Xm + (long long)kb * NB * Mp + c * NB, Mp, tid);
    cp_async_commit();
  };
  if (j > 0) issue(0);

  // diagonal update: lower tiles q = warp, warp + 8, ... < 36 with q = ti (ti + 1) / 2 + tj
  int sti[5], stj[5];
  double s0[5], s1[5], r0[4], r1[4];
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    const int q = warp + 8 * u;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
    sti[u] = ti;
    stj[u] = q - ti * (ti + 1) / 2;
    s0[u] = s1[u] = 0.0;
    if (q < 36) {
      const double2 v = *reinterpret_cast<const double2*>(Dk + (8 * sti[u] + lr) * NB + 8 * stj[u] + 2 * lc);
      s0[u] = v.x;
      s1[u] = v.y;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double2 rv = make_double2(0.0, 0.0);
    if (role == 1)
      rv = *reinterpret_cast<const double2*>(Lm + ((long long)c * NB + 8 * rt + lr) * Mp + j0 + 8 * (t0 + t) + 2 * lc);
    r0[t] = rv.x;
    r1[t] = rv.y;
  }
  for (int e = tid; e < NB * NB; e += 256) {
    S[(e >> 6) * LDS64 + (e & 63)] = 0.0;
    X[(e >> 6) * LDS64 + (e & 63)] = 0.0;
  }
  __syncthreads();  // the zero fill must not overtake the accumulator write-back below when j == 0
#pragma unroll 1
  for (int kb = 0; kb < j; ++kb) {
    if (kb + 1 < j) {
      issue(kb + 1);  // its buffer was last read in iteration kb - 1, which ended with a barrier
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* pa = Pa + (kb & 1) * NB * LDS64;
    const double* pb = Pb + (kb & 1) * NB * LDS64;
#pragma unroll 2
    for (int k4 = 0; k4 < NB; k4 += 4) {  // S -= L_jk L_jk^T on this warp's tiles
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        if (warp + 8 * u < 36) {
          const double av = -pa[(8 * sti[u] + lr) * LDS64 + k4 + lc];
          const double bv = pa[(8 * stj[u] + lr) * LDS64 + k4 + lc];
          dmma884(s0[u], s1[u], av, bv);
        }
      }
    }
    if (role == 1) row_mma4<true>(r0, r1, pb, rt, pa, t0, -1.0, lane);                   // R -= L_ck L_jk^T
    else if (role == 2 && kb >= c) row_mma4<false>(r0, r1, pa, rt, pb, t0, 1.0, lane);  // R += L_jk Linv_kc
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    if (warp + 8 * u < 36) {
      double* sp = S + (8 * sti[u] + lr) * LDS64 + 8 * stj[u] + 2 * lc;
      sp[0] = s0[u];
      sp[1] = s1[u];
    }
  }
  if (role != 0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      double* rp = R + (8 * rt + lr) * LDS64 + 8 * (t0 + t) + 2 * lc;
      rp[0] = r0[t];
      rp[1] = r1[t];
    }
  }
  __syncthreads();

  factor_and_invert_smem(S, X, T, colS, dinv, role == 0 ? info + k : nullptr, j0);

  if (role == 0) {
    double* Lk = Lm + (long long)j0 * Mp + j0;
    double* Xk = Xm + (long long)j0 * Mp + j0;
    double lv[16], xv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = tid + 256 * u;
      const int rr = e >> 6, cc = e & 63;
      lv[u] = (cc <= rr) ? S[rr * LDS64 + cc] : 0.0;
      xv[u] = (cc <= rr) ? X[rr * LDS64 + cc] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = tid + 256 * u;
      const int rr = e >> 6, cc = e & 63;
      Lk[(long long)rr * Mp + cc] = lv[u];
      Xk[(long long)rr * Mp + cc] = xv[u];
    }
    return;
  }
  double o0[4], o1[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) o0[t] = o1[t] = 0.0;
  if (role == 1) row_mma4<true>(o0, o1, R, rt, X, t0, 1.0, lane);   // L_cj (row half) = R Linv_jj^T
  else row_mma4<false>(o0, o1, X, rt, R, t0, -1.0, lane);           // Linv_jc (column half) = -Linv_jj R
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double* op = O + (8 * rt + lr) * LDS64 + 8 * (t0 + t) + 2 * lc;
    op[0] = o0[t];
    op[1] = o1[t];
  }
  __syncthreads();
  double* dst = (role == 1) ? Lm + (long long)c * NB * Mp + j0 : Xm + (long long)j0 * Mp + c * NB;
  double ov[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {  // this CTA's 32 x 64 (role 1) or 64 x 32 (role 2) half of the block
    const int e = tid + 256 * u;
    const int rr = (role == 1) ? 32 * half + (e >> 6) : (e >> 5);
    const int cc = (role == 1) ? (e & 63) : 32 * half + (e & 31);
    ov[u] = O[rr * LDS64 + cc];
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int e = tid + 256 * u;
    const int rr = (role == 1) ? 32 * half + (e >> 6) : (e >> 5);
    const int cc = (role == 1) ? (e & 63) : 32 * half + (e & 31);
    dst[(long long)rr * Mp + cc] = ov[u];
  }
}

// Linv <- 0 everywhere and L <- 0 strictly above the 64x64 diagonal blocks (the factorisation only
// fills the lower block triangle; consumers rely on zeros to the right of the diagonal blocks).  One
// wide launch instead of a memset node plus a serial zero-fill inside every diagonal-block kernel.
__global__ void __launch_bounds__(256) chol_init_kernel(double* __restrict__ L,
                                                        double* __restrict__ Linv, int Mp,
                                                        double* __restrict__ Dg) {
  const long long k = blockIdx.z;
  const int r = blockIdx.y;
  double* Lr = L + (k * Mp + r) * Mp;
  double* Xr = Linv + (k * Mp + r) * Mp;
  const int first_right = (r / NB + 1) * NB;  // first column right of this row's diagonal block
  if (Dg && blockIdx.x == 0 && threadIdx.x < NB / 2) {  // copy of the diagonal blocks (fused column path)
    const int jb = r / NB;
    const double2 v = *reinterpret_cast<const double2*>(Lr + jb * NB + 2 * threadIdx.x);
    *reinterpret_cast<double2*>(Dg + ((k * (Mp / NB) + jb) * NB + (r % NB)) * NB + 2 * threadIdx.x) = v;
  }
  for (int c = (blockIdx.x * 256 + threadIdx.x) * 2; c < Mp; c += gridDim.x * 512) {
    *reinterpret_cast<double2*>(Xr + c) = make_double2(0.0, 0.0);
    if (c >= first_right) *reinterpret_cast<double2*>(Lr + c) = make_double2(0.0, 0.0);
  }
}

}  // namespace

int chol_padded_dim(int M) {
  int Mp = NB;
  while (Mp < M) Mp *= 2;
  return Mp;
}

// recursive-doubling temporaries (Mp^2 / 4) or the copy of the diagonal blocks of the fused column
// path (Mp * NB), whichever is larger
size_t chol_scratch_doubles(int K, int Mp) {
  const size_t a = (size_t)K * Mp * Mp / 4, b = (size_t)K * Mp * NB;
  return a > b ? a : b;
}

int potrf_trtri_batched(double* L, double* Linv, double* T, int K, int Mp, int* info,
                        cudaStream_t stream) {
  const long long sL = (long long)Mp * Mp;
  const int nblk = Mp / NB;
  static const bool fused_on = [] {
    const char* e = getenv("GPFX_CHOL_FUSED");
    return !(e && e[0] == '0');
  }();
  const bool fused = fused_on && nblk >= 2 && nblk <= 4;
  {
    dim3 grid(ceil_div(Mp, 512), Mp, K);
    chol_init_kernel<<<grid, 256, 0, stream>>>(L, Linv, Mp, fused ? T : nullptr);
    GPFX_LAUNCH_CHECK();
  }
  // info[k] is cleared by the first diagonal-block kernel (a 4-byte memset would be a copy-engine
  // node that serialises parallel branches of a captured CUDA graph)
  if (fused) {
    static SmemAttrOnce once;
    GPFX_TRY(ensure_dyn_smem(potrf_column_fused_kernel, FUSED_SMEM, once));
    for (int j = 0; j < nblk; ++j) {
      potrf_column_fused_kernel<<<dim3(K, 2 * nblk - 1), 256, FUSED_SMEM, stream>>>(L, Linv, T, Mp, sL, j, info);
      GPFX_LAUNCH_CHECK();
    }
    return GPFX_OK;
  }
  for (int j = 0; j < nblk; ++j) {
    const int j0 = j * NB;
    if (j > 0) {
      GemmArgs g;
      g.A = L + (long long)j0 * Mp;
      g.B = L + (long long)j0 * Mp;
      g.C = L + (long long)j0 * Mp + j0;
      g.M = Mp - j0; g.N = NB; g.K = j0;
      g.lda = g.ldb = g.ldc = Mp;
      g.transB = 1;
      g.batch1 = K; g.sA1 = g.sB1 = g.sC1 = sL;
      g.alpha = -1.0; g.beta = 1.0;
      GPFX_TRY(launch_gemm(g, stream));
    }
    {
      constexpr size_t smem = ((size_t)(2 * NB + 32) * LDS64 + 3 * NB) * sizeof(double);
      static SmemAttrOnce once;
      GPFX_TRY(ensure_dyn_smem(potrf_diag_panel_kernel, smem, once));
      static const bool use_old = getenv("GPFX_POTRF_OLD") != nullptr;
      if (use_old) potrf_diag_kernel<<<K, 256, 0, stream>>>(L, Linv, Mp, sL, j0, info);
      else potrf_diag_panel_kernel<<<K, 256, smem, stream>>>(L, Linv, Mp, sL, j0, info);
    }
    GPFX_LAUNCH_CHECK();
    if (j < nblk - 1) {
      GemmArgs g;
      g.A = L + (long long)(j0 + NB) * Mp + j0;
      g.B = Linv + (long long)j0 * Mp + j0;
      g.C = L + (long long)(j0 + NB) * Mp + j0;  // in place: every CTA tile spans all NB columns
      g.cfg = 1;                                 // ... which needs BN >= NB (64x64 tiles)
      g.M = Mp - j0 - NB; g.N = NB; g.K = NB;
      g.lda = g.ldb = g.ldc = Mp;
      g.transB = 1;
      g.batch1 = K; g.sA1 = g.sB1 = g.sC1 = sL;
      GPFX_TRY(launch_gemm(g, stream));
    }
  }
  // L^-1 by recursive doubling: Linv21 = -Linv22 * (L21 * Linv11)
  for (int s = NB; s < Mp; s *= 2) {
    const int nq = Mp / (2 * s);
    const long long pair = 2LL * s * (Mp + 1);
    GemmArgs g1;
    g1.A = L + (long long)s * Mp; g1.lda = Mp;
    g1.B = Linv; g1.ldb = Mp;
    g1.C = T; g1.ldc = s;
    g1.M = g1.N = g1.K = s;
    g1.batch1 = K; g1.batch2 = nq;
    g1.sA1 = sL; g1.sA2 = pair;
    g1.sB1 = sL; g1.sB2 = pair;
    g1.sC1 = (long long)nq * s * s; g1.sC2 = (long long)s * s;
    g1.tri = TRI_LOWER_B;
    GPFX_TRY(launch_gemm(g1, stream));
    GemmArgs g2;
    g2.A = Linv + (long long)s * (Mp + 1); g2.lda = Mp;
    g2.B = T; g2.ldb = s;
    g2.C = Linv + (long long)s * Mp; g2.ldc = Mp;
    g2.M = g2.N = g2.K = s;
    g2.batch1 = K; g2.batch2 = nq;
    g2.sA1 = sL; g2.sA2 = pair;
    g2.sB1 = g1.sC1; g2.sB2 = g1.sC2;
    g2.sC1 = sL; g2.sC2 = pair;
    g2.tri = TRI_LOWER_A;
    g2.alpha = -1.0;
    GPFX_TRY(launch_gemm(g2, stream));
  }
  return GPFX_OK;
}

// dK = L^-T Phi(L^T dL) L^-1  (un-symmetrised; consumers contract it with a symmetric dK/dtheta)
int chol_backward_batched(const double* L, const double* Linv, const double* dL, double* T1,
                          double* T2, double* dK, int K, int Mp, cudaStream_t stream) {
  const long long sL = (long long)Mp * Mp;
  GemmArgs g;
  g.M = g.N = g.K = Mp;
  g.lda = g.ldb = g.ldc = Mp;
  g.batch1 = K; g.sA1 = g.sB1 = g.sC1 = sL;
  // T1 = Phi(L^T dL)
  g.A = L; g.transA = 1; g.B = dL; g.C = T1;
  g.tri = TRI_UPPER_A | TRI_LOWER_B; g.tril_out = 1; g.diag_scale = 0.5;
  GPFX_TRY(launch_gemm(g, stream));
  // T2 = T1 Linv
  g.A = T1; g.transA = 0; g.B = Linv; g.C = T2;
  g.tri = TRI_LOWER_A | TRI_LOWER_B; g.tril_out = 0; g.diag_scale = 1.0;
  GPFX_TRY(launch_gemm(g, stream));
  // dK = Linv^T T2
  g.A = Linv; g.transA = 1; g.B = T2; g.C = dK;
  g.tri = TRI_UPPER_A | TRI_LOWER_B;
  GPFX_TRY(launch_gemm(g, stream));
  return GPFX_OK;
}

}  // namespace gpfx
