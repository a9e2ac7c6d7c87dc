// FP64 DMMA GEMM, Blackwell operand feed: tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) into a
// 128B-swizzled shared-memory ring, one producer warp + DMMA consumer warps, per-stage full/empty
// mbarriers (no CTA-wide barrier in the mainloop).  Same contract as gemm_f64_kernel (GemmArgs):
// triangular k-range skipping, tril output, split-K, fused column reductions.
//
// Shared-memory layouts (TBK = 16 doubles = one 128-byte swizzle row):
//   k-major operand  (element (r,k) at g[r*ld + k]):  3-D map {K, rows, batch}, box {16, ROWS, 1}
//        smem row r = 128 bytes of k; 16-byte chunk c of row r is stored at chunk c ^ (r & 7)
//   mn-major operand (element (r,k) at g[k*ld + r]):  4-D map {16, K, rows/16, batch},
//        box {16, 16, ROWS/16, 1}: slab s = r/16 holds 16 k-rows of 128 bytes (16 consecutive r);
//        chunk c of k-row k is stored at chunk c ^ (k & 7)
// Fragment ownership is chosen so that every DMMA fragment load is bank-conflict free under the
// swizzle (see frag_row): 8x8 fragments take interleaved row blocks of the CTA tile, which also
// balances triangular / tril-output work between the warps of a CTA.
//
// The kernel is persistent (except when one tile alone lasts > 250 us: those grids run one tile per
// CTA, `tmax`): CTAs take tiles from a longest-first list
// through an atomic counter (first tile = blockIdx.x; the producer thread fetches the next tile id one
// tile ahead and hands it to the consumer warps through a small mbarrier queue), the shared memory
// ring and its mbarrier phases run on across tile boundaries, so the producer is already loading the
// next tile while the consumers store the current one.
//
// Replaces the tf.linalg.triangular_solve / tf.matmul calls of GPflow's base_conditional that
// GPLayer.predict reaches (reference gpflux/layers/gp_layer.py:257-266)
// and their autodiff transposes (row a14), on the large shapes (c2 / c3 of BASELINE.json).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "../../include/gpfx.h"
#include "context.h"
#include "gemm_f64.h"

namespace gpfx {

namespace {

constexpr int TBK = 16;

__device__ __forceinline__ unsigned sm_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(sm_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(sm_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(sm_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];\n" ::"r"(sm_u32(dst)),
      "l"(map), "r"(sm_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];\n" ::"r"(sm_u32(dst)),
      "l"(map), "r"(sm_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Row (inside the CTA tile) of DMMA fragment `f`, lane-row `lr` (= lane >> 2 for the A operand and
// for C rows; = 2*(lane & 3) + e for C columns), for warp `w` of `WARPS` along that dimension.
//   k-major : fragment f is the 8-row block f*WARPS + w
//   mn-major: fragments 2s, 2s+1 share the 16-row slab s*WARPS + w; lane-row lr of sub-fragment h
//             takes slab row (lr&1) + 8*((lr>>1)&1) + 2*(lr>>2) + 4*h   (conflict-free under the swizzle)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   sm_u32(dst)),
               "l"(src), "r"(bytes), "r"(sm_u32(bar))
               : "memory");
}

template <bool KMAJOR, int WARPS>
__host__ __device__ constexpr int frag_row(int w, int f, int lr) {
  if (KMAJOR) return (f * WARPS + w) * 8 + lr;
  return ((f >> 1) * WARPS + w) * 16 + (lr & 1) + 8 * ((lr >> 1) & 1) + 2 * (lr >> 2) + 4 * (f & 1);
}
// smallest / largest tile row of fragment f
template <bool KMAJOR, int WARPS>
__host__ __device__ constexpr int frag_lo(int w, int f) {
  if (KMAJOR) return (f * WARPS + w) * 8;
  return ((f >> 1) * WARPS + w) * 16 + 4 * (f & 1);
}
template <bool KMAJOR, int WARPS>
__host__ __device__ constexpr int frag_hi(int w, int f) {
  if (KMAJOR) return (f * WARPS + w) * 8 + 7;
  return ((f >> 1) * WARPS + w) * 16 + 11 + 4 * (f & 1);
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int MINB>
struct TmaCfg {
  static constexpr int NC = WARPS_M * WARPS_N;  // consumer warps
  // one extra warpgroup: its first warp is the TMA producer.  Register allocation is per warpgroup
  // (setmaxnreg): the producer group shrinks to 40 registers, the consumers grow to CONS_REGS.
  static constexpr int NT = (NC + 4) * 32;
  static constexpr int LAUNCH_REGS = (65536 / (NT * MINB)) / 8 * 8;                             // what ptxas assigns under the launch bound
  static constexpr int CONS_REGS_RAW = LAUNCH_REGS + (LAUNCH_REGS - 40) * 4 / NC;
  static constexpr int CONS_REGS = (CONS_REGS_RAW > 248 ? 248 : CONS_REGS_RAW) / 8 * 8;
  static_assert(NC % 4 == 0, "consumer warps come in warpgroups");
  static constexpr int MI = BM / WARPS_M / 8, NI = BN / WARPS_N / 8;
  static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // scratch of the fused epilogues: column reductions [WARPS_M][BN], row reductions [WARPS_N][BM]
  static constexpr int RED_BYTES = (WARPS_M * BN > WARPS_N * BM ? WARPS_M * BN : WARPS_N * BM) * 8;
  static constexpr int SCHED_Q = 2;                  // depth of the tile-id queue producer -> consumers
  static constexpr int KSCALE_BYTES = STAGES * TBK * 8;  // per-stage slice of GemmArgs.kscale
  static constexpr size_t SMEM =
      (size_t)STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8 + RED_BYTES + 3 * SCHED_Q * 8 + 16 + KSCALE_BYTES;
  static_assert(MI * NI <= 32, "live mask is one 32-bit word");
  static_assert(MI % 2 == 0 && NI % 2 == 0, "mn-major operands pair fragments in 16-row slabs");
};

// Tile counters of the dynamic scheduler, one slot per launch in flight (the host rotates through the
// slots; the last CTA to finish resets its slot, so a slot is clean whenever a launch picks it up).
struct SchedSlot {
  unsigned next, done;
};
constexpr int kSchedSlots = 1024;
__device__ SchedSlot g_sched[kSchedSlots];

// AKM: A is k-major (GemmArgs.transA == 0); BKM: B is k-major (GemmArgs.transB == 1)
// ADDIN: the instantiation that carries the add-in epilogue terms (kept out of the others: its extra live
// values cost the plain kernels ~3 % through spills)
// KSCALE: the instantiation that multiplies the B fragments by GemmArgs.kscale[k] (a 128-byte slice per stage,
// brought in by the producer with a 1-D bulk copy on the stage's mbarrier)
template <int BM, int BN, int WARPS_M, int WARPS_N, bool AKM, bool BKM, int STAGES, int MINB, bool ADDIN,
          bool KSCALE>
__global__ void __launch_bounds__((WARPS_M * WARPS_N + 4) * 32, MINB)
    gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                        const GemmArgs p, const int bmulA, const int bmulB, const int sw32, const int slot,
                        const int tmax) {
  using Cfg = TmaCfg<BM, BN, WARPS_M, WARPS_N, STAGES, MINB>;
  constexpr int NC = Cfg::NC, MI = Cfg::MI, NI = Cfg::NI;
  extern __shared__ char smem_raw[];
  char* tiles = smem_raw + ((1024u - (sm_u32(smem_raw) & 1023u)) & 1023u);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(tiles + STAGES * Cfg::STAGE_BYTES);
  unsigned long long* empty = full + STAGES;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler as well
  double* red = reinterpret_cast<double*>(empty + STAGES);  // [WARPS_M][BN]
  constexpr int SQ = Cfg::SCHED_Q;
  unsigned long long* sq_full = reinterpret_cast<unsigned long long*>(red + Cfg::RED_BYTES / 8);
  unsigned long long* sq_empty = sq_full + SQ;
  volatile int* sq_tile = reinterpret_cast<volatile int*>(sq_empty + SQ);
  char* ksc_raw = reinterpret_cast<char*>(sq_empty + 2 * SQ);
  double* ksc = reinterpret_cast<double*>(ksc_raw + ((16u - (sm_u32(ksc_raw) & 15u)) & 15u));  // [STAGES][TBK]

  // ---- tile list (identical in every warp of every CTA) ----------------------------------------
  const int tiles_x = (p.N + BN - 1) / BN, tiles_y = (p.M + BM - 1) / BM;
  const int nbz = p.batch1 * p.splits;
  // tril output with square tiles: the strictly-lower tiles first, then the (shorter) diagonal tiles,
  // then the tiles above the diagonal (no k-loop: they only store beta*C); otherwise the panel raster
  // described in decode()
  const bool tril_list = p.tril_out && BM == BN && tiles_x == tiles_y;
  const int n_strict = tiles_y * (tiles_y - 1) / 2;
  const int total_tiles = tiles_x * tiles_y * nbz;
  struct Tile {
    int m0, n0, by, b, split, klo, t_begin, t_end;
  };
  auto decode = [&](int tile) -> Tile {
    int bx, by, bz;
    if (tril_list) {
      const int n_live = (n_strict + tiles_y) * nbz;
      if (tile >= n_strict * nbz && tile < n_live) {
        const int d = tile - n_strict * nbz;
        bz = d / tiles_y;
        by = bx = d - bz * tiles_y;
      } else {
        const bool upper = tile >= n_live;
        const int e = upper ? tile - n_live : tile;
        bz = e / n_strict;
        const int u = e - bz * n_strict;
        by = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)u)) * 0.5f);
        while (by * (by - 1) / 2 > u) --by;
        while ((by + 1) * by / 2 <= u) ++by;
        bx = u - by * (by - 1) / 2;
        if (upper) {
          const int sw = by;
          by = bx;
          bx = sw;
        }
      }
    } else {
      // L2-friendly raster: batch / split slowest; inside it panels of kPanel column tiles, and inside a
      // panel the row blocks (longest k-range first) sweep the panel's columns.  The tiles in flight
      // then share one batch's operands: a panel of op(B) (kPanel x 128 columns) stays in L2 while every
      // row block of op(A) passes over it.  (Ordering the row-block classes across batches instead costs
      // 7.7x the algorithmic DRAM traffic on the c3 products: profiles/ncu_gemm_tma_r02b.txt.)
      constexpr int kPanel = 8;
      const int per_bz = tiles_x * tiles_y;
      bz = tile / per_bz;
      int r = tile - bz * per_bz;
      const int panel = r / (kPanel * tiles_y);
      r -= panel * kPanel * tiles_y;
      const int pw = min(kPanel, tiles_x - panel * kPanel);  // width of this (possibly last, narrower) panel
      const int yy = r / pw;
      bx = panel * kPanel + (r - yy * pw);
      by = (p.tri & TRI_LOWER_A) ? tiles_y - 1 - yy : yy;
    }
    Tile t;
    t.by = by;
    t.m0 = by * BM;
    t.n0 = bx * BN;
    t.split = bz % p.splits;
    t.b = bz / p.splits;
    int klo = 0, khi = p.K;
    if (p.tri & TRI_UPPER_A) klo = max(klo, t.m0);
    if (p.tri & TRI_LOWER_B) klo = max(klo, t.n0);
    if (p.tri & TRI_LOWER_A) khi = min(khi, t.m0 + BM);
    if (p.tri & TRI_UPPER_B) khi = min(khi, t.n0 + BN);
    klo = (klo / TBK) * TBK;
    const bool dead = p.tril_out && (t.n0 >= t.m0 + BM);
    const int T = (khi > klo && !dead) ? (khi - klo + TBK - 1) / TBK : 0;
    t.klo = klo;
    t.t_begin = 0;
    t.t_end = T;
    if (p.splits > 1) {
      const int cps = (T + p.splits - 1) / p.splits;
      t.t_begin = min(T, t.split * cps);
      t.t_end = min(T, t.t_begin + cps);
    }
    return t;
  };

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mapB) : "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NC);
    }
#pragma unroll
    for (int s = 0; s < SQ; ++s) {
      mbar_init(&sq_full[s], 1);
      mbar_init(&sq_empty[s], NC);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NC) {
    // ------------------------------------------------------------------ TMA producer warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == NC && lane == 0) {
      int stage = 0;
      unsigned phase = 0;
      int q = 0;
      unsigned qphase = 0;
      int tile = blockIdx.x;
      int ntiles = 0;
      while (true) {
        // hand the tile id (or the end marker) to the consumer warps
        mbar_wait(&sq_empty[q], qphase ^ 1u);
        sq_tile[q] = tile;
        mbar_arrive(&sq_full[q]);
        if (++q == SQ) {
          q = 0;
          qphase ^= 1u;
        }
        if (tile >= total_tiles) break;
        // the id of the tile after this one is fetched now and used when this tile's loads are issued
        // (a CTA retires after tmax tiles so that kernels of other, higher-priority streams get SM slots)
        const int next_tile = (++ntiles < tmax) ? (int)gridDim.x + (int)atomicAdd(&g_sched[slot].next, 1u) : 0x7fffffff;
        const Tile tl = decode(tile);
        const int bA = tl.b * bmulA, bB = tl.b * bmulB;
        const double* ksrc = KSCALE ? p.kscale + (long long)tl.b * p.sKscale : nullptr;
        for (int t = tl.t_begin; t < tl.t_end; ++t) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES + (KSCALE ? TBK * 8 : 0));
          char* as = tiles + stage * Cfg::STAGE_BYTES;
          char* bs = as + Cfg::A_BYTES;
          const int k0 = tl.klo + t * TBK;
          if (KSCALE) bulk_load_1d(ksc + stage * TBK, ksrc + k0, TBK * 8, &full[stage]);
          if (AKM) tma_load_3d(as, &mapA, &full[stage], k0, tl.m0, bA);
          else tma_load_4d(as, &mapA, &full[stage], 0, k0, tl.m0 >> 4, bA);
          if (BKM) tma_load_3d(bs, &mapB, &full[stage], k0, tl.n0, bB);
          else tma_load_4d(bs, &mapB, &full[stage], 0, k0, tl.n0 >> 4, bB);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tile = next_tile;
      }
      // every CTA stops drawing from the counter before it reports in: the last one leaves the slot clean
      __threadfence();
      if (atomicAdd(&g_sched[slot].done, 1u) == gridDim.x - 1) {
        g_sched[slot].next = 0u;
        g_sched[slot].done = 0u;
        __threadfence();
      }
    }
    return;
  }

  // -------------------------------------------------------------------- DMMA consumers
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONS_REGS));
  // odd warp rows run right to left: the two warps of a scheduler (warp id mod 4) then sit at mirrored
  // columns, which evens out the live fragments of a tril tile on the diagonal between the schedulers
  const int wm = warp / WARPS_N;
  const int wn = (wm & 1) ? WARPS_N - 1 - warp % WARPS_N : warp % WARPS_N;
  const int lr = lane >> 2, lc = lane & 3;

  // per-thread fragment addressing (bytes inside the stage's operand tile)
  //   k-major : thr + f*(WARPS*1024) + xk[k4],           xk[k4] = (((lc>>1) ^ lr) << 4) ^ (k4 << 5)
  //   mn-major: thr + (f>>1)*(WARPS*2048) + k4*512 + xm[f&1][k4&1],
  //             xm[h][kp] = ((4*((lr>>1)&1) + (lr>>2)) ^ lc) << 4) ^ (h << 5) ^ (kp << 6)
  // k-major operands use the 128B swizzle with 32-byte atoms when the driver accepts it for FP64
  // (sw32): 32-byte atom a of row r is stored at atom a ^ (r & 3), which makes the 8 rows x 32 bytes of
  // a fragment conflict free per half-warp:   thr + f*(WARPS*1024) + xk[k4],
  //   xk[k4] = ((k4 ^ (lr & 3)) << 5) + lc*8
  const unsigned xm0 = (unsigned)(((4 * ((lr >> 1) & 1) + (lr >> 2)) ^ lc) << 4);
  const unsigned kthr = sw32 ? 0u : (unsigned)((lc & 1) * 8);
  const unsigned a_thr = AKM ? (unsigned)((wm * 8 + lr) * 128) + kthr
                             : (unsigned)(wm * (TBK * 128) + lc * 128 + (lr & 1) * 8);
  const unsigned b_thr = BKM ? (unsigned)((wn * 8 + lr) * 128) + kthr
                             : (unsigned)(wn * (TBK * 128) + lc * 128 + (lr & 1) * 8);
  unsigned xk[4], xm[2][2];
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4)
    xk[k4] = sw32 ? (unsigned)(((k4 ^ (lr & 3)) << 5) + lc * 8) : (unsigned)((((lc >> 1) ^ lr) << 4) ^ (k4 << 5));
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) xm[h][kp] = xm0 ^ (unsigned)(h << 5) ^ (unsigned)(kp << 6);

  // Triangular structure at fragment granularity.  Skipping is an optimisation only: the skipped
  // operand entries are stored zeros (and the skipped outputs of a tril product are never stored), so a
  // path without a mask stays exact.
  //   A-side (lower-A / upper-A): fragment i is live at k4 step kk iff  a_lo_i <= kk + 3  and
  //     kk <= a_hi_i.  CTAs whose warps all own the same rows (WARPS_M == 1) run the k-tiles of the
  //     diagonal block through bodies whose live set is a compile-time constant; otherwise every
  //     quantity is warp-uniform (the warp index comes from a shuffle) and the test is a uniform branch
  //     around the NI DMMAs of that fragment row.
  //   tril output, square tile on the diagonal, both operands k-major: (i, j) is live iff
  //     i*WARPS_M - j*WARPS_N >= wn - wm; the mainloop is instantiated per value of wn - wm, the mask
  //     is a compile-time constant there.
  //   B-side triangular operands only shorten the CTA's k-range (klo / khi).
  const bool upperA = (p.tri & TRI_UPPER_A) != 0, lowerA = (p.tri & TRI_LOWER_A) != 0;
  constexpr int A_STEP = AKM ? WARPS_M * 8 : 0;  // k-major: row range of fragment i = fragment 0 + i*A_STEP
  constexpr int kNoTril = -1000;

  auto a_off = [&](int i, int k4) -> unsigned {
    return AKM ? a_thr + (unsigned)(i * WARPS_M * 1024) + xk[k4]
               : a_thr + (unsigned)((i >> 1) * WARPS_M * (TBK * 128) + k4 * 512) + xm[i & 1][k4 & 1];
  };
  auto b_off = [&](int j, int k4) -> unsigned {
    return BKM ? b_thr + (unsigned)(j * WARPS_N * 1024) + xk[k4]
               : b_thr + (unsigned)((j >> 1) * WARPS_N * (TBK * 128) + k4 * 512) + xm[j & 1][k4 & 1];
  };
  auto consumer_sync = [] { asm volatile("bar.sync 1, %0;\n" ::"n"(NC * 32) : "memory"); };

  int stage = 0;
  unsigned phase = 0;
  int q = 0;
  unsigned qphase = 0;
  while (true) {
    mbar_wait(&sq_full[q], qphase);
    const int tile = sq_tile[q];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sq_empty[q]);
    if (++q == SQ) {
      q = 0;
      qphase ^= 1u;
    }
    if (tile >= total_tiles) break;
    const Tile tl = decode(tile);
    const int m0 = tl.m0, n0 = tl.n0, b = tl.b;
    const bool diag_tile = p.tril_out && AKM && BKM && BM == BN && (n0 == m0);
    const int a_lo0 = m0 + frag_lo<AKM, WARPS_M>(wm, 0), a_hi0 = m0 + frag_hi<AKM, WARPS_M>(wm, 0);
    // k-range in which every A fragment of this warp is live
    int full_lo = 0, full_hi = 0x7fffffff;
    if (upperA) full_lo = max(full_lo, m0 + frag_hi<AKM, WARPS_M>(wm, MI - 1) + 1);
    if (lowerA) full_hi = min(full_hi, m0 + frag_lo<AKM, WARPS_M>(wm, 0));

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (ADDIN && p.addmat != nullptr && m0 + BM <= p.M && n0 + BN <= p.N) {
      // the epilogue reads this tile of addmat: pull it into L2 while the mainloop runs
      const char* at0 = reinterpret_cast<const char*>(p.addmat + (long long)b * p.sAdd + (long long)m0 * p.ldadd + n0);
      constexpr int LPR = BN * 8 / 128;  // 128-byte lines per tile row
      for (int idx = tid; idx < BM * LPR; idx += NC * 32) {
        const char* line = at0 + (long long)(idx / LPR) * p.ldadd * 8 + (idx % LPR) * 128;
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(line));
      }
    }

    // DELTA = kNoTril: no static mask; otherwise the tril mask of a diagonal tile for wn - wm == DELTA
    auto mainloop = [&](auto delta_c) {
      constexpr int DELTA = decltype(delta_c)::value;
      auto live = [](int i, int j) constexpr -> bool {
        return DELTA == kNoTril || (i * WARPS_M - j * WARPS_N >= DELTA);
      };
      // one k-tile.  AV = -1: every A fragment live; AV = td (lower-A) / 100 + td (upper-A): k-tile td
      // of the diagonal block of a CTA whose warps all own the same rows (WARPS_M == 1)
      auto ktile = [&](const char* as, const char* bs, const double* sc, auto av_c) {
        constexpr int AV = decltype(av_c)::value;
        auto a_on = [](int i, int k4) constexpr -> bool {
          if (AV < 0) return true;
          if (AV < 100) return 16 * AV + 4 * k4 <= frag_hi<AKM, 1>(0, i);
          return frag_lo<AKM, 1>(0, i) <= 16 * (AV - 100) + 4 * k4 + 3;
        };
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          double af[MI], bf[NI];
#pragma unroll
          for (int i = 0; i < MI; ++i) af[i] = *reinterpret_cast<const double*>(as + a_off(i, k4));
#pragma unroll
          for (int j = 0; j < NI; ++j) bf[j] = *reinterpret_cast<const double*>(bs + b_off(j, k4));
          if (KSCALE) {
            const double sk = sc[k4 * 4 + lc];  // this lane's k of the 8x8x4 fragment
#pragma unroll
            for (int j = 0; j < NI; ++j) bf[j] *= sk;
          }
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j)
              if (live(i, j) && a_on(i, k4)) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      };
      for (int t = tl.t_begin; t < tl.t_end; ++t) {
        mbar_wait(&full[stage], phase);
        const char* as = tiles + stage * Cfg::STAGE_BYTES;
        const char* bs = as + Cfg::A_BYTES;
        const int k0t = tl.klo + t * TBK;
        const double* sc = ksc + stage * TBK;
        if (k0t >= full_lo && k0t + TBK - 1 <= full_hi) {
          ktile(as, bs, sc, std::integral_constant<int, -1>{});
        } else if (WARPS_M == 1 && DELTA == kNoTril && (lowerA != upperA)) {
          // k-tile td of the diagonal block: the live A fragments are known at compile time
          const int td = (k0t - m0) >> 4;
          const int av = lowerA ? td : 100 + td;
          bool done = false;
#define GPFX_AV_CASE(V)                                                      \
  if constexpr (((V) % 100) < BM / TBK) {                                    \
    if (!done && av == (V)) {                                                \
      ktile(as, bs, sc, std::integral_constant<int, (V)>{});                 \
      done = true;                                                           \
    }                                                                        \
  }
          GPFX_AV_CASE(0) GPFX_AV_CASE(1) GPFX_AV_CASE(2) GPFX_AV_CASE(3)
          GPFX_AV_CASE(4) GPFX_AV_CASE(5) GPFX_AV_CASE(6) GPFX_AV_CASE(7)
          GPFX_AV_CASE(100) GPFX_AV_CASE(101) GPFX_AV_CASE(102) GPFX_AV_CASE(103)
          GPFX_AV_CASE(104) GPFX_AV_CASE(105) GPFX_AV_CASE(106) GPFX_AV_CASE(107)
#undef GPFX_AV_CASE
          if (!done) ktile(as, bs, sc, std::integral_constant<int, -1>{});
        } else {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kk = k0t + k4 * 4;
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = *reinterpret_cast<const double*>(as + a_off(i, k4));
#pragma unroll
            for (int j = 0; j < NI; ++j) bf[j] = *reinterpret_cast<const double*>(bs + b_off(j, k4));
            if (KSCALE) {
              const double sk = sc[k4 * 4 + lc];
#pragma unroll
              for (int j = 0; j < NI; ++j) bf[j] *= sk;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i) {
              const int lo_i = AKM ? a_lo0 + i * A_STEP : m0 + frag_lo<AKM, WARPS_M>(wm, i);
              const int hi_i = AKM ? a_hi0 + i * A_STEP : m0 + frag_hi<AKM, WARPS_M>(wm, i);
              const bool on = (!upperA || lo_i <= kk + 3) && (!lowerA || kk <= hi_i);
              if (on) {
#pragma unroll
                for (int j = 0; j < NI; ++j)
                  if (live(i, j)) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    };
    if (diag_tile) {
      const int delta = wn - wm;  // -(WARPS_M-1) .. WARPS_N-1
      bool done = false;
#define GPFX_TRIL_CASE(D)                                                      \
  if constexpr ((D) >= -(WARPS_M - 1) && (D) <= WARPS_N - 1) {                  \
    if (!done && delta == (D)) {                                               \
      mainloop(std::integral_constant<int, (D)>{});                            \
      done = true;                                                             \
    }                                                                          \
  }
      GPFX_TRIL_CASE(-3) GPFX_TRIL_CASE(-2) GPFX_TRIL_CASE(-1) GPFX_TRIL_CASE(0)
      GPFX_TRIL_CASE(1) GPFX_TRIL_CASE(2) GPFX_TRIL_CASE(3)
#undef GPFX_TRIL_CASE
      if (!done) mainloop(std::integral_constant<int, kNoTril>{});
    } else {
      mainloop(std::integral_constant<int, kNoTril>{});
    }

    // ---------------------------------------------------------------- epilogue (consumer warps)
    if (p.splits > 1) {
      const long long nb = (long long)p.batch1 * p.batch2;
      double* W = p.splitws + ((long long)tl.split * nb + b) * (long long)p.M * p.N;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = m0 + frag_row<AKM, WARPS_M>(wm, i, lr);
        if (r >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int c = n0 + frag_row<BKM, WARPS_N>(wn, j, 2 * lc);
          if (c < p.N) W[(long long)r * p.N + c] = acc[i][j][0];
          if (c + 1 < p.N) W[(long long)r * p.N + c + 1] = acc[i][j][1];
        }
      }
      continue;
    }

    double* Cb = p.C + (long long)b * p.sC1;
    const double* cs = p.colscale ? p.colscale + b * p.sColscale : nullptr;
    const bool vecC = ((p.ldc & 1) == 0) && ((((uintptr_t)Cb) & 15) == 0);
    // interior tile of a plain product: C = alpha * acc, 16-byte stores, no per-element tests
    const bool plain = !p.tril_out && (cs == nullptr || (ADDIN && p.addmat != nullptr)) && p.diag_scale == 1.0 &&
                       vecC && m0 + BM <= p.M && n0 + BN <= p.N;
    if (plain) {
      const long long roff = (long long)(m0 + frag_row<AKM, WARPS_M>(wm, 0, lr));
      const int coff = n0 + frag_row<BKM, WARPS_N>(wn, 0, 2 * lc);
      double* cthr = Cb + roff * p.ldc + coff;
      if (ADDIN && p.addmat != nullptr) {
        // C = alpha*acc*colscale[n] + addscale*addcol[n]*addmat[m][n] + r1u[m]*r1v[n];  rowdot = sum_n addmat[m][n] r1v[n]
        const double* athr = p.addmat + (long long)b * p.sAdd + roff * p.ldadd + coff;
        const double* colv = p.addcol + (long long)b * p.sAddcol + coff;
        const double* rv = p.r1v + (long long)b * p.sR1v + coff;
        const double* ru = p.r1u + (long long)b * p.sR1u + roff * p.ldr1u;
        const bool do_rowdot = p.rowdot != nullptr;
        if (do_rowdot) consumer_sync();  // earlier readers of `red` are done
        // column blocks outermost: only one block's column terms are live next to the accumulators
        double rsum[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) rsum[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int dj = frag_row<BKM, WARPS_N>(0, j, 0) - frag_row<BKM, WARPS_N>(0, 0, 0);
          double2 cs2 = *reinterpret_cast<const double2*>(colv + dj);
          cs2.x *= p.addscale;
          cs2.y *= p.addscale;
          const double2 rv2 = *reinterpret_cast<const double2*>(rv + dj);
          double2 al2 = make_double2(p.alpha, p.alpha);
          if (cs != nullptr) {
            const double2 c2 = *reinterpret_cast<const double2*>(cs + coff + dj);
            al2.x *= c2.x;
            al2.y *= c2.y;
          }
#pragma unroll
          for (int i = 0; i < MI; ++i) {
            const int di = frag_row<AKM, WARPS_M>(0, i, 0) - frag_row<AKM, WARPS_M>(0, 0, 0);
            const double u = ru[(long long)di * p.ldr1u];
            const double2 a2 = *reinterpret_cast<const double2*>(athr + (long long)di * p.ldadd + dj);
            double v0 = al2.x * acc[i][j][0], v1 = al2.y * acc[i][j][1];
            v0 += cs2.x * a2.x;
            v1 += cs2.y * a2.y;
            v0 += u * rv2.x;
            v1 += u * rv2.y;
            rsum[i] += a2.x * rv2.x + a2.y * rv2.y;
            acc[i][j][0] = v0;
            acc[i][j][1] = v1;
            *reinterpret_cast<double2*>(cthr + (long long)di * p.ldc + dj) = make_double2(v0, v1);
          }
        }
        if (do_rowdot) {
#pragma unroll
          for (int i = 0; i < MI; ++i) {
            // fixed-order row reduction: the 4 lanes of a row here, the WARPS_N warps through shared memory below
            double r = rsum[i];
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            if (lc == 0) red[wn * BM + frag_row<AKM, WARPS_M>(wm, i, lr)] = r;
          }
        }
        if (do_rowdot) {
          consumer_sync();
          for (int r = tid; r < BM; r += NC * 32) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS_N; ++w) tot += red[w * BM + r];
            p.rowdot[((long long)(n0 / BN) * p.batch1 + b) * p.M + m0 + r] = tot;
          }
        }
      } else {
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        // rows / columns of the other fragments are compile-time offsets from fragment 0
        double* crow = cthr + (long long)(frag_row<AKM, WARPS_M>(0, i, 0) - frag_row<AKM, WARPS_M>(0, 0, 0)) * p.ldc;
        if (p.beta != 0.0) {
          double2 o[NI];
#pragma unroll
          for (int j = 0; j < NI; ++j)
            o[j] = *reinterpret_cast<const double2*>(crow + (frag_row<BKM, WARPS_N>(0, j, 0) - frag_row<BKM, WARPS_N>(0, 0, 0)));
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];  // same rounding as the general path
            v0 += p.beta * o[j].x;
            v1 += p.beta * o[j].y;
            acc[i][j][0] = v0;
            acc[i][j][1] = v1;
          }
        } else {
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            acc[i][j][0] *= p.alpha;
            acc[i][j][1] *= p.alpha;
          }
        }
#pragma unroll
        for (int j = 0; j < NI; ++j)
          *reinterpret_cast<double2*>(crow + (frag_row<BKM, WARPS_N>(0, j, 0) - frag_row<BKM, WARPS_N>(0, 0, 0))) =
              make_double2(acc[i][j][0], acc[i][j][1]);
      }
      }
    } else
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int c = n0 + frag_row<BKM, WARPS_N>(wn, j, 2 * lc);
      double s0 = 1.0, s1 = 1.0;
      if (cs) {
        s0 = (c < p.N) ? cs[c] : 0.0;
        s1 = (c + 1 < p.N) ? cs[c + 1] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = m0 + frag_row<AKM, WARPS_M>(wm, i, lr);
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

    if (p.sumsq == nullptr && p.wsum == nullptr) continue;
    // fused column reductions over this tile's BM rows, through the dedicated scratch (the ring already
    // receives the next tile)
    consumer_sync();  // the previous tile's readers of `red` are done
    if (p.sumsq) {
#pragma unroll
      for (int j = 0; j < NI; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double sq = 0.0;
#pragma unroll
          for (int i = 0; i < MI; ++i) sq += acc[i][j][e] * acc[i][j][e];
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          sq += __shfl_xor_sync(0xffffffffu, sq, 8);
          sq += __shfl_xor_sync(0xffffffffu, sq, 16);
          if (lr == 0) red[wm * BN + frag_row<BKM, WARPS_N>(wn, j, 2 * lc) + e] = sq;
        }
      }
      consumer_sync();
      for (int c = tid; c < BN; c += NC * 32) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS_M; ++w) tot += red[w * BN + c];
        if (n0 + c < p.N) p.sumsq[b * p.sSumsq + (long long)tl.by * p.N + n0 + c] = tot;
      }
      consumer_sync();
    }

    if (p.wsum) {
      const double* wv = p.wvec + b * p.sW;
      for (int q = 0; q < p.nw; ++q) {
        double wgt[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int r = m0 + frag_row<AKM, WARPS_M>(wm, i, lr);
          wgt[i] = (r < p.M) ? wv[(long long)r * p.ldw + q] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double ws = 0.0;
#pragma unroll
            for (int i = 0; i < MI; ++i) ws += acc[i][j][e] * wgt[i];
            ws += __shfl_xor_sync(0xffffffffu, ws, 4);
            ws += __shfl_xor_sync(0xffffffffu, ws, 8);
            ws += __shfl_xor_sync(0xffffffffu, ws, 16);
            if (lr == 0) red[wm * BN + frag_row<BKM, WARPS_N>(wn, j, 2 * lc) + e] = ws;
          }
        }
        consumer_sync();
        for (int c = tid; c < BN; c += NC * 32) {
          double tot = 0.0;
#pragma unroll
          for (int w = 0; w < WARPS_M; ++w) tot += red[w * BN + c];
          if (n0 + c < p.N) p.wsum[b * p.sWsum + ((long long)tl.by * p.nw + q) * p.N + n0 + c] = tot;
        }
        consumer_sync();
      }
    }
  }
}

// ---------------------------------------------------------------------------- host side
PFN_cuTensorMapEncodeTiled encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    cudaGetLastError();
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
  }();
  return fn;
}

// tensor map of one operand: rows = non-contracted extent, kmajor as in the kernel
int make_map(CUtensorMap* map, const double* base, bool kmajor, int rows, int K, int ld, int batch, long long sbatch,
             int box_rows, bool atom32) {
  PFN_cuTensorMapEncodeTiled enc = encode_fn();
  if (!enc) {
    set_last_error("gemm_tma: cuTensorMapEncodeTiled is not available from the driver");
    return GPFX_E_CUDA;
  }
  const cuuint64_t nb = (batch > 1 && sbatch != 0) ? (cuuint64_t)batch : 1;
  const cuuint64_t sb = (nb > 1) ? (cuuint64_t)sbatch * 8 : (cuuint64_t)16;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  int rank;
  if (kmajor) {
    rank = 3;
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; dims[2] = nb;
    strides[0] = (cuuint64_t)ld * 8; strides[1] = sb;
    box[0] = TBK; box[1] = (cuuint32_t)box_rows; box[2] = 1;
  } else {
    rank = 4;
    dims[0] = 16; dims[1] = (cuuint64_t)K; dims[2] = (cuuint64_t)(rows / 16); dims[3] = nb;
    strides[0] = (cuuint64_t)ld * 8; strides[1] = 128; strides[2] = sb;
    box[0] = 16; box[1] = TBK; box[2] = (cuuint32_t)(box_rows / 16); box[3] = 1;
  }
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<double*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         (kmajor && atom32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("gemm_tma: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d batch=%d", (int)r, rows, K, ld,
                   batch);
    return GPFX_E_CUDA;
  }
  return GPFX_OK;
}

int sm_count(int* out) {
  static std::atomic<int> cached[64];
  int dev = 0;
  GPFX_CUDA_CHECK(cudaGetDevice(&dev));
  int v = cached[dev & 63].load(std::memory_order_relaxed);
  if (v == 0) {
    GPFX_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    cached[dev & 63].store(v, std::memory_order_relaxed);
  }
  *out = v;
  return GPFX_OK;
}

// k-major operands: 128B swizzle with 32-byte atoms (conflict-free fragment reads) when the driver
// encodes it for FP64, the plain 128B swizzle otherwise.  GPFX_GEMM_TMA_ATOM32=0 forces the latter.
bool use_atom32(const void* probe) {
  static const bool ok = [probe] {
    const char* e = getenv("GPFX_GEMM_TMA_ATOM32");
    if (e && atoi(e) == 0) return false;
    PFN_cuTensorMapEncodeTiled enc = encode_fn();
    if (!enc) return false;
    alignas(64) CUtensorMap m;
    cuuint64_t dims[3] = {64, 64, 1}, strides[2] = {512, 512 * 64};
    cuuint32_t box[3] = {16, 64, 1}, estr[3] = {1, 1, 1};
    // (encoding touches no memory: any 16-byte-aligned device address will do)
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(probe), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
  }();
  return ok;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int MINB>
int launch_tma_cfg(const GemmArgs& a, cudaStream_t stream) {
  using Cfg = TmaCfg<BM, BN, WARPS_M, WARPS_N, STAGES, MINB>;
  alignas(64) CUtensorMap mapA, mapB;
  const bool akm = !a.transA, bkm = a.transB != 0;
  const bool atom32 = use_atom32(a.A);
  GPFX_TRY(make_map(&mapA, a.A, akm, a.M, a.K, a.lda, a.batch1, a.sA1, BM, atom32));
  GPFX_TRY(make_map(&mapB, a.B, bkm, a.N, a.K, a.ldb, a.batch1, a.sB1, BN, atom32));
  const int bmulA = (a.batch1 > 1 && a.sA1 != 0) ? 1 : 0;
  const int bmulB = (a.batch1 > 1 && a.sB1 != 0) ? 1 : 0;
  const int sw32 = atom32 ? 1 : 0;
  const int tx = ceil_div(a.N, BN), ty = ceil_div(a.M, BM);
  const long long nbz = (long long)a.batch1 * a.splits;
  const long long tiles = (long long)tx * ty * nbz;
  if (tiles > 0x7fffffffLL) {
    set_last_error("gemm_tma: too many tiles");
    return GPFX_E_SHAPE;
  }
  int sms = 0;
  GPFX_TRY(sm_count(&sms));
  // CTA lifetime ~ 250 us: a k-tile takes ~2.1 us of an SM's DMMA pipe (128x128x16, or two co-resident
  // 64x128x16), so a CTA keeps its SM for tmax tiles and then makes room (GPFX_GEMM_TMA_TMAX overrides)
  static const int env_tmax = [] {
    const char* e = getenv("GPFX_GEMM_TMA_TMAX");
    return e ? atoi(e) : 0;
  }();
  const double ktiles = std::max(1.0, (double)a.K / TBK * ((a.tri & (TRI_LOWER_A | TRI_UPPER_A)) ? 0.55 : 1.0) / a.splits);
  // measured (tools/gemm_tma_check.py, c3 step): one tile per CTA pays when a tile alone lasts that long
  // (the K = 8192 tril products: +3 %, the hardware block scheduler then balances 1 ms tiles); shorter
  // tiles lose 4 % to the CTA restarts, so those grids stay resident to the end
  int tmax = env_tmax > 0 ? env_tmax : ((2.1 * ktiles >= 250.0) ? 1 : 0x3fffffff);
  const long long resident = std::min<long long>(tiles, (long long)sms * MINB);
  dim3 grid((unsigned)std::max<long long>(resident, ceil_div_ll(tiles, tmax)));
  dim3 block(Cfg::NT);
  static std::atomic<unsigned> launch_seq{0};
  const int slot = (int)(launch_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
#define GPFX_TMA_LAUNCH(AKM, BKM, ADDIN, KSCALE)                                                     \
  do {                                                                                               \
    auto kern = gemm_f64_tma_kernel<BM, BN, WARPS_M, WARPS_N, AKM, BKM, STAGES, MINB, ADDIN, KSCALE>; \
    static SmemAttrOnce once;                                                                \
    GPFX_TRY(ensure_dyn_smem(kern, Cfg::SMEM, once));                                        \
    kern<<<grid, block, Cfg::SMEM, stream>>>(mapA, mapB, a, bmulA, bmulB, sw32, slot, tmax); \
  } while (0)
  if (a.addmat != nullptr) {
    // add-in epilogue: the layer backward's dA product (k-major A, mn-major B, 64x128 tiles) only
    if constexpr (BM == 64) {
      if (akm && !bkm) GPFX_TMA_LAUNCH(true, false, true, false);
      else {
        set_last_error("gemm_tma: add-in terms need an NN product");
        return GPFX_E_UNSUPPORTED;
      }
    } else {
      set_last_error("gemm_tma: add-in terms need the 64x128 variant");
      return GPFX_E_UNSUPPORTED;
    }
  } else if (a.kscale != nullptr) {
    // per-k scale: the layer backward's dLq product (both operands k-major, 128x128 tiles) only
    if constexpr (BM == 128) {
      if (akm && bkm) GPFX_TMA_LAUNCH(true, true, false, true);
      else {
        set_last_error("gemm_tma: a per-k scale needs an NT product");
        return GPFX_E_UNSUPPORTED;
      }
    } else {
      set_last_error("gemm_tma: a per-k scale needs the 128x128 variant");
      return GPFX_E_UNSUPPORTED;
    }
  } else if (akm) {
    if (bkm) GPFX_TMA_LAUNCH(true, true, false, false);
    else GPFX_TMA_LAUNCH(true, false, false, false);
  } else {
    if (bkm) GPFX_TMA_LAUNCH(false, true, false, false);
    else GPFX_TMA_LAUNCH(false, false, false, false);
  }
#undef GPFX_TMA_LAUNCH
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

// 0: off; 1..: kernel variant.  gpfx_tune_gemm_config(100 + v) overrides the GPFX_GEMM_TMA default;
// forcing one of the cp.async tile configurations (0..7) switches the TMA kernel off as well.
int tma_env_cfg() {
  static const int v = [] {
    const char* e = getenv("GPFX_GEMM_TMA");
    return e ? atoi(e) : 1;
  }();
  const int fc = forced_gemm_cfg();
  if (fc >= 100) return fc - 100;
  if (fc >= 0) return 0;
  return v;
}

}  // namespace

// Kernel variant for a problem, from its shape only (no pointers: callers size their column-partial
// buffers from the tile height).  0: stays on the cp.async kernel.
//   1 = 128x128 CTA tile, 8 consumer warps (64x32 warp tiles), 6 stages, one CTA per SM
//       dense products and tril-output products (static masks on the diagonal tiles)
//   2 = 64x128 CTA tile, 4 consumer warps (64x32), 4 stages, two CTAs per SM
//       triangular-A products (64-row k-ranges, compile-time masks in the diagonal block; the second
//       CTA covers the other one's epilogue)
// Small grids (less than two tiles per resident CTA) stay on the cp.async kernel, whose 32-row tiles
// balance better there (measured: tools/gemm_tma_check.py, c2 "A = Linv Kuf").
static int tma_variant(const GemmArgs& a) {
  const int mode = tma_env_cfg();
  if (mode <= 0) return 0;
  if (a.cfg >= 0 || a.small_tile) return 0;
  if (a.batch2 != 1 || a.nseg != 1) return 0;
  if (a.M < 128 || a.N < 128 || a.K < 64) return 0;
  if ((a.lda & 1) || (a.ldb & 1) || (a.sA1 & 1) || (a.sB1 & 1)) return 0;
  if (a.transA && (a.M & 15)) return 0;   // mn-major A: whole 16-row slabs
  if (!a.transB && (a.N & 15)) return 0;  // mn-major B
  if (forced_gemm_cfg() >= 101) return mode;   // tuning hook: this variant for every eligible problem
  const int v = (a.tri & (TRI_LOWER_A | TRI_UPPER_A)) ? 2 : 1;
  const long long bm = (v == 2) ? 64 : 128;
  const long long tiles = ceil_div_ll(a.M, bm) * ceil_div_ll(a.N, 128) * a.batch1 * (a.splits < 1 ? 1 : a.splits);
  if (tiles < 2LL * 148 * (v == 2 ? 2 : 1)) return 0;
  return v;
}

int gemm_tma_block_m(const GemmArgs& a) {
  const int v = tma_variant(a);
  return v == 0 ? 0 : (v == 2 ? 64 : 128);
}

int launch_gemm_tma(const GemmArgs& a, cudaStream_t stream) {
  switch (tma_variant(a)) {
    case 2: return launch_tma_cfg<64, 128, 1, 4, 4, 2>(a, stream);
    case 1: return launch_tma_cfg<128, 128, 2, 4, 6, 1>(a, stream);
    default: set_last_error("launch_gemm_tma: problem is not eligible"); return GPFX_E_UNSUPPORTED;
  }
}

bool gemm_addin_supported(const GemmArgs& a) {
  const int v = tma_variant(a);
  if (v != 2 || a.transA || a.transB || !gemm_tma_pointers_ok(a)) return false;  // the instantiation that exists
  const int bm = 64;
  if (a.M % bm != 0 || a.N % 128 != 0) return false;  // every tile takes the plain epilogue
  if (a.tril_out || a.diag_scale != 1.0 || a.splits > 1 || a.beta != 0.0 || a.kscale != nullptr) return false;
  if (a.colscale != nullptr && ((a.sColscale & 1) || ((uintptr_t)a.colscale & 15))) return false;
  if ((a.ldc & 1) || (a.sC1 & 1) || ((uintptr_t)a.C & 15)) return false;
  if (a.addmat == nullptr || a.addcol == nullptr || a.r1u == nullptr || a.r1v == nullptr) return false;
  if ((a.ldadd & 1) || (a.sAdd & 1) || ((uintptr_t)a.addmat & 15)) return false;
  if ((a.sAddcol & 1) || ((uintptr_t)a.addcol & 15) || (a.sR1v & 1) || ((uintptr_t)a.r1v & 15)) return false;
  return true;
}

bool gemm_kscale_supported(const GemmArgs& a) {
  if (tma_variant(a) != 1 || a.transA || !a.transB || !gemm_tma_pointers_ok(a)) return false;  // the instantiation that exists
  if (a.kscale == nullptr || a.addmat != nullptr) return false;
  if ((a.K % TBK) != 0) return false;  // whole 128-byte slices of the scale vector
  if ((a.sKscale & 1) || ((uintptr_t)a.kscale & 15)) return false;
  return true;
}

bool gemm_tma_pointers_ok(const GemmArgs& a) {
  return ((((uintptr_t)a.A) | ((uintptr_t)a.B)) & 15) == 0;
}

}  // namespace gpfx
