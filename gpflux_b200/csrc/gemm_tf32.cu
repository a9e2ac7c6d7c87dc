// float32 opt-in contraction path: 3xTF32 GEMM on the 5th-generation tensor cores.
//
//   C[b] (f64) = alpha * op(A[b]) op(B[b]) + beta * C[b]          (optionally lower triangle only)
//
// The float64 operands are first split into two TF32-representable float32 terms each
// (x ~ hi + lo, hi = tf32(x), lo = tf32(x - hi): 22 significand bits) and laid out k-major
// ([rows][K], transposing where the source is mn-major) by split_tf32_kernel.  The product is then
// hi*hi + hi*lo + lo*hi with float32 accumulation in tensor memory - the classical 3xTF32 scheme,
// ~5e-7 relative per product, inside the 1e-4 the float32 opt-in promises - at three
// tcgen05.mma.kind::tf32 instructions per k step instead of FP64 DMMA.
//
// Kernel anatomy (one CTA per SM, persistent over 128 x 128 output tiles, 192 threads):
//   warp 0      TMA producer: four boxes per stage (A_hi, A_lo, B_hi, B_lo; 128 rows x 32 tf32 = 128 B,
//               SWIZZLE_128B) into a 3-stage ring, full / empty mbarriers
//   warp 1      MMA issuer: one elected lane issues 12 tcgen05.mma (4 k-steps of 8 x 3 products) per
//               stage from shared-memory descriptors (k-major, 128B swizzle, SBO = 1024 B), accumulating
//               into one of two 128-column TMEM accumulator stages; tcgen05.commit releases the smem
//               stage / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns per warp and call) -> float64, alpha / beta /
//               tril, stores to C; overlaps the next tile's MMAs through the second accumulator stage
//
// Replaces, for the float32 opt-in only, the same tf.matmul / triangular_solve calls as gemm_f64*.cu
// (reference gpflux/layers/gp_layer.py:257-266 under gpflow.config.set_default_float(np.float32)).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <atomic>

#include "../../include/gpfx.h"
#include "gemm_f64.h"
#include "gemm_tf32.h"

#include "context.h"

namespace gpfx {
namespace {

constexpr int TM = 128, TN = 128, TK = 32;  // CTA tile; TK floats = one 128-byte swizzle row
constexpr int TSTAGES = 3;
constexpr int OP_BYTES = TM * TK * 4;             // one operand term of a stage: 16 KB
constexpr int STAGE_BYTES = 4 * OP_BYTES;         // A_hi, A_lo, B_hi, B_lo
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * TN;        // 256 columns of 128 lanes x 32 bit
constexpr int EPI_STAGE_FLOATS = 32 * 33;           // per epilogue warp: a 32 x 32 float tile, padded rows
constexpr size_t TF32_SMEM = (size_t)TSTAGES * STAGE_BYTES + 1024 + (2 * TSTAGES + 2 * ACC_STAGES) * 8 + 16 +
                             4 * EPI_STAGE_FLOATS * sizeof(float);

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(s32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                       int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];\n" ::"r"(s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// shared-memory matrix descriptor: k-major operand tile of 128-byte rows under the 128B swizzle
// (canonical ((8,n),2):((8,SBO),1) in 16-byte units): start address >> 4, LBO = 1 (unused), SBO = 1024 B
// between 8-row groups, descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ unsigned long long make_desc(unsigned smem_addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)(1024 >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)2 << 61;
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both k-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr unsigned kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TN >> 3) << 17) | ((unsigned)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(bar))
               : "memory");
}

struct Tf32Args {
  double* C;
  long long sC;
  int ldc;
  int M, N, K, batch;
  double alpha, beta;
  int tril_out;
  int tri;  // TRI_* flags of gemm_f64.h: they only shorten the k-range per tile here
  int k0, k1;  // this launch contracts k in [k0, k1) (multiples of TK): long contractions are cut into chunks
               // whose float32 partial sums are added in float64 (beta = 1) - the tensor core accumulates
               // with truncation, so the error of one chunk grows linearly with its length
};

__global__ void __launch_bounds__(192, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                       const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
                       const Tf32Args p) {
  extern __shared__ char smem_raw[];
  char* tiles = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(tiles + TSTAGES * STAGE_BYTES);
  unsigned long long* empty = full + TSTAGES;
  unsigned long long* acc_full = empty + TSTAGES;
  unsigned long long* acc_empty = acc_full + ACC_STAGES;
  unsigned* tmem_base_slot = reinterpret_cast<unsigned*>(acc_empty + ACC_STAGES);
  float* epi_stage = reinterpret_cast<float*>(tmem_base_slot + 4);  // [4 warps][32][33]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  const int tiles_x = (p.N + TN - 1) / TN, tiles_y = (p.M + TM - 1) / TM;
  const int total = tiles_x * tiles_y * p.batch;
  struct Tile {
    int m0, n0, b, kb0, kb1;
  };
  auto decode = [&](int t) -> Tile {
    Tile tl;
    const int per_b = tiles_x * tiles_y;
    tl.b = t / per_b;
    const int r = t - tl.b * per_b;
    const int by = r / tiles_x, bx = r - by * tiles_x;
    tl.m0 = by * TM;
    tl.n0 = bx * TN;
    int klo = p.k0, khi = min(p.K, p.k1);
    if (p.tri & TRI_UPPER_A) klo = max(klo, tl.m0);
    if (p.tri & TRI_LOWER_B) klo = max(klo, tl.n0);
    if (p.tri & TRI_LOWER_A) khi = min(khi, tl.m0 + TM);
    if (p.tri & TRI_UPPER_B) khi = min(khi, tl.n0 + TN);
    const bool dead = p.tril_out && (tl.n0 >= tl.m0 + TM);
    tl.kb0 = klo / TK;
    tl.kb1 = (khi > klo && !dead) ? (khi + TK - 1) / TK : tl.kb0;
    return tl;
  };

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mAhi) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&mBlo) : "memory");
    for (int s = 0; s < TSTAGES; ++s) {
      mb_init(&full[s], 1);
      mb_init(&empty[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mb_init(&acc_full[s], 1);
      mb_init(&acc_empty[s], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (this warp also frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      unsigned phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const Tile tl = decode(t);
        for (int kb = tl.kb0; kb < tl.kb1; ++kb) {
          mb_wait(&empty[stage], phase ^ 1u);
          mb_expect_tx(&full[stage], STAGE_BYTES);
          char* st = tiles + stage * STAGE_BYTES;
          tma_3d(st, &mAhi, &full[stage], kb * TK, tl.m0, tl.b);
          tma_3d(st + OP_BYTES, &mAlo, &full[stage], kb * TK, tl.m0, tl.b);
          tma_3d(st + 2 * OP_BYTES, &mBhi, &full[stage], kb * TK, tl.n0, tl.b);
          tma_3d(st + 3 * OP_BYTES, &mBlo, &full[stage], kb * TK, tl.n0, tl.b);
          if (++stage == TSTAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0, as = 0;
      unsigned phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const Tile tl = decode(t);
        mb_wait(&acc_empty[as], aphase ^ 1u);  // the epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const unsigned tmem_d = tmem_base + (unsigned)(as * TN);
        if (tl.kb1 == tl.kb0) {
          mb_arrive(&acc_full[as]);  // nothing to contract: the epilogue takes acc = 0
        } else {
          for (int kb = tl.kb0; kb < tl.kb1; ++kb) {
            mb_wait(&full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const unsigned st = s32(tiles + stage * STAGE_BYTES);
            const unsigned long long dAhi = make_desc(st), dAlo = make_desc(st + OP_BYTES);
            const unsigned long long dBhi = make_desc(st + 2 * OP_BYTES), dBlo = make_desc(st + 3 * OP_BYTES);
#pragma unroll
            for (int k8 = 0; k8 < TK / 8; ++k8) {
              const unsigned long long adv = (unsigned long long)(k8 * 32 >> 4);  // 8 tf32 = 32 bytes along k
              const unsigned first = (kb == tl.kb0 && k8 == 0) ? 0u : 1u;
              umma_tf32(tmem_d, dAlo + adv, dBhi + adv, first);  // small terms first
              umma_tf32(tmem_d, dAhi + adv, dBlo + adv, 1u);
              umma_tf32(tmem_d, dAhi + adv, dBhi + adv, 1u);
            }
            umma_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
            if (++stage == TSTAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&acc_full[as]);  // accumulator complete
        }
        if (++as == ACC_STAGES) {
          as = 0;
          aphase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    const int quad = warp & 3;  // TMEM lanes 32*quad .. 32*quad+31 are the ones this warp may touch
    int as = 0;
    unsigned aphase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const Tile tl = decode(t);
      mb_wait(&acc_full[as], aphase);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // TMEM hands lane l its row (32 consecutive columns per load); the tile is turned through shared memory so
      // that each store instruction of the warp writes 32 consecutive columns of ONE row (256 contiguous bytes)
      float* stg = epi_stage + (warp - 2) * EPI_STAGE_FLOATS;
      double* cbase = p.C + (long long)tl.b * p.sC;
      const bool have = tl.kb1 > tl.kb0;
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 32) {
        unsigned v[32];
        if (have) {
          const unsigned taddr = tmem_base + ((unsigned)(quad * 32) << 16) + (unsigned)(as * TN + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
              "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                "=r"(v[30]), "=r"(v[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int c = tl.n0 + c0 + lane;
        if (c < p.N) {
          const int rbase = tl.m0 + quad * 32;
          double* dcol = cbase + (long long)rbase * p.ldc + c;
#pragma unroll 1
          for (int r0 = 0; r0 < 32; r0 += 16) {
            double old[16];
            if (p.beta != 0.0) {  // the 16 reads of the beta term go out together (independent L2 round trips)
#pragma unroll
              for (int u = 0; u < 16; ++u) old[u] = (rbase + r0 + u < p.M) ? dcol[(long long)(r0 + u) * p.ldc] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int r = rbase + r0 + u;
              if (r >= p.M) break;
              double x = p.alpha * (double)stg[(r0 + u) * 33 + lane];
              if (p.tril_out && c > r) x = 0.0;
              if (p.beta != 0.0) x += p.beta * old[u];
              dcol[(long long)(r0 + u) * p.ldc] = x;
            }
          }
        }
        __syncwarp();
      }
      // this warp is done reading the accumulator stage
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive(&acc_empty[as]);
      if (++as == ACC_STAGES) {
        as = 0;
        aphase ^= 1u;
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- operand split: f64 [rows][cols] (ld) -> hi, lo float32, k-major [R][Kp] ---------------------------------
// TRANSPOSE == false: out[r][c] = split(in[r][c]);  true: out[c][r] = split(in[r][c]).  32 x 32 tiles through
// shared memory keep both the reads and the writes coalesced.  Columns in [K, Kp) of the output are zero.
template <bool TRANSPOSE>
__global__ void __launch_bounds__(256) split_tf32_kernel(const double* __restrict__ in, long long s_in, int ld_in, int rows,
                                                         int cols, float* __restrict__ hi, float* __restrict__ lo,
                                                         long long s_out, int ld_out) {
  __shared__ float sh[32][33], sl[32][33];
  const int b = blockIdx.z;
  const double* src = in + (long long)b * s_in;
  float* H = hi + (long long)b * s_out;
  float* L = lo + (long long)b * s_out;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float h = 0.f, l = 0.f;
    if (r < rows && c < cols) {
      const double x = src[(long long)r * ld_in + c];
      unsigned hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hb) : "f"((float)x));
      h = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lb) : "f"((float)(x - (double)h)));
      l = __uint_as_float(lb);
    }
    if (TRANSPOSE) {
      sh[ty + 8 * i][tx] = h;
      sl[ty + 8 * i][tx] = l;
    } else if (r < rows && c < ld_out) {
      H[(long long)r * ld_out + c] = h;
      L[(long long)r * ld_out + c] = l;
    }
  }
  if (TRANSPOSE) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;  // output row = input column
      if (c < cols && r < ld_out) {
        H[(long long)c * ld_out + r] = sh[tx][ty + 8 * i];
        L[(long long)c * ld_out + r] = sl[tx][ty + 8 * i];
      }
    }
  }
}

PFN_cuTensorMapEncodeTiled tf32_encode_fn() {
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

int make_map_f32(CUtensorMap* map, const float* base, int rows, int K, int ld, int batch, long long sbatch) {
  PFN_cuTensorMapEncodeTiled enc = tf32_encode_fn();
  if (!enc) {
    set_last_error("gemm_tf32: cuTensorMapEncodeTiled is not available from the driver");
    return GPFX_E_CUDA;
  }
  const cuuint64_t nb = (batch > 1 && sbatch != 0) ? (cuuint64_t)batch : 1;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, nb};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, nb > 1 ? (cuuint64_t)sbatch * 4 : (cuuint64_t)16};
  cuuint32_t box[3] = {TK, TM, 1}, estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("gemm_tf32: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d batch=%d", (int)r, rows, K, ld, batch);
    return GPFX_E_CUDA;
  }
  return GPFX_OK;
}

}  // namespace

size_t gemm_tf32_workspace_floats(int batch, int M, int N, int K) {
  const size_t Kp = (size_t)(K + 3) / 4 * 4;
  return 2 * (size_t)batch * ((size_t)M + (size_t)N) * Kp + 64;
}

int launch_gemm_tf32x3(const GemmArgs& a, float* ws, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.batch1 <= 0) return GPFX_OK;
  if (a.batch2 != 1 || a.nseg != 1 || a.splits > 1 || a.colscale || a.kscale || a.sumsq || a.wsum || a.addmat || a.diag_scale != 1.0) {
    set_last_error("gemm_tf32: only plain (alpha, beta, tril) products are available on the float32 path");
    return GPFX_E_UNSUPPORTED;
  }
  const int batch = a.batch1, M = a.M, N = a.N, K = a.K;
  const int Kp = (K + 3) / 4 * 4;
  const int nA = (a.sA1 == 0 && batch > 1) ? 1 : batch, nB = (a.sB1 == 0 && batch > 1) ? 1 : batch;
  // (a shared operand is split once; its tensor map has a single batch entry and the kernel's batch
  // coordinate must then be 0 - not wired up: the float32 path is used with per-batch operands)
  if ((nA == 1 || nB == 1) && batch > 1) {
    set_last_error("gemm_tf32: shared (stride-0) operands with batch > 1 are not supported");
    return GPFX_E_UNSUPPORTED;
  }
  float* Ahi = ws;
  float* Alo = Ahi + (size_t)nA * M * Kp;
  float* Bhi = Alo + (size_t)nA * M * Kp;
  float* Blo = Bhi + (size_t)nB * N * Kp;
  {  // op(A)[m][k]: stored [M][K] (k-major) or, transA, [K][M]
    if (a.transA) {
      dim3 g(ceil_div(M, 32), ceil_div(K, 32), nA);
      split_tf32_kernel<true><<<g, 256, 0, stream>>>(a.A, a.sA1, a.lda, K, M, Ahi, Alo, (long long)M * Kp, Kp);
    } else {
      dim3 g(ceil_div(Kp, 32), ceil_div(M, 32), nA);
      split_tf32_kernel<false><<<g, 256, 0, stream>>>(a.A, a.sA1, a.lda, M, K, Ahi, Alo, (long long)M * Kp, Kp);
    }
    GPFX_LAUNCH_CHECK();
  }
  {  // op(B)[k][n] is needed as [N][K]: stored [N][K] when transB, else [K][N]
    if (a.transB) {
      dim3 g(ceil_div(Kp, 32), ceil_div(N, 32), nB);
      split_tf32_kernel<false><<<g, 256, 0, stream>>>(a.B, a.sB1, a.ldb, N, K, Bhi, Blo, (long long)N * Kp, Kp);
    } else {
      dim3 g(ceil_div(N, 32), ceil_div(K, 32), nB);
      split_tf32_kernel<true><<<g, 256, 0, stream>>>(a.B, a.sB1, a.ldb, K, N, Bhi, Blo, (long long)N * Kp, Kp);
    }
    GPFX_LAUNCH_CHECK();
  }
  alignas(64) CUtensorMap mAhi, mAlo, mBhi, mBlo;
  const long long sA = (nA > 1) ? (long long)M * Kp : 0, sB = (nB > 1) ? (long long)N * Kp : 0;
  GPFX_TRY(make_map_f32(&mAhi, Ahi, M, Kp, Kp, batch, sA));
  GPFX_TRY(make_map_f32(&mAlo, Alo, M, Kp, Kp, batch, sA));
  GPFX_TRY(make_map_f32(&mBhi, Bhi, N, Kp, Kp, batch, sB));
  GPFX_TRY(make_map_f32(&mBlo, Blo, N, Kp, Kp, batch, sB));
  Tf32Args p;
  p.C = a.C; p.sC = a.sC1; p.ldc = a.ldc;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.alpha = a.alpha; p.beta = a.beta; p.tril_out = a.tril_out; p.tri = a.tri;
  static SmemAttrOnce once;
  GPFX_TRY(ensure_dyn_smem(gemm_tf32x3_kernel, TF32_SMEM, once));
  int dev = 0, sms = 0;
  GPFX_CUDA_CHECK(cudaGetDevice(&dev));
  GPFX_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (long long)ceil_div(M, TM) * ceil_div(N, TN) * batch;
  const int grid = (int)std::min<long long>(tiles, sms);
  // k-chunks of kChunk: float32 accumulation inside a chunk, float64 across chunks
  constexpr int kChunk = 1024;
  for (int k0 = 0; k0 < K; k0 += kChunk) {
    p.k0 = k0;
    p.k1 = std::min(K, k0 + kChunk);
    if (k0 > 0) p.beta = 1.0;
    gemm_tf32x3_kernel<<<grid, 192, TF32_SMEM, stream>>>(mAhi, mAlo, mBhi, mBlo, p);
    GPFX_LAUNCH_CHECK();
  }
  return GPFX_OK;
}

bool gemm_tf32_eligible(const GemmArgs& a) {
  if (a.batch2 != 1 || a.nseg != 1 || a.colscale || a.kscale || a.diag_scale != 1.0) return false;
  if (a.M < 128 || a.N < 128 || a.K < 32) return false;
  if (a.batch1 > 1 && (a.sA1 == 0 || a.sB1 == 0)) return false;  // shared operands: not wired up
  return true;
}

bool gemm_tf32_active(const GemmArgs& a) {
  Context* c = current_context();
  if (c->contraction != 1 || c->tf32_ws == nullptr || !gemm_tf32_eligible(a)) return false;
  return gemm_tf32_workspace_floats(a.batch1, a.M, a.N, a.K) * sizeof(float) <= c->tf32_ws_bytes;
}

int launch_gemm_auto(const GemmArgs& a_in, cudaStream_t stream) {
  if (!gemm_tf32_active(a_in) || a_in.sumsq || a_in.wsum || a_in.addmat) return launch_gemm(a_in, stream);
  GemmArgs a = a_in;
  a.splits = 1;  // the k-chunks of the 3xTF32 kernel replace split-K
  a.splitws = nullptr;
  return launch_gemm_tf32x3(a, current_context()->tf32_ws, stream);
}

}  // namespace gpfx
