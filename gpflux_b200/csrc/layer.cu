// GPLayer forward / backward orchestration: stream-ordered, no host syncs, caller-owned buffers.
//
// The layer is cut into a PARAMETER-ONLY part (factor) and a DATA part (cond) so that a caller
// can run the factor parts of all layers concurrently with the data path (they do not depend on
// the minibatch), on other streams or as parallel branches of a CUDA graph:
//
//   factor forward  (gpflux/layers/gp_layer.py:303-311 and the Kuu/Cholesky half of :257-266)
//     Lq = tril(q_sqrt), KL                       kl_tril            (a9)
//     Kuu + jitter I  -> L, L^-1                  kmat_fwd, potrf    (a1, a3)
//   cond forward    (gp_layer.py:226-268, 339-363)
//     Kuf                                         kmat_fwd           (a2)
//     A   = L^-1 Kuf      (+ sum A^2, A^T q_mu)   DMMA GEMM, lower-A (a4, a5, a6)
//     B_p = Lq_p^T A      (+ sum B^2)             DMMA GEMM, upper-A (a7)
//     mean, var, f = mean + sqrt(var) eps         finalize           (a8)
//   cond backward   (a14)  dLq_p = tril(A Bs_p^T), dA = sum_p Lq_p Bs_p + q_mu gm^T - 2 A diag(gv),
//                          dKuf = L^-T dA, dL = -tril(dKuf A^T), Kuf kernel-matrix backward
//   factor backward (a14)  dKuu = L^-T Phi(L^T dL) L^-1, Kuu kernel-matrix backward, KL backward
//
// gpfx_layer_forward / gpfx_layer_backward run both parts back to back on one stream.
#include "layer.h"
#include "context.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "chol.h"
#include "elementwise.h"
#include "gemm_f64.h"
#include "gemm_tf32.h"
#include "kmat.h"
#include "trsm.h"

namespace gpfx {

// Which products of the cond path the 3xTF32 contraction mode takes (bit 0: A = L^-1 Kuf, 1: B_p = Lq_p^T A,
// 2: dLq, 3: dA, 4: dKuf = L^-T dA, 5: dL).  Default 14 = B_p, dLq, dA.  The products with L^-1 as an operand
// (A, dKuf) and the one that feeds the Cholesky adjoint (dL) stay on the FP64 pipe: their rounding error is
// amplified by cond(Kuu) - measured at cond(Kuu) = 2e7 (tools/tf32_probe.py): A in 3xTF32 puts dZ off by 1.6e-3
// and dL by 9e-4, while B_p / dLq / dA together keep every output and gradient within 6e-6 of float64.
// GPFX_TF32_PRODUCTS overrides the mask (diagnostics).
static int tf32_product_mask() {
  static const int m = [] {
    const char* e = getenv("GPFX_TF32_PRODUCTS");
    return e ? atoi(e) : 14;
  }();
  return m;
}

namespace {

struct Carver {
  size_t off = 0;
  size_t take(size_t doubles) {
    const size_t o = off;
    off += align_up(doubles * sizeof(double), 256);
    return o;
  }
};

struct Plan {
  int N, M, D, P, K, Mp;
  // factor ctx
  size_t f_Lq, f_L, f_Linv, f_info, f_alpha, fctx_bytes;
  // cond ctx
  size_t c_Kuf, c_A, c_B, cctx_bytes;
  // factor workspace
  size_t fw_T, fw_kl, fw_T1, fw_T2, fw_dKuu, fw_part, fw_Lqtmp, fw_Wa, fw_dL2, fws_bytes;
  // cond workspace
  size_t cw_sqA, cw_wsum, cw_sqB, cw_gmT, cw_gv2T, cw_gvsum, cw_dA, cw_dKuf, cw_split, cw_part,
      cws_bytes;
  size_t dL_bytes;
  int gridM_A, gridM_B;
  bool tf32A = false, tf32B = false;  // the A / B_p products run on the 3xTF32 kernel (unfused column reductions)
  int splits_q, splits_l, splits_m;
  size_t cw_dqmu;
  GemmArgs gA, gB;  // forward GEMM templates (pointers filled later)
};

// split-K factor for the [M,M] = [M,N] x [N,M] lower-triangular-output contractions: enough CTAs
// for ~2 waves (live lower tiles only), at least 512 k per split
int pick_splits(int batch, int M, int N) {
  const int edge = (M <= 64) ? 64 : 128;
  const int t = ceil_div(M, edge);
  const long long live = (long long)batch * t * (t + 1) / 2;
  if (live >= 148) return 1;
  long long s = ceil_div_ll(296, live > 0 ? live : 1);
  const long long smax = N / 512 > 1 ? N / 512 : 1;
  if (s > smax) s = smax;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return (int)s;
}

int make_plan(const gpfx_layer_desc* d, Plan& pl) {
  if (d == nullptr) return GPFX_E_SHAPE;
  if (d->N <= 0 || d->M <= 0 || d->D <= 0 || d->P <= 0 || !(d->K == 1 || d->K == d->P)) {
    set_last_error("gpfx_layer: bad shape N=%d M=%d D=%d P=%d K=%d (K must be 1 or P)", d->N,
                   d->M, d->D, d->P, d->K);
    return GPFX_E_SHAPE;
  }
  if (d->solver != GPFX_SOLVER_INVERSE_GEMM && d->solver != GPFX_SOLVER_TRSM) {
    set_last_error("gpfx_layer: unknown solver id %d", d->solver);
    return GPFX_E_UNSUPPORTED;
  }
  if (d->kernel < GPFX_KERNEL_SE || d->kernel > GPFX_KERNEL_MATERN52) {
    set_last_error("gpfx_layer: unknown kernel id %d", d->kernel);
    return GPFX_E_UNSUPPORTED;
  }
  pl.N = d->N; pl.M = d->M; pl.D = d->D; pl.P = d->P; pl.K = d->K;
  pl.Mp = chol_padded_dim(d->M);
  const size_t N = d->N, M = d->M, P = d->P, K = d->K, Mp = pl.Mp;

  {
    Carver c;
    pl.f_Lq = c.take(P * M * M);
    pl.f_L = c.take(K * Mp * Mp);
    pl.f_Linv = c.take(K * Mp * Mp);
    pl.f_info = c.take((K + 1) / 2 + 1);
    pl.f_alpha = c.take(d->whiten ? 0 : M * P);  // non-whitened: alpha = L^-1 q_mu (and f_Lq holds L^-1 Lq)
    pl.fctx_bytes = c.off;
  }
  {
    Carver c;
    pl.c_Kuf = c.take(K * M * N);
    pl.c_A = c.take(K * M * N);
    pl.c_B = c.take(P * M * N);
    pl.cctx_bytes = c.off;
  }

  // forward GEMM shapes decide the tile size, hence the number of column partials
  GemmArgs& a = pl.gA;
  a = GemmArgs();
  a.M = d->M; a.N = d->N; a.K = d->M;
  a.lda = pl.Mp; a.ldb = d->N; a.ldc = d->N;
  a.batch1 = d->K;
  a.sA1 = (long long)Mp * Mp; a.sB1 = (long long)M * N; a.sC1 = (long long)M * N;
  a.tri = TRI_LOWER_A;
  // solver 1 (true TRSM) and the 3xTF32 contraction mode: the column reductions come from one pass over
  // all M rows
  pl.tf32A = (d->solver != GPFX_SOLVER_TRSM) && (tf32_product_mask() & 1) && gemm_tf32_active(a);
  pl.gridM_A = (d->solver == GPFX_SOLVER_TRSM || pl.tf32A) ? 1 : gemm_grid_m(a);
  GemmArgs& b = pl.gB;
  b = GemmArgs();
  b.M = d->M; b.N = d->N; b.K = d->M;
  b.lda = d->M; b.ldb = d->N; b.ldc = d->N;
  b.transA = 1;
  b.batch1 = d->P;
  b.sA1 = (long long)M * M; b.sB1 = (d->K > 1) ? (long long)M * N : 0; b.sC1 = (long long)M * N;
  b.tri = TRI_UPPER_A;
  pl.tf32B = (tf32_product_mask() & 2) && gemm_tf32_active(b);
  pl.gridM_B = pl.tf32B ? 1 : gemm_grid_m(b);

  pl.splits_q = pick_splits(d->P, d->M, d->N);
  pl.splits_l = pick_splits(d->K, d->M, d->N);
  {
    // dq_mu = A gm: few output tiles, long contraction over N
    const long long tiles = (long long)ceil_div(d->M, 64) * (d->K > 1 ? d->P : 1);
    long long s = ceil_div_ll(296, tiles);
    const long long smax = d->N / 512 > 1 ? d->N / 512 : 1;
    pl.splits_m = (int)(s < 1 ? 1 : (s > smax ? smax : s));
  }
  const size_t nw = (d->K > 1) ? 1 : P;

  {
    Carver w;
    pl.fw_T = w.take(chol_scratch_doubles(d->K, pl.Mp));
    pl.fw_kl = w.take(kl_scratch_doubles(d->M, d->P));
    pl.fw_T1 = w.take(K * Mp * Mp);
    pl.fw_T2 = w.take(K * Mp * Mp);
    pl.fw_dKuu = w.take(K * Mp * Mp);
    pl.fw_part = w.take(kmat_bwd_scratch_doubles(d->K, d->M, d->M, d->D));
    // non-whitened parameterisation: scratch for the change of variables (alpha, C) <-> (q_mu, Lq)
    pl.fw_Lqtmp = w.take(d->whiten ? 0 : P * M * M);
    pl.fw_Wa = w.take(d->whiten ? 0 : M * P);
    pl.fw_dL2 = w.take(d->whiten ? 0 : K * Mp * Mp);
    pl.fws_bytes = w.off;
  }
  {
    Carver w;
    pl.cw_sqA = w.take(K * pl.gridM_A * N);
    pl.cw_wsum = w.take(K * pl.gridM_A * nw * N);
    pl.cw_sqB = w.take(P * pl.gridM_B * N);
    pl.cw_gmT = w.take(P * N);
    pl.cw_gv2T = w.take(P * N);
    pl.cw_gvsum = w.take(K * N);
    pl.cw_dA = w.take(K * M * N);
    pl.cw_dKuf = w.take(K * M * N);
    const size_t sq = (pl.splits_q > 1) ? (size_t)pl.splits_q * P * M * M : 0;
    const size_t sl = (pl.splits_l > 1) ? (size_t)pl.splits_l * K * M * M : 0;
    const size_t sm = std::max((pl.splits_m > 1) ? (size_t)pl.splits_m * M * P : (size_t)0,
                               dqmu_scratch_doubles(d->M, d->N, d->P, d->K));
    pl.cw_split = w.take(std::max(sq, std::max(sl, sm)));
    pl.cw_dqmu = w.take(std::max(dqmu_scratch_doubles(d->M, d->N, d->P, d->K),
                                 (size_t)ceil_div(d->N, 128) * d->K * d->M));  // may run beside the split-K GEMMs
    pl.cw_part = w.take(kmat_bwd_scratch_doubles(d->K, d->M, d->N, d->D));
    pl.cws_bytes = w.off;
  }
  pl.dL_bytes = align_up(K * Mp * Mp * sizeof(double), 256);
  return GPFX_OK;
}

inline double* at(void* base, size_t off) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off);
}
inline const double* at(const void* base, size_t off) {
  return reinterpret_cast<const double*>(reinterpret_cast<const char*>(base) + off);
}

// ---- auxiliary stream for the backward pass -------------------------------------------------
// The cond backward has two independent strands once the cotangents are prepared: the HBM/latency
// bound side work (dq_mu, the initial value of dA, the Kuf kernel-matrix backward) and the chain of
// DMMA GEMMs.  The side strand runs on a library-owned non-blocking stream forked from and joined
// back into the caller's stream with events, so the caller still sees stream-ordered semantics (and
// a CUDA graph captures both strands as parallel branches).  One context per caller stream; it is
// created outside of stream capture only (the first eager call) and can be switched off with
// GPFX_AUX_STREAM=0.
// (AuxCtx and its pool live in the library context: context.h / gpfx_create)
AuxCtx* get_aux(cudaStream_t main) { return context_aux(main); }

}  // namespace

size_t layer_tf32_workspace_bytes(const gpfx_layer_desc* d) {
  if (d == nullptr || d->N <= 0 || d->M <= 0 || d->P <= 0 || d->K <= 0) return 0;
  // the products of the cond path: [M x M] x [M x N] (batch K or P) and [M x N] x [N x M] (batch P or K)
  const int nb = d->P > d->K ? d->P : d->K;
  const size_t a = gemm_tf32_workspace_floats(nb, d->M, d->N, d->M);
  const size_t b = gemm_tf32_workspace_floats(nb, d->M, d->M, d->N);
  return (a > b ? a : b) * sizeof(float);
}

size_t layer_factor_ctx_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.fctx_bytes : 0;
}
size_t layer_cond_ctx_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.cctx_bytes : 0;
}
size_t layer_factor_workspace_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.fws_bytes : 0;
}
size_t layer_cond_workspace_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.cws_bytes : 0;
}
size_t layer_dL_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.dL_bytes : 0;
}
size_t layer_ctx_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.fctx_bytes + pl.cctx_bytes : 0;
}
size_t layer_workspace_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.fws_bytes + pl.cws_bytes + pl.dL_bytes : 0;
}

// ------------------------------------------------------------------------------------ factor
int layer_factor_forward(const gpfx_layer_desc* d, const double* Z, const double* ls,
                         const double* var, const double* q_mu, const double* q_sqrt, double* kl,
                         void* fctx, void* fws, cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!Z || !ls || !var || !q_mu || !q_sqrt || !kl || !fctx || !fws) {
    set_last_error("gpfx_layer_factor_forward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int M = pl.M, D = pl.D, P = pl.P, K = pl.K, Mp = pl.Mp;
  double* Lq = at(fctx, pl.f_Lq);
  double* L = at(fctx, pl.f_L);
  double* Linv = at(fctx, pl.f_Linv);
  int* info = reinterpret_cast<int*>(at(fctx, pl.f_info));

  // a9: KL + clean lower-triangular copy of q_sqrt (non-whitened: the copy goes to scratch and the
  // KL is taken after the change of variables below)
  double* Lq_raw = d->whiten ? Lq : at(fws, pl.fw_Lqtmp);
  GPFX_TRY(launch_kl_white(q_mu, q_sqrt, Lq_raw, kl, at(fws, pl.fw_kl), M, P, stream));
  // a1: Kuu + jitter, identity-padded to Mp
  {
    KmatArgs k;
    k.U = Z; k.V = Z; k.ls = ls; k.var = var; k.out = L;
    k.K = K; k.nu = M; k.nv = M; k.D = D;
    k.sU = k.sV = (long long)M * D; k.sOut = (long long)Mp * Mp;
    k.ldo = Mp; k.rows_out = Mp; k.cols_out = Mp;
    k.kind = d->kernel; k.add_jitter = 1; k.jitter = d->jitter;
    GPFX_TRY(launch_kmat_fwd(k, stream));
  }
  // a3: L, L^-1
  GPFX_TRY(potrf_trtri_batched(L, Linv, at(fws, pl.fw_T), K, Mp, info, stream));
  if (!d->whiten) {
    // a9': q(u) = N(q_mu, Lq Lq^T) is the whitened q(v) = N(alpha, C C^T) with alpha = L^-1 q_mu,
    // C_p = L^-1 Lq_p (u = L v).  Conditional and KL are then exactly the whitened ones in
    // (alpha, C): KL_white(alpha, C) = 1/2 [|alpha|^2 - MP - log|C C^T| + |C|_F^2] equals gauss_kl's
    // non-white branch because log diag(C)^2 = log diag(Lq)^2 - log diag(L)^2.  Costs O(P M^3)
    // instead of the reference's extra M^2 N triangular solve.
    const long long MpMp = (long long)Mp * Mp, MM = (long long)M * M;
    double* alpha = at(fctx, pl.f_alpha);
    {
      GemmArgs g;  // alpha = L^-1 q_mu
      g.A = Linv; g.lda = Mp; g.tri = TRI_LOWER_A;
      g.B = q_mu; g.ldb = P;
      g.C = alpha; g.ldc = P;
      g.M = M; g.K = M;
      if (K > 1) { g.N = 1; g.batch1 = P; g.sA1 = MpMp; g.sB1 = 1; g.sC1 = 1; }
      else { g.N = P; g.batch1 = 1; }
      GPFX_TRY(launch_gemm(g, stream));
    }
    {
      GemmArgs g;  // C_p = L_k^-1 Lq_p (lower x lower = lower)
      g.A = Linv; g.lda = Mp; g.sA1 = (K > 1) ? MpMp : 0;
      g.B = Lq_raw; g.ldb = M; g.sB1 = MM;
      g.C = Lq; g.ldc = M; g.sC1 = MM;
      g.M = M; g.N = M; g.K = M;
      g.batch1 = P;
      g.tri = TRI_LOWER_A | TRI_LOWER_B;
      GPFX_TRY(launch_gemm(g, stream));
    }
    GPFX_TRY(launch_kl_white(alpha, Lq, nullptr, kl, at(fws, pl.fw_kl), M, P, stream));
  }
  return GPFX_OK;
}

int layer_factor_backward(const gpfx_layer_desc* d, const double* Z, const double* ls,
                          const double* var, const double* q_mu, const double* g_kl,
                          const double* dL, double* dZ, double* dls, double* dvar, double* dq_mu,
                          double* dq_sqrt, int accumulate, const void* fctx, void* fws,
                          cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!Z || !ls || !var || !q_mu || !dL || !dZ || !dls || !dvar || !dq_mu || !dq_sqrt || !fctx ||
      !fws) {
    set_last_error("gpfx_layer_factor_backward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int M = pl.M, D = pl.D, P = pl.P, K = pl.K, Mp = pl.Mp;
  const long long MpMp = (long long)Mp * Mp;
  const double* Lq = at(fctx, pl.f_Lq);
  const double* L = at(fctx, pl.f_L);
  const double* Linv = at(fctx, pl.f_Linv);
  double* dKuu = at(fws, pl.fw_dKuu);

  const int acc_h = accumulate & 1;  // Z / lengthscale / variance gradients already hold the data part
  const int acc_q = accumulate & 2;  // dq_mu / dq_sqrt already hold the data-part cotangents
  const double* q_eff = d->whiten ? q_mu : at(fctx, pl.f_alpha);
  if (g_kl != nullptr) {
    GPFX_TRY(launch_kl_white_bwd(q_eff, Lq, g_kl, dq_mu, dq_sqrt, M, P, acc_q ? 1 : 0, stream));
  } else if (!acc_q) {
    GPFX_CUDA_CHECK(cudaMemsetAsync(dq_mu, 0, sizeof(double) * M * P, stream));
    GPFX_CUDA_CHECK(cudaMemsetAsync(dq_sqrt, 0, sizeof(double) * P * M * M, stream));
  }
  if (!d->whiten) {
    // a9' backward: dq_mu / dq_sqrt hold the cotangents of (alpha, C) = (L^-1 q_mu, L^-1 Lq).
    // For Y = L^-1 X: dX = L^-T dY and dL -= tril(L^-T dY Y^T).
    const long long MM = (long long)M * M;
    const double* alpha = at(fctx, pl.f_alpha);
    double* Wa = at(fws, pl.fw_Wa);
    double* Wc = at(fws, pl.fw_Lqtmp);
    double* dL2 = at(fws, pl.fw_dL2);
    GPFX_CUDA_CHECK(cudaMemcpyAsync(dL2, dL, sizeof(double) * MpMp * K, cudaMemcpyDeviceToDevice, stream));
    {
      GemmArgs g;  // Wa = L^-T d alpha
      g.A = Linv; g.lda = Mp; g.transA = 1; g.tri = TRI_UPPER_A;
      g.B = dq_mu; g.ldb = P;
      g.C = Wa; g.ldc = P;
      g.M = M; g.K = M;
      if (K > 1) { g.N = 1; g.batch1 = P; g.sA1 = MpMp; g.sB1 = 1; g.sC1 = 1; }
      else { g.N = P; g.batch1 = 1; }
      GPFX_TRY(launch_gemm(g, stream));
    }
    {
      GemmArgs g;  // Wc_p = L_k^-T dC_p
      g.A = Linv; g.lda = Mp; g.sA1 = (K > 1) ? MpMp : 0; g.transA = 1; g.tri = TRI_UPPER_A;
      g.B = dq_sqrt; g.ldb = M; g.sB1 = MM;
      g.C = Wc; g.ldc = M; g.sC1 = MM;
      g.M = M; g.N = M; g.K = M;
      g.batch1 = P;
      GPFX_TRY(launch_gemm(g, stream));
    }
    {
      GemmArgs g;  // dL -= tril(Wa alpha^T)
      g.A = Wa; g.lda = P;
      g.B = alpha; g.ldb = P; g.transB = 1;
      g.C = dL2; g.ldc = Mp;
      g.M = M; g.N = M;
      g.alpha = -1.0; g.beta = 1.0; g.tril_out = 1;
      if (K > 1) { g.K = 1; g.batch1 = P; g.sA1 = 1; g.sB1 = 1; g.sC1 = MpMp; }
      else { g.K = P; g.batch1 = 1; }
      GPFX_TRY(launch_gemm(g, stream));
    }
    {
      GemmArgs g;  // dL_k -= sum_{p in k} tril(Wc_p C_p^T)
      g.A = Wc; g.lda = M;
      g.B = Lq; g.ldb = M; g.transB = 1; g.tri = TRI_UPPER_B;
      g.C = dL2; g.ldc = Mp;
      g.M = M; g.N = M; g.K = M;
      g.alpha = -1.0; g.beta = 1.0; g.tril_out = 1;
      if (K > 1) { g.batch1 = P; g.sA1 = MM; g.sB1 = MM; g.sC1 = MpMp; }
      else { g.batch1 = 1; g.nseg = P; g.gA = MM; g.gB = MM; }
      GPFX_TRY(launch_gemm(g, stream));
    }
    GPFX_CUDA_CHECK(cudaMemcpyAsync(dq_mu, Wa, sizeof(double) * M * P, cudaMemcpyDeviceToDevice, stream));
    GPFX_TRY(launch_tril_copy(Wc, dq_sqrt, P, M, stream));
    dL = dL2;
  }
  GPFX_TRY(chol_backward_batched(L, Linv, dL, at(fws, pl.fw_T1), at(fws, pl.fw_T2), dKuu, K, Mp,
                                 stream));
  {
    KmatBwdArgs k;
    k.U = Z; k.V = Z; k.ls = ls; k.var = var; k.dK = dKuu;
    k.K = K; k.nu = M; k.nv = M; k.D = D;
    k.sU = (long long)M * D; k.sV = (long long)M * D; k.sK = MpMp; k.ldk = Mp;
    k.kind = d->kernel;
    k.dU = dZ; k.sdU = (long long)M * D; k.accumulate_dU = acc_h;
    k.dV = dZ; k.sdV = (long long)M * D; k.accumulate_dV = 1;
    k.dls = dls; k.dvar = dvar; k.accumulate_hyp = acc_h;
    k.scratch = at(fws, pl.fw_part);
    // fused single pass (operands centred on Z[0]) for M >= 128; tiny Kuu keeps the direct passes
    k.gemm_form = (M >= 128 && D + 1 <= 40) ? 2 : 0;
    GPFX_TRY(launch_kmat_bwd(k, stream));
  }
  return GPFX_OK;
}

// ------------------------------------------------------------------------------------ cond
namespace {

// a2: Kuf = k(Z, X) into the cond context
int launch_kuf(const Plan& pl, const gpfx_layer_desc* d, const double* X, const double* Z,
               const double* ls, const double* var, void* cctx, cudaStream_t stream) {
  KmatArgs k;
  k.U = Z; k.V = X; k.ls = ls; k.var = var; k.out = at(cctx, pl.c_Kuf);
  k.K = pl.K; k.nu = pl.M; k.nv = pl.N; k.D = pl.D;
  k.sU = (long long)pl.M * pl.D; k.sV = 0; k.sOut = (long long)pl.M * pl.N;
  k.ldo = pl.N; k.rows_out = pl.M; k.cols_out = pl.N;
  k.kind = d->kernel;
  return launch_kmat_fwd(k, stream);
}

int cond_forward_impl(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                      const double* var, const double* q_mu, const void* fctx,
                      const double* mean_add, const double* eps, double* fmean, double* fvar,
                      double* fsample, void* cctx, void* cws, cudaStream_t stream, bool kuf_done);

}  // namespace

int layer_cond_forward(const gpfx_layer_desc* d, const double* X, const double* Z,
                       const double* ls, const double* var, const double* q_mu, const void* fctx,
                       const double* mean_add, const double* eps, double* fmean, double* fvar,
                       double* fsample, void* cctx, void* cws, cudaStream_t stream) {
  return cond_forward_impl(d, X, Z, ls, var, q_mu, fctx, mean_add, eps, fmean, fvar, fsample, cctx, cws,
                           stream, false);
}

namespace {

int cond_forward_impl(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                      const double* var, const double* q_mu, const void* fctx,
                      const double* mean_add, const double* eps, double* fmean, double* fvar,
                      double* fsample, void* cctx, void* cws, cudaStream_t stream, bool kuf_done) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!X || !Z || !ls || !var || !q_mu || !fctx || !fmean || !fvar || !cctx || !cws) {
    set_last_error("gpfx_layer_cond_forward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int N = pl.N, P = pl.P, K = pl.K;
  const double* Lq = at(fctx, pl.f_Lq);
  const double* Linv = at(fctx, pl.f_Linv);
  if (!d->whiten) q_mu = at(fctx, pl.f_alpha);  // effective (whitened) variational mean
  double* Kuf = at(cctx, pl.c_Kuf);
  double* A = at(cctx, pl.c_A);
  double* B = at(cctx, pl.c_B);

  // a2: Kuf (already under way on the side strand when called from the fused layer_forward)
  if (!kuf_done) GPFX_TRY(launch_kuf(pl, d, X, Z, ls, var, cctx, stream));
  // a4-a6: A = L^-1 Kuf with fused column partials of A^2 and A^T q_mu
  if (d->solver == GPFX_SOLVER_TRSM) {
    // reference-faithful form: blocked forward substitution with L itself, no inverse
    const long long MN = (long long)pl.M * N;
    GPFX_CUDA_CHECK(cudaMemcpyAsync(A, Kuf, sizeof(double) * MN * K, cudaMemcpyDeviceToDevice, stream));
    GPFX_TRY(trsm_lower_batched(at(fctx, pl.f_L), (long long)pl.Mp * pl.Mp, pl.Mp, A, MN, K, pl.M, N,
                                false, stream));
    GPFX_TRY(launch_col_reduce(A, MN, K, pl.M, N, q_mu, K > 1 ? 1 : 0, P, K > 1 ? 1 : P,
                               at(cws, pl.cw_sqA), at(cws, pl.cw_wsum), stream));
  } else if (pl.tf32A) {
    // float32 opt-in: plain 3xTF32 product, then the column reductions as one streaming pass
    GemmArgs g = pl.gA;
    g.A = Linv; g.B = Kuf; g.C = A;
    GPFX_TRY(launch_gemm_auto(g, stream));
    GPFX_TRY(launch_col_reduce(A, (long long)pl.M * N, K, pl.M, N, q_mu, K > 1 ? 1 : 0, P, K > 1 ? 1 : P,
                               at(cws, pl.cw_sqA), at(cws, pl.cw_wsum), stream));
  } else {
    GemmArgs g = pl.gA;
    g.A = Linv; g.B = Kuf; g.C = A;
    g.sumsq = at(cws, pl.cw_sqA); g.sSumsq = (long long)pl.gridM_A * N;
    g.wvec = q_mu; g.ldw = P;
    if (K > 1) { g.nw = 1; g.sW = 1; } else { g.nw = P; g.sW = 0; }
    g.wsum = at(cws, pl.cw_wsum); g.sWsum = (long long)pl.gridM_A * g.nw * N;
    GPFX_TRY(launch_gemm(g, stream));
  }
  // a7: B_p = Lq_p^T A with fused column partials of B^2
  {
    GemmArgs g = pl.gB;
    g.A = Lq; g.B = A; g.C = B;
    if (pl.tf32B) {
      GPFX_TRY(launch_gemm_auto(g, stream));
      GPFX_TRY(launch_col_reduce(B, (long long)pl.M * N, P, pl.M, N, nullptr, 0, 0, 0, at(cws, pl.cw_sqB), nullptr,
                                 stream));
    } else {
      g.sumsq = at(cws, pl.cw_sqB); g.sSumsq = (long long)pl.gridM_B * N;
      GPFX_TRY(launch_gemm(g, stream));
    }
  }
  // a5-a8
  {
    FinalizeArgs f;
    f.N = N; f.P = P; f.K = K;
    f.gridM_A = pl.gridM_A; f.gridM_B = pl.gridM_B;
    f.wsum = at(cws, pl.cw_wsum); f.sqA = at(cws, pl.cw_sqA); f.sqB = at(cws, pl.cw_sqB);
    f.variance = var; f.mean_add = mean_add; f.eps = eps;
    f.fmean = fmean; f.fvar = fvar; f.fsample = fsample;
    GPFX_TRY(launch_finalize(f, stream));
  }
  return GPFX_OK;
}

}  // namespace

int layer_cond_backward(const gpfx_layer_desc* d, const double* X, const double* Z,
                        const double* ls, const double* var, const double* q_mu,
                        const double* eps, const double* fvar, const double* g_f,
                        const double* g_mean, const double* g_var, double* dX, double* dZ,
                        double* dls, double* dvar, double* dq_mu, double* dq_sqrt, double* dL,
                        double* dmean, const void* fctx, void* cctx, void* cws,
                        cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!X || !Z || !ls || !var || !q_mu || !fvar || !dX || !dZ || !dls || !dvar || !dq_mu ||
      !dq_sqrt || !dL || !fctx || !cctx || !cws) {
    set_last_error("gpfx_layer_cond_backward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int N = pl.N, M = pl.M, D = pl.D, P = pl.P, K = pl.K, Mp = pl.Mp;
  const long long MN = (long long)M * N, MM = (long long)M * M, MpMp = (long long)Mp * Mp;
  const double* Lq = at(fctx, pl.f_Lq);
  const double* Linv = at(fctx, pl.f_Linv);
  if (!d->whiten) q_mu = at(fctx, pl.f_alpha);  // gradients below are w.r.t. the effective (alpha, C)
  double* A = at(cctx, pl.c_A);
  double* B = at(cctx, pl.c_B);
  double* gmT = at(cws, pl.cw_gmT);
  double* gv2T = at(cws, pl.cw_gv2T);
  double* gvsum = at(cws, pl.cw_gvsum);
  double* dA = at(cws, pl.cw_dA);
  double* dKuf = at(cws, pl.cw_dKuf);
  double* split = at(cws, pl.cw_split);

  {
    BwdPrepArgs b;
    b.N = N; b.P = P; b.K = K;
    b.g_f = g_f; b.g_mean = g_mean; b.g_var = g_var; b.eps = eps; b.fvar = fvar;
    b.dmean = dmean; b.gmT = gmT; b.gv2T = gv2T; b.gvsum = gvsum;
    GPFX_TRY(launch_bwd_prep(b, stream));
  }
  // dA_k = q_mu gm^T - 2 A diag(gvsum) + sum_{p in k} Lq_p Bs_p.  Separate kernels (K == P) on the TMA-fed
  // GEMM: the first two terms and dq_mu = A gm ride in that GEMM's epilogue (A is read there once, dA is
  // written once); otherwise they are separate streaming kernels on the side strand.
  GemmArgs gdA;
  gdA.A = Lq; gdA.lda = M;
  gdA.B = B; gdA.ldb = N;
  gdA.C = dA; gdA.ldc = N;
  gdA.M = M; gdA.N = N; gdA.K = M;
  gdA.tri = TRI_LOWER_A;
  bool fused_dA = false;
  if (K > 1) {
    gdA.batch1 = P; gdA.sA1 = MM; gdA.sB1 = MN; gdA.sC1 = MN;
    GemmArgs t = gdA;
    t.addmat = A; t.ldadd = N; t.sAdd = MN;
    t.addcol = gvsum; t.sAddcol = N; t.addscale = -2.0;
    t.r1u = q_mu; t.ldr1u = P; t.sR1u = 1;
    t.r1v = gmT; t.sR1v = N;
    t.rowdot = at(cws, pl.cw_dqmu);
    if (!((tf32_product_mask() & 8) && gemm_tf32_active(gdA)) && gemm_addin_supported(t)) {  // (the 3xTF32 product takes the unfused form)
      gdA = t;
      fused_dA = true;
    }
  } else {
    gdA.batch1 = 1; gdA.nseg = P; gdA.gA = MM; gdA.gB = MN;
  }
  if (!fused_dA) gdA.beta = 1.0;
  // side strand (see AuxCtx): dq_mu and the initial value of dA only need the prepared cotangents
  // (the GEMM fallback of dq_mu shares the split-K workspace with the main strand: no fork then)
  const bool dq_ok = dqmu_supported(A, gmT, N);
  AuxCtx* ax = dq_ok ? get_aux(stream) : nullptr;
  cudaStream_t side = ax ? ax->aux : stream;
  if (ax) {
    GPFX_CUDA_CHECK(cudaEventRecord(ax->ev[0], stream));
    GPFX_CUDA_CHECK(cudaStreamWaitEvent(side, ax->ev[0], 0));
  }
  if (!fused_dA) {
    // dq_mu = A gm: dedicated streaming kernel (A read once); split-K DMMA GEMM [M,N] x [N,P] otherwise
    if (dq_ok) {
      GPFX_TRY(launch_dqmu(A, gmT, dq_mu, at(cws, pl.cw_dqmu), M, N, P, K, side));
    } else {
      GemmArgs g;
      g.A = A; g.lda = N;
      g.B = gmT; g.ldb = N; g.transB = 1;
      g.C = dq_mu; g.ldc = P;
      g.M = M; g.K = N;
      if (K > 1) {
        g.N = 1; g.batch1 = P; g.sA1 = MN; g.sB1 = N; g.sC1 = 1;
      } else {
        g.N = P; g.batch1 = 1;
      }
      g.splits = pl.splits_m; g.splitws = split;
      GPFX_TRY(launch_gemm(g, stream));
    }
    GPFX_TRY(launch_dA_init(q_mu, gmT, A, gvsum, dA, M, N, P, K, side));
  }
  if (ax) GPFX_CUDA_CHECK(cudaEventRecord(ax->ev[1], side));
  // dLq_p = tril(A_k Bs_p^T),  Bs_p = B_p diag(2 gv_p)
  {
    GemmArgs g;
    g.A = A; g.lda = N; g.sA1 = (K > 1) ? MN : 0;
    g.B = B; g.ldb = N; g.sB1 = MN; g.transB = 1;
    g.C = dq_sqrt; g.ldc = M; g.sC1 = MM;
    g.M = M; g.N = M; g.K = N;
    g.batch1 = P;
    g.tril_out = 1;
    g.splits = pl.splits_q; g.splitws = split;
    // Bs_p is never stored when both of its consumers can scale on the fly: this product multiplies its B
    // fragments by 2 gv_p[k], the fused dA product scales its accumulator columns
    GemmArgs gs = g;
    gs.kscale = gv2T; gs.sKscale = N;
    GemmArgs ts = gdA;
    ts.colscale = gv2T; ts.sColscale = N;
    const bool tf32_q = (tf32_product_mask() & 4) && gemm_tf32_active(g);
    if (fused_dA && !tf32_q && gemm_kscale_supported(gs) && gemm_addin_supported(ts)) {
      g = gs;
      gdA = ts;
    } else {
      GPFX_TRY(launch_scale_cols(B, gv2T, P, M, N, stream));
    }
    GPFX_TRY((tf32_product_mask() & 4) ? launch_gemm_auto(g, stream) : launch_gemm(g, stream));
  }
  // dA_k (see above)
  if (ax) GPFX_CUDA_CHECK(cudaStreamWaitEvent(stream, ax->ev[1], 0));
  GPFX_TRY((tf32_product_mask() & 8) ? launch_gemm_auto(gdA, stream) : launch_gemm(gdA, stream));
  if (fused_dA) GPFX_TRY(launch_dqmu_finish(at(cws, pl.cw_dqmu), dq_mu, M, P, K, ceil_div(N, 128), stream));
  // dKuf_k = L_k^-T dA_k
  if (d->solver == GPFX_SOLVER_TRSM) {
    GPFX_CUDA_CHECK(cudaMemcpyAsync(dKuf, dA, sizeof(double) * MN * K, cudaMemcpyDeviceToDevice, stream));
    GPFX_TRY(trsm_lower_batched(at(fctx, pl.f_L), MpMp, Mp, dKuf, MN, K, M, N, true, stream));
  } else {
    GemmArgs g;
    g.A = Linv; g.lda = Mp; g.sA1 = MpMp; g.transA = 1;
    g.B = dA; g.ldb = N; g.sB1 = MN;
    g.C = dKuf; g.ldc = N; g.sC1 = MN;
    g.M = M; g.N = N; g.K = M;
    g.batch1 = K;
    g.tri = TRI_UPPER_A;
    GPFX_TRY((tf32_product_mask() & 16) ? launch_gemm_auto(g, stream) : launch_gemm(g, stream));
  }
  // the Kuf kernel-matrix backward and dL both consume dKuf; only the fused single-pass form leaves
  // dKuf untouched (the other forms overwrite it with R), so only that one may run beside dL
  const bool km_side = ax && N >= 128 && D <= 15;
  if (km_side) {
    GPFX_CUDA_CHECK(cudaEventRecord(ax->ev[2], stream));
    GPFX_CUDA_CHECK(cudaStreamWaitEvent(side, ax->ev[2], 0));
  }
  // dL_k = -tril(dKuf_k A_k^T), zero-padded to Mp
  GPFX_CUDA_CHECK(cudaMemsetAsync(dL, 0, sizeof(double) * MpMp * K, stream));
  {
    GemmArgs g;
    g.A = dKuf; g.lda = N; g.sA1 = MN;
    g.B = A; g.ldb = N; g.sB1 = MN; g.transB = 1;
    g.C = dL; g.ldc = Mp; g.sC1 = MpMp;
    g.M = M; g.N = M; g.K = N;
    g.batch1 = K;
    g.alpha = -1.0;
    g.tril_out = 1;
    g.splits = pl.splits_l; g.splitws = split;
    GPFX_TRY((tf32_product_mask() & 32) ? launch_gemm_auto(g, stream) : launch_gemm(g, stream));
  }
  // kernel-matrix backward through Kuf
  {
    KmatBwdArgs k;
    k.U = Z; k.V = X; k.ls = ls; k.var = var; k.dK = dKuf;
    k.K = K; k.nu = M; k.nv = N; k.D = D;
    k.sU = (long long)M * D; k.sV = 0; k.sK = MN; k.ldk = N;
    k.kind = d->kernel;
    k.dU = dZ; k.sdU = (long long)M * D; k.accumulate_dU = 0;
    k.dV = dX; k.sdV = 0; k.accumulate_dV = 0;
    k.dls = dls; k.dvar = dvar; k.accumulate_hyp = 0;
    k.scratch = at(cws, pl.cw_part);
    k.Kmat = at(cctx, pl.c_Kuf); k.sKmat = MN; k.ldkmat = N;
    k.gemm_form = (N >= 128) ? (D <= 15 ? 2 : 1) : 0;  // small D: fused single pass; else GEMM form
    GPFX_TRY(launch_kmat_bwd(k, km_side ? side : stream));
  }
  if (ax) {  // join: everything the side strand produced is ordered before what follows on `stream`
    GPFX_CUDA_CHECK(cudaEventRecord(ax->ev[3], side));
    GPFX_CUDA_CHECK(cudaStreamWaitEvent(stream, ax->ev[3], 0));
  }
  // Knn = variance: d variance_k += sum_n sum_{p in k} gv
  GPFX_TRY(launch_row_sums(gvsum, K, N, dvar, 1, stream));
  return GPFX_OK;
}

// ------------------------------------------------------------------------------------ fused
int layer_forward(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                  const double* var, const double* q_mu, const double* q_sqrt,
                  const double* mean_add, const double* eps, double* fmean, double* fvar,
                  double* fsample, double* kl, void* ctx, void* ws, cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!ctx || !ws) {
    set_last_error("gpfx_layer_forward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  char* c = reinterpret_cast<char*>(ctx);
  char* w = reinterpret_cast<char*>(ws);
  // Kuf = k(Z, X) does not depend on the factorisation: it runs on the side strand while the
  // latency-bound Cholesky chain occupies a handful of SMs
  AuxCtx* ax = (X && Z && ls && var) ? get_aux(stream) : nullptr;
  if (ax) {
    GPFX_CUDA_CHECK(cudaEventRecord(ax->ev[0], stream));
    GPFX_CUDA_CHECK(cudaStreamWaitEvent(ax->aux, ax->ev[0], 0));
    GPFX_TRY(launch_kuf(pl, d, X, Z, ls, var, c + pl.fctx_bytes, ax->aux));
    GPFX_CUDA_CHECK(cudaEventRecord(ax->ev[1], ax->aux));
  }
  const int st = layer_factor_forward(d, Z, ls, var, q_mu, q_sqrt, kl, c, w, stream);
  if (ax) GPFX_CUDA_CHECK(cudaStreamWaitEvent(stream, ax->ev[1], 0));  // join even on error
  GPFX_TRY(st);
  return cond_forward_impl(d, X, Z, ls, var, q_mu, c, mean_add, eps, fmean, fvar, fsample,
                           c + pl.fctx_bytes, w + pl.fws_bytes, stream, ax != nullptr);
}

int layer_backward(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                   const double* var, const double* q_mu, const double* eps, const double* fvar,
                   const double* g_f, const double* g_mean, const double* g_var,
                   const double* g_kl, double* dX, double* dZ, double* dls, double* dvar,
                   double* dq_mu, double* dq_sqrt, double* dmean, void* ctx, void* ws,
                   cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!ctx || !ws) {
    set_last_error("gpfx_layer_backward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  char* c = reinterpret_cast<char*>(ctx);
  char* w = reinterpret_cast<char*>(ws);
  double* dL = reinterpret_cast<double*>(w + pl.fws_bytes + pl.cws_bytes);
  GPFX_TRY(layer_cond_backward(d, X, Z, ls, var, q_mu, eps, fvar, g_f, g_mean, g_var, dX, dZ, dls,
                               dvar, dq_mu, dq_sqrt, dL, dmean, c, c + pl.fctx_bytes,
                               w + pl.fws_bytes, stream));
  return layer_factor_backward(d, Z, ls, var, q_mu, g_kl, dL, dZ, dls, dvar, dq_mu, dq_sqrt, 3, c,
                               w, stream);
}

int layer_status_async(const gpfx_layer_desc* d, const void* fctx, int32_t* info_dev,
                       cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!fctx || !info_dev) {
    set_last_error("gpfx_layer_status_async: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int* info = reinterpret_cast<const int*>(reinterpret_cast<const char*>(fctx) + pl.f_info);
  GPFX_CUDA_CHECK(cudaMemcpyAsync(info_dev, info, sizeof(int) * pl.K, cudaMemcpyDeviceToDevice, stream));
  return GPFX_OK;
}

int layer_status(const gpfx_layer_desc* d, const void* fctx, int32_t* info_host,
                 cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  const int* info = reinterpret_cast<const int*>(reinterpret_cast<const char*>(fctx) + pl.f_info);
  GPFX_CUDA_CHECK(cudaMemcpyAsync(info_host, info, sizeof(int) * pl.K, cudaMemcpyDeviceToHost,
                                  stream));
  GPFX_CUDA_CHECK(cudaStreamSynchronize(stream));
  for (int k = 0; k < pl.K; ++k)
    if (info_host[k] != 0) {
      set_last_error("Kuu[%d] is not positive definite (pivot %d)", k, info_host[k] - 1);
      return GPFX_E_NOT_PD;
    }
  return GPFX_OK;
}

}  // namespace gpfx
