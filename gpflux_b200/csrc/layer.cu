// GPLayer forward / backward orchestration: one stream, no host syncs, caller-owned buffers.
//
// Forward  (gpflux/layers/gp_layer.py:226-268, 303-311, 339-363):
//   Lq = tril(q_sqrt), KL                       kl_tril            (a9)
//   Kuu + jitter I  -> L, L^-1                  kmat_fwd, potrf    (a1, a3)
//   Kuf                                         kmat_fwd           (a2)
//   A   = L^-1 Kuf      (+ sum A^2, A^T q_mu)   DMMA GEMM, lower-A (a4, a5, a6)
//   B_p = Lq_p^T A      (+ sum B^2)             DMMA GEMM, upper-A (a7)
//   mean, var, f = mean + sqrt(var) eps         finalize           (a8)
// Backward (a14), 3x the forward's triangular flops:
//   dLq_p = tril(A Bs_p^T), dA = sum_p Lq_p Bs_p + q_mu gm^T - 2 A diag(gv),
//   dKuf = L^-T dA, dL = -tril(dKuf A^T), dKuu = L^-T Phi(L^T dL) L^-1, kernel-matrix backward.
#include "layer.h"

#include <algorithm>

#include "chol.h"
#include "elementwise.h"
#include "gemm_f64.h"
#include "kmat.h"

namespace gpfx {
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
  // ctx
  size_t c_Lq, c_L, c_Linv, c_Kuf, c_A, c_B, c_info, ctx_bytes;
  // workspace (forward)
  size_t w_T, w_kl, w_sqA, w_wsum, w_sqB;
  // workspace (backward)
  size_t w_gmT, w_gv2T, w_gvsum, w_dA, w_dKuf, w_dL, w_T1, w_T2, w_dKuu, w_split, w_part;
  size_t ws_bytes;
  int gridM_A, gridM_B;
  int splits_q, splits_l, splits_m;
  GemmArgs gA, gB;  // forward GEMM templates (pointers filled later)
};

// split-K factor for the [M,M] = [M,N] x [N,M] lower-triangular-output contractions: enough CTAs
// for ~2 waves of 128x128 tiles (live lower tiles only), at least 512 k per split
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
  if (!d->whiten) {
    set_last_error("gpfx_layer: only the whitened parameterisation is implemented");
    return GPFX_E_UNSUPPORTED;
  }
  if (d->kernel < GPFX_KERNEL_SE || d->kernel > GPFX_KERNEL_MATERN52) {
    set_last_error("gpfx_layer: unknown kernel id %d", d->kernel);
    return GPFX_E_UNSUPPORTED;
  }
  pl.N = d->N; pl.M = d->M; pl.D = d->D; pl.P = d->P; pl.K = d->K;
  pl.Mp = chol_padded_dim(d->M);
  const size_t N = d->N, M = d->M, P = d->P, K = d->K, Mp = pl.Mp, D = d->D;

  Carver c;
  pl.c_Lq = c.take(P * M * M);
  pl.c_L = c.take(K * Mp * Mp);
  pl.c_Linv = c.take(K * Mp * Mp);
  pl.c_Kuf = c.take(K * M * N);
  pl.c_A = c.take(K * M * N);
  pl.c_B = c.take(P * M * N);
  pl.c_info = c.take((K + 1) / 2 + 1);
  pl.ctx_bytes = c.off;

  // forward GEMM shapes decide the tile size, hence the number of column partials
  GemmArgs& a = pl.gA;
  a = GemmArgs();
  a.M = d->M; a.N = d->N; a.K = d->M;
  a.lda = pl.Mp; a.ldb = d->N; a.ldc = d->N;
  a.batch1 = d->K;
  a.sA1 = (long long)Mp * Mp; a.sB1 = (long long)M * N; a.sC1 = (long long)M * N;
  a.tri = TRI_LOWER_A;
  pl.gridM_A = gemm_grid_m(a);
  GemmArgs& b = pl.gB;
  b = GemmArgs();
  b.M = d->M; b.N = d->N; b.K = d->M;
  b.lda = d->M; b.ldb = d->N; b.ldc = d->N;
  b.transA = 1;
  b.batch1 = d->P;
  b.sA1 = (long long)M * M; b.sB1 = (d->K > 1) ? (long long)M * N : 0; b.sC1 = (long long)M * N;
  b.tri = TRI_UPPER_A;
  pl.gridM_B = gemm_grid_m(b);

  pl.splits_q = pick_splits(d->P, d->M, d->N);
  pl.splits_l = pick_splits(d->K, d->M, d->N);
  const size_t nw = (d->K > 1) ? 1 : P;

  Carver w;
  pl.w_T = w.take(chol_scratch_doubles(d->K, pl.Mp));
  pl.w_kl = w.take(kl_scratch_doubles(d->M, d->P));
  pl.w_sqA = w.take(K * pl.gridM_A * N);
  pl.w_wsum = w.take(K * pl.gridM_A * nw * N);
  pl.w_sqB = w.take(P * pl.gridM_B * N);
  pl.w_gmT = w.take(P * N);
  pl.w_gv2T = w.take(P * N);
  pl.w_gvsum = w.take(K * N);
  pl.w_dA = w.take(K * M * N);
  pl.w_dKuf = w.take(K * M * N);
  pl.w_dL = w.take(K * Mp * Mp);
  pl.w_T1 = w.take(K * Mp * Mp);
  pl.w_T2 = w.take(K * Mp * Mp);
  pl.w_dKuu = w.take(K * Mp * Mp);
  {
    // dq_mu = A gm: few output tiles, long contraction over N
    const long long tiles = (long long)ceil_div(d->M, 64) * (d->K > 1 ? d->P : 1);
    long long s = ceil_div_ll(296, tiles);
    const long long smax = d->N / 512 > 1 ? d->N / 512 : 1;
    pl.splits_m = (int)(s < 1 ? 1 : (s > smax ? smax : s));
  }
  const size_t sq = (pl.splits_q > 1) ? (size_t)pl.splits_q * P * M * M : 0;
  const size_t sl = (pl.splits_l > 1) ? (size_t)pl.splits_l * K * M * M : 0;
  const size_t sm = (pl.splits_m > 1) ? (size_t)pl.splits_m * M * P : 0;
  pl.w_split = w.take(std::max(sq, std::max(sl, sm)));
  {
    const size_t s1 = kmat_bwd_scratch_doubles(d->K, d->M, d->N, d->D);
    const size_t s2 = kmat_bwd_scratch_doubles(d->K, d->M, d->M, d->D);
    pl.w_part = w.take(s1 > s2 ? s1 : s2);
  }
  pl.ws_bytes = w.off;
  return GPFX_OK;
}

inline double* at(void* base, size_t off) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off);
}

}  // namespace

size_t layer_ctx_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.ctx_bytes : 0;
}
size_t layer_workspace_bytes(const gpfx_layer_desc* d) {
  Plan pl;
  return make_plan(d, pl) == GPFX_OK ? pl.ws_bytes : 0;
}

int layer_forward(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                  const double* var, const double* q_mu, const double* q_sqrt,
                  const double* mean_add, const double* eps, double* fmean, double* fvar,
                  double* fsample, double* kl, void* ctx, void* ws, cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!X || !Z || !ls || !var || !q_mu || !q_sqrt || !fmean || !fvar || !kl || !ctx || !ws) {
    set_last_error("gpfx_layer_forward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int N = pl.N, M = pl.M, D = pl.D, P = pl.P, K = pl.K, Mp = pl.Mp;
  double* Lq = at(ctx, pl.c_Lq);
  double* L = at(ctx, pl.c_L);
  double* Linv = at(ctx, pl.c_Linv);
  double* Kuf = at(ctx, pl.c_Kuf);
  double* A = at(ctx, pl.c_A);
  double* B = at(ctx, pl.c_B);
  int* info = reinterpret_cast<int*>(at(ctx, pl.c_info));

  // a9: KL + clean lower-triangular copy of q_sqrt
  GPFX_TRY(launch_kl_white(q_mu, q_sqrt, Lq, kl, at(ws, pl.w_kl), M, P, stream));

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
  GPFX_TRY(potrf_trtri_batched(L, Linv, at(ws, pl.w_T), K, Mp, info, stream));
  // a2: Kuf
  {
    KmatArgs k;
    k.U = Z; k.V = X; k.ls = ls; k.var = var; k.out = Kuf;
    k.K = K; k.nu = M; k.nv = N; k.D = D;
    k.sU = (long long)M * D; k.sV = 0; k.sOut = (long long)M * N;
    k.ldo = N; k.rows_out = M; k.cols_out = N;
    k.kind = d->kernel;
    GPFX_TRY(launch_kmat_fwd(k, stream));
  }
  // a4-a6: A = L^-1 Kuf with fused column partials of A^2 and A^T q_mu
  {
    GemmArgs g = pl.gA;
    g.A = Linv; g.B = Kuf; g.C = A;
    g.sumsq = at(ws, pl.w_sqA); g.sSumsq = (long long)pl.gridM_A * N;
    g.wvec = q_mu; g.ldw = P;
    if (K > 1) { g.nw = 1; g.sW = 1; } else { g.nw = P; g.sW = 0; }
    g.wsum = at(ws, pl.w_wsum); g.sWsum = (long long)pl.gridM_A * g.nw * N;
    GPFX_TRY(launch_gemm(g, stream));
  }
  // a7: B_p = Lq_p^T A with fused column partials of B^2
  {
    GemmArgs g = pl.gB;
    g.A = Lq; g.B = A; g.C = B;
    g.sumsq = at(ws, pl.w_sqB); g.sSumsq = (long long)pl.gridM_B * N;
    GPFX_TRY(launch_gemm(g, stream));
  }
  // a5-a8
  {
    FinalizeArgs f;
    f.N = N; f.P = P; f.K = K;
    f.gridM_A = pl.gridM_A; f.gridM_B = pl.gridM_B;
    f.wsum = at(ws, pl.w_wsum); f.sqA = at(ws, pl.w_sqA); f.sqB = at(ws, pl.w_sqB);
    f.variance = var; f.mean_add = mean_add; f.eps = eps;
    f.fmean = fmean; f.fvar = fvar; f.fsample = fsample;
    GPFX_TRY(launch_finalize(f, stream));
  }
  return GPFX_OK;
}

int layer_backward(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                   const double* var, const double* q_mu, const double* eps, const double* fvar,
                   const double* g_f, const double* g_mean, const double* g_var,
                   const double* g_kl, double* dX, double* dZ, double* dls, double* dvar,
                   double* dq_mu, double* dq_sqrt, double* dmean, void* ctx, void* ws,
                   cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  if (!X || !Z || !ls || !var || !q_mu || !fvar || !g_kl || !dX || !dZ || !dls || !dvar ||
      !dq_mu || !dq_sqrt || !ctx || !ws) {
    set_last_error("gpfx_layer_backward: null pointer argument");
    return GPFX_E_SHAPE;
  }
  const int N = pl.N, M = pl.M, D = pl.D, P = pl.P, K = pl.K, Mp = pl.Mp;
  const long long MN = (long long)M * N, MM = (long long)M * M, MpMp = (long long)Mp * Mp;
  double* Lq = at(ctx, pl.c_Lq);
  double* L = at(ctx, pl.c_L);
  double* Linv = at(ctx, pl.c_Linv);
  double* A = at(ctx, pl.c_A);
  double* B = at(ctx, pl.c_B);
  double* gmT = at(ws, pl.w_gmT);
  double* gv2T = at(ws, pl.w_gv2T);
  double* gvsum = at(ws, pl.w_gvsum);
  double* dA = at(ws, pl.w_dA);
  double* dKuf = at(ws, pl.w_dKuf);
  double* dL = at(ws, pl.w_dL);
  double* dKuu = at(ws, pl.w_dKuu);
  double* split = at(ws, pl.w_split);

  {
    BwdPrepArgs b;
    b.N = N; b.P = P; b.K = K;
    b.g_f = g_f; b.g_mean = g_mean; b.g_var = g_var; b.eps = eps; b.fvar = fvar;
    b.dmean = dmean; b.gmT = gmT; b.gv2T = gv2T; b.gvsum = gvsum;
    GPFX_TRY(launch_bwd_prep(b, stream));
  }
  // dq_mu = A gm  (+ KL term below): split-K DMMA GEMM [M,N] x [N,P]
  {
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
  // Bs_p = B_p diag(2 gv_p)
  GPFX_TRY(launch_scale_cols(B, gv2T, P, M, N, stream));
  // dLq_p = tril(A_k Bs_p^T)
  {
    GemmArgs g;
    g.A = A; g.lda = N; g.sA1 = (K > 1) ? MN : 0;
    g.B = B; g.ldb = N; g.sB1 = MN; g.transB = 1;
    g.C = dq_sqrt; g.ldc = M; g.sC1 = MM;
    g.M = M; g.N = M; g.K = N;
    g.batch1 = P;
    g.tril_out = 1;
    g.splits = pl.splits_q; g.splitws = split;
    GPFX_TRY(launch_gemm(g, stream));
  }
  GPFX_TRY(launch_kl_white_bwd(q_mu, Lq, g_kl, dq_mu, dq_sqrt, M, P, 1, stream));
  // dA_k = q_mu gm^T - 2 A diag(gvsum) + sum_{p in k} Lq_p Bs_p
  GPFX_TRY(launch_dA_init(q_mu, gmT, A, gvsum, dA, M, N, P, K, stream));
  {
    GemmArgs g;
    g.A = Lq; g.lda = M;
    g.B = B; g.ldb = N;
    g.C = dA; g.ldc = N;
    g.M = M; g.N = N; g.K = M;
    g.tri = TRI_LOWER_A;
    g.beta = 1.0;
    if (K > 1) {
      g.batch1 = P; g.sA1 = MM; g.sB1 = MN; g.sC1 = MN;
    } else {
      g.batch1 = 1; g.nseg = P; g.gA = MM; g.gB = MN;
    }
    GPFX_TRY(launch_gemm(g, stream));
  }
  // dKuf_k = L_k^-T dA_k
  {
    GemmArgs g;
    g.A = Linv; g.lda = Mp; g.sA1 = MpMp; g.transA = 1;
    g.B = dA; g.ldb = N; g.sB1 = MN;
    g.C = dKuf; g.ldc = N; g.sC1 = MN;
    g.M = M; g.N = N; g.K = M;
    g.batch1 = K;
    g.tri = TRI_UPPER_A;
    GPFX_TRY(launch_gemm(g, stream));
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
    GPFX_TRY(launch_gemm(g, stream));
  }
  GPFX_TRY(chol_backward_batched(L, Linv, dL, at(ws, pl.w_T1), at(ws, pl.w_T2), dKuu, K, Mp,
                                 stream));
  // kernel-matrix backward: Kuf then Kuu
  {
    KmatBwdArgs k;
    k.U = Z; k.V = X; k.ls = ls; k.var = var; k.dK = dKuf;
    k.K = K; k.nu = M; k.nv = N; k.D = D;
    k.sU = (long long)M * D; k.sV = 0; k.sK = MN; k.ldk = N;
    k.kind = d->kernel;
    k.dU = dZ; k.sdU = (long long)M * D; k.accumulate_dU = 0;
    k.dV = dX; k.sdV = 0; k.accumulate_dV = 0;
    k.dls = dls; k.dvar = dvar; k.accumulate_hyp = 0;
    k.scratch = at(ws, pl.w_part);
    k.Kmat = at(ctx, pl.c_Kuf); k.sKmat = MN; k.ldkmat = N;
    GPFX_TRY(launch_kmat_bwd(k, stream));
  }
  {
    KmatBwdArgs k;
    k.U = Z; k.V = Z; k.ls = ls; k.var = var; k.dK = dKuu;
    k.K = K; k.nu = M; k.nv = M; k.D = D;
    k.sU = (long long)M * D; k.sV = (long long)M * D; k.sK = MpMp; k.ldk = Mp;
    k.kind = d->kernel;
    k.dU = dZ; k.sdU = (long long)M * D; k.accumulate_dU = 1;
    k.dV = dZ; k.sdV = (long long)M * D; k.accumulate_dV = 1;
    k.dls = dls; k.dvar = dvar; k.accumulate_hyp = 1;
    k.scratch = at(ws, pl.w_part);
    GPFX_TRY(launch_kmat_bwd(k, stream));
  }
  // Knn = variance: d variance_k += sum_n sum_{p in k} gv
  GPFX_TRY(launch_row_sums(gvsum, K, N, dvar, 1, stream));
  return GPFX_OK;
}

int layer_status(const gpfx_layer_desc* d, const void* ctx, int32_t* info_host,
                 cudaStream_t stream) {
  Plan pl;
  GPFX_TRY(make_plan(d, pl));
  const int* info = reinterpret_cast<const int*>(reinterpret_cast<const char*>(ctx) + pl.c_info);
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
