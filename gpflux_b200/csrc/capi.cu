// extern "C" surface of libgpflux_b200.so (declared in include/gpfx.h).
#include "../../include/gpfx.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "chol.h"
#include "common.cuh"
#include "context.h"
#include "elementwise.h"
#include "gemm_f64.h"
#include "gemm_tf32.h"
#include "kmat.h"
#include "layer.h"
#include "natgrad.h"
#include "rff.h"

namespace gpfx {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  Context* c = current_context();  // also kept per handle (gpfx_handle_last_error)
  std::lock_guard<std::mutex> lock(c->mu);
  snprintf(c->err, sizeof c->err, "%s", g_err);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace gpfx

using namespace gpfx;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int gpfx_version(void) { return GPFX_VERSION; }
const char* gpfx_last_error(void) { return g_err; }

int gpfx_create(int device, gpfx_handle* out) {
  Context* c = nullptr;
  const int st = context_create(device, &c);
  if (st == GPFX_OK) *out = reinterpret_cast<gpfx_handle>(c);
  return st;
}
int gpfx_destroy(gpfx_handle h) { return context_destroy(reinterpret_cast<Context*>(h)); }
int gpfx_set_current(gpfx_handle h) {
  context_set_current(reinterpret_cast<Context*>(h));
  return GPFX_OK;
}
gpfx_handle gpfx_get_current(void) { return reinterpret_cast<gpfx_handle>(current_context()); }
const char* gpfx_handle_last_error(gpfx_handle h) {
  Context* c = h ? reinterpret_cast<Context*>(h) : current_context();
  return c->err;
}
void gpfx_tune_gemm_config(int cfg) { gemm_force_cfg(cfg); }
long long gpfx_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void gpfx_gemm_profile_enable(int on) { gemm_profile_enable(on); }
int gpfx_gemm_profile_read(gpfx_gemm_sample* out, int max) { return gemm_profile_read(out, max); }

size_t gpfx_layer_ctx_bytes(const gpfx_layer_desc* d) { return layer_ctx_bytes(d); }
size_t gpfx_layer_workspace_bytes(const gpfx_layer_desc* d) { return layer_workspace_bytes(d); }

int gpfx_layer_forward(const gpfx_layer_desc* d, const double* X, const double* Z,
                       const double* ls, const double* var, const double* q_mu,
                       const double* q_sqrt, const double* mean_add, const double* eps,
                       double* fmean, double* fvar, double* fsample, double* kl, void* ctx,
                       void* ws, void* stream) {
  return layer_forward(d, X, Z, ls, var, q_mu, q_sqrt, mean_add, eps, fmean, fvar, fsample, kl,
                       ctx, ws, S(stream));
}

int gpfx_layer_backward(const gpfx_layer_desc* d, const double* X, const double* Z,
                        const double* ls, const double* var, const double* q_mu,
                        const double* eps, const double* fvar, const double* g_f,
                        const double* g_mean, const double* g_var, const double* g_kl, double* dX,
                        double* dZ, double* dls, double* dvar, double* dq_mu, double* dq_sqrt,
                        double* dmean, void* ctx, void* ws, void* stream) {
  return layer_backward(d, X, Z, ls, var, q_mu, eps, fvar, g_f, g_mean, g_var, g_kl, dX, dZ, dls,
                        dvar, dq_mu, dq_sqrt, dmean, ctx, ws, S(stream));
}

size_t gpfx_layer_factor_ctx_bytes(const gpfx_layer_desc* d) { return layer_factor_ctx_bytes(d); }
size_t gpfx_layer_cond_ctx_bytes(const gpfx_layer_desc* d) { return layer_cond_ctx_bytes(d); }
size_t gpfx_layer_factor_workspace_bytes(const gpfx_layer_desc* d) {
  return layer_factor_workspace_bytes(d);
}
size_t gpfx_layer_cond_workspace_bytes(const gpfx_layer_desc* d) {
  return layer_cond_workspace_bytes(d);
}
size_t gpfx_layer_dL_bytes(const gpfx_layer_desc* d) { return layer_dL_bytes(d); }

int gpfx_layer_factor_forward(const gpfx_layer_desc* d, const double* Z, const double* ls,
                              const double* var, const double* q_mu, const double* q_sqrt,
                              double* kl, void* fctx, void* ws, void* stream) {
  return layer_factor_forward(d, Z, ls, var, q_mu, q_sqrt, kl, fctx, ws, S(stream));
}

int gpfx_layer_factor_backward(const gpfx_layer_desc* d, const double* Z, const double* ls,
                               const double* var, const double* q_mu, const double* g_kl,
                               const double* dL, double* dZ, double* dls, double* dvar,
                               double* dq_mu, double* dq_sqrt, int accumulate, const void* fctx,
                               void* ws, void* stream) {
  return layer_factor_backward(d, Z, ls, var, q_mu, g_kl, dL, dZ, dls, dvar, dq_mu, dq_sqrt,
                               accumulate, fctx, ws, S(stream));
}

int gpfx_layer_cond_forward(const gpfx_layer_desc* d, const double* X, const double* Z,
                            const double* ls, const double* var, const double* q_mu,
                            const void* fctx, const double* mean_add, const double* eps,
                            double* fmean, double* fvar, double* fsample, void* cctx, void* ws,
                            void* stream) {
  return layer_cond_forward(d, X, Z, ls, var, q_mu, fctx, mean_add, eps, fmean, fvar, fsample, cctx,
                            ws, S(stream));
}

int gpfx_layer_cond_backward(const gpfx_layer_desc* d, const double* X, const double* Z,
                             const double* ls, const double* var, const double* q_mu,
                             const double* eps, const double* fvar, const double* g_f,
                             const double* g_mean, const double* g_var, double* dX, double* dZ,
                             double* dls, double* dvar, double* dq_mu, double* dq_sqrt, double* dL,
                             double* dmean, const void* fctx, void* cctx, void* ws, void* stream) {
  return layer_cond_backward(d, X, Z, ls, var, q_mu, eps, fvar, g_f, g_mean, g_var, dX, dZ, dls,
                             dvar, dq_mu, dq_sqrt, dL, dmean, fctx, cctx, ws, S(stream));
}

int gpfx_layer_status(const gpfx_layer_desc* d, const void* ctx, int32_t* info_host,
                      void* stream) {
  return layer_status(d, ctx, info_host, S(stream));
}

int gpfx_layer_status_async(const gpfx_layer_desc* d, const void* ctx, int32_t* info_dev,
                            void* stream) {
  return layer_status_async(d, ctx, info_dev, S(stream));
}

int gpfx_kernel_matrix(int kernel, int K, int nu, int nv, int D, const double* U, int u_shared,
                       const double* V, int v_shared, const double* ls, const double* var,
                       int add_jitter, double jitter, double* out, void* stream) {
  if (K <= 0 || nu <= 0 || nv <= 0 || D <= 0) {
    set_last_error("gpfx_kernel_matrix: bad shape");
    return GPFX_E_SHAPE;
  }
  KmatArgs k;
  k.U = U; k.V = V; k.ls = ls; k.var = var; k.out = out;
  k.K = K; k.nu = nu; k.nv = nv; k.D = D;
  k.sU = u_shared ? 0 : (long long)nu * D;
  k.sV = v_shared ? 0 : (long long)nv * D;
  k.sOut = (long long)nu * nv;
  k.ldo = nv; k.rows_out = nu; k.cols_out = nv;
  k.kind = kernel; k.add_jitter = add_jitter; k.jitter = jitter;
  return launch_kmat_fwd(k, S(stream));
}

size_t gpfx_kernel_matrix_bwd_scratch_bytes(int K, int nu, int nv, int D) {
  return kmat_bwd_scratch_doubles(K, nu, nv, D) * sizeof(double);
}

int gpfx_kernel_matrix_bwd(int kernel, int K, int nu, int nv, int D, const double* U,
                           const double* V, int v_shared, const double* ls, const double* var,
                           double* dK, double* dU, double* dV, double* dls, double* dvar,
                           void* scratch, void* stream) {
  KmatBwdArgs k;
  k.U = U; k.V = V; k.ls = ls; k.var = var; k.dK = dK;
  k.K = K; k.nu = nu; k.nv = nv; k.D = D;
  k.sU = (long long)nu * D; k.sV = v_shared ? 0 : (long long)nv * D;
  k.sK = (long long)nu * nv; k.ldk = nv;
  k.kind = kernel;
  k.dU = dU; k.sdU = (long long)nu * D;
  k.dV = dV; k.sdV = v_shared ? 0 : (long long)nv * D;
  k.dls = dls; k.dvar = dvar;
  k.scratch = reinterpret_cast<double*>(scratch);
  // same selection as the layer's Kuf backward; D == 5 keeps the GEMM form covered at small D by the tests
  k.gemm_form = (nv >= 128) ? ((D <= 15 && D != 5) ? 2 : 1) : 0;
  return launch_kmat_bwd(k, S(stream));
}

size_t gpfx_natgrad_workspace_bytes(int M, int P) { return natgrad_workspace_bytes(M, P); }

int gpfx_natgrad_step(int M, int P, double gamma, const double* q_mu, const double* q_sqrt,
                      const double* dq_mu, const double* dq_sqrt, double* q_mu_new,
                      double* q_sqrt_new, int32_t* info, void* workspace, void* stream) {
  return natgrad_step(M, P, gamma, q_mu, q_sqrt, dq_mu, dq_sqrt, q_mu_new, q_sqrt_new, info,
                      workspace, S(stream));
}

size_t gpfx_cholesky_workspace_bytes(int K, int M) {
  const size_t Mp = chol_padded_dim(M);
  return (2 * (size_t)K * Mp * Mp + chol_scratch_doubles(K, (int)Mp)) * sizeof(double);
}

namespace {
__global__ void pad_in_kernel(const double* A, double* Lp, int M, int Mp) {
  const long long k = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)Mp * Mp;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / Mp), j = (int)(e % Mp);
    Lp[k * Mp * Mp + e] = (i < M && j < M) ? A[k * M * M + (long long)i * M + j] : (i == j ? 1.0 : 0.0);
  }
}
__global__ void pad_out_kernel(const double* Lp, const double* Lip, double* L, double* Linv, int M,
                               int Mp) {
  const long long k = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * M;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    L[k * M * M + e] = Lp[k * Mp * Mp + (long long)i * Mp + j];
    Linv[k * M * M + e] = Lip[k * Mp * Mp + (long long)i * Mp + j];
  }
}
// dLp = tril(dL) zero-padded to Mp
__global__ void pad_tril_zero_kernel(const double* dL, double* dLp, int M, int Mp) {
  const long long k = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)Mp * Mp;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / Mp), j = (int)(e % Mp);
    dLp[k * Mp * Mp + e] = (i < M && j <= i) ? dL[k * M * M + (long long)i * M + j] : 0.0;
  }
}
// dA = (dKp + dKp^T) / 2 on the leading M x M block
__global__ void unpad_sym_kernel(const double* dKp, double* dA, int M, int Mp) {
  const long long k = blockIdx.y;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * M;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    const double* b = dKp + k * Mp * Mp;
    dA[k * M * M + e] = 0.5 * (b[(long long)i * Mp + j] + b[(long long)j * Mp + i]);
  }
}
}  // namespace

size_t gpfx_cholesky_backward_workspace_bytes(int K, int M) {
  const size_t Mp = (size_t)chol_padded_dim(M);
  return 6 * (size_t)K * Mp * Mp * sizeof(double);
}

int gpfx_cholesky_backward(int K, int M, const double* L, const double* Linv, const double* dL, double* dA,
                           void* workspace, void* stream) {
  if (K <= 0 || M <= 0 || !L || !Linv || !dL || !dA || !workspace) {
    set_last_error("gpfx_cholesky_backward: bad arguments");
    return GPFX_E_SHAPE;
  }
  const int Mp = chol_padded_dim(M);
  const size_t n = (size_t)K * Mp * Mp;
  double* Lp = reinterpret_cast<double*>(workspace);
  double *Lip = Lp + n, *dLp = Lip + n, *T1 = dLp + n, *T2 = T1 + n, *dKp = T2 + n;
  dim3 grid(ceil_div(Mp * Mp, 256) < 1024 ? ceil_div(Mp * Mp, 256) : 1024, K);
  pad_in_kernel<<<grid, 256, 0, S(stream)>>>(L, Lp, M, Mp);
  GPFX_LAUNCH_CHECK();
  pad_in_kernel<<<grid, 256, 0, S(stream)>>>(Linv, Lip, M, Mp);
  GPFX_LAUNCH_CHECK();
  pad_tril_zero_kernel<<<grid, 256, 0, S(stream)>>>(dL, dLp, M, Mp);
  GPFX_LAUNCH_CHECK();
  GPFX_TRY(chol_backward_batched(Lp, Lip, dLp, T1, T2, dKp, K, Mp, S(stream)));
  unpad_sym_kernel<<<grid, 256, 0, S(stream)>>>(dKp, dA, M, Mp);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int gpfx_cholesky_inverse(int K, int M, const double* A, double* L, double* Linv, int32_t* info,
                          void* workspace, void* stream) {
  if (K <= 0 || M <= 0) {
    set_last_error("gpfx_cholesky_inverse: bad shape");
    return GPFX_E_SHAPE;
  }
  const int Mp = chol_padded_dim(M);
  double* Lp = reinterpret_cast<double*>(workspace);
  double* Lip = Lp + (size_t)K * Mp * Mp;
  double* T = Lip + (size_t)K * Mp * Mp;
  dim3 grid(ceil_div(Mp * Mp, 256) < 1024 ? ceil_div(Mp * Mp, 256) : 1024, K);
  pad_in_kernel<<<grid, 256, 0, S(stream)>>>(A, Lp, M, Mp);
  GPFX_LAUNCH_CHECK();
  GPFX_TRY(potrf_trtri_batched(Lp, Lip, T, K, Mp, info, S(stream)));
  pad_out_kernel<<<grid, 256, 0, S(stream)>>>(Lp, Lip, L, Linv, M, Mp);
  GPFX_LAUNCH_CHECK();
  return GPFX_OK;
}

int gpfx_gemm(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
              long long strideA, int transA, const double* B, int ldb, long long strideB,
              int transB, double beta, double* C, int ldc, long long strideC, int tri,
              int tril_out, void* stream) {
  GemmArgs g;
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = Kdim;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.batch1 = batch; g.sA1 = strideA; g.sB1 = strideB; g.sC1 = strideC;
  g.transA = transA; g.transB = transB;
  g.alpha = alpha; g.beta = beta;
  g.tri = tri; g.tril_out = tril_out;
  return launch_gemm(g, S(stream));
}

int gpfx_gemm_splitk(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
                     long long strideA, int transA, const double* B, int ldb, long long strideB,
                     int transB, double beta, double* C, int ldc, long long strideC, int tri,
                     int tril_out, int splits, void* workspace, void* stream) {
  GemmArgs g;
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = Kdim;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.batch1 = batch; g.sA1 = strideA; g.sB1 = strideB; g.sC1 = strideC;
  g.transA = transA; g.transB = transB;
  g.alpha = alpha; g.beta = beta;
  g.tri = tri; g.tril_out = tril_out;
  g.splits = splits; g.splitws = reinterpret_cast<double*>(workspace);
  return launch_gemm(g, S(stream));
}

int gpfx_gemm_kscaled(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
                      long long strideA, const double* B, int ldb, long long strideB,
                      const double* kscale, long long strideScale, double* C, int ldc,
                      long long strideC, int tril_out, int splits, void* workspace, void* stream) {
  GemmArgs g;
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = Kdim;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.batch1 = batch; g.sA1 = strideA; g.sB1 = strideB; g.sC1 = strideC;
  g.transA = 0; g.transB = 1;
  g.alpha = alpha;
  g.tril_out = tril_out;
  g.splits = splits; g.splitws = reinterpret_cast<double*>(workspace);
  g.kscale = kscale; g.sKscale = strideScale;
  return launch_gemm(g, S(stream));
}

int gpfx_reparam_sample(const double* mean, const double* var, const double* eps, double* f,
                        long long n, void* stream) {
  return launch_reparam_sample(mean, var, eps, f, n, S(stream));
}

int gpfx_philox_raw(unsigned long long seed, unsigned long long offset, uint32_t* out,
                    long long n_pairs, void* stream) {
  return launch_philox_raw(seed, offset, out, n_pairs, S(stream));
}

int gpfx_randn_philox(unsigned long long seed, unsigned long long offset, double* out, long long n,
                      void* stream) {
  return launch_philox_normal(seed, offset, nullptr, nullptr, out, nullptr, n, S(stream));
}

int gpfx_reparam_sample_philox(const double* mean, const double* var, unsigned long long seed,
                               unsigned long long offset, double* f, double* eps_out, long long n,
                               void* stream) {
  if (!mean || !var) {
    set_last_error("reparam_sample_philox: mean and var are required");
    return GPFX_E_SHAPE;
  }
  return launch_philox_normal(seed, offset, mean, var, f, eps_out, n, S(stream));
}

size_t gpfx_kl_scratch_bytes(int M, int P) { return kl_scratch_doubles(M, P) * sizeof(double); }

int gpfx_kl_white(int M, int P, const double* q_mu, const double* q_sqrt, double* kl,
                  void* scratch, void* stream) {
  return launch_kl_white(q_mu, q_sqrt, nullptr, kl, reinterpret_cast<double*>(scratch), M, P,
                         S(stream));
}

int gpfx_kl_white_tril(int M, int P, const double* q_mu, const double* q_sqrt, double* Lq, double* kl,
                       void* scratch, void* stream) {
  if (M <= 0 || P <= 0 || !q_mu || !q_sqrt || !Lq || !kl || !scratch) {
    set_last_error("gpfx_kl_white_tril: bad arguments");
    return GPFX_E_SHAPE;
  }
  return launch_kl_white(q_mu, q_sqrt, Lq, kl, reinterpret_cast<double*>(scratch), M, P, S(stream));
}

int gpfx_kl_white_bwd(int M, int P, const double* q_mu, const double* q_sqrt, const double* g_kl,
                      double* dq_mu, double* dq_sqrt, void* stream) {
  return launch_kl_white_bwd(q_mu, q_sqrt, g_kl, dq_mu, dq_sqrt, M, P, 0, S(stream));
}

size_t gpfx_gaussian_ve_scratch_bytes(long long n) {
  return gaussian_ve_scratch_doubles(n) * sizeof(double);
}

int gpfx_gaussian_ve(const double* Fmu, const double* Fvar, const double* Y,
                     const double* noise_variance, long long n, double* ve_sum,
                     const double* scale, double* dFmu, double* dFvar, double* dnoise,
                     void* scratch, void* stream) {
  return launch_gaussian_ve(Fmu, Fvar, Y, noise_variance, n, ve_sum, scale, dFmu, dFvar, dnoise,
                            reinterpret_cast<double*>(scratch), S(stream));
}

size_t gpfx_latent_scratch_bytes(long long N, int W) {
  return latent_scratch_doubles(N, W) * sizeof(double);
}

int gpfx_latent_forward(long long N, int D, int W, const double* X, const double* mean,
                        const double* std, int per_row, const double* prior_mean,
                        const double* prior_std, const double* eps, double* out, double* kl_sum,
                        void* scratch, void* stream) {
  if (N <= 0 || D < 0 || W <= 0) {
    set_last_error("gpfx_latent_forward: bad shape");
    return GPFX_E_SHAPE;
  }
  return launch_latent_forward(X, mean, std, per_row ? W : 0, prior_mean, prior_std, eps, out, kl_sum,
                               reinterpret_cast<double*>(scratch), N, D, W, S(stream));
}

int gpfx_latent_backward(long long N, int D, int W, const double* mean, const double* std,
                         const double* prior_mean, const double* prior_std, const double* eps,
                         const double* g_out, const double* g_kl, double* dX, double* dmean,
                         double* dstd, void* stream) {
  return launch_latent_backward(mean, std, prior_mean, prior_std, eps, g_out, g_kl, dX, dmean, dstd,
                                N, D, W, S(stream));
}

int gpfx_rff_features(int K, int N, int D, int L, int concat, const double* X, const double* W,
                      const double* bias, const double* ls, const double* var, double* out,
                      void* stream) {
  if (K <= 0 || N <= 0 || D <= 0 || L <= 0 || (!concat && bias == nullptr)) {
    set_last_error("gpfx_rff_features: bad arguments");
    return GPFX_E_SHAPE;
  }
  KmatArgs k;
  k.U = X; k.V = W; k.ls = ls; k.var = var; k.out = out;
  k.K = K; k.nu = N; k.nv = L; k.D = D;
  k.sU = 0; k.sV = (long long)L * D;
  const int width = concat ? 2 * L : L;
  k.sOut = (long long)N * width; k.ldo = width; k.rows_out = N; k.cols_out = width;
  k.mode = concat ? 2 : 1;
  k.bias = bias; k.sBias = L;
  return launch_kmat_fwd(k, S(stream));
}

size_t gpfx_rff_backward_scratch_bytes(int K, int N, int D, int L) {
  return rff_bwd_scratch_doubles(K, N, D, L) * sizeof(double);
}

int gpfx_rff_backward(int K, int N, int D, int L, int concat, const double* X, const double* W,
                      const double* bias, const double* ls, const double* var, const double* dPhi,
                      double* dX, double* dls, double* dvar, void* scratch, void* stream) {
  if (K <= 0 || N <= 0 || D <= 0 || L <= 0 || (!concat && bias == nullptr) || !X || !W || !ls || !var || !dPhi ||
      !dX || !dls || !dvar || !scratch) {
    set_last_error("gpfx_rff_backward: bad arguments");
    return GPFX_E_SHAPE;
  }
  RffBwdArgs a;
  a.K = K; a.N = N; a.D = D; a.L = L; a.concat = concat;
  a.X = X; a.W = W; a.bias = bias; a.ls = ls; a.var = var; a.dPhi = dPhi;
  a.dX = dX; a.dls = dls; a.dvar = dvar; a.scratch = reinterpret_cast<double*>(scratch);
  return launch_rff_bwd(a, S(stream));
}

int gpfx_wilson_eval(int kernel, int Sn, int N, int D, int P, int L, int M, const double* X,
                     int x_shared, const double* W, const double* b, const double* ls,
                     const double* var, const double* Z, const double* w, const double* v,
                     const double* mean_add, double* out, void* stream) {
  WilsonArgs a;
  a.kind = kernel; a.S = Sn; a.N = N; a.D = D; a.P = P; a.L = L; a.M = M;
  a.X = X; a.sX = x_shared ? 0 : (long long)N * D;
  a.W = W; a.b = b; a.ls = ls; a.var = var; a.Z = Z; a.w = w; a.v = v;
  a.mean_add = mean_add; a.out = out;
  return launch_wilson_eval(a, S(stream));
}

size_t gpfx_wilson_setup_scratch_bytes(int Sn, int M, int D, int P, int L) {
  return wilson_setup_scratch_doubles(Sn, M, D, P, L) * sizeof(double);
}

int gpfx_wilson_setup(int kernel, int whiten, int Sn, int M, int D, int P, int L, double jitter,
                      const double* Z, const double* ls, const double* var, const double* q_mu,
                      const double* q_sqrt, const double* W, const double* b, const double* coeffs,
                      const double* eps_w, const double* eps_u, double* w, double* v, int32_t* info,
                      void* scratch, void* stream) {
  if (!Z || !ls || !var || !q_mu || !q_sqrt || !W || !b || !coeffs || !eps_w || !eps_u || !w || !v || !info ||
      !scratch) {
    set_last_error("gpfx_wilson_setup: null pointer argument");
    return GPFX_E_SHAPE;
  }
  WilsonSetupArgs a;
  a.kind = kernel; a.whiten = whiten; a.S = Sn; a.M = M; a.D = D; a.P = P; a.L = L; a.jitter = jitter;
  a.Z = Z; a.ls = ls; a.var = var; a.q_mu = q_mu; a.q_sqrt = q_sqrt; a.W = W; a.b = b; a.coeffs = coeffs;
  a.eps_w = eps_w; a.eps_u = eps_u; a.w = w; a.v = v; a.info = info;
  a.scratch = reinterpret_cast<double*>(scratch);
  return launch_wilson_setup(a, S(stream));
}

size_t gpfx_gemm_tf32x3_workspace_bytes(int batch, int M, int N, int Kdim) {
  return gemm_tf32_workspace_floats(batch, M, N, Kdim) * sizeof(float);
}

int gpfx_gemm_tf32x3(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
                     long long strideA, int transA, const double* B, int ldb, long long strideB,
                     int transB, double beta, double* C, int ldc, long long strideC, int tri,
                     int tril_out, void* workspace, void* stream) {
  if (!A || !B || !C || !workspace) {
    set_last_error("gpfx_gemm_tf32x3: null pointer argument");
    return GPFX_E_SHAPE;
  }
  GemmArgs g;
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = Kdim;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.batch1 = batch; g.sA1 = strideA; g.sB1 = strideB; g.sC1 = strideC;
  g.transA = transA; g.transB = transB;
  g.alpha = alpha; g.beta = beta;
  g.tri = tri; g.tril_out = tril_out;
  return launch_gemm_tf32x3(g, reinterpret_cast<float*>(workspace), S(stream));
}

int gpfx_set_contraction(int mode, void* workspace, size_t workspace_bytes) {
  if (mode != 0 && mode != 1) {
    set_last_error("gpfx_set_contraction: mode must be 0 (FP64 DMMA) or 1 (3xTF32)");
    return GPFX_E_UNSUPPORTED;
  }
  Context* c = current_context();
  c->contraction = mode;
  c->tf32_ws = reinterpret_cast<float*>(workspace);
  c->tf32_ws_bytes = workspace_bytes;
  return GPFX_OK;
}
int gpfx_get_contraction(void) { return current_context()->contraction; }
size_t gpfx_layer_tf32_workspace_bytes(const gpfx_layer_desc* d) { return layer_tf32_workspace_bytes(d); }

}  // extern "C"
