/*
 * gpfx.h - C ABI of libgpflux_b200.so: the B200 (sm_100a) implementation of GPflux's per-layer
 * sparse-variational GP hot path.
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer to C-contiguous float64 unless it says "host";
 *   - the caller owns all memory, including ctx (saved-for-backward) and workspace (scratch)
 *     buffers whose sizes come from the *_bytes queries; the library never allocates;
 *   - all work is ordered on `stream` (a cudaStream_t passed as void*); no hidden syncs, so the
 *     entry points are CUDA-graph capturable.  The layer entry points may fork part of their work
 *     onto a library-owned non-blocking stream and join it back with events before they return
 *     (parallel branches in a captured graph); GPFX_AUX_STREAM=0 in the environment disables that;
 *   - return value: 0 on success, < 0 = GPFX_E_*; gpfx_last_error() gives the message of the last
 *     failure on the calling thread.
 *
 * Each entry point names the reference interface it replaces (paths relative to the GPflux
 * repository, v0.4.4).  INTEGRATION.md shows the binding a GPflux maintainer would add.
 */
#ifndef GPFX_H_
#define GPFX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPFX_VERSION 100 /* 0.1.0 */

#define GPFX_OK 0
#define GPFX_E_SHAPE (-1)
#define GPFX_E_DTYPE (-2)
#define GPFX_E_LAYOUT (-3)
#define GPFX_E_NOT_PD (-4)
#define GPFX_E_CUDA (-5)
#define GPFX_E_NCCL (-6)
#define GPFX_E_UNSUPPORTED (-7)
#define GPFX_E_WORKSPACE (-8)

/* gpflow.kernels.{SquaredExponential, Matern12, Matern32, Matern52} */
#define GPFX_KERNEL_SE 0
#define GPFX_KERNEL_MATERN12 1
#define GPFX_KERNEL_MATERN32 2
#define GPFX_KERNEL_MATERN52 3

/* A = L^-1 Kuf (gpflow base_conditional's triangular_solve, reached from gp_layer.py:257-266):
 *   INVERSE_GEMM  L^-1 is formed once per step (M^3/3, parameter-only) and applied as a DMMA GEMM with the
 *                 column reductions fused into its epilogue - the fast default;
 *   TRSM          blocked forward / back substitution with L itself (substitution kernels on the 64-row
 *                 diagonal blocks, DMMA GEMM updates off the diagonal), no inverse anywhere on the
 *                 N-sized path - the reference-faithful form, for badly conditioned Kuu. */
#define GPFX_SOLVER_INVERSE_GEMM 0
#define GPFX_SOLVER_TRSM 1

/* One GPLayer evaluated on one minibatch shard. */
typedef struct gpfx_layer_desc {
  int32_t N;      /* rows of the minibatch on this GPU                                   */
  int32_t M;      /* inducing points                                                     */
  int32_t D;      /* layer input dimension                                               */
  int32_t P;      /* latent GPs (q_mu [M,P], q_sqrt [P,M,M])                             */
  int32_t K;      /* distinct (kernel, Z) pairs: 1 = SharedIndependent, P = SeparateInd. */
  int32_t kernel; /* GPFX_KERNEL_*                                                       */
  int32_t whiten; /* 1 = whitened q(v) (GPLayer default); 0 = q(u) = N(q_mu, q_sqrt q_sqrt^T) */
  int32_t solver; /* GPFX_SOLVER_*: how A = L^-1 Kuf (and its transpose in the backward) is evaluated */
  double jitter;  /* gpflow.config.default_jitter() = 1e-6                               */
} gpfx_layer_desc;

int gpfx_version(void);
const char* gpfx_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Library handle (SURVEY §8b "gpfx_create(device, stream?) -> handle / gpfx_destroy").
 * A handle owns what the entry points need besides the caller's buffers: the pool of auxiliary
 * streams and events that the layer backward / the overlapped factorisation fork side strands
 * onto (created HERE, outside of any stream capture), the GEMM tuning override and the last
 * error text.  Entry points act on the calling thread's CURRENT handle; a thread that never calls
 * gpfx_set_current shares the process default handle (created lazily, never destroyed), which is
 * what the reference's single-threaded Python host needs.  Different host threads may drive
 * different handles concurrently; gpfx_destroy must not race with calls that use the handle.
 * The library never allocates device memory on behalf of a handle.
 * ------------------------------------------------------------------------------------------ */
typedef struct gpfx_context_s* gpfx_handle;
int gpfx_create(int device, gpfx_handle* out);
int gpfx_destroy(gpfx_handle h);
int gpfx_set_current(gpfx_handle h); /* NULL: back to the process default handle */
gpfx_handle gpfx_get_current(void);
const char* gpfx_handle_last_error(gpfx_handle h); /* NULL: the current handle */
/* tuning hook: force the DMMA GEMM tile configuration used for large problems
 * (0: 128x128/8 warps, 2: 128x128/16 warps, 3: 64x256/8 warps; -1 = automatic). */
void gpfx_tune_gemm_config(int cfg);
/* number of CUDA kernels this library has launched in this process (diagnostics / bench.py) */
long long gpfx_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GPLayer forward: GPLayer.predict + prior_kl + the reparameterised sample.
 * Replaces gpflux/layers/gp_layer.py:226-268 (predict -> gpflow.conditionals.conditional),
 * :303-311 (prior_kl -> gauss_kl) and :339-363 (MultivariateNormalDiag(...).sample()).
 *
 *   X [N,D]; Z [K,M,D]; lengthscales [K,D]; variance [K]; q_mu [M,P]; q_sqrt [P,M,M]
 *   mean_add [N,P] or NULL : mean_function(X), added to the conditional mean (:268)
 *   eps      [N,P] or NULL : the N(0,1) draw; NULL -> fsample = fmean
 *   fmean, fvar [N,P]; fsample [N,P] or NULL; kl [1]
 * ------------------------------------------------------------------------------------------ */
size_t gpfx_layer_ctx_bytes(const gpfx_layer_desc* desc);
size_t gpfx_layer_workspace_bytes(const gpfx_layer_desc* desc);
int gpfx_layer_forward(const gpfx_layer_desc* desc, const double* X, const double* Z,
                       const double* lengthscales, const double* variance, const double* q_mu,
                       const double* q_sqrt, const double* mean_add, const double* eps,
                       double* fmean, double* fvar, double* fsample, double* kl, void* ctx,
                       void* workspace, void* stream);

/* GPLayer backward: what TF autodiff produces for the ops above (GradientTape in
 * gpflux/optimization/keras_natgrad.py:207-211 / Keras train_step).
 *   g_fsample, g_fmean, g_fvar [N,P] or NULL : upstream gradients; g_kl [1] (device)
 *   dX [N,D] (through the kernel only), dZ [K,M,D], dlengthscales [K,D], dvariance [K],
 *   dq_mu [M,P], dq_sqrt [P,M,M] (lower triangle, zeros above), dmean [N,P] = dLoss/d fmean
 *   (what the host back-propagates into the mean function).
 * ctx must be the buffer filled by the matching gpfx_layer_forward; it is consumed (modified). */
int gpfx_layer_backward(const gpfx_layer_desc* desc, const double* X, const double* Z,
                        const double* lengthscales, const double* variance, const double* q_mu,
                        const double* eps, const double* fvar, const double* g_fsample,
                        const double* g_fmean, const double* g_fvar, const double* g_kl,
                        double* dX, double* dZ, double* dlengthscales, double* dvariance,
                        double* dq_mu, double* dq_sqrt, double* dmean, void* ctx,
                        void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * The same layer, cut into its parameter-only part ("factor": tril(q_sqrt), KL, Kuu, Cholesky,
 * L^-1) and its data part ("cond": Kuf, A, B_p, mean/variance/sample).  The factor part does not
 * depend on the minibatch, so a caller may run the factor parts of all layers on other streams
 * (or as parallel CUDA-graph branches) while the data path of earlier layers is executing.
 * gpfx_layer_forward == factor_forward; cond_forward and
 * gpfx_layer_backward == cond_backward; factor_backward(accumulate = 1) on one stream.
 *
 *   fctx (gpfx_layer_factor_ctx_bytes): Lq, L, L^-1, Cholesky status - written by factor_forward,
 *        read-only afterwards;  cctx (gpfx_layer_cond_ctx_bytes): Kuf, A, B - consumed by
 *        cond_backward;  dL [K,Mp,Mp] (gpfx_layer_dL_bytes): dLoss/dL, produced by cond_backward and
 *        consumed by factor_backward.  The two parts need separate workspaces when they may overlap.
 *   factor_backward: `accumulate` is a bit mask.  Bit 0: dZ, dlengthscales, dvariance already hold the
 *        data part and are added to (else overwritten).  Bit 1: dq_mu / dq_sqrt hold, on entry, the
 *        data-part cotangents w.r.t. the EFFECTIVE variational parameters cond_backward produced
 *        (whiten = 1: q_mu and tril(q_sqrt) themselves; whiten = 0: alpha = L^-1 q_mu and
 *        C_p = L^-1 tril(q_sqrt_p)) and hold, on exit, the total gradients w.r.t. q_mu / q_sqrt; without
 *        bit 1 they are overwritten with the KL part alone.  g_kl may be NULL (no KL gradient).
 *   Non-whitened layers are evaluated as the whitened layer in (alpha, C): conditional and KL are
 *   identical (u = L v), at O(P M^3) extra cost instead of the reference's extra M^2 N solve.
 * ------------------------------------------------------------------------------------------ */
size_t gpfx_layer_factor_ctx_bytes(const gpfx_layer_desc* desc);
size_t gpfx_layer_cond_ctx_bytes(const gpfx_layer_desc* desc);
size_t gpfx_layer_factor_workspace_bytes(const gpfx_layer_desc* desc);
size_t gpfx_layer_cond_workspace_bytes(const gpfx_layer_desc* desc);
size_t gpfx_layer_dL_bytes(const gpfx_layer_desc* desc);
int gpfx_layer_factor_forward(const gpfx_layer_desc* desc, const double* Z,
                              const double* lengthscales, const double* variance,
                              const double* q_mu, const double* q_sqrt, double* kl, void* fctx,
                              void* workspace, void* stream);
int gpfx_layer_factor_backward(const gpfx_layer_desc* desc, const double* Z,
                               const double* lengthscales, const double* variance,
                               const double* q_mu, const double* g_kl, const double* dL,
                               double* dZ, double* dlengthscales, double* dvariance,
                               double* dq_mu, double* dq_sqrt, int accumulate, const void* fctx,
                               void* workspace, void* stream);
int gpfx_layer_cond_forward(const gpfx_layer_desc* desc, const double* X, const double* Z,
                            const double* lengthscales, const double* variance,
                            const double* q_mu, const void* fctx, const double* mean_add,
                            const double* eps, double* fmean, double* fvar, double* fsample,
                            void* cctx, void* workspace, void* stream);
int gpfx_layer_cond_backward(const gpfx_layer_desc* desc, const double* X, const double* Z,
                             const double* lengthscales, const double* variance,
                             const double* q_mu, const double* eps, const double* fvar,
                             const double* g_fsample, const double* g_fmean, const double* g_fvar,
                             double* dX, double* dZ, double* dlengthscales, double* dvariance,
                             double* dq_mu, double* dq_sqrt, double* dL, double* dmean,
                             const void* fctx, void* cctx, void* workspace, void* stream);

/* Synchronises `stream` and copies the Cholesky status out of ctx (or fctx): info_host[k] = 0, or
 * 1 + index of the first non-positive pivot of Kuu_k (TF raises InvalidArgumentError there).
 * Returns GPFX_E_NOT_PD if any entry is non-zero. */
int gpfx_layer_status(const gpfx_layer_desc* desc, const void* ctx, int32_t* info_host,
                      void* stream);
/* The same status copied to a caller-owned DEVICE buffer info_dev [K], ordered on `stream`, without any
 * synchronisation (CUDA-graph capturable): the host reads it whenever it next synchronises anyway (e.g.
 * when the loss is read back) and raises there - how a non-PD Kuu surfaces on the training path. */
int gpfx_layer_status_async(const gpfx_layer_desc* desc, const void* ctx, int32_t* info_dev,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks (also the units the parity tests exercise one by one)
 * ------------------------------------------------------------------------------------------ */

/* gpflow.covariances.Kuu / Kuf for InducingPoints (gp_layer.py:257-266):
 * out[k] = k_k(U[k], V[k]) (+ jitter * I if add_jitter).  U [K,nu,D] (or [nu,D] with
 * u_shared), V likewise; out [K,nu,nv]. */
int gpfx_kernel_matrix(int kernel, int K, int nu, int nv, int D, const double* U, int u_shared,
                       const double* V, int v_shared, const double* lengthscales,
                       const double* variance, int add_jitter, double jitter, double* out,
                       void* stream);

/* Backward of gpfx_kernel_matrix.  dK [K,nu,nv] is consumed (overwritten).  dV is [nv,D] when
 * v_shared (summed over k) else [K,nv,D].  scratch: gpfx_kernel_matrix_bwd_scratch_bytes. */
size_t gpfx_kernel_matrix_bwd_scratch_bytes(int K, int nu, int nv, int D);
int gpfx_kernel_matrix_bwd(int kernel, int K, int nu, int nv, int D, const double* U,
                           const double* V, int v_shared, const double* lengthscales,
                           const double* variance, double* dK, double* dU, double* dV,
                           double* dlengthscales, double* dvariance, void* scratch,
                           void* stream);

/* tf.linalg.cholesky + L^-1 (gpflux/math.py:25-62, base_conditional).  A [K,M,M] symmetric PD
 * (lower triangle read).  L, Linv [K,M,M] out.  info [K] int32 device. */
size_t gpfx_cholesky_workspace_bytes(int K, int M);
int gpfx_cholesky_inverse(int K, int M, const double* A, double* L, double* Linv, int32_t* info,
                          void* workspace, void* stream);

/* Backward of gpfx_cholesky_inverse w.r.t. A, given dLoss/dL (lower triangle used):
 *   dA = sym(L^-T Phi(L^T dL) L^-1),  Phi = lower triangle with halved diagonal
 * (the adjoint tf.linalg.cholesky registers; a gradient w.r.t. L^-1 is folded in by the caller as
 * dL -= tril(L^-T dLinv L^-T)).  workspace: gpfx_cholesky_backward_workspace_bytes(K, M). */
size_t gpfx_cholesky_backward_workspace_bytes(int K, int M);
int gpfx_cholesky_backward(int K, int M, const double* L, const double* Linv, const double* dL,
                           double* dA, void* workspace, void* stream);

/* General FP64 tensor-core GEMM used by every contraction of the path:
 * C[b] = alpha * op(A[b]) op(B[b]) + beta * C[b]; row-major; op = transpose if trans? != 0;
 * tri = bit mask of GPFX_TRI_* describing known-zero triangles that are skipped. */
#define GPFX_TRI_LOWER_A 1
#define GPFX_TRI_UPPER_A 2
#define GPFX_TRI_LOWER_B 4
#define GPFX_TRI_UPPER_B 8
int gpfx_gemm(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
              long long strideA, int transA, const double* B, int ldb, long long strideB,
              int transB, double beta, double* C, int ldc, long long strideC, int tri,
              int tril_out, void* stream);

/* Same, with split-K: the contraction is cut into `splits` slices whose partial products go to
 * `workspace` (splits * batch * M * N doubles) and are then reduced in a fixed order. */
int gpfx_gemm_splitk(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
                     long long strideA, int transA, const double* B, int ldb, long long strideB,
                     int transB, double beta, double* C, int ldc, long long strideC, int tri,
                     int tril_out, int splits, void* workspace, void* stream);

/* C[b] = alpha * A[b] diag(kscale[b]) B[b]^T (optionally only its lower triangle): A [M, Kdim] and
 * B [N, Kdim] row-major, kscale [batch][Kdim].  The scale is applied to the operand fragments inside
 * the TMA-fed kernel, no scaled copy is stored; this is how the layer backward forms
 * dq_sqrt_p = tril(A diag(2 gv_p) B_p^T) (autodiff transpose of the fvar term of GPflow's
 * base_conditional, reached from gpflux/layers/gp_layer.py:257-266).  Returns GPFX_E_UNSUPPORTED when the
 * shape is not one the TMA kernel takes (M, N >= 128 and at least 296 output tiles of 128 x 128 over
 * batch * splits, Kdim a multiple of 16, even strides, 16-byte aligned pointers); split-K as in
 * gpfx_gemm_splitk. */
int gpfx_gemm_kscaled(int batch, int M, int N, int Kdim, double alpha, const double* A, int lda,
                      long long strideA, const double* B, int ldb, long long strideB,
                      const double* kscale, long long strideScale, double* C, int ldc,
                      long long strideC, int tril_out, int splits, void* workspace, void* stream);

/* a8: f = mean + sqrt(var) * eps over n elements (gp_layer.py:339-363). */
int gpfx_reparam_sample(const double* mean, const double* var, const double* eps, double* f,
                        long long n, void* stream);

/* a8 with the noise drawn inside the kernel.  NOT a parity path: GPflux's noise comes from TFP's
 * generator (gp_layer.py:358-363), whose stream is not reproducible outside TF; this variant is
 * equal in distribution only and exists for throughput (no eps array to produce, read or keep - the
 * backward pass can regenerate it).  Generator: Philox4x32-10, key = seed, counter = (pair index i,
 * offset); the four words give two 53-bit uniforms and Box-Muller gives elements 2i, 2i+1.
 *   gpfx_philox_raw             out [n_pairs,4] uint32 (16-byte aligned): the generator's words;
 *   gpfx_randn_philox           out [n] = eps;
 *   gpfx_reparam_sample_philox  f [n] = mean + sqrt(var) eps; eps_out [n] or NULL. */
int gpfx_philox_raw(unsigned long long seed, unsigned long long offset, uint32_t* out,
                    long long n_pairs, void* stream);
int gpfx_randn_philox(unsigned long long seed, unsigned long long offset, double* out, long long n,
                      void* stream);
int gpfx_reparam_sample_philox(const double* mean, const double* var, unsigned long long seed,
                               unsigned long long offset, double* f, double* eps_out, long long n,
                               void* stream);

/* a9: whitened KL (gp_layer.py:303-311).  kl [1].  scratch: gpfx_kl_scratch_bytes(M, P). */
size_t gpfx_kl_scratch_bytes(int M, int P);
int gpfx_kl_white(int M, int P, const double* q_mu, const double* q_sqrt, double* kl,
                  void* scratch, void* stream);

/* The same pass as the layer runs it (first kernel of gpfx_layer_factor_forward): besides the KL it
 * writes Lq [P,M,M] = tril(q_sqrt) (zeros above the diagonal), the operand of the B_p = Lq_p^T A GEMM. */
int gpfx_kl_white_tril(int M, int P, const double* q_mu, const double* q_sqrt, double* Lq, double* kl,
                       void* scratch, void* stream);

/* Backward of gpfx_kl_white: dq_mu [M,P] = g_kl * q_mu; dq_sqrt [P,M,M] = g_kl * (Lq - diag(1/Lq_ii))
 * on the lower triangle, zero above.  g_kl [1] device. */
int gpfx_kl_white_bwd(int M, int P, const double* q_mu, const double* q_sqrt, const double* g_kl,
                      double* dq_mu, double* dq_sqrt, void* stream);

/* f-1: Gaussian likelihood (gpflux/layers/likelihood_layer.py:84-93 ->
 * gpflow.likelihoods.Gaussian.variational_expectations), summed over the n = N*Q entries.
 * If scale (device [1]) != NULL also writes dFmu, dFvar [n] and dnoise [1] = scale * d(sum VE)/d. */
size_t gpfx_gaussian_ve_scratch_bytes(long long n);
int gpfx_gaussian_ve(const double* Fmu, const double* Fvar, const double* Y,
                     const double* noise_variance, long long n, double* ve_sum,
                     const double* scale, double* dFmu, double* dFvar, double* dnoise,
                     void* scratch, void* stream);

/* f-2: LatentVariableLayer for diagonal Gaussians (gpflux/layers/latent_variable_layer.py:140-228):
 * out [N, D+W] = concat(X [N,D], mean + std * eps [N,W]) (the default Concatenate compositor, :109-113)
 * and kl_sum [1] = sum_{n,w} KL[N(mean, std^2) || N(prior_mean, prior_std^2)] (:221-228).
 * per_row != 0: mean/std are [N,W] (posterior from the encoder, training); per_row == 0: they are [W]
 * vectors and kl_sum = 0 (sampling the prior at prediction time, :206-219).
 * scratch: gpfx_latent_scratch_bytes(N, W). */
size_t gpfx_latent_scratch_bytes(long long N, int W);
int gpfx_latent_forward(long long N, int D, int W, const double* X, const double* mean,
                        const double* std, int per_row, const double* prior_mean,
                        const double* prior_std, const double* eps, double* out, double* kl_sum,
                        void* scratch, void* stream);
/* Backward: g_out [N, D+W], g_kl [1] device (or NULL) -> dX [N,D], dmean, dstd [N,W]. */
int gpfx_latent_backward(long long N, int D, int W, const double* mean, const double* std,
                         const double* prior_mean, const double* prior_std, const double* eps,
                         const double* g_out, const double* g_kl, double* dX, double* dmean,
                         double* dstd, void* stream);

/* ------------------------------------------------------------------------------------------
 * Efficient posterior sampling
 * ------------------------------------------------------------------------------------------ */

/* a11: FourierFeaturesBase.call (gpflux/layers/basis_functions/fourier_features/base.py:59-76 with
 * utils.py:37-54): features of X [N,D] for K feature sets W [K,L,D] (lengthscales [K,D], variance [K]).
 *   concat == 0: RandomFourierFeaturesCosine  out [K,N,L]  = sqrt(2 var/L)  cos((X/l) W^T + bias),
 *                bias [K,L];
 *   concat == 1: RandomFourierFeatures        out [K,N,2L] = sqrt(2 var/2L) [sin, cos]((X/l) W^T). */
int gpfx_rff_features(int K, int N, int D, int L, int concat, const double* X, const double* W,
                      const double* bias, const double* lengthscales, const double* variance,
                      double* out, void* stream);

/* float32 opt-in contraction (north star "1e-4 for an opt-in fp32 path"): the same product as gpfx_gemm
 * evaluated as a 3xTF32 product on the 5th-generation tensor cores (tcgen05.mma kind::tf32, float32
 * accumulation in tensor memory): the float64 operands are split into two TF32 terms each (hi + lo, 22
 * significand bits) and the product is hi*hi + hi*lo + lo*hi - ~5e-7 relative per product.  Plain products
 * only (alpha, beta, tri flags shorten the k-range, tril_out).  C stays float64.
 * workspace: gpfx_gemm_tf32x3_workspace_bytes(batch, M, N, K) (the split operands). */
size_t gpfx_gemm_tf32x3_workspace_bytes(int batch, int M, int N, int K);
int gpfx_gemm_tf32x3(int batch, int M, int N, int K, double alpha, const double* A, int lda,
                     long long strideA, int transA, const double* B, int ldb, long long strideB,
                     int transB, double beta, double* C, int ldc, long long strideC, int tri,
                     int tril_out, void* workspace, void* stream);

/* Contraction precision of the layer entry points, per handle (gpfx_set_current).
 *   mode 0 (default): every contraction on the FP64 tensor pipe (DMMA).
 *   mode 1 (the float32 opt-in): the large well-conditioned products of gpfx_layer_cond_forward / _backward
 *          (B_p = Lq_p^T A, dLq = tril(A Bs^T), dA = sum Lq_p Bs_p) run as 3xTF32 products on tcgen05
 *          (gpfx_gemm_tf32x3); their fused column reductions / add-in terms become separate streaming passes.
 *          The products with L^-1 as an operand (A = L^-1 Kuf, dKuf = L^-T dA) and dL, which feeds the
 *          Cholesky adjoint, stay on the FP64 tensor pipe: their error is amplified by cond(Kuu).
 *          Everything else (kernel matrices, Cholesky, KL, the elementwise kernels) is float64 as well.
 *          Results are within ~1e-5 of mode 0 (tests: 1e-4).  `workspace` (float32 split operands,
 *          caller-owned, >= gpfx_layer_tf32_workspace_bytes(desc) for every layer that will be evaluated)
 *          must stay valid until the last launch that used it has finished; products that do not fit it or
 *          that are not eligible (M, N < 128, shared operands, k-segments) keep the FP64 path.
 * Set the mode BEFORE querying the ctx / workspace sizes of a layer: the column-partial layout depends
 * on it. */
int gpfx_set_contraction(int mode, void* workspace, size_t workspace_bytes);
int gpfx_get_contraction(void);
size_t gpfx_layer_tf32_workspace_bytes(const gpfx_layer_desc* desc);

/* a11 backward: gradients of the feature map w.r.t. X [N,D] (shared by the K feature sets: summed),
 * the lengthscales [K,D] and the variance [K], given dPhi (same shape as the forward's out).  W and the
 * bias are not trainable in the reference (fourier_features/random/base.py:60-80).
 * scratch: gpfx_rff_backward_scratch_bytes(K, N, D, L). */
size_t gpfx_rff_backward_scratch_bytes(int K, int N, int D, int L);
int gpfx_rff_backward(int K, int N, int D, int L, int concat, const double* X, const double* W,
                      const double* bias, const double* lengthscales, const double* variance,
                      const double* dPhi, double* dX, double* dlengthscales, double* dvariance,
                      void* scratch, void* stream);

/* a13: the one-off set-up of S Matheron-rule draws (gpflux/sampling/sample.py:158-181 with
 * compute_A_inv_b, gpflux/math.py:42-62), for a single-output cosine feature map:
 *   w_s = sqrt(coeffs) * eps_w[s]                      prior weights            [S,L,P]
 *   u_s = q_mu + (q_sqrt eps_u[s])^T ;  u_s <- L_uu u_s when whiten != 0
 *   v_s = Kuu^-1 (u_s - Phi(Z) w_s),  Kuu = k(Z,Z) + jitter I                   [S,M,P]
 * Z [M,D]; lengthscales [D]; variance [1]; q_mu [M,P]; q_sqrt [P,M,M] (lower triangular, zeros above);
 * W [L,D]; b [L]; coeffs [L] (KernelWithFeatureDecomposition.feature_coefficients);
 * eps_w [S,L,P], eps_u [S,P,M]: the N(0,1) draws; info [1]: Cholesky status of Kuu (0 = ok).
 * scratch: gpfx_wilson_setup_scratch_bytes(S, M, D, P, L).  The outputs feed gpfx_wilson_eval. */
size_t gpfx_wilson_setup_scratch_bytes(int S, int M, int D, int P, int L);
int gpfx_wilson_setup(int kernel, int whiten, int S, int M, int D, int P, int L, double jitter,
                      const double* Z, const double* lengthscales, const double* variance,
                      const double* q_mu, const double* q_sqrt, const double* W, const double* b,
                      const double* coeffs, const double* eps_w, const double* eps_u, double* w,
                      double* v, int32_t* info, void* scratch, void* stream);

/* a12: WilsonSample.__call__ (gpflux/sampling/sample.py:184-198), fused: for S independent draws
 *   out[s] = Phi(X[s]) w[s] + K_{X[s] Z} v[s] (+ mean_add[s]),   Phi = cosine features (W [L,D], b [L]).
 * X is [S,N,D], or [N,D] shared by all draws when x_shared != 0; w [S,L,P]; v [S,M,P];
 * lengthscales [D]; variance [1]; Z [M,D]; out [S,N,P].  Neither Phi nor K_XZ is materialised. */
int gpfx_wilson_eval(int kernel, int S, int N, int D, int P, int L, int M, const double* X,
                     int x_shared, const double* W, const double* b, const double* lengthscales,
                     const double* variance, const double* Z, const double* w, const double* v,
                     const double* mean_add, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Natural-gradient step for q(u)
 * ------------------------------------------------------------------------------------------ */

/* gpflow.optimizers.NaturalGradient._natgrad_apply_gradients with the default XiNat
 * parameterisation, as called per GPLayer by NatGradModel._apply_backwards_pass
 * (gpflux/optimization/keras_natgrad.py:162-193):  given the ORDINARY gradients dq_mu [M,P] and
 * dq_sqrt [P,M,M] (lower triangular) of the loss, moves the natural parameters of
 * N(q_mu[:,p], q_sqrt_p q_sqrt_p^T) by -gamma * dLoss/d(expectation parameters) and converts back.
 * q_mu_new / q_sqrt_new may alias q_mu / q_sqrt.  info [3*P] int32 device: Cholesky status of the
 * current covariance, the new precision and the new covariance (0 = ok). */
size_t gpfx_natgrad_workspace_bytes(int M, int P);
int gpfx_natgrad_step(int M, int P, double gamma, const double* q_mu, const double* q_sqrt,
                      const double* dq_mu, const double* dq_sqrt, double* q_mu_new,
                      double* q_sqrt_new, int32_t* info, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: data-parallel gradient exchange.  The reference has no distributed code;
 * the minibatch rows shard over GPUs (one process per GPU), parameters and the M x M work are
 * replicated, and the flat FP64 gradient buffer is reduced over ranks.
 *
 *   gpfx_comm_unique_id_bytes  size of the rendezvous token (sizeof(ncclUniqueId) = 128)
 *   gpfx_comm_unique_id        rank 0 creates the token (host memory) and hands it to the other ranks
 *                              by any means the host has (MPI, torch.distributed, a file ...)
 *   gpfx_comm_init             collective: every rank calls it with the same token, its rank and the
 *                              world size, with its CUDA device current -> communicator handle
 *   gpfx_allreduce_grads       in-place sum (average != 0: mean) of flat_buf[0..count) (device, f64)
 *                              over the ranks, ordered on `stream`; CUDA-graph capturable; a no-op for
 *                              a world of one.  Call it once per step, or once per layer bucket as the
 *                              layer's backward finishes so that it overlaps the next layer's backward.
 * NCCL is bound at run time (dlopen of libnccl.so.2): GPFX_E_NCCL if it is absent or a call fails.
 * ------------------------------------------------------------------------------------------ */
typedef void* gpfx_comm_t;
int gpfx_comm_unique_id_bytes(void);
int gpfx_comm_nccl_version(void); /* e.g. 22809; 0 if NCCL could not be loaded */
int gpfx_comm_unique_id(void* id_out_host);
int gpfx_comm_init(const void* id_host, int rank, int world_size, gpfx_comm_t* comm_out);
int gpfx_comm_destroy(gpfx_comm_t comm);
int gpfx_allreduce_grads(gpfx_comm_t comm, double* flat_buf, long long count, int average,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Diagnostics: per-launch device time of the DMMA GEMM (bench.py's roofline leg).  While enabled,
 * every GEMM launched outside stream capture is bracketed by a pair of CUDA events on its stream;
 * gpfx_gemm_profile_read synchronises those events, writes up to `max` samples (oldest first) and
 * clears the log.  Off by default; costs two event records per launch while on.
 * ------------------------------------------------------------------------------------------ */
typedef struct gpfx_gemm_sample {
  int32_t M, N, K, batch, nseg, tri, transA, transB, tril_out, splits, cfg, reserved;
  double ms;
} gpfx_gemm_sample;
void gpfx_gemm_profile_enable(int on);
int gpfx_gemm_profile_read(gpfx_gemm_sample* out_host, int max);

#ifdef __cplusplus
}
#endif
#endif /* GPFX_H_ */
