// GPLayer forward / backward orchestration (layer.cu).
#pragma once
#include "../../include/gpfx.h"
#include "common.cuh"

namespace gpfx {

size_t layer_ctx_bytes(const gpfx_layer_desc* d);
size_t layer_workspace_bytes(const gpfx_layer_desc* d);
size_t layer_factor_ctx_bytes(const gpfx_layer_desc* d);
size_t layer_cond_ctx_bytes(const gpfx_layer_desc* d);
size_t layer_factor_workspace_bytes(const gpfx_layer_desc* d);
size_t layer_cond_workspace_bytes(const gpfx_layer_desc* d);
size_t layer_dL_bytes(const gpfx_layer_desc* d);

int layer_factor_forward(const gpfx_layer_desc* d, const double* Z, const double* ls,
                         const double* var, const double* q_mu, const double* q_sqrt, double* kl,
                         void* fctx, void* fws, cudaStream_t stream);
int layer_factor_backward(const gpfx_layer_desc* d, const double* Z, const double* ls,
                          const double* var, const double* q_mu, const double* g_kl,
                          const double* dL, double* dZ, double* dls, double* dvar, double* dq_mu,
                          double* dq_sqrt, int accumulate, const void* fctx, void* fws,
                          cudaStream_t stream);
int layer_cond_forward(const gpfx_layer_desc* d, const double* X, const double* Z,
                       const double* ls, const double* var, const double* q_mu, const void* fctx,
                       const double* mean_add, const double* eps, double* fmean, double* fvar,
                       double* fsample, void* cctx, void* cws, cudaStream_t stream);
int layer_cond_backward(const gpfx_layer_desc* d, const double* X, const double* Z,
                        const double* ls, const double* var, const double* q_mu,
                        const double* eps, const double* fvar, const double* g_f,
                        const double* g_mean, const double* g_var, double* dX, double* dZ,
                        double* dls, double* dvar, double* dq_mu, double* dq_sqrt, double* dL,
                        double* dmean, const void* fctx, void* cctx, void* cws,
                        cudaStream_t stream);

int layer_forward(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                  const double* var, const double* q_mu, const double* q_sqrt,
                  const double* mean_add, const double* eps, double* fmean, double* fvar,
                  double* fsample, double* kl, void* ctx, void* ws, cudaStream_t stream);
int layer_backward(const gpfx_layer_desc* d, const double* X, const double* Z, const double* ls,
                   const double* var, const double* q_mu, const double* eps, const double* fvar,
                   const double* g_f, const double* g_mean, const double* g_var,
                   const double* g_kl, double* dX, double* dZ, double* dls, double* dvar,
                   double* dq_mu, double* dq_sqrt, double* dmean, void* ctx, void* ws,
                   cudaStream_t stream);
// float32 split-operand workspace of the 3xTF32 contraction mode (largest product of the layer)
size_t layer_tf32_workspace_bytes(const gpfx_layer_desc* d);
int layer_status(const gpfx_layer_desc* d, const void* fctx, int32_t* info_host,
                 cudaStream_t stream);
int layer_status_async(const gpfx_layer_desc* d, const void* fctx, int32_t* info_dev,
                       cudaStream_t stream);

}  // namespace gpfx
