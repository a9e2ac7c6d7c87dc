// TensorFlow custom-op shim over the C ABI of include/gpfx.h (SURVEY §8b "Host bindings (ii)").
//
// WRITTEN BUT NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT: TensorFlow is not installable here
// (no network, not in the wheelhouse), so this file is built only when `tensorflow.sysconfig` is
// importable (gpflux_b200/tf_shim/build.py).  It is the binding a GPflux maintainer would add so
// that gpflux.layers.GPLayer (gpflux/layers/gp_layer.py:226-368) reaches the sm_100a kernels with
// zero copies: TF GPU tensors already are raw device memory, the op passes their pointers and
// TF's own compute stream straight through.
//
//   GpfxLayerForward  : GPLayer.predict + prior_kl + reparameterised sample  -> gpfx_layer_forward
//   GpfxLayerBackward : the GradientTape transpose of the above                -> gpfx_layer_backward
//
// Python side (registration of the gradient, DLPack helper): gpfx_tf_ops.py next to this file.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "../../include/gpfx.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;

namespace {

// [N,D], [K,M,D], [K,D], [K], [M,P], [P,M,M] -> desc
tf::Status make_desc(const tf::Tensor& x, const tf::Tensor& z, const tf::Tensor& q_mu, int kernel,
                     int whiten, double jitter, gpfx_layer_desc* d) {
  if (x.dims() != 2 || z.dims() != 3 || q_mu.dims() != 2 || x.dim_size(1) != z.dim_size(2) ||
      q_mu.dim_size(0) != z.dim_size(1))
    return tf::errors::InvalidArgument("GpfxLayer: expected x [N,D], z [K,M,D], q_mu [M,P]");
  d->N = static_cast<int32_t>(x.dim_size(0));
  d->D = static_cast<int32_t>(x.dim_size(1));
  d->K = static_cast<int32_t>(z.dim_size(0));
  d->M = static_cast<int32_t>(z.dim_size(1));
  d->P = static_cast<int32_t>(q_mu.dim_size(1));
  d->kernel = kernel;
  d->whiten = whiten;
  d->solver = GPFX_SOLVER_INVERSE_GEMM;
  d->jitter = jitter;
  if (d->K != 1 && d->K != d->P)
    return tf::errors::InvalidArgument("GpfxLayer: K must be 1 (SharedIndependent) or P (SeparateIndependent)");
  return tf::OkStatus();
}

const double* ptr_or_null(const tf::Tensor& t) {
  return t.NumElements() == 0 ? nullptr : t.flat<double>().data();
}

}  // namespace

REGISTER_OP("GpfxLayerForward")
    .Input("x: double")             // [N,D]
    .Input("z: double")             // [K,M,D]
    .Input("lengthscales: double")  // [K,D]
    .Input("variance: double")      // [K]
    .Input("q_mu: double")          // [M,P]
    .Input("q_sqrt: double")        // [P,M,M] (lower triangle read)
    .Input("mean_add: double")      // [N,P] or empty: mean_function(x)
    .Input("eps: double")           // [N,P] or empty: reparameterisation noise
    .Attr("kernel: int = 0")
    .Attr("whiten: bool = true")
    .Attr("jitter: float = 1e-6")
    .Output("fmean: double")
    .Output("fvar: double")
    .Output("fsample: double")
    .Output("kl: double")
    .Output("ctx: uint8")  // saved-for-backward buffer (Kuf, A, B, L, L^-1, tril(q_sqrt))
    .SetShapeFn([](InferenceContext* c) {
      auto n = c->Dim(c->input(0), 0);
      auto p = c->Dim(c->input(4), 1);
      for (int i = 0; i < 3; ++i) c->set_output(i, c->Matrix(n, p));
      c->set_output(3, c->Scalar());
      c->set_output(4, c->Vector(InferenceContext::kUnknownDim));
      return tf::OkStatus();
    });

class GpfxLayerForwardOp : public tf::OpKernel {
 public:
  explicit GpfxLayerForwardOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    float j;
    OP_REQUIRES_OK(c, c->GetAttr("kernel", &kernel_));
    OP_REQUIRES_OK(c, c->GetAttr("whiten", &whiten_));
    OP_REQUIRES_OK(c, c->GetAttr("jitter", &j));
    jitter_ = j;
  }
  void Compute(tf::OpKernelContext* c) override {
    const tf::Tensor &x = c->input(0), &z = c->input(1), &ls = c->input(2), &var = c->input(3),
                     &q_mu = c->input(4), &q_sqrt = c->input(5), &mean_add = c->input(6),
                     &eps = c->input(7);
    gpfx_layer_desc d;
    OP_REQUIRES_OK(c, make_desc(x, z, q_mu, kernel_, whiten_ ? 1 : 0, jitter_, &d));
    tf::Tensor *fmean, *fvar, *fsample, *kl, *ctx, ws;
    const tf::TensorShape np({d.N, d.P});
    OP_REQUIRES_OK(c, c->allocate_output(0, np, &fmean));
    OP_REQUIRES_OK(c, c->allocate_output(1, np, &fvar));
    OP_REQUIRES_OK(c, c->allocate_output(2, np, &fsample));
    OP_REQUIRES_OK(c, c->allocate_output(3, tf::TensorShape({}), &kl));
    OP_REQUIRES_OK(c, c->allocate_output(
                          4, tf::TensorShape({static_cast<tf::int64>(gpfx_layer_ctx_bytes(&d))}), &ctx));
    OP_REQUIRES_OK(c, c->allocate_temp(
                          tf::DT_UINT8,
                          tf::TensorShape({static_cast<tf::int64>(gpfx_layer_workspace_bytes(&d))}), &ws));
    auto stream = c->eigen_device<Eigen::GpuDevice>().stream();  // TF's compute stream: no extra sync
    const int st = gpfx_layer_forward(
        &d, x.flat<double>().data(), z.flat<double>().data(), ls.flat<double>().data(),
        var.flat<double>().data(), q_mu.flat<double>().data(), q_sqrt.flat<double>().data(),
        ptr_or_null(mean_add), ptr_or_null(eps), fmean->flat<double>().data(),
        fvar->flat<double>().data(), fsample->flat<double>().data(), kl->flat<double>().data(),
        ctx->flat<tf::uint8>().data(), ws.flat<tf::uint8>().data(), stream);
    OP_REQUIRES(c, st == GPFX_OK, tf::errors::Internal("gpfx_layer_forward: ", gpfx_last_error()));
  }

 private:
  int kernel_;
  bool whiten_;
  double jitter_;
};
REGISTER_KERNEL_BUILDER(Name("GpfxLayerForward").Device(tf::DEVICE_GPU), GpfxLayerForwardOp);

REGISTER_OP("GpfxLayerBackward")
    .Input("x: double").Input("z: double").Input("lengthscales: double").Input("variance: double")
    .Input("q_mu: double").Input("eps: double").Input("fvar: double")
    .Input("g_fsample: double").Input("g_fmean: double").Input("g_fvar: double").Input("g_kl: double")
    .Input("ctx: uint8")
    .Attr("kernel: int = 0")
    .Attr("whiten: bool = true")
    .Attr("jitter: float = 1e-6")
    .Output("dx: double").Output("dz: double").Output("dlengthscales: double")
    .Output("dvariance: double").Output("dq_mu: double").Output("dq_sqrt: double")
    .Output("dmean_add: double")
    .SetShapeFn([](InferenceContext* c) {
      for (int i = 0; i < 5; ++i) c->set_output(i, c->input(i));
      auto m = c->Dim(c->input(4), 0);
      auto p = c->Dim(c->input(4), 1);
      c->set_output(5, c->MakeShape({p, m, m}));
      c->set_output(6, c->input(8));
      return tf::OkStatus();
    });

class GpfxLayerBackwardOp : public tf::OpKernel {
 public:
  explicit GpfxLayerBackwardOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    float j;
    OP_REQUIRES_OK(c, c->GetAttr("kernel", &kernel_));
    OP_REQUIRES_OK(c, c->GetAttr("whiten", &whiten_));
    OP_REQUIRES_OK(c, c->GetAttr("jitter", &j));
    jitter_ = j;
  }
  void Compute(tf::OpKernelContext* c) override {
    const tf::Tensor &x = c->input(0), &z = c->input(1), &ls = c->input(2), &var = c->input(3),
                     &q_mu = c->input(4), &eps = c->input(5), &fvar = c->input(6),
                     &g_f = c->input(7), &g_mean = c->input(8), &g_var = c->input(9),
                     &g_kl = c->input(10), &ctx = c->input(11);
    gpfx_layer_desc d;
    OP_REQUIRES_OK(c, make_desc(x, z, q_mu, kernel_, whiten_ ? 1 : 0, jitter_, &d));
    tf::Tensor *dx, *dz, *dls, *dvar, *dq_mu, *dq_sqrt, *dmean, ws;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(c, c->allocate_output(1, z.shape(), &dz));
    OP_REQUIRES_OK(c, c->allocate_output(2, ls.shape(), &dls));
    OP_REQUIRES_OK(c, c->allocate_output(3, var.shape(), &dvar));
    OP_REQUIRES_OK(c, c->allocate_output(4, q_mu.shape(), &dq_mu));
    OP_REQUIRES_OK(c, c->allocate_output(5, tf::TensorShape({d.P, d.M, d.M}), &dq_sqrt));
    OP_REQUIRES_OK(c, c->allocate_output(6, tf::TensorShape({d.N, d.P}), &dmean));
    OP_REQUIRES_OK(c, c->allocate_temp(
                          tf::DT_UINT8,
                          tf::TensorShape({static_cast<tf::int64>(gpfx_layer_workspace_bytes(&d))}), &ws));
    auto stream = c->eigen_device<Eigen::GpuDevice>().stream();
    // the saved context is consumed in place (dK scratch); TF hands the forward op's output buffer
    void* ctx_ptr = const_cast<tf::uint8*>(ctx.flat<tf::uint8>().data());
    const int st = gpfx_layer_backward(
        &d, x.flat<double>().data(), z.flat<double>().data(), ls.flat<double>().data(),
        var.flat<double>().data(), q_mu.flat<double>().data(), ptr_or_null(eps),
        fvar.flat<double>().data(), ptr_or_null(g_f), ptr_or_null(g_mean), ptr_or_null(g_var),
        g_kl.flat<double>().data(), dx->flat<double>().data(), dz->flat<double>().data(),
        dls->flat<double>().data(), dvar->flat<double>().data(), dq_mu->flat<double>().data(),
        dq_sqrt->flat<double>().data(), dmean->flat<double>().data(), ctx_ptr,
        ws.flat<tf::uint8>().data(), stream);
    OP_REQUIRES(c, st == GPFX_OK, tf::errors::Internal("gpfx_layer_backward: ", gpfx_last_error()));
  }

 private:
  int kernel_;
  bool whiten_;
  double jitter_;
};
REGISTER_KERNEL_BUILDER(Name("GpfxLayerBackward").Device(tf::DEVICE_GPU), GpfxLayerBackwardOp);
