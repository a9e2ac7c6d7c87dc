"""Python half of the TensorFlow shim (SURVEY §8b "Host bindings (ii)"): loads ``_gpfx_tf_ops.so``,
registers the gradient of ``GpfxLayerForward`` and shows the two call sites that change in
``gpflux/layers/gp_layer.py``.  NOT importable here (TensorFlow is absent); kept as the source a
GPflux maintainer would drop into ``gpflux/ops/``.  ``tests/test_cpu_host.py`` only checks that the
C++ file references entry points that ``include/gpfx.h`` really declares."""
from __future__ import annotations

import os

KERNEL_ID = {"SquaredExponential": 0, "Matern12": 1, "Matern32": 2, "Matern52": 3}


def load():
    import tensorflow as tf  # deferred: this module must stay importable without TensorFlow

    lib = tf.load_op_library(os.path.join(os.path.dirname(__file__), "_gpfx_tf_ops.so"))

    @tf.RegisterGradient("GpfxLayerForward")
    def _gpfx_layer_grad(op, g_fmean, g_fvar, g_fsample, g_kl, _g_ctx):
        x, z, ls, var, q_mu, _q_sqrt, _mean_add, eps = op.inputs
        _fmean, fvar, _fsample, _kl, ctx = op.outputs
        dx, dz, dls, dvar, dq_mu, dq_sqrt, dmean = lib.gpfx_layer_backward(
            x, z, ls, var, q_mu, eps, fvar, g_fsample, g_fmean, g_fvar, g_kl, ctx,
            kernel=op.get_attr("kernel"), whiten=op.get_attr("whiten"), jitter=op.get_attr("jitter"))
        return dx, dz, dls, dvar, dq_mu, dq_sqrt, dmean, None  # eps is not differentiated

    return lib


def from_dlpack_zero_copy(capsule):
    """Tensors produced by another framework enter TF without a copy (tf.experimental.dlpack);
    TF-native tensors need nothing: the op reads ``tensor.flat<double>().data()``."""
    import tensorflow as tf

    return tf.experimental.dlpack.from_dlpack(capsule)


# gpflux/layers/gp_layer.py, the two call sites that change:
#
#   def predict(self, inputs, *, full_cov=False, full_output_cov=False):      # gp_layer.py:226-268
#       fmean, fvar, _, _, _ = ops.gpfx_layer_forward(
#           inputs, Z, lengthscales, variance, self.q_mu, self.q_sqrt, self.mean_function(inputs),
#           tf.zeros([0, 0], inputs.dtype), kernel=KERNEL_ID[type(base_kernel).__name__], whiten=self.whiten)
#       return fmean, fvar
#
#   def call(self, inputs, *args, **kwargs):                                   # gp_layer.py:270-301
#       eps = tf.random.normal([N, P], dtype=inputs.dtype)
#       fmean, fvar, fsample, kl, _ = ops.gpfx_layer_forward(..., eps, ...)
#       # DistributionLambda's convert_to_tensor_fn returns `fsample`; prior_kl() returns `kl`
