"""Mirror of gpflux/architectures/constant_input_dim_deep_gp.py:42-168 (host-only builder)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
from scipy.cluster.vq import kmeans2

from ..gpflow_compat import Gaussian, SquaredExponential, Zero
from ..helpers import construct_basic_inducing_variables, construct_basic_kernel, construct_mean_function
from ..layers import GPLayer, LikelihoodLayer
from ..models import DeepGP


@dataclass
class Config:
    """Reference constant_input_dim_deep_gp.py:42-72."""

    num_inducing: int
    inner_layer_qsqrt_factor: float
    likelihood_noise_variance: float
    whiten: bool = True


def _construct_kernel(input_dim: int, is_last_layer: bool) -> SquaredExponential:
    """ARD lengthscales 2, variance 1e-6 (hidden) / 1.0 (last) - reference :75-90."""
    variance = 1e-6 if not is_last_layer else 1.0
    return SquaredExponential(lengthscales=[2.0] * input_dim, variance=variance)


def build_constant_input_dim_deep_gp(X: np.ndarray, num_layers: int, config: Config, z_init=None) -> DeepGP:
    """Reference :93-168.  ``z_init`` (optional, [M, D]) skips the k-means initialisation (the
    benchmark uses the first M rows so that start-up is not dominated by scipy)."""
    if X.dtype != np.float64:
        raise ValueError(
            f"X needs to have dtype according to gpflow.default_float() = float64 however got X with {X.dtype} dtype."
        )
    num_data, input_dim = X.shape
    X_running = X
    gp_layers = []
    if z_init is None:
        centroids, _ = kmeans2(X, k=config.num_inducing, minit="points")
    else:
        centroids = np.asarray(z_init, dtype=np.float64)
    for i_layer in range(num_layers):
        is_last_layer = i_layer == num_layers - 1
        D_in = input_dim
        D_out = 1 if is_last_layer else input_dim
        inducing_var = construct_basic_inducing_variables(
            num_inducing=config.num_inducing, input_dim=D_in, share_variables=True, z_init=centroids
        )
        kernel = construct_basic_kernel(kernels=_construct_kernel(D_in, is_last_layer), output_dim=D_out,
                                        share_hyperparams=True)
        assert config.whiten is True, "non-whitened case not implemented yet"
        if is_last_layer:
            mean_function = Zero()
            q_sqrt_scaling = 1.0
        else:
            mean_function = construct_mean_function(X_running, D_in, D_out)
            with torch.no_grad():
                out = mean_function(torch.as_tensor(X_running))
            X_running = out.numpy() if out is not None else X_running
            q_sqrt_scaling = config.inner_layer_qsqrt_factor
        layer = GPLayer(kernel, inducing_var, num_data, mean_function=mean_function, name=f"gp_{i_layer}")
        layer.q_sqrt.assign(layer.q_sqrt.value() * q_sqrt_scaling)
        gp_layers.append(layer)
    likelihood = Gaussian(config.likelihood_noise_variance)
    return DeepGP(gp_layers, LikelihoodLayer(likelihood))
