"""Builder helpers mirroring gpflux/helpers.py:42-293 (host-only; no GPU work)."""
from __future__ import annotations

import warnings
from copy import deepcopy
from typing import List, Optional, Type, Union

import numpy as np

from .gpflow_compat import (
    Identity,
    InducingPoints,
    Kernel,
    Linear,
    MeanFunction,
    SeparateIndependent,
    SeparateIndependentInducingVariables,
    SharedIndependent,
    SharedIndependentInducingVariables,
    SquaredExponential,
    Zero,
    set_trainable,
)
from .layers.gp_layer import GPLayer


def construct_basic_kernel(kernels: Union[Kernel, List[Kernel]], output_dim: Optional[int] = None,
                           share_hyperparams: bool = False):
    """Reference helpers.py:42-73."""
    if isinstance(kernels, list):
        return SeparateIndependent(kernels)
    if not share_hyperparams:
        return SeparateIndependent([deepcopy(kernels) for _ in range(output_dim)])
    return SharedIndependent(kernels, output_dim)


def construct_basic_inducing_variables(num_inducing: Union[int, List[int]], input_dim: int,
                                       output_dim: Optional[int] = None, share_variables: bool = False,
                                       z_init: Optional[np.ndarray] = None):
    """Reference helpers.py:76-171."""
    if z_init is None:
        warnings.warn(
            "No `z_init` has been specified in `construct_basic_inducing_variables`. "
            "Default initialization using random normal N(0, 1) will be used."
        )
    z_init_is_given = z_init is not None
    if isinstance(num_inducing, list):
        if output_dim is not None:
            assert output_dim == len(num_inducing)
        assert share_variables is False
        ivs = []
        for i, m in enumerate(num_inducing):
            if z_init_is_given:
                assert len(z_init[i]) == m
                z_i = z_init[i]
            else:
                z_i = np.random.randn(m, input_dim)
            assert z_i.shape == (m, input_dim)
            ivs.append(InducingPoints(z_i))
        return SeparateIndependentInducingVariables(ivs)
    if not share_variables:
        ivs = []
        for o in range(output_dim):
            if z_init_is_given:
                if z_init.shape != (output_dim, num_inducing, input_dim):
                    raise ValueError(
                        "When not sharing variables, z_init must have shape"
                        "[output_dim, num_inducing, input_dim]"
                    )
                z_o = z_init[o]
            else:
                z_o = np.random.randn(num_inducing, input_dim)
            ivs.append(InducingPoints(z_o))
        return SeparateIndependentInducingVariables(ivs)
    z_init = z_init if z_init_is_given else np.random.randn(num_inducing, input_dim)
    return SharedIndependentInducingVariables(InducingPoints(z_init))


def construct_mean_function(X: np.ndarray, D_in: int, D_out: int) -> MeanFunction:
    """Identity when D_in == D_out, else a fixed PCA / zero-padding Linear (reference
    helpers.py:174-208)."""
    assert X.shape[-1] == D_in
    if D_in == D_out:
        return Identity()
    if D_in > D_out:
        _, _, V = np.linalg.svd(X, full_matrices=False)
        W = V[:D_out, :].T
    else:
        W = np.concatenate([np.eye(D_in), np.zeros((D_in, D_out - D_in))], axis=1)
    assert W.shape == (D_in, D_out)
    mean_function = Linear(W, np.zeros(D_out))
    set_trainable(mean_function, False)
    return mean_function


def construct_gp_layer(num_data: int, num_inducing: int, input_dim: int, output_dim: int,
                       kernel_class: Type = SquaredExponential, z_init: Optional[np.ndarray] = None,
                       name: Optional[str] = None) -> GPLayer:
    """Vanilla GP layer: shared kernel, shared inducing points, zero mean (reference
    helpers.py:211-258)."""
    lengthscale = float(input_dim) ** 0.5
    base_kernel = kernel_class(lengthscales=np.full(input_dim, lengthscale))
    kernel = construct_basic_kernel(base_kernel, output_dim=output_dim, share_hyperparams=True)
    inducing_variable = construct_basic_inducing_variables(
        num_inducing, input_dim, output_dim=output_dim, share_variables=True, z_init=z_init
    )
    return GPLayer(kernel=kernel, inducing_variable=inducing_variable, num_data=num_data,
                   mean_function=Zero(), name=name)


def xavier_initialization_numpy(input_dim: int, output_dim: int) -> np.ndarray:
    """Reference helpers.py:280-293."""
    return np.random.randn(input_dim, output_dim) * (2.0 / (input_dim + output_dim)) ** 0.5
