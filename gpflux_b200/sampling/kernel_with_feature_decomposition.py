"""Mirror of gpflux/sampling/kernel_with_feature_decomposition.py:93-178: a kernel together with a
finite feature decomposition k(x, x') = sum_i lambda_i phi_i(x) phi_i(x')."""
from __future__ import annotations

import torch

from ..gpflow_compat import Kernel


class KernelWithFeatureDecomposition(Kernel):
    def __init__(self, kernel, feature_functions, feature_coefficients):
        """:param kernel: the exact kernel (``None`` - approximate-only kernels - is not supported on
            the B200 path: the conditional needs the closed-form stationary kernel).
        :param feature_functions: module whose call evaluates the L features, ``[N, D] -> [N, L]``.
        :param feature_coefficients: ``[L, 1]`` coefficients lambda_i."""
        super().__init__()
        if kernel is None:
            raise NotImplementedError("KernelWithFeatureDecomposition(kernel=None) is not supported")
        self._kernel = kernel
        self._feature_functions = feature_functions
        fc = torch.as_tensor(feature_coefficients, dtype=torch.float64)
        if fc.dim() != 2 or fc.shape[1] != 1:
            raise ValueError("feature_coefficients must have shape [L, 1]")
        self.register_buffer("_feature_coefficients", fc)

    @property
    def feature_functions(self):
        return self._feature_functions

    @property
    def feature_coefficients(self) -> torch.Tensor:
        return self._feature_coefficients

    @property
    def trainable_parameters(self):
        return self._kernel.trainable_parameters
