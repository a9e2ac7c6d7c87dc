"""Random Fourier feature layers: mirror of
gpflux/layers/basis_functions/fourier_features/{base.py:30-127, random/base.py:40-295, utils.py}.

``call`` (the GEMM + sin/cos epilogue) runs on the device through
``gpfx_rff_features``; weight / phase initialisation is one-off host work (torch RNG), as in the
reference (``_weights_init``: N(0,1) for SquaredExponential, Student-t(2p+1) for Matern-(p+1/2);
``_bias_init``: U(0, 2 pi))."""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .... import ops
from ....gpflow_compat import (
    Matern12,
    Matern32,
    Matern52,
    MultioutputKernel,
    SeparateIndependent,
    SharedIndependent,
    SquaredExponential,
    default_float,
)

RFF_SUPPORTED_KERNELS = (SquaredExponential, Matern12, Matern32, Matern52)
RFF_SUPPORTED_MULTIOUTPUTS = (SeparateIndependent, SharedIndependent)


def _matern_number(kernel) -> int:
    """Reference utils.py:24-34."""
    if isinstance(kernel, Matern52):
        return 2
    if isinstance(kernel, Matern32):
        return 1
    if isinstance(kernel, Matern12):
        return 0
    raise NotImplementedError("Not a recognized Matern kernel")


def _sample_students_t(nu: float, shape, generator=None) -> torch.Tensor:
    """Reference random/base.py:51-77: X = N(0,1) / sqrt(Gamma(nu/2, rate nu/2))."""
    normal = torch.randn(shape, dtype=default_float(), generator=generator)
    gshape = tuple(shape[:-1]) + (1,)
    gamma = torch.distributions.Gamma(torch.tensor(0.5 * nu, dtype=default_float()),
                                      torch.tensor(0.5 * nu, dtype=default_float())).sample(gshape)
    return normal * torch.rsqrt(gamma)


class RandomFourierFeaturesBase(nn.Module):
    def __init__(self, kernel, n_components: int, input_dim: Optional[int] = None, **kwargs):
        assert isinstance(kernel, RFF_SUPPORTED_KERNELS + RFF_SUPPORTED_MULTIOUTPUTS), (
            f"Unsupported Kernel: only the following kernel types are supported: "
            f"{[k.__name__ for k in RFF_SUPPORTED_MULTIOUTPUTS + RFF_SUPPORTED_KERNELS]}"
        )
        if isinstance(kernel, RFF_SUPPORTED_MULTIOUTPUTS):
            for k in kernel.latent_kernels:
                assert isinstance(k, RFF_SUPPORTED_KERNELS), (
                    f"Unsupported Kernel within the multioutput kernel; only the following"
                    f"kernel types are supported: {[k.__name__ for k in RFF_SUPPORTED_KERNELS]}"
                )
        super().__init__()
        self.kernel = kernel
        self.n_components = n_components
        self.is_multioutput = isinstance(kernel, MultioutputKernel)
        self.num_latent_gps = kernel.num_latent_gps if self.is_multioutput else 1
        self._input_dim = input_dim
        self.register_buffer("W", None)
        self.register_buffer("b", None)
        if input_dim:
            self.build((input_dim,))

    # ---- one-off host-side initialisation
    def _weights_init_individual(self, kernel, shape) -> torch.Tensor:
        if isinstance(kernel, SquaredExponential):
            return torch.randn(shape, dtype=default_float())
        p = _matern_number(kernel)
        return _sample_students_t(2.0 * p + 1.0, shape)

    def build(self, input_shape) -> None:
        input_dim = int(input_shape[-1])
        self._input_dim = input_dim
        shape = (self.n_components, input_dim)
        if self.is_multioutput:
            if isinstance(self.kernel, SharedIndependent):
                ws = [self._weights_init_individual(self.kernel.latent_kernels[0], shape) for _ in range(self.num_latent_gps)]
            else:
                ws = [self._weights_init_individual(k, shape) for k in self.kernel.latent_kernels]
            self.W = torch.stack(ws, 0)  # [P, M, D]
        else:
            self.W = self._weights_init_individual(self.kernel, shape)  # [M, D]

    @staticmethod
    def rff_constant(variance, output_dim: int):
        """Reference random/base.py:147-152."""
        return torch.sqrt(2.0 * variance / output_dim)

    def _latent(self):
        if not self.is_multioutput:
            return [self.kernel]
        ks = list(self.kernel.latent_kernels)
        return ks * self.num_latent_gps if len(ks) == 1 else ks

    def _packed(self, device):
        ks = self._latent()
        D = self._input_dim
        ls = torch.stack([k.lengthscales_vector(D) for k in ks]).to(device)
        var = torch.stack([k.variance.value().reshape(()) for k in ks]).to(device)
        W = self.W if self.is_multioutput else self.W[None]
        return W.to(device), ls.detach(), var.detach()

    def compute_output_shape(self, input_shape):
        out = (int(input_shape[0]), self._compute_output_dim(input_shape))
        return (self.num_latent_gps,) + out if self.is_multioutput else out

    def call(self, inputs: torch.Tensor) -> torch.Tensor:
        """[N, D] -> [N, L'] or [P, N, L'] (multi-output) - reference base.py:59-76."""
        if self.W is None:
            self.build(inputs.shape)
        W, ls, var = self._packed(inputs.device)
        out = self._features(inputs, W, ls, var)
        return out if self.is_multioutput else out[0]

    forward = call


class RandomFourierFeatures(RandomFourierFeaturesBase):
    """[sin, cos] features, 2 n_components outputs (reference random/base.py:155-213)."""

    def _compute_output_dim(self, input_shape) -> int:
        return 2 * self.n_components

    def _features(self, X, W, ls, var):
        return ops.rff_features(X, W, None, ls, var, concat=True)


class RandomFourierFeaturesCosine(RandomFourierFeaturesBase):
    """Phase-shifted cosine features, n_components outputs (reference random/base.py:216-295)."""

    def build(self, input_shape) -> None:
        shape = (self.num_latent_gps, 1, self.n_components) if self.is_multioutput else (1, self.n_components)
        self.b = 2.0 * math.pi * torch.rand(shape, dtype=default_float())
        super().build(input_shape)

    def _compute_output_dim(self, input_shape) -> int:
        return self.n_components

    def _features(self, X, W, ls, var):
        b = self.b.to(X.device).reshape(W.shape[0], self.n_components)
        return ops.rff_features(X, W, b, ls, var, concat=False)


class OrthogonalRandomFeatures(RandomFourierFeatures):
    """Orthogonal random features (reference random/orthogonal.py:59-87): the weight matrix is a
    stack of Haar-orthogonal D x D blocks with chi(D)-distributed row norms.  SquaredExponential only.
    Initialisation is one-off host work; ``call`` runs on the device like RandomFourierFeatures."""

    def __init__(self, kernel, n_components: int, **kwargs):
        assert isinstance(kernel, SquaredExponential), "Unsupported Kernel"
        super().__init__(kernel, n_components, **kwargs)

    def _weights_init_individual(self, kernel, shape) -> torch.Tensor:
        n_components, input_dim = shape
        n_reps = -(-n_components // input_dim)  # smallest K with K * D >= M
        Wn = torch.randn((n_reps, input_dim, input_dim), dtype=default_float())
        Q, _ = torch.linalg.qr(Wn)
        chi2 = torch.distributions.Gamma(torch.tensor(0.5 * input_dim, dtype=default_float()),
                                         torch.tensor(0.5, dtype=default_float())).sample((n_reps, input_dim))
        U = torch.sqrt(chi2)[..., None] * Q  # diag(s) @ Q
        return U.reshape(-1, input_dim)[:n_components]


class QuadratureFourierFeatures(nn.Module):
    """Gauss-Hermite quadrature Fourier features (reference quadrature/gaussian.py:44-92):
    ``[sin, cos](X/l @ abscissa^T) * sqrt(variance * factors)`` with the ``n^D`` tensor-product GH
    nodes GPflow's ``ndgh_points_and_weights`` returns (nodes * sqrt(2), weights / sqrt(pi)).
    SquaredExponential only.  The sin/cos GEMM runs on the device (``gpfx_rff_features``); the
    per-feature quadrature weights are applied on the result (this class is not on the north-star
    path: n^D features only make sense for very small D)."""

    def __init__(self, kernel, n_components: int, input_dim: Optional[int] = None, **kwargs):
        assert isinstance(kernel, SquaredExponential), "Unsupported Kernel"
        super().__init__()
        self.kernel = kernel
        self.n_components = n_components
        self.is_multioutput = False
        self._input_dim = input_dim
        self.register_buffer("abscissa", None)
        self.register_buffer("factors", None)
        if input_dim:
            self.build((input_dim,))

    def build(self, input_shape) -> None:
        import itertools

        import numpy as np

        D = int(input_shape[-1])
        self._input_dim = D
        x, w = np.polynomial.hermite.hermgauss(self.n_components)
        x, w = x * np.sqrt(2.0), w / np.sqrt(np.pi)
        nodes = np.array(list(itertools.product(*([x] * D))))  # [n^D, D]
        weights = np.prod(np.array(list(itertools.product(*([w] * D)))), axis=1)  # [n^D]
        self.abscissa = torch.as_tensor(nodes, dtype=default_float())
        self.factors = torch.as_tensor(weights, dtype=default_float())

    def compute_output_shape(self, input_shape):
        return (int(input_shape[0]), 2 * self.n_components ** int(input_shape[-1]))

    def call(self, inputs: torch.Tensor) -> torch.Tensor:
        if self.abscissa is None:
            self.build(inputs.shape)
        D = self._input_dim
        L = self.abscissa.shape[0]
        ls = self.kernel.lengthscales_vector(D).detach().to(inputs.device)[None]
        # ask the kernel for unit-constant features: sqrt(2 * var / 2L) = 1  <=>  var = L
        unit = torch.full((1,), float(L), dtype=default_float(), device=inputs.device)
        bases = ops.rff_features(inputs, self.abscissa.to(inputs.device)[None], None, ls, unit, concat=True)[0]
        const = torch.sqrt(self.kernel.variance.value().detach().to(inputs.device) * self.factors.to(inputs.device))
        return bases * const.repeat(2)

    forward = call
