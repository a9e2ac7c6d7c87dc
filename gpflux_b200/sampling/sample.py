"""Efficient posterior sampling: mirror of gpflux/sampling/sample.py:35-200 (Matheron-rule path).

``efficient_sample(inducing_variable, kernel, q_mu, q_sqrt=..., whiten=...)`` returns a ``Sample``
- a deterministic function draw.  The one-off set-up (sample.py:158-181) is a handful of small
device calls (kernel matrix, batched Cholesky/inverse, DMMA GEMMs, RFF features); evaluating the
draw (sample.py:184-198) is ONE fused kernel (``gpfx_wilson_eval``) that never materialises
Phi(X) or K_XZ.  ``num_draws=S`` batches S independent draws (the reference draws one at a time).

The conditional-Gaussian fallback sampler (sample.py:85-134, O(N^3), grows state) is out of scope
(SURVEY §2 row 5).
"""
from __future__ import annotations

import abc
from typing import Callable, Optional, Union

import torch

from .. import ops
from ..gpflow_compat import InducingPoints, MultioutputInducingVariables, Stationary, default_jitter
from .kernel_with_feature_decomposition import KernelWithFeatureDecomposition


class Sample(abc.ABC):
    """A function draw from a GP that can be evaluated consistently at new inputs
    (reference sample.py:39-82)."""

    @abc.abstractmethod
    def __call__(self, X: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def __add__(self, other: Union["Sample", Callable[[torch.Tensor], torch.Tensor]]) -> "Sample":
        this = self.__call__

        class AddSample(Sample):
            def __call__(self, X):
                extra = other(X)
                out = this(X)
                return out if extra is None else out + extra

        return AddSample()


def _mm(A, B, **kw):
    """2-D GEMM on the FP64 tensor pipe."""
    return ops.gemm(A[None].contiguous(), B[None].contiguous(), **kw)[0]


def efficient_sample(inducing_variable, kernel, q_mu, *, q_sqrt: Optional[torch.Tensor] = None,
                     whiten: bool = False, num_draws: Optional[int] = None, eps_w=None, eps_u=None) -> Sample:
    """Dispatch on the kernel type like the reference's ``Dispatcher`` (sample.py:35,85,137)."""
    if isinstance(kernel, KernelWithFeatureDecomposition):
        return _efficient_sample_matheron_rule(inducing_variable, kernel, q_mu, q_sqrt=q_sqrt, whiten=whiten,
                                               num_draws=num_draws, eps_w=eps_w, eps_u=eps_u)
    raise NotImplementedError(
        "efficient_sample is implemented for KernelWithFeatureDecomposition kernels only "
        "(Matheron rule); the O(N^3) conditional-Gaussian fallback is out of scope"
    )


def _efficient_sample_matheron_rule(inducing_variable, kernel, q_mu, *, q_sqrt, whiten, num_draws=None,
                                    eps_w=None, eps_u=None) -> Sample:
    """Reference sample.py:137-200.  ``eps_w [S,L,P]`` / ``eps_u [S,P,M,1]`` pin the N(0,1) draws
    (parity runs); by default they come from torch.randn on the device."""
    from ..layers.basis_functions.fourier_features import RandomFourierFeaturesCosine

    if isinstance(inducing_variable, MultioutputInducingVariables):
        ivs = inducing_variable.inducing_variables
        if len(ivs) != 1:
            raise NotImplementedError("Matheron-rule sampling needs a single set of inducing points (Kmm is [M, M])")
        inducing_variable = ivs[0]
    if not isinstance(inducing_variable, InducingPoints):
        raise NotImplementedError("only InducingPoints are supported")
    base = kernel._kernel
    if not isinstance(base, Stationary):
        raise NotImplementedError("the wrapped kernel must be a stationary kernel")
    feats = kernel.feature_functions
    if not isinstance(feats, RandomFourierFeaturesCosine) or feats.is_multioutput:
        raise NotImplementedError("feature_functions must be a single-output RandomFourierFeaturesCosine")

    squeeze = num_draws is None
    S = 1 if squeeze else int(num_draws)
    q_mu = q_mu.detach()
    q_sqrt = q_sqrt.detach()
    dev = q_mu.device
    M, P = q_mu.shape
    Z = inducing_variable.Z.value().detach().to(dev)
    D = Z.shape[1]
    if feats.W is None:
        feats.build((D,))
    W = feats.W.to(dev)
    b = feats.b.to(dev).reshape(-1)
    L = W.shape[0]
    ls = base.lengthscales_vector(D).detach().to(dev)
    var = base.variance.value().detach().reshape(1).to(dev)
    coeffs = kernel.feature_coefficients.to(dev)  # [L, 1]

    if eps_w is None:
        eps_w = torch.randn((S, L, P), dtype=torch.float64, device=dev)
    if eps_u is None:
        eps_u = torch.randn((S, P, M, 1), dtype=torch.float64, device=dev)
    prior_weights = torch.sqrt(coeffs)[None] * eps_w  # [S, L, P]

    Kmm = ops.kernel_matrix(base.kind, Z[None], Z[None], ls[None], var, jitter=default_jitter())  # [1, M, M]
    Luu, Linv, info = ops.cholesky_inverse(Kmm)
    phi_Z = ops.rff_features(Z, W[None], b[None], ls[None], var)[0]  # [M, L]

    vs = []
    for s in range(S):
        u_noise = ops.gemm(q_sqrt.contiguous(), eps_u[s].contiguous())  # [P, M, 1]
        u = q_mu + u_noise[..., 0].T  # [M, P]
        if whiten:
            u = _mm(Luu[0], u.contiguous())
        diff = u - _mm(phi_Z, prior_weights[s])  # [M, P]
        t = _mm(Linv[0], diff.contiguous())  # L^-1 diff   (compute_A_inv_b, reference math.py:42-62)
        vs.append(_mm(Linv[0], t.contiguous(), transA=True))  # L^-T L^-1 diff = Kmm^-1 diff
    v = torch.stack(vs)  # [S, M, P]

    class WilsonSample(Sample):
        """f(X) = Phi(X) w + K_XZ v, evaluated by one fused kernel."""

        num_draws = S

        def __call__(self, X: torch.Tensor) -> torch.Tensor:
            if X.dim() == 3 and squeeze:
                raise ValueError("a single draw expects X of shape [N, D]")
            out = ops.wilson_eval(base.kind, X, W, b, ls, var, Z, prior_weights, v)
            return out[0] if squeeze else out

    return WilsonSample()
