"""Efficient posterior sampling: mirror of gpflux/sampling/sample.py:35-200 (Matheron-rule path).

``efficient_sample(inducing_variable, kernel, q_mu, q_sqrt=..., whiten=...)`` returns a ``Sample``
- a deterministic function draw.  The one-off set-up (sample.py:158-181) is one C-ABI call
(``gpfx_wilson_setup``: kernel matrix, Cholesky/inverse, RFF features, DMMA GEMMs); evaluating the
draw (sample.py:184-198) is ONE fused kernel (``gpfx_wilson_eval``) that never materialises
Phi(X) or K_XZ.  ``num_draws=S`` batches S independent draws (the reference draws one at a time).

The conditional-Gaussian fallback sampler for plain kernels (sample.py:85-134: O(N^3), keeps and
grows the set of evaluated points) is composed from the same device calls plus the full-covariance
conditional (``gpflux_b200/conditionals.py``) and ``draw_conditional_sample`` (sampling/utils.py:27-78).
"""
from __future__ import annotations

import abc
from typing import Callable, Optional, Union

import torch

from .. import _cabi, ops
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
    return _efficient_sample_conditional_gaussian(inducing_variable, kernel, q_mu, q_sqrt=q_sqrt, whiten=whiten)


def _cholesky_with_jitter(cov: torch.Tensor) -> torch.Tensor:
    """gpflux/math.py:25-39: chol(cov + default_jitter() I) for a stack [B, n, n]."""
    n = cov.shape[-1]
    covj = cov + default_jitter() * torch.eye(n, dtype=cov.dtype, device=cov.device)
    return ops.cholesky_inverse(covj.contiguous())


def draw_conditional_sample(mean: torch.Tensor, cov: torch.Tensor, f_old: torch.Tensor,
                            eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A draw from p(f_new | f_old) of the joint Gaussian N(mean [P, N+M], cov [P, N+M, N+M]) over
    [f_old, f_new] (reference gpflux/sampling/utils.py:27-78; ``sample_mvn(..., full_cov=True)`` =
    mean + chol(cov + jitter I) eps).  ``eps`` [P, M] pins the N(0,1) draw.  -> [P, M]."""
    P, N = f_old.shape
    M = mean.shape[-1] - N
    if eps is None:
        eps = torch.randn((P, M), dtype=mean.dtype, device=mean.device)
    cov_new = cov[:, N:, N:].contiguous()
    mean_new = mean[:, N:]
    if N > 0:
        L_old, L_old_inv, _ = _cholesky_with_jitter(cov[:, :N, :N])
        cov_cross = cov[:, :N, N:].contiguous()
        A = ops.gemm(L_old_inv, cov_cross, tri=_cabi.TRI_LOWER_A)  # [P, N, M]
        var_new = ops.gemm(A, A, transA=True, alpha=-1.0, beta=1.0, C_in=cov_new)
        diff = (f_old - mean[:, :N])[:, :, None].contiguous()  # [P, N, 1]
        AM = ops.gemm(L_old_inv, diff, tri=_cabi.TRI_LOWER_A)
        mean_new = mean_new + ops.gemm(A, AM, transA=True)[:, :, 0]
    else:
        var_new = cov_new
    chol, _, _ = _cholesky_with_jitter(var_new)
    return mean_new + ops.gemm(chol, eps[:, :, None].contiguous(), tri=_cabi.TRI_LOWER_A)[:, :, 0]


def _efficient_sample_conditional_gaussian(inducing_variable, kernel, q_mu, *, q_sqrt, whiten) -> Sample:
    """Reference sample.py:85-134: the most costly way to a consistent sample, valid for any
    (stationary, here) kernel: every call conditions the new points on all points evaluated so far."""
    from ..conditionals import full_cov_conditional
    from ..gpflow_compat import MultioutputKernel

    kernels = tuple(kernel.latent_kernels) if isinstance(kernel, MultioutputKernel) else (kernel,)
    ivs = tuple(inducing_variable.inducing_variables) if isinstance(inducing_variable, MultioutputInducingVariables) \
        else (inducing_variable,)
    for k in kernels:
        if not isinstance(k, Stationary):
            raise NotImplementedError(f"kernel {type(k).__name__} is not supported by the B200 path")
    if len({k.kind for k in kernels}) != 1:
        raise NotImplementedError("mixed kernel families are not supported")
    for iv in ivs:
        if not isinstance(iv, InducingPoints):
            raise NotImplementedError("only InducingPoints are supported")
    P = q_mu.shape[-1]
    K = max(len(kernels), len(ivs))
    if K > 1 and K != P:
        raise ValueError(f"{K} separate kernels / inducing variables for {P} latent GPs")
    if q_sqrt is None:
        q_sqrt = torch.zeros((P, q_mu.shape[0], q_mu.shape[0]), dtype=q_mu.dtype, device=q_mu.device)

    def packed(D):
        Zs = [iv.Z.value() for iv in ivs] * (K if len(ivs) == 1 else 1)
        ls = [k.lengthscales_vector(D) for k in kernels] * (K if len(kernels) == 1 else 1)
        var = [k.variance.value().reshape(()) for k in kernels] * (K if len(kernels) == 1 else 1)
        return torch.stack(Zs[:K]), torch.stack(ls[:K]), torch.stack(var[:K])

    class SampleConditional(Sample):
        def __init__(self):
            self.X = None  # [N_old, D]
            self.P = P
            self.f = torch.zeros((0, P), dtype=q_mu.dtype, device=q_mu.device)  # [N_old, P]

        def __call__(self, X_new: torch.Tensor, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
            N_new = X_new.shape[0]
            self.X = X_new if self.X is None else torch.cat([self.X, X_new], dim=0)
            Z, ls, var = packed(X_new.shape[1])
            with torch.no_grad():
                mean, cov = full_cov_conditional(self.X.contiguous(), kernels[0].kind, Z, ls, var, q_mu.detach(),
                                                 q_sqrt.detach(), whiten=whiten, jitter=default_jitter())
                f_new = draw_conditional_sample(mean.T.contiguous(), cov, self.f.T.contiguous(), eps).T.contiguous()
            self.f = torch.cat([self.f, f_new], dim=0)
            assert tuple(f_new.shape) == (N_new, self.P)
            return f_new

    return SampleConditional()


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
    # a13 in one C-ABI call: w_s = sqrt(lambda) eps_w, u_s = q_mu + (q_sqrt eps_u)^T (-> L_uu u_s when whitened),
    # v_s = Kuu^-1 (u_s - Phi(Z) w_s)  (reference sample.py:158-181; compute_A_inv_b, math.py:42-62)
    prior_weights, v, _info = ops.wilson_setup(
        base.kind, Z.contiguous(), ls, var, q_mu.contiguous(), q_sqrt.contiguous(), W.contiguous(), b.contiguous(),
        coeffs.reshape(-1).contiguous(), eps_w.contiguous(), eps_u.reshape(S, P, M).contiguous(), whiten=whiten,
        jitter=default_jitter())

    class WilsonSample(Sample):
        """f(X) = Phi(X) w + K_XZ v, evaluated by one fused kernel."""

        num_draws = S

        def __call__(self, X: torch.Tensor) -> torch.Tensor:
            if X.dim() == 3 and squeeze:
                raise ValueError("a single draw expects X of shape [N, D]")
            out = ops.wilson_eval(base.kind, X, W, b, ls, var, Z, prior_weights, v)
            return out[0] if squeeze else out

    return WilsonSample()
