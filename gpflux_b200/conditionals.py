"""Full-covariance conditional: ``gpflow.conditionals.conditional(..., full_cov=True)`` as reached
from ``GPLayer.predict`` (reference gpflux/layers/gp_layer.py:257-266) and from the
conditional-Gaussian sampler (gpflux/sampling/sample.py:113-121), composed from the C-ABI kernel-matrix,
Cholesky and DMMA GEMM entry points.  Each building block is a torch.autograd.Function whose backward is
again C-ABI calls (transposed GEMMs, gpfx_kernel_matrix_bwd, gpfx_cholesky_backward), so the branch is
differentiable end to end, as in the reference."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _cabi, ops


def full_cov_conditional(X, kind: str, Z, lengthscales, variance, q_mu, q_sqrt, *, whiten: bool,
                         jitter: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> mean [N,P] (without the mean function), cov [P,N,N] = Knn - A^T A + B_p^T B_p with
    A = L^-1 Kuf and B_p = tril(q_sqrt_p)^T A (whitened variables; a non-whitened q(u) is first
    mapped to them, (q_mu, q_sqrt) -> (L^-1 q_mu, L^-1 q_sqrt), since u = L v).
    Z [K,M,D], lengthscales [K,D], variance [K] with K = 1 (shared) or P (separate)."""
    N, D = X.shape
    P = q_mu.shape[1]
    Z, ls, var = Z.contiguous(), lengthscales.contiguous(), variance.contiguous()
    K = Z.shape[0]
    Kuu = ops.kernel_matrix(kind, Z, Z, ls, var, jitter=jitter)
    _, Linv, _ = ops.cholesky_inverse(Kuu)
    Kuf = ops.kernel_matrix(kind, Z, X, ls, var)  # [K,M,N]
    Knn = ops.kernel_matrix(kind, X.expand(K, N, D).contiguous(), X, ls, var)  # [K,N,N]
    A = ops.gemm(Linv, Kuf, tri=_cabi.TRI_LOWER_A)
    base = ops.gemm(A, A, transA=True, alpha=-1.0, beta=1.0, C_in=Knn)
    if K == 1 and P > 1:
        A = A.expand(P, -1, -1).contiguous()
        base = base.expand(P, -1, -1).contiguous()
        Linv = Linv.expand(P, -1, -1).contiguous()
    Lq = torch.tril(q_sqrt)
    q_mu_t = q_mu.T.contiguous()[:, :, None]  # [P,M,1]
    if not whiten:
        Lq = ops.gemm(Linv, Lq, tri=_cabi.TRI_LOWER_A | _cabi.TRI_LOWER_B)
        q_mu_t = ops.gemm(Linv, q_mu_t, tri=_cabi.TRI_LOWER_A)
    B = ops.gemm(Lq, A, transA=True, tri=_cabi.TRI_UPPER_A)  # [P,M,N]
    cov = ops.gemm(B, B, transA=True, beta=1.0, C_in=base)  # [P,N,N]
    mean = ops.gemm(A, q_mu_t, transA=True)[:, :, 0].T.contiguous()  # [N,P]
    return mean, cov
