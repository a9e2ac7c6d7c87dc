"""torch-CPU float64 restatement of GPflow's natural-gradient step for q(u).

TEST INFRASTRUCTURE ONLY - PARITY UNPINNED (see ``oracle/__init__.py``).

Reference call site: ``NatGradModel._apply_backwards_pass`` calls
``natgrad_optimizer._natgrad_apply_gradients(q_mu_grad, q_sqrt_grad, q_mu, q_sqrt)`` per layer
(gpflux/optimization/keras_natgrad.py:187-191).  The arithmetic is GPflow 2.9.x
``gpflow/optimizers/natgrad.py`` (not vendored): the default ``XiNat`` parameterisation, restated
here the way GPflow evaluates it - by autodiff through ``expectation_to_meanvarsqrt`` - so that it
is independent of the closed form the CUDA path uses.

Shapes follow GPflux: q_mu [M, P], q_sqrt [P, M, M] (lower triangular).
"""
from __future__ import annotations

import torch


def meanvarsqrt_to_expectation(m: torch.Tensor, v_sqrt: torch.Tensor):
    """gpflow.optimizers.natgrad.meanvarsqrt_to_expectation: eta1 = m, eta2 = S + m m^T."""
    v = v_sqrt @ v_sqrt.transpose(-1, -2)
    mt = m.T  # [P, M]
    return m, v + mt[:, :, None] * mt[:, None, :]


def expectation_to_meanvarsqrt(eta1: torch.Tensor, eta2: torch.Tensor):
    """gpflow ...expectation_to_meanvarsqrt: var = eta2 - eta1 eta1^T, sqrt = cholesky(var)."""
    mt = eta1.T
    var = eta2 - mt[:, :, None] * mt[:, None, :]
    return eta1, torch.linalg.cholesky(var)


def meanvarsqrt_to_natural(mu: torch.Tensor, s_sqrt: torch.Tensor):
    """gpflow ...meanvarsqrt_to_natural: nat1 = S^-1 mu, nat2 = -S^-1 / 2."""
    eye = torch.eye(s_sqrt.shape[-1], dtype=s_sqrt.dtype).expand_as(s_sqrt)
    s_sqrt_inv = torch.linalg.solve_triangular(s_sqrt, eye, upper=False)
    s_inv = s_sqrt_inv.transpose(-1, -2) @ s_sqrt_inv
    nat1 = (s_inv @ mu.T[:, :, None])[:, :, 0].T
    return nat1, -0.5 * s_inv


def natural_to_meanvarsqrt(nat1: torch.Tensor, nat2: torch.Tensor):
    """gpflow ...natural_to_meanvarsqrt: chol(-2 nat2) -> inverse -> S = C^-T C^-1 -> chol(S)."""
    var_sqrt_inv = torch.linalg.cholesky(-2.0 * nat2)
    eye = torch.eye(nat2.shape[-1], dtype=nat2.dtype).expand_as(nat2)
    var_sqrt = torch.linalg.solve_triangular(var_sqrt_inv, eye, upper=False)
    S = var_sqrt.transpose(-1, -2) @ var_sqrt
    mu = (S @ nat1.T[:, :, None])[:, :, 0].T
    return mu, torch.linalg.cholesky(S)


def natgrad_step(q_mu, q_sqrt, dq_mu, dq_sqrt, gamma: float):
    """NaturalGradient._natgrad_apply_gradients with xi_transform = XiNat():

        dL/deta = d(mean, varsqrt)/d(eta)^T [dL/dmean, dL/dvarsqrt]      (tape.gradient, output_gradients)
        nat_new = nat - gamma * dL/deta ;  (q_mu, q_sqrt) <- natural_to_meanvarsqrt(nat_new)

    ``dq_sqrt`` is the gradient with respect to the constrained (lower-triangular) q_sqrt, which is
    what gpflow's ``_to_constrained`` hands over for the FillTriangular transform."""
    q_mu = q_mu.detach().double()
    q_sqrt = torch.tril(q_sqrt.detach().double())
    eta1, eta2 = meanvarsqrt_to_expectation(q_mu, q_sqrt)
    eta1 = eta1.clone().requires_grad_(True)
    eta2 = eta2.clone().requires_grad_(True)
    mean, varsqrt = expectation_to_meanvarsqrt(eta1, eta2)
    d1, d2 = torch.autograd.grad([mean, varsqrt], [eta1, eta2],
                                 grad_outputs=[dq_mu.detach().double(), torch.tril(dq_sqrt.detach().double())])
    # tf.linalg.cholesky's registered gradient is symmetric (grad_a = (g + g^T)/2); torch agrees,
    # and the explicit symmetrisation keeps this restatement independent of that convention
    d2 = 0.5 * (d2 + d2.transpose(-1, -2))
    nat1, nat2 = meanvarsqrt_to_natural(q_mu, q_sqrt)
    return natural_to_meanvarsqrt(nat1 - gamma * d1, nat2 - gamma * d2)
