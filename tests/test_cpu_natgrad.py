"""CPU checks of the natural-gradient oracle (oracle/natgrad_torch.py): the autodiff restatement of
GPflow's XiNat step against the closed form the CUDA path implements, and the textbook invariant
that ONE gamma = 1 step on the un-scaled negative ELBO of a Gaussian-likelihood SVGP lands on the
optimal q(u) (the reference's natgrad equivalence test, tests/integration/test_svgp_equivalence.py:
215-262, relies on the same property over 20 steps)."""
import numpy as np
import torch

from oracle import natgrad_torch as ng
from oracle import svgp_torch as ot


def _closed_form(q_mu, q_sqrt, gm, G, gamma):
    """numpy restatement of gpflux_b200/csrc/natgrad.cu's derivation."""
    M, P = q_mu.shape
    mu_new, sqrt_new = np.empty_like(q_mu), np.empty_like(q_sqrt)
    for p in range(P):
        L = np.tril(q_sqrt[p])
        Li = np.linalg.inv(L)
        T = np.tril(L.T @ np.tril(G[p]))
        T[np.diag_indices(M)] *= 0.5
        dS = Li.T @ T @ Li
        Shat = 0.5 * (dS + dS.T)
        Sinv = Li.T @ Li
        lam = Sinv + 2.0 * gamma * Shat
        nat1 = lam @ q_mu[:, p] - gamma * gm[:, p]
        Ci = np.linalg.inv(np.linalg.cholesky(lam))
        S = Ci.T @ Ci
        mu_new[:, p] = S @ nat1
        sqrt_new[p] = np.linalg.cholesky(S)
    return mu_new, sqrt_new


def test_autodiff_step_matches_closed_form():
    rng = np.random.default_rng(0)
    for M, P in [(5, 1), (17, 3)]:
        q_mu = rng.standard_normal((M, P))
        q_sqrt = np.stack([np.tril(0.2 * rng.standard_normal((M, M))) + np.eye(M) for _ in range(P)])
        gm = rng.standard_normal((M, P))
        # a PSD-ish dL/dS keeps the new precision positive definite at gamma = 0.3
        G = np.stack([np.tril(0.1 * rng.standard_normal((M, M))) for _ in range(P)])
        mu_a, sq_a = ng.natgrad_step(*(torch.from_numpy(a) for a in (q_mu, q_sqrt, gm, G)), 0.3)
        mu_c, sq_c = _closed_form(q_mu, q_sqrt, gm, G, 0.3)
        np.testing.assert_allclose(mu_a.numpy(), mu_c, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(sq_a.numpy(), sq_c, rtol=1e-10, atol=1e-12)


def test_one_unit_step_reaches_the_optimal_q():
    rng = np.random.default_rng(1)
    N, M, D = 120, 15, 2
    X = torch.from_numpy(rng.standard_normal((N, D)))
    Y = torch.from_numpy(np.sin(X.numpy().sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1)))
    noise = torch.tensor(0.05, dtype=torch.float64)
    spec = ot.make_spec(rng, M, D, 1)

    def neg_elbo(s):
        loss, _ = ot.deep_gp_loss([s], X, Y, [None], noise, N)
        return loss * N

    s = ot.spec_requires_grad(spec)
    neg_elbo(s).backward()
    mu1, sq1 = ng.natgrad_step(s["q_mu"], s["q_sqrt"], s["q_mu"].grad, s["q_sqrt"].grad, 1.0)
    s1 = dict(spec, q_mu=mu1, q_sqrt=sq1)
    s1 = ot.spec_requires_grad(s1)
    neg_elbo(s1).backward()
    assert float(s1["q_mu"].grad.abs().max()) < 1e-7
    assert float(torch.tril(s1["q_sqrt"].grad).abs().max()) < 1e-7
    # whitened optimum in closed form: precision I + A A^T / noise, mean = precision^-1 A y / noise
    L = torch.linalg.cholesky(ot.Kuu(spec, 0))
    A = torch.linalg.solve_triangular(L, ot.Kuf(spec, 0, X), upper=False)
    prec = torch.eye(M, dtype=torch.float64) + A @ A.T / noise
    mean = torch.linalg.solve(prec, A @ Y / noise)
    np.testing.assert_allclose(mu1.numpy(), mean.numpy(), rtol=1e-8, atol=1e-10)
    S1 = sq1[0] @ sq1[0].T
    np.testing.assert_allclose(S1.numpy(), torch.linalg.inv(prec).numpy(), rtol=1e-8, atol=1e-10)
