"""GPU parity of gpfx_natgrad_step and the NatGradModel mirror against the oracle's
autodiff restatement of GPflow's XiNat natural-gradient step."""
import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,P,gamma", [(20, 1, 0.3), (50, 3, 1.0), (64, 2, 0.1), (130, 2, 0.5), (256, 8, 0.2)])
def test_natgrad_step_matches_oracle(M, P, gamma):
    from gpflux_b200 import ops
    from oracle import natgrad_torch as ng

    rng = np.random.default_rng(M + P)
    q_mu = torch.from_numpy(rng.standard_normal((M, P)))
    q_sqrt = torch.from_numpy(np.stack([np.tril(0.1 * rng.standard_normal((M, M))) / np.sqrt(M) + 0.7 * np.eye(M) for _ in range(P)]))
    gm = torch.from_numpy(rng.standard_normal((M, P)))
    G = torch.from_numpy(np.stack([np.tril(0.1 * rng.standard_normal((M, M))) / np.sqrt(M) for _ in range(P)]))
    mu_ref, sq_ref = ng.natgrad_step(q_mu, q_sqrt, gm, G, gamma)
    mu, sq, info = ops.natgrad_step(q_mu.cuda(), q_sqrt.cuda(), gm.cuda(), G.cuda(), gamma)
    assert int(info.abs().sum()) == 0
    assert_close(mu, mu_ref, what="q_mu after natgrad step")
    assert_close(sq, sq_ref, what="q_sqrt after natgrad step")
    assert float(torch.triu(sq, 1).abs().max()) == 0.0


def test_natgrad_flags_indefinite_precision():
    from gpflux_b200 import ops

    M = 16
    q_mu = torch.zeros((M, 1), dtype=torch.float64, device="cuda")
    q_sqrt = torch.eye(M, dtype=torch.float64, device="cuda")[None]
    # dL/dL = -10 L  ->  dL/dS = -5 I  ->  new precision I - 10 gamma I is negative definite at gamma = 1
    G = -10.0 * q_sqrt
    _, _, info = ops.natgrad_step(q_mu, q_sqrt, torch.zeros_like(q_mu), G, 1.0)
    assert int(info[0, 0]) == 0 and int(info[1, 0]) != 0


def test_natgrad_model_one_unit_step_is_optimal():
    """gamma = num_data on the per-datapoint loss == gamma = 1 on the negative ELBO: a single step
    must land on the optimal q(u) of a Gaussian-likelihood SVGP (gradient of q_mu, q_sqrt ~ 0),
    while the hyperparameters move by the regular optimiser (keras_natgrad.py:165-193)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.models import DeepGP
    from gpflux_b200.optimization import NatGradModel, NatGradWrapper

    rng = np.random.default_rng(3)
    N, M = 200, 20
    X = torch.from_numpy(rng.uniform(0, 6, (N, 1))).cuda()
    Y = (torch.sin(X) + 0.1 * torch.from_numpy(rng.standard_normal((N, 1))).cuda())
    kernel = gf.SharedIndependent(gf.SquaredExponential(variance=0.7, lengthscales=0.6), 1)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.linspace(0, 6, M)[:, None]))
    layer = GPLayer(kernel, iv, num_data=N, mean_function=gf.Zero())
    layer.q_mu.assign(0.3 * rng.standard_normal((M, 1)))
    model = DeepGP([layer], gf.Gaussian(0.08)).cuda()

    wrapper = NatGradWrapper(model.as_training_model())
    assert isinstance(wrapper, NatGradModel)
    with pytest.raises(AttributeError):
        wrapper.natgrad_layers
    wrapper.natgrad_layers = True
    assert wrapper.natgrad_layers == [layer]
    with pytest.raises(TypeError):
        wrapper.compile([gf.NaturalGradient(1.0), gf.NaturalGradient(1.0)])
    with pytest.raises(TypeError):
        wrapper.compile(torch.optim.SGD(model.parameters(), lr=0.0))
    ls_before = layer.kernel.kernel.lengthscales.value().clone()
    wrapper.compile([gf.NaturalGradient(gamma=float(N)), lambda params: torch.optim.SGD(params, lr=0.0)])
    loss0 = float(wrapper.train_step(X, Y))
    loss1 = model.loss(X, Y)
    assert float(loss1) < loss0
    g_mu, g_sqrt = torch.autograd.grad(loss1, [layer.q_mu.unconstrained_variable, layer.q_sqrt.unconstrained_variable])
    assert float(g_mu.abs().max()) * N < 1e-6
    assert float(torch.tril(g_sqrt).abs().max()) * N < 1e-6
    assert torch.equal(layer.kernel.kernel.lengthscales.value(), ls_before)  # lr = 0 optimiser saw them
    # a second natural-gradient step from the optimum stays there
    q_before = layer.q_mu.value().clone()
    wrapper.train_step(X, Y)
    assert_close(layer.q_mu.value(), q_before, tol=1e-7, what="fixed point")
