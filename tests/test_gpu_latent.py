"""GPU parity of the LatentVariableLayer arithmetic and a conditional-density DGP
(BASELINE config 4 shape: LatentVariableLayer + GPLayer) end to end against the oracle."""
import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,W", [(1, 1, 1), (300, 1, 2), (1000, 3, 5), (257, 0, 2)])
def test_latent_forward_backward(N, D, W):
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(20)
    X = torch.from_numpy(rng.standard_normal((N, D))).requires_grad_(True)
    mean = torch.from_numpy(rng.standard_normal((N, W))).requires_grad_(True)
    std = torch.from_numpy(rng.uniform(0.1, 2.0, (N, W))).requires_grad_(True)
    pm = torch.from_numpy(rng.standard_normal(W))
    ps = torch.from_numpy(rng.uniform(0.5, 1.5, W))
    eps = torch.from_numpy(rng.standard_normal((N, W)))
    wgt = torch.from_numpy(rng.standard_normal((N, D + W)))
    out, kl = ot.latent_layer(X, mean, std, pm, ps, eps)
    ((out * wgt).sum() + 0.3 * kl).backward()
    # reference tests/gpflux/layers/test_latent_variable_layer.py:82-99: KL == the library KL
    ref_kl = torch.distributions.kl_divergence(torch.distributions.Normal(mean, std), torch.distributions.Normal(pm, ps)).sum()
    assert_close(kl.detach(), ref_kl.detach(), tol=1e-12, what="oracle KL vs torch.distributions")

    gX, gm, gs = (t.detach().cuda().requires_grad_(True) for t in (X, mean, std))
    gout, gkl = ops.latent_sample_and_kl(gX, gm, gs, pm.cuda(), ps.cuda(), eps.cuda())
    ((gout * wgt.cuda()).sum() + 0.3 * gkl).backward()
    assert_close(gout, out, scale=1.0, what="composed output")
    assert_close(gkl, kl, what="kl")
    if D > 0:
        assert_close(gX.grad, X.grad, what="dX")
    assert_close(gm.grad, mean.grad, what="dmean")
    assert_close(gs.grad, std.grad, what="dstd")


def test_cde_model_trains_and_matches_oracle():
    """LatentVariableLayer + GPLayer conditional-density model (cfg 4 shape, reduced): loss parity
    with the oracle and the reference's integration invariant (loss decreases;
    tests/integration/test_latent_variable_integration.py:94-124)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.helpers import construct_basic_inducing_variables, construct_basic_kernel
    from gpflux_b200.layers import GPLayer, LikelihoodLayer
    from gpflux_b200.layers.latent_variable_layer import (DirectlyParameterizedNormalDiag, LatentVariableLayer,
                                                          MultivariateNormalDiagPrior)
    from gpflux_b200.models import DeepGP
    from oracle import svgp_torch as ot
    from oracle.from_model import spec_from_layer

    rng = np.random.default_rng(21)
    N, W, M = 200, 2, 30
    X = rng.uniform(0, 1, (N, 1))
    Y = np.sin(6 * X) + rng.standard_normal((N, 1)) * (0.1 + 0.5 * X)
    prior = MultivariateNormalDiagPrior(np.zeros(W), np.ones(W))
    enc = DirectlyParameterizedNormalDiag(N, W, means=0.01 * rng.standard_normal((N, W)))
    lv = LatentVariableLayer(prior, enc)
    kern = construct_basic_kernel(gf.SquaredExponential(lengthscales=np.ones(1 + W)), output_dim=1, share_hyperparams=True)
    iv = construct_basic_inducing_variables(M, 1 + W, share_variables=True, output_dim=1, z_init=rng.standard_normal((M, 1 + W)))
    gp = GPLayer(kern, iv, N, mean_function=gf.Zero())
    gp.q_mu.assign(0.1 * rng.standard_normal((M, 1)))
    spec = spec_from_layer(gp)
    model = DeepGP([lv, gp], LikelihoodLayer(gf.Gaussian(0.1))).cuda()
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    eps_w = torch.from_numpy(rng.standard_normal((N, W)))
    loss = model.loss(Xd, Yd, eps=[eps_w.cuda(), None])
    # oracle
    means, stds = enc.means.value().detach().cpu(), enc.stds.value().detach().cpu()
    h, kl_w = ot.latent_layer(torch.from_numpy(X), means, stds, prior.loc, prior.scale_diag, eps_w)
    m, v = ot.predict(spec, h)
    ve = ot.gaussian_variational_expectations(m, v, torch.from_numpy(Y), torch.tensor(0.1, dtype=torch.float64))
    ref = kl_w / N + ot.prior_kl(spec) / N - ve.mean()
    assert_close(loss, ref, what="CDE loss")
    # prediction mode samples the prior and adds no loss
    out = model(Xd)
    assert float(lv.losses[0]) == 0.0 and tuple(out.f_mean.shape) == (N, 1)
    # training decreases the loss
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    first = last = None
    for it in range(40):
        opt.zero_grad(set_to_none=True)
        l = model.loss(Xd, Yd)
        l.backward()
        opt.step()
        first = float(l) if first is None else first
        last = float(l)
    assert last < first
