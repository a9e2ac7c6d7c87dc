"""SeparateIndependent with MIXED kernel families in one GPLayer - the reference builds
``SeparateIndependent([SquaredExponential, Matern32])`` in
tests/gpflux/layers/basis_functions/fourier_features/test_random.py:209-228 - and the lazy non-PD
status (where the reference raises InvalidArgumentError from tf.linalg.cholesky)."""
import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu


def test_mixed_kernel_families_match_oracle_per_output():
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(3)
    N, M, D, P = 150, 30, 3, 3
    kinds = ["se", "matern32", "se"]
    cls = {"se": gf.SquaredExponential, "matern32": gf.Matern32}
    ls = rng.uniform(1.0, 2.0, (P, D))
    var = rng.uniform(0.8, 1.2, P)
    Z = rng.standard_normal((P, M, D))
    kern = gf.SeparateIndependent([cls[k](variance=var[p], lengthscales=ls[p]) for p, k in enumerate(kinds)])
    iv = gf.SeparateIndependentInducingVariables([gf.InducingPoints(Z[p]) for p in range(P)])
    layer = GPLayer(kern, iv, num_data=1000, mean_function=gf.Identity())
    q_mu = 0.3 * rng.standard_normal((M, P))
    q_sqrt = np.stack([0.5 * np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M)), -1) for _ in range(P)])
    layer.q_mu.assign(q_mu)
    layer.q_sqrt.assign(q_sqrt)
    layer = layer.cuda()
    X = torch.from_numpy(rng.standard_normal((N, D)))
    eps = torch.from_numpy(rng.standard_normal((N, P)))
    w = torch.from_numpy(rng.standard_normal((N, P)))

    Xg = X.cuda().requires_grad_(True)
    dist = layer(Xg, training=True, eps=eps.cuda())
    loss = (dist.sample() * w.cuda()).sum() + (dist.variance() * w.cuda()).sum() + layer.losses[0] * 1000
    loss.backward()

    # oracle: each output is its own single-kernel layer (Identity mean adds X[:, p])
    Xo = X.clone().requires_grad_(True)
    o_loss = 0.0
    specs = []
    for p, kind in enumerate(kinds):
        spec = ot.spec_requires_grad({
            "kernel": kind, "Z": torch.from_numpy(Z[p : p + 1]), "lengthscales": torch.from_numpy(ls[p : p + 1]),
            "variance": torch.from_numpy(var[p : p + 1]), "q_mu": torch.from_numpy(q_mu[:, p : p + 1]),
            "q_sqrt": torch.from_numpy(q_sqrt[p : p + 1]), "mean": "zero", "whiten": True, "jitter": 1e-6})
        specs.append(spec)
        m, v = ot.conditional(spec, Xo)
        m = m + Xo[:, p : p + 1]
        f = m + torch.sqrt(v) * eps[:, p : p + 1]
        o_loss = o_loss + (f * w[:, p : p + 1]).sum() + (v * w[:, p : p + 1]).sum() + ot.prior_kl(spec)
    o_loss.backward()
    assert_close(loss, o_loss, what="loss")
    assert_close(Xg.grad, Xo.grad, what="dX")
    assert_close(layer.q_mu.unconstrained_variable.grad, torch.cat([s["q_mu"].grad for s in specs], 1), what="dq_mu")
    assert_close(torch.tril(layer.q_sqrt.unconstrained_variable.grad), torch.cat([s["q_sqrt"].grad for s in specs], 0),
                 what="dq_sqrt")
    for p in range(P):
        Zp = layer.inducing_variable.inducing_variables[p].Z.unconstrained_variable
        assert_close(Zp.grad, specs[p]["Z"].grad[0], what=f"dZ[{p}]")
        u = layer.kernel.latent_kernels[p].lengthscales.unconstrained_variable
        assert_close(u.grad, specs[p]["lengthscales"].grad[0] * torch.sigmoid(u.detach().cpu()), what=f"dls[{p}]")


def test_non_pd_kuu_is_reported_lazily_on_the_training_path():
    """Duplicate inducing points with a NEGATIVE jitter make Kuu indefinite: the step itself does not
    sync or raise (NaNs flow), ``check_status`` raises GpfxError(E_NOT_PD) afterwards, and fit() raises
    as soon as the loss it reads back is non-finite."""
    from gpflux_b200 import _cabi
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.models import DeepGP

    rng = np.random.default_rng(0)
    M, D = 16, 2
    Zv = rng.standard_normal((M, D))
    Zv[1] = Zv[0]  # exactly singular without jitter

    def build():
        layer = GPLayer(gf.SharedIndependent(gf.SquaredExponential(), 1),
                        gf.SharedIndependentInducingVariables(gf.InducingPoints(Zv)), num_data=100,
                        mean_function=gf.Zero())
        return DeepGP([layer], gf.Gaussian(0.1)).cuda()

    model = build()
    X = torch.from_numpy(rng.standard_normal((64, D))).cuda()
    Y = torch.from_numpy(rng.standard_normal((64, 1))).cuda()
    old = gf.default_jitter()
    try:
        gf.set_default_jitter(-1e-3)
        loss = model.loss(X, Y)  # no exception here: nothing on this path synchronises
        with pytest.raises(_cabi.GpfxError, match="GPFX_E_NOT_PD"):
            model.check_status()
        assert not torch.isfinite(loss)
        tm = model.as_training_model()
        tm.compile(torch.optim.SGD(model.parameters(), lr=1e-3))
        with pytest.raises(_cabi.GpfxError, match="not positive definite"):
            tm.fit({"inputs": X, "targets": Y}, epochs=1)
    finally:
        gf.set_default_jitter(old)
    # a healthy evaluation reports nothing (fresh model: the failed fit() above stepped on NaN gradients)
    model = build()
    assert torch.isfinite(model.loss(X, Y))
    model.check_status()
