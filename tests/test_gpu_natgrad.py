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


def _oracle_natgrad_adam_run(specs, state, u_noise0, batches, num_data, gamma, lr):
    """The reference's NatGradModel step on the oracle (keras_natgrad.py:165-193): all gradients are
    taken at the same point; (q_mu, q_sqrt) of every layer move by the XiNat natural-gradient step,
    everything else by Adam in its unconstrained variable."""
    from oracle import natgrad_torch as ng
    from oracle import svgp_torch as ot

    sp = torch.nn.functional.softplus
    leaves = [{k: v.clone().requires_grad_(True) for k, v in st.items()} for st in state]
    u_noise = u_noise0.clone().requires_grad_(True)
    others = [lv[k] for lv in leaves for k in ("Z", "u_ls", "u_var")] + [u_noise]
    opt = torch.optim.Adam(others, lr=lr)
    losses = []
    for Xb, Yb, eps in batches:
        cur = []
        for spec, lv in zip(specs, leaves):
            s = dict(spec)
            s.update(Z=lv["Z"].unsqueeze(0), lengthscales=sp(lv["u_ls"]).reshape(1, -1), variance=sp(lv["u_var"]).reshape(1),
                     q_mu=lv["q_mu"], q_sqrt=torch.tril(lv["q_sqrt"]))
            cur.append(s)
        loss, _ = ot.deep_gp_loss(cur, Xb, Yb, eps, sp(u_noise), num_data)
        variational = [lv[k] for lv in leaves for k in ("q_mu", "q_sqrt")]
        grads = torch.autograd.grad(loss, variational + others)
        for i, lv in enumerate(leaves):
            mu, sq = ng.natgrad_step(lv["q_mu"], lv["q_sqrt"], grads[2 * i], grads[2 * i + 1], gamma)
            with torch.no_grad():
                lv["q_mu"].copy_(mu)
                lv["q_sqrt"].copy_(sq)
        for v, g in zip(others, grads[len(variational):]):
            v.grad = g
        opt.step()
        losses.append(float(loss.detach()))
    return losses, leaves, u_noise


@pytest.mark.parametrize("case", ["single_layer", "two_layer_philox_noise"])
def test_natgrad_plus_adam_trajectory_matches_oracle(case):
    """20 NatGradModel.train_step calls (natural gradient on every q(u), Adam on the rest) against the
    same loop on the oracle - the keras-natgrad arm of the reference's
    tests/integration/test_svgp_equivalence.py:215-283 (atol 1e-7 there).  The 2-layer case lets the
    hidden layer draw its own noise through GPLayer.seed_noise; the oracle regenerates that stream
    with oracle/philox_numpy.py (call c of the layer = stream offset c)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.models import DeepGP
    from gpflux_b200.optimization import NatGradWrapper
    from oracle import philox_numpy as px
    from oracle.from_model import specs_from_model

    steps, lr, seed = 20, 0.01, 77
    rng = np.random.default_rng(11)
    if case == "single_layer":
        N, M = 200, 20
        X = rng.uniform(0, 6, (N, 1))
        Y = np.sin(X) + 0.1 * rng.standard_normal((N, 1))
        kernel = gf.SharedIndependent(gf.SquaredExponential(variance=0.7, lengthscales=0.6), 1)
        iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.linspace(0, 6, M)[:, None]))
        layer = GPLayer(kernel, iv, num_data=N, mean_function=gf.Zero())
        layer.q_mu.assign(0.3 * rng.standard_normal((M, 1)))
        model = DeepGP([layer], gf.Gaussian(0.08))
        gamma = 0.1 * N  # 0.1 on the ELBO scale (the loss is per datapoint)
        batches = [(torch.from_numpy(X), torch.from_numpy(Y), [None])] * steps
    else:
        D, M, B = 3, 30, 128
        X = rng.standard_normal((1000, D))
        Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((1000, 1))
        model = build_constant_input_dim_deep_gp(X, 2, Config(M, 1e-1, 0.05), z_init=X[:M])
        for layer in model.f_layers:
            layer.q_mu.assign(0.2 * rng.standard_normal(tuple(layer.q_mu.shape)))
        gamma = 0.02 * model.num_data
        batches = []
        for t in range(steps):
            idx = rng.choice(X.shape[0], B, replace=False)
            eps = torch.from_numpy(px.randn(seed, t, B * D).reshape(B, D))
            batches.append((torch.from_numpy(X[idx]), torch.from_numpy(Y[idx]), [eps, None]))

    specs = specs_from_model(model)
    state = []
    for layer in model.f_layers:
        kern = layer.kernel.kernel
        state.append({
            "Z": layer.inducing_variable.inducing_variable.Z.unconstrained_variable.detach().clone(),
            "u_ls": kern.lengthscales.unconstrained_variable.detach().clone().reshape(-1),
            "u_var": kern.variance.unconstrained_variable.detach().clone().reshape(()),
            "q_mu": layer.q_mu.unconstrained_variable.detach().clone(),
            "q_sqrt": layer.q_sqrt.unconstrained_variable.detach().clone(),
        })
    u_noise0 = model.likelihood_layer.likelihood.variance.unconstrained_variable.detach().clone()
    o_losses, o_leaves, o_noise = _oracle_natgrad_adam_run(specs, state, u_noise0, batches, model.num_data, gamma, lr)
    assert np.all(np.isfinite(o_losses)) and o_losses[-1] < o_losses[0]
    model = model.cuda()
    if case == "two_layer_philox_noise":
        model.f_layers[0].seed_noise(seed)
    wrapper = NatGradWrapper(model.as_training_model())
    wrapper.natgrad_layers = True
    wrapper.compile([gf.NaturalGradient(gamma=gamma) for _ in model.f_layers] + [lambda params: torch.optim.Adam(params, lr=lr)])
    g_losses = [float(wrapper.train_step(Xb.cuda(), Yb.cuda())) for Xb, Yb, _ in batches]

    dev = np.max(np.abs(np.array(g_losses) - np.array(o_losses)))
    print(f"{case}: max |loss deviation| over {steps} steps = {dev:.3g}; loss {o_losses[0]:.6f} -> {o_losses[-1]:.6f}")
    np.testing.assert_allclose(g_losses, o_losses, rtol=1e-7, atol=1e-7)  # the reference's tolerances
    for layer, lv in zip(model.f_layers, o_leaves):
        assert_close(layer.q_mu.value(), lv["q_mu"], tol=1e-6, what="q_mu after training")
        assert_close(layer.q_sqrt.value(), torch.tril(lv["q_sqrt"]), tol=1e-6, what="q_sqrt after training")
        assert_close(layer.inducing_variable.inducing_variable.Z.unconstrained_variable, lv["Z"], tol=1e-6, what="Z after training")
    assert_close(model.likelihood_layer.likelihood.variance.unconstrained_variable.reshape(()), o_noise.reshape(()), tol=1e-6, what="noise")
