"""GPU tests of the host-side mirror (GPLayer / LikelihoodLayer / DeepGP) against the oracle and
against the invariants the reference's own tests assert."""
import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu


def _build(D=3, M=40, n_layers=2, seed=0, factor=1e-2):
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp

    rng = np.random.default_rng(seed)
    X = rng.standard_normal((1500, D))
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((1500, 1))
    model = build_constant_input_dim_deep_gp(X, n_layers, Config(M, factor, 0.05), z_init=X[:M])
    for layer in model.f_layers:
        layer.q_mu.assign(0.2 * rng.standard_normal(tuple(layer.q_mu.shape)))
    return model, X, Y, rng


@pytest.mark.parametrize("n_layers,D,M,N", [(1, 1, 20, 200), (2, 3, 40, 300), (3, 5, 70, 257), (2, 8, 256, 1024)])
def test_deep_gp_loss_and_grads_match_oracle(n_layers, D, M, N):
    from oracle import svgp_torch as ot
    from oracle.from_model import noise_variance_from_model, specs_from_model

    model, X, Y, rng = _build(D, M, n_layers)
    specs, noise = specs_from_model(model), noise_variance_from_model(model)
    model = model.cuda()
    Xb, Yb = torch.from_numpy(X[:N]), torch.from_numpy(Y[:N])
    eps = [torch.from_numpy(rng.standard_normal((N, D))) for _ in range(n_layers - 1)] + [None]
    loss = model.loss(Xb.cuda(), Yb.cuda(), eps=[e.cuda() if e is not None else None for e in eps])
    loss.backward()

    o_specs = [ot.spec_requires_grad(s) for s in specs]
    o_noise = noise.clone().requires_grad_(True)
    o_loss, _ = ot.deep_gp_loss(o_specs, Xb, Yb, eps, o_noise, model.num_data)
    o_loss.backward()
    assert_close(loss, o_loss, what="loss")
    for i, layer in enumerate(model.f_layers):
        assert_close(layer.q_mu.unconstrained_variable.grad, o_specs[i]["q_mu"].grad, what=f"L{i} dq_mu")
        assert_close(layer.q_sqrt.unconstrained_variable.grad, o_specs[i]["q_sqrt"].grad, what=f"L{i} dq_sqrt")
        Zp = layer.inducing_variable.inducing_variable.Z.unconstrained_variable
        assert_close(Zp.grad, o_specs[i]["Z"].grad[0], what=f"L{i} dZ")
        # softplus chain rule on the host: d/du = d/dtheta * sigmoid(u)
        kern = layer.kernel.kernel
        u = kern.lengthscales.unconstrained_variable
        assert_close(u.grad, o_specs[i]["lengthscales"].grad[0] * torch.sigmoid(u.detach().cpu()), what=f"L{i} dls")
        u = kern.variance.unconstrained_variable
        assert_close(u.grad, o_specs[i]["variance"].grad[0] * torch.sigmoid(u.detach().cpu()), what=f"L{i} dvar")
    u = model.likelihood_layer.likelihood.variance.unconstrained_variable
    assert_close(u.grad, o_noise.grad * torch.sigmoid(u.detach().cpu()), what="dnoise")
    # ELBO = -loss * num_data (reference deep_gp.py:226-231)
    elbo = model.elbo((Xb.cuda(), Yb.cuda()), eps=[e.cuda() if e is not None else None for e in eps])
    assert_close(elbo, -o_loss.detach() * model.num_data, what="elbo")


def test_snelson_anchor():
    """SURVEY §8c restatement-derived anchors: snelson, SE var=0.7 ls=0.6, noise 0.08, Z=linspace(0,6,20)."""
    import os

    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.models import DeepGP

    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "snelson1d.npz"))
    X, Y = torch.from_numpy(d["X"]).cuda(), torch.from_numpy(d["Y"]).cuda()
    M = 20
    kernel = gf.SharedIndependent(gf.SquaredExponential(variance=0.7, lengthscales=0.6), 1)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.linspace(0, 6, M)[:, None]))
    layer = GPLayer(kernel, iv, num_data=200, mean_function=gf.Zero())
    model = DeepGP([layer], gf.Gaussian(0.08)).cuda()
    # at init: fmean == 0, fvar == 0.7, KL == 0, sum VE anchor
    elbo0 = model.elbo((X, Y))
    assert_close(elbo0, torch.tensor(-1840.588157487725, dtype=torch.float64), what="ELBO at init")
    layer.q_mu.assign(np.sin(np.arange(20.0))[:, None])
    layer.q_sqrt.assign((0.1 * np.tril(np.cos(np.arange(400.0)).reshape(20, 20)) + np.eye(20))[None])
    elbo1 = model.elbo((X, Y))
    assert_close(elbo1, torch.tensor(-2866.7514905888424, dtype=torch.float64), what="ELBO perturbed")
    assert_close(layer.prior_kl(), torch.tensor(5.312033758504997, dtype=torch.float64), what="KL perturbed")


def test_layer_losses_follow_reference_semantics():
    """reference tests/gpflux/layers/test_gp_layer.py:159-180: training loss == prior_kl()/num_data,
    eval loss == 0, one loss per layer after repeated calls."""
    model, X, Y, _ = _build()
    model = model.cuda()
    layer = model.f_layers[0]
    Xb = torch.from_numpy(X[:100]).cuda()
    layer(Xb, training=True)
    assert len(layer.losses) == 1
    assert float(layer.losses[0]) == float(layer.prior_kl()) / layer.num_data
    layer(Xb, training=True)
    assert len(layer.losses) == 1
    layer(Xb, training=False)
    assert float(layer.losses[0]) == 0.0
    layer(Xb)
    assert float(layer.losses[0]) == 0.0


def test_prior_kl_zero_at_init_and_positive_after():
    """reference tests/gpflux/layers/test_gp_layer.py:67-85."""
    from gpflux_b200.helpers import construct_gp_layer

    layer = construct_gp_layer(100, 20, 2, 3, z_init=np.random.default_rng(0).standard_normal((20, 2))).cuda()
    assert tuple(layer.q_mu.shape) == (20, 3) and tuple(layer.q_sqrt.shape) == (3, 20, 20)
    assert float(layer.prior_kl()) == 0.0
    layer.q_mu.assign(np.ones((20, 3)))
    assert float(layer.prior_kl()) > 0.0
    layer.q_mu.assign(np.zeros((20, 3)))
    layer.q_sqrt.assign(2.0 * np.stack([np.eye(20)] * 3))
    assert float(layer.prior_kl()) > 0.0
    kl = layer.prior_kl()
    kl.backward()
    assert float(layer.q_sqrt.unconstrained_variable.grad.abs().sum()) > 0.0


def test_shapes_num_samples_and_predict():
    """reference tests/gpflux/layers/test_gp_layer.py:123-156."""
    from gpflux_b200.helpers import construct_gp_layer

    Z = np.random.default_rng(1).standard_normal((16, 2))
    layer = construct_gp_layer(100, 16, 2, 3, z_init=Z).cuda()
    X = torch.from_numpy(np.random.default_rng(2).standard_normal((31, 2))).cuda()
    mean, var = layer.predict(X)
    assert tuple(mean.shape) == (31, 3) and tuple(var.shape) == (31, 3)
    dist = layer(X)
    assert tuple(dist.sample().shape) == (31, 3)
    layer.num_samples = 5
    dist = layer(X)
    assert tuple(layer._convert_to_tensor_fn(dist).shape) == (5, 31, 3)
    layer.num_samples = None
    layer.full_cov = layer.full_output_cov = True
    with pytest.raises(NotImplementedError):
        layer(X)


def test_dtype_errors():
    """reference tests/gpflux/models/test_deep_gp.py:160-174: non-float64 inputs raise ValueError."""
    model, X, Y, _ = _build()
    model = model.cuda()
    for dt in (torch.float32, torch.float16, torch.int32):
        with pytest.raises(ValueError):
            model(torch.from_numpy(X[:10]).to(dt).cuda())
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp

    with pytest.raises(ValueError):
        build_constant_input_dim_deep_gp(X.astype(np.float32), 2, Config(10, 1e-2, 0.1))


def test_prediction_model_and_training_fit_decreases_loss():
    model, X, Y, _ = _build(D=2, M=30, n_layers=2, factor=1e-3)
    model = model.cuda()
    Xd, Yd = torch.from_numpy(X[:600]).cuda(), torch.from_numpy(Y[:600]).cuda()
    out = model.as_prediction_model()(Xd)
    assert out.y_mean is not None and tuple(out.f_mean.shape) == (600, 1)
    assert_close(out.y_var, out.f_var + 0.05, what="y_var = f_var + noise")
    tm = model.as_training_model()
    tm.compile(torch.optim.Adam(model.parameters(), lr=0.02))
    hist = tm.fit({"inputs": Xd, "targets": Yd}, epochs=30, batch_size=300)
    assert hist["loss"][-1] < hist["loss"][0]


def test_overlapped_factorization_matches_inline():
    """The factor part of every layer on side streams (DeepGP default) must give bit-identical
    loss and gradients to running everything in line on one stream."""
    model, X, Y, rng = _build(D=4, M=70, n_layers=3)
    model = model.cuda()
    N = 300
    Xb, Yb = torch.from_numpy(X[:N]).cuda(), torch.from_numpy(Y[:N]).cuda()
    eps = [torch.from_numpy(rng.standard_normal((N, 4))).cuda() for _ in range(2)] + [None]
    out = {}
    for overlap in (True, False):
        model.overlap_factorization = overlap
        for p in model.parameters():
            p.grad = None
        loss = model.loss(Xb, Yb, eps=eps)
        loss.backward()
        torch.cuda.synchronize()
        out[overlap] = (loss.detach().clone(), [p.grad.clone() for p in model.parameters() if p.grad is not None])
    assert torch.equal(out[True][0], out[False][0])
    assert len(out[True][1]) == len(out[False][1])
    for a, b in zip(out[True][1], out[False][1]):
        assert_close(a, b, tol=1e-13, what="grad overlap vs inline")


def test_nonwhitened_deep_gp_split_path_matches_oracle():
    """whiten=False through DeepGP (layer 0 fused node, layers >= 1 factor/cond nodes on side streams):
    loss, ELBO gradients w.r.t. q_mu / q_sqrt / Z against the oracle."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.helpers import construct_basic_inducing_variables, construct_basic_kernel
    from gpflux_b200.layers import GPLayer, LikelihoodLayer
    from gpflux_b200.models import DeepGP
    from oracle import svgp_torch as ot
    from oracle.from_model import noise_variance_from_model, specs_from_model

    rng = np.random.default_rng(31)
    D, M, N = 3, 40, 250
    X = rng.standard_normal((600, D))
    Y = np.sin(X.sum(1, keepdims=True))
    layers = []
    for i, P in enumerate((D, D, 1)):
        kern = construct_basic_kernel(gf.SquaredExponential(lengthscales=np.full(D, 1.5), variance=0.5 + 0.3 * i), output_dim=P,
                                      share_hyperparams=True)
        iv = construct_basic_inducing_variables(M, D, share_variables=True, output_dim=P, z_init=rng.standard_normal((M, D)))
        layer = GPLayer(kern, iv, 600, mean_function=gf.Identity() if P == D else gf.Zero(), whiten=False)
        layer.q_mu.assign(0.2 * rng.standard_normal((M, P)))
        layer.q_sqrt.assign(np.stack([0.3 * np.eye(M) + 0.02 * np.tril(rng.standard_normal((M, M)), -1) for _ in range(P)]))
        layers.append(layer)
    model = DeepGP(layers, LikelihoodLayer(gf.Gaussian(0.05)))
    specs, noise = specs_from_model(model), noise_variance_from_model(model)
    assert all(s["whiten"] is False for s in specs)
    model = model.cuda()
    Xb, Yb = torch.from_numpy(X[:N]), torch.from_numpy(Y[:N])
    eps = [torch.from_numpy(rng.standard_normal((N, D))) for _ in range(2)] + [None]
    loss = model.loss(Xb.cuda(), Yb.cuda(), eps=[e.cuda() if e is not None else None for e in eps])
    loss.backward()
    o_specs = [ot.spec_requires_grad(s) for s in specs]
    o_loss, _ = ot.deep_gp_loss(o_specs, Xb, Yb, eps, noise.clone(), 600)
    o_loss.backward()
    assert_close(loss, o_loss, what="loss")
    for i, layer in enumerate(model.f_layers):
        assert_close(layer.q_mu.unconstrained_variable.grad, o_specs[i]["q_mu"].grad, what=f"L{i} dq_mu")
        assert_close(layer.q_sqrt.unconstrained_variable.grad, o_specs[i]["q_sqrt"].grad, what=f"L{i} dq_sqrt")
        Zp = layer.inducing_variable.inducing_variable.Z.unconstrained_variable
        assert_close(Zp.grad, o_specs[i]["Z"].grad[0], what=f"L{i} dZ")
        assert_close(layer.prior_kl(), ot.prior_kl(specs[i]), what=f"L{i} prior_kl")


def test_factor_then_cond_equals_fused_layer():
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot
    from _util import spec_to_cuda

    rng = np.random.default_rng(3)
    spec = spec_to_cuda(ot.make_spec(rng, 50, 3, 2, K=2))
    X = torch.from_numpy(rng.standard_normal((120, 3))).cuda()
    eps = torch.from_numpy(rng.standard_normal((120, 2))).cuda()
    args = (spec["Z"], spec["lengthscales"], spec["variance"], spec["q_mu"], spec["q_sqrt"])
    f0, m0, v0, kl0 = ops.svgp_layer(X, *args, eps=eps)
    factor = ops.svgp_factor(*args)
    f1, m1, v1 = ops.svgp_cond(X, *args, factor, eps=eps)
    assert torch.equal(f0, f1) and torch.equal(m0, m1) and torch.equal(v0, v1) and torch.equal(kl0, factor[2])


def test_likelihood_layer_and_likelihood_loss_give_equal_results():
    """Mirror of reference tests/gpflux/test_losses.py:11-32, plus the tensor branch (-log p(y | f))."""
    import math

    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import LikelihoodLayer
    from gpflux_b200.layers.gp_layer import MultivariateNormalDiag
    from gpflux_b200.losses import LikelihoodLoss

    rng = np.random.default_rng(123)
    f_mean = torch.from_numpy(rng.standard_normal((7, 1))).cuda()
    f_var = torch.from_numpy(rng.standard_normal((7, 1)) ** 4).cuda()
    targets = torch.from_numpy(rng.standard_normal((7, 1))).cuda()
    f_dist = MultivariateNormalDiag(f_mean, f_var)
    likelihood = gf.Gaussian(0.123).cuda()
    layer = LikelihoodLayer(likelihood)
    layer(f_dist, targets=targets, training=True)
    [layer_loss] = layer.losses
    loss_loss = LikelihoodLoss(likelihood)(targets, f_dist)
    assert_close(loss_loss, layer_loss, tol=1e-15, what="LikelihoodLoss vs LikelihoodLayer")
    ref = 0.5 * math.log(2 * math.pi * 0.123) + 0.5 * ((targets - f_mean) ** 2 + f_var) / 0.123
    assert_close(loss_loss, ref.mean(), tol=1e-12, what="closed form")
    sample_loss = LikelihoodLoss(likelihood)(targets, f_mean)
    ref = 0.5 * math.log(2 * math.pi * 0.123) + 0.5 * (targets - f_mean) ** 2 / 0.123
    assert_close(sample_loss, ref.mean(), tol=1e-12, what="-log p(y | f)")


def test_cuda_graph_replay_matches_eager():
    """The whole step (both layers, the likelihood, every backward kernel, the library's side strand
    and DeepGP's factor side streams) captured in ONE CUDA graph replays to the eager result - the
    entry points make no hidden synchronisation and re-join every stream they fork."""
    model, X, Y, rng = _build(D=4, M=64, n_layers=2, seed=7)
    model = model.cuda()
    N = 512
    sX = torch.from_numpy(X[:N]).cuda()
    sY = torch.from_numpy(Y[:N]).cuda()
    eps = torch.from_numpy(rng.standard_normal((N, 4))).cuda()
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        loss = model.loss(sX, sY, eps=[eps, None])
        loss.backward()
        return loss

    # everything runs on a non-default stream from the start (autograd nodes remember the stream they
    # first ran on, and a captured graph must not depend on the legacy stream)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
        loss_eager = step().detach().clone()
        grads_eager = [p.grad.detach().clone() for p in params]
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss_static = step()
    grads_static = [p.grad for p in params]
    for t in grads_static:
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert_close(loss_static, loss_eager, tol=1e-13, what="graph loss")
    for a, b in zip(grads_static, grads_eager):
        assert_close(a, b, tol=1e-12, what="graph grads")


def test_mixed_architecture_deep_gp_matches_oracle():
    """A hand-built 3-layer DGP that exercises the combinations the builder never produces: separate
    Matern-3/2 kernels with separate inducing points + identity mean, a shared Matern-5/2 layer with
    P != D and a zero mean, a non-whitened squared-exponential output layer, num_data passed at the
    model level."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.helpers import construct_basic_inducing_variables, construct_basic_kernel
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.models import DeepGP
    from oracle import svgp_torch as ot
    from oracle.from_model import noise_variance_from_model, specs_from_model

    rng = np.random.default_rng(77)
    Ndata, N, D = 900, 260, 3
    X = rng.standard_normal((Ndata, D))
    Y = np.cos(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((Ndata, 1))
    k1 = construct_basic_kernel([gf.Matern32(lengthscales=rng.uniform(1.0, 2.0, D), variance=0.3) for _ in range(D)])
    iv1 = construct_basic_inducing_variables([33] * D, D, z_init=[rng.standard_normal((33, D)) for _ in range(D)])
    l1 = GPLayer(k1, iv1, Ndata, mean_function=gf.Identity())
    k2 = construct_basic_kernel(gf.Matern52(lengthscales=np.full(D, 1.5)), output_dim=2, share_hyperparams=True)
    iv2 = construct_basic_inducing_variables(40, D, output_dim=2, share_variables=True, z_init=rng.standard_normal((40, D)))
    l2 = GPLayer(k2, iv2, Ndata, mean_function=gf.Zero())
    k3 = construct_basic_kernel(gf.SquaredExponential(lengthscales=np.full(2, 1.2)), output_dim=1, share_hyperparams=True)
    iv3 = construct_basic_inducing_variables(25, 2, output_dim=1, share_variables=True, z_init=rng.standard_normal((25, 2)))
    l3 = GPLayer(k3, iv3, Ndata, mean_function=gf.Zero(), whiten=False)
    for layer in (l1, l2, l3):
        M, P = layer.q_mu.shape
        layer.q_mu.assign(0.2 * rng.standard_normal((M, P)))
        layer.q_sqrt.assign(np.stack([0.4 * np.eye(M) + 0.03 * np.tril(rng.standard_normal((M, M)), -1) for _ in range(P)]))
    model = DeepGP([l1, l2, l3], gf.Gaussian(0.05), num_data=Ndata)
    specs, noise = specs_from_model(model), noise_variance_from_model(model)
    model = model.cuda()
    Xb, Yb = torch.from_numpy(X[:N]), torch.from_numpy(Y[:N])
    eps = [torch.from_numpy(rng.standard_normal((N, D))), torch.from_numpy(rng.standard_normal((N, 2))), None]
    loss = model.loss(Xb.cuda(), Yb.cuda(), eps=[e.cuda() if e is not None else None for e in eps])
    loss.backward()
    o_specs = [ot.spec_requires_grad(s) for s in specs]
    o_noise = noise.clone().requires_grad_(True)
    o_loss, _ = ot.deep_gp_loss(o_specs, Xb, Yb, eps, o_noise, Ndata)
    o_loss.backward()
    assert_close(loss, o_loss, what="loss")
    for i, layer in enumerate(model.f_layers):
        tol = 1e-9 if layer.whiten else 1e-7  # the non-whitened output layer divides by Kuu twice
        assert_close(layer.q_mu.unconstrained_variable.grad, o_specs[i]["q_mu"].grad, tol=tol, what=f"L{i} dq_mu")
        assert_close(layer.q_sqrt.unconstrained_variable.grad, o_specs[i]["q_sqrt"].grad, tol=tol, what=f"L{i} dq_sqrt")
    ivs = model.f_layers[0].inducing_variable.inducing_variables
    for k, iv in enumerate(ivs):
        assert_close(iv.Z.unconstrained_variable.grad, o_specs[0]["Z"].grad[k], what=f"L0 dZ[{k}]")


def _oracle_training_run(model_cpu_state, specs, u_noise0, batches, num_data, lr, steps):
    """Adam on the oracle in the SAME unconstrained variables the GPU model trains (softplus for the
    positive parameters, tril for q_sqrt): returns the loss trajectory and the final leaves."""
    from oracle import svgp_torch as ot

    leaves = []
    for st in model_cpu_state:
        leaves.append({k: v.clone().requires_grad_(True) for k, v in st.items()})
    u_noise = u_noise0.clone().requires_grad_(True)
    params = [v for lv in leaves for v in lv.values()] + [u_noise]
    opt = torch.optim.Adam(params, lr=lr)
    sp = torch.nn.functional.softplus
    losses = []
    for t in range(steps):
        Xb, Yb, eps = batches[t]
        cur = []
        for spec, lv in zip(specs, leaves):
            s = dict(spec)
            s["Z"] = lv["Z"].unsqueeze(0)
            s["lengthscales"] = sp(lv["u_ls"]).reshape(1, -1)
            s["variance"] = sp(lv["u_var"]).reshape(1)
            s["q_mu"] = lv["q_mu"]
            s["q_sqrt"] = torch.tril(lv["q_sqrt"])
            cur.append(s)
        opt.zero_grad()
        loss, _ = ot.deep_gp_loss(cur, Xb, Yb, eps, sp(u_noise), num_data)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses, leaves, u_noise


@pytest.mark.parametrize("case", ["snelson_svgp", "two_layer_minibatch"])
def test_adam_training_trajectory_matches_oracle(case):
    """The training-loop equivalence the reference pins with live GPflow
    (tests/integration/test_svgp_equivalence.py:265-283: 20 optimiser steps, rtol 1e-7): here the other
    side is the CPU oracle trained with the same optimiser in the same unconstrained variables.  Every
    step's loss and the final parameters must agree, so gradient errors would have to stay below the
    tolerance after being fed back 20 times."""
    import os

    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.models import DeepGP
    from oracle.from_model import specs_from_model

    steps, lr = 20, 0.01
    rng = np.random.default_rng(5)
    if case == "snelson_svgp":
        d = np.load(os.path.join(os.path.dirname(__file__), "golden", "snelson1d.npz"))
        X, Y = d["X"], d["Y"]
        kernel = gf.SharedIndependent(gf.SquaredExponential(variance=0.7, lengthscales=0.6), 1)
        iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.linspace(0, 6, 20)[:, None]))
        model = DeepGP([GPLayer(kernel, iv, num_data=200, mean_function=gf.Zero())], gf.Gaussian(0.08))
        batches = [(torch.from_numpy(X), torch.from_numpy(Y), [None])] * steps  # full batch, as the reference test
    else:
        model, X, Y, _ = _build(D=5, M=25, n_layers=2, seed=3)
        batches = []
        for t in range(steps):
            idx = rng.choice(X.shape[0], 100, replace=False)
            batches.append((torch.from_numpy(X[idx]), torch.from_numpy(Y[idx]), [torch.from_numpy(rng.standard_normal((100, 5))), None]))

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
    # ARD lengthscales are one unconstrained value per input dimension; a scalar one is broadcast by the kernel
    D = specs[0]["Z"].shape[-1]
    for st in state:
        assert st["u_ls"].numel() in (1, D)
    o_losses, o_leaves, o_noise = _oracle_training_run(state, specs, u_noise0, batches, model.num_data, lr, steps)

    model = model.cuda()
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    g_losses = []
    for Xb, Yb, eps in batches:
        opt.zero_grad()
        loss = model.loss(Xb.cuda(), Yb.cuda(), eps=[e.cuda() if e is not None else None for e in eps])
        loss.backward()
        opt.step()
        g_losses.append(float(loss.detach()))

    assert o_losses[-1] < o_losses[0]  # it did train
    np.testing.assert_allclose(g_losses, o_losses, rtol=1e-7, atol=0)  # the reference's own tolerance
    print(f"{case}: max relative loss deviation over {steps} steps = "
          f"{np.max(np.abs(np.array(g_losses) / np.array(o_losses) - 1.0)):.3g}; loss {o_losses[0]:.6f} -> {o_losses[-1]:.6f}")
    for layer, lv in zip(model.f_layers, o_leaves):
        kern = layer.kernel.kernel
        assert_close(layer.q_mu.unconstrained_variable, lv["q_mu"], tol=1e-6, what="q_mu after training")
        assert_close(torch.tril(layer.q_sqrt.unconstrained_variable), torch.tril(lv["q_sqrt"]), tol=1e-6, what="q_sqrt after training")
        assert_close(layer.inducing_variable.inducing_variable.Z.unconstrained_variable, lv["Z"], tol=1e-6, what="Z after training")
        assert_close(kern.lengthscales.unconstrained_variable.reshape(-1), lv["u_ls"], tol=1e-6, what="lengthscales after training")
        assert_close(kern.variance.unconstrained_variable.reshape(()), lv["u_var"], tol=1e-6, what="variance after training")
    assert_close(model.likelihood_layer.likelihood.variance.unconstrained_variable.reshape(()), o_noise.reshape(()), tol=1e-6, what="noise after training")


def test_chunked_loss_and_gradients_equal_the_whole_minibatch():
    """DeepGP.loss_and_grad_chunked: row chunks (ragged last chunk) reproduce loss and every gradient of the
    one-piece evaluation - what lets a 65536-row minibatch of config 3 run on one GPU."""
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp

    rng = np.random.default_rng(12)
    N, D, M = 700, 3, 32
    X = rng.standard_normal((1500, D))
    Y = np.cos(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((1500, 1))
    model = build_constant_input_dim_deep_gp(X, 2, Config(M, 1e-2, 1e-2), z_init=X[:M])
    for layer in model.f_layers:
        layer.q_mu.assign(0.1 * rng.standard_normal(tuple(layer.q_mu.shape)))
    model = model.to("cuda")
    Xb, Yb = torch.from_numpy(X[:N]).cuda(), torch.from_numpy(Y[:N]).cuda()
    eps = [torch.from_numpy(rng.standard_normal((N, D))).cuda(), None]

    loss = model.loss(Xb, Yb, eps=eps)
    loss.backward()
    want = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    got = model.loss_and_grad_chunked(Xb, Yb, rows_per_chunk=256, eps=eps)
    assert abs(float(got) - float(loss)) <= 1e-12 * abs(float(loss))
    assert set(want) == {n for n, p in model.named_parameters() if p.grad is not None}
    for n, p in model.named_parameters():
        if p.grad is not None:
            scale = max(float(want[n].abs().max()), 1e-300)
            assert float((p.grad - want[n]).abs().max()) <= 1e-11 * scale, n
