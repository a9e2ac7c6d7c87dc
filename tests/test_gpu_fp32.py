"""The float32 opt-in (north star: "1e-4 for an opt-in fp32 path"; reference switch
gpflow.config.set_default_float(np.float32), read at gpflux/layers/gp_layer.py:163,213,220 and enforced on
inputs at gpflux/models/deep_gp.py:137-149): float32 parameters / inputs / outputs / gradients, float64
contraction on the DMMA pipe.  Parity tolerance 1e-4 relative, written here."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


@pytest.fixture
def fp32():
    from gpflux_b200 import gpflow_compat as gf

    gf.set_default_float(np.float32)
    try:
        yield
    finally:
        gf.set_default_float(np.float64)


def _close(a, b, what, tol=FP32_TOL):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(float(b.abs().max()), 1e-30)
    err = float((a - b).abs().max())
    assert err <= tol * scale, f"{what}: {err:.3e} vs scale {scale:.3e}"


def _build(rng, N=600, D=4, M=48, layers=2):
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp

    X = rng.standard_normal((2000, D))
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((2000, 1))
    model = build_constant_input_dim_deep_gp(X, layers, Config(M, 1e-2, 1e-2), z_init=X[:M])
    for layer in model.f_layers:
        layer.q_mu.assign(0.1 * rng.standard_normal(tuple(layer.q_mu.shape)))
    return model, X[:N], Y[:N]


def test_fp32_deep_gp_loss_and_gradients_within_1e4_of_the_float64_oracle(fp32):
    from oracle import svgp_torch as ot
    from oracle.from_model import noise_variance_from_model, specs_from_model

    rng = np.random.default_rng(3)
    model, X, Y = _build(rng)
    for p in model.parameters():
        assert p.dtype == torch.float32
    specs = [{k: (v.double() if isinstance(v, torch.Tensor) else v) for k, v in s.items()} for s in specs_from_model(model)]
    noise = noise_variance_from_model(model).double()
    model = model.to("cuda")
    eps = torch.from_numpy(rng.standard_normal((X.shape[0], X.shape[1])))

    Xb, Yb = torch.from_numpy(X).float().cuda(), torch.from_numpy(Y).float().cuda()
    loss = model.loss(Xb, Yb, eps=[eps.float().cuda(), None])
    assert loss.dtype == torch.float32
    loss.backward()

    # the oracle runs in float64 on the float32-rounded values the model actually holds
    o_specs = [ot.spec_requires_grad(s) for s in specs]
    o_noise = noise.clone().requires_grad_(True)
    o_loss, _ = ot.deep_gp_loss(o_specs, Xb.double().cpu(), Yb.double().cpu(), [eps.float().double(), None], o_noise,
                                model.num_data)
    o_loss.backward()
    _close(loss, o_loss, "loss")
    for i, layer in enumerate(model.f_layers):
        for name, par, ref in [("q_mu", layer.q_mu, o_specs[i]["q_mu"]), ("q_sqrt", layer.q_sqrt, o_specs[i]["q_sqrt"])]:
            g = par.unconstrained_variable.grad
            assert g.dtype == torch.float32
            _close(g, ref.grad, f"layer {i} d{name}")
        Zp = layer.inducing_variable.inducing_variable.Z.unconstrained_variable
        _close(Zp.grad, o_specs[i]["Z"].grad[0], f"layer {i} dZ")


def test_fp32_predict_dtypes_and_dtype_check(fp32):
    rng = np.random.default_rng(4)
    model, X, _ = _build(rng, layers=1)
    model = model.to("cuda")
    Xb = torch.from_numpy(X).float().cuda()
    m, v = model.predict_f(Xb)
    assert m.dtype == torch.float32 and v.dtype == torch.float32 and bool((v > 0).all())
    layer = model.f_layers[0]
    assert layer.prior_kl().dtype == torch.float32
    with pytest.raises(ValueError):  # float64 inputs are refused while the opt-in is on (deep_gp.py:137-149)
        model.predict_f(Xb.double())


def test_float64_is_the_default_again():
    from gpflux_b200 import gpflow_compat as gf

    assert gf.default_float() is torch.float64
    with pytest.raises(ValueError):
        gf.set_default_float(np.float16)
