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


def test_fp32_uses_the_tf32_tensor_core_path_and_stays_within_1e4(fp32):
    """Shapes large enough for the 3xTF32 tcgen05 products (M, N >= 128): the opt-in switches the library's
    contraction mode, outputs and every gradient stay within 1e-4 of the float64 oracle."""
    from gpflux_b200 import _cabi, ops
    from oracle import svgp_torch as ot

    lib = _cabi.load()
    rng = np.random.default_rng(9)
    N, M, D, P = 1024, 128, 4, 4
    spec = ot.make_spec(rng, M, D, P, K=P, kernel="se", mean="identity", whiten=True)
    spec32 = {k: (v.float().double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in spec.items()}
    X = torch.from_numpy(rng.standard_normal((N, D))).float()
    eps = torch.from_numpy(rng.standard_normal((N, P))).float()
    w_f, w_m, w_v = (torch.from_numpy(rng.standard_normal((N, P))).float() for _ in range(3))

    names = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")
    so = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in spec32.items()}
    Xo = X.double().clone().requires_grad_(True)
    f_ref, m_ref, v_ref = ot.layer_forward(so, Xo, eps.double())
    kl_ref = ot.prior_kl(so)
    ((f_ref * w_f.double()).sum() + (m_ref * w_m.double()).sum() + (v_ref * w_v.double()).sum() + 0.37 * kl_ref).backward()

    p = {k: spec32[k].float().cuda().requires_grad_(True) for k in names}
    Xg = X.cuda().requires_grad_(True)
    f, m, v, kl = ops.svgp_layer(Xg, p["Z"], p["lengthscales"], p["variance"], p["q_mu"], p["q_sqrt"], mean_add=Xg,
                                 eps=eps.cuda(), kernel="se", whiten=True, jitter=spec["jitter"])
    assert lib.gpfx_get_contraction() == 1
    assert f.dtype == torch.float32
    ((f * w_f.cuda()).sum() + (m * w_m.cuda()).sum() + (v * w_v.cuda()).sum() + 0.37 * kl).backward()
    _close(m, m_ref, "fmean")
    _close(v, v_ref, "fvar")
    _close(f, f_ref, "fsample")
    _close(kl, kl_ref, "kl")
    for k in names:
        _close(p[k].grad, so[k].grad, f"grad {k}")
    _close(Xg.grad, Xo.grad, "grad X")


def test_contraction_mode_returns_to_fp64_with_the_default_float():
    from gpflux_b200 import _cabi, ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(10)
    spec = ot.make_spec(rng, 32, 2, 2)
    X = torch.from_numpy(rng.standard_normal((64, 2))).cuda()
    ops.svgp_layer(X, spec["Z"].cuda(), spec["lengthscales"].cuda(), spec["variance"].cuda(), spec["q_mu"].cuda(),
                   spec["q_sqrt"].cuda(), kernel="se", whiten=True, jitter=spec["jitter"])
    assert _cabi.load().gpfx_get_contraction() == 0
