"""CPU tests: the oracle against the committed golden vectors, its two independent restatements
against each other, the SURVEY §8c anchors and the invariants the reference's own tests assert
(SURVEY §4).  No GPU needed."""
import math
import os

import numpy as np
import pytest
import torch

from _util import assert_close
from oracle import svgp_numpy as on
from oracle import svgp_torch as ot

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, "layer_cases.npz"))
    N, M, D, P, K = [int(v) for v in z[f"{name}/meta"]]
    spec = {k: torch.from_numpy(z[f"{name}/in/{k}"]) for k in ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")}
    spec["kernel"] = str(z[f"{name}/kernel"])
    spec["mean"] = str(z[f"{name}/mean"])
    spec["whiten"] = True
    spec["jitter"] = 1e-6
    if spec["mean"] == "linear":
        spec["mean_A"] = torch.from_numpy(z[f"{name}/in/mean_A"])
        spec["mean_b"] = torch.from_numpy(z[f"{name}/in/mean_b"])
    return z, spec


CASE_NAMES = ["se_shared", "se_separate_identity", "m32_shared_linear", "m52_separate", "m12_shared"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_golden(name):
    z, spec = load_case(name)
    X = torch.from_numpy(z[f"{name}/in/X"])
    eps = torch.from_numpy(z[f"{name}/in/eps"])
    w = torch.from_numpy(z[f"{name}/in/w"])
    s = ot.spec_requires_grad(spec)
    Xr = X.clone().requires_grad_(True)
    f, m, v = ot.layer_forward(s, Xr, eps)
    kl = ot.prior_kl(s)
    ((f * w[0]).sum() + (m * w[1]).sum() + (v * w[2]).sum() + 0.37 * kl).backward()
    assert_close(f, z[f"{name}/out/fsample"], tol=1e-12, what="fsample")
    assert_close(m, z[f"{name}/out/fmean"], tol=1e-12, what="fmean")
    assert_close(v, z[f"{name}/out/fvar"], tol=1e-12, what="fvar")
    assert_close(kl, z[f"{name}/out/kl"], tol=1e-12, what="kl")
    for k in ("Z", "lengthscales", "variance", "q_mu", "q_sqrt"):
        assert_close(s[k].grad, z[f"{name}/grad/{k}"], tol=1e-11, what=f"grad {k}")
    assert_close(Xr.grad, z[f"{name}/grad/X"], tol=1e-11, what="grad X")


@pytest.mark.parametrize("name", CASE_NAMES)
def test_numpy_and_torch_restatements_agree(name):
    z, spec = load_case(name)
    X = torch.from_numpy(z[f"{name}/in/X"])
    m_t, v_t = ot.predict(spec, X)
    m_n, v_n = on.predict(spec, X)
    assert_close(m_n, m_t, scale=1.0, tol=1e-11, what="mean")
    assert_close(v_n, v_t, scale=1.0, tol=1e-11, what="var")
    assert_close(np.asarray(on.prior_kl(spec)), ot.prior_kl(spec), tol=1e-11, what="kl (dense closed form vs gauss_kl)")


def _snelson_spec(M=20):
    return dict(
        kernel="se",
        Z=torch.linspace(0, 6, M, dtype=torch.float64)[None, :, None],
        lengthscales=torch.tensor([[0.6]], dtype=torch.float64),
        variance=torch.tensor([0.7], dtype=torch.float64),
        q_mu=torch.zeros(M, 1, dtype=torch.float64),
        q_sqrt=torch.eye(M, dtype=torch.float64)[None],
        mean="zero", whiten=True, jitter=1e-6,
    )


def test_survey_anchors_on_snelson():
    """numbers computed in the survey session with an independent numpy script."""
    d = np.load(os.path.join(GOLDEN, "snelson1d.npz"))
    X, Y = torch.from_numpy(d["X"]), torch.from_numpy(d["Y"])
    assert tuple(X.shape) == (200, 1) and tuple(Y.shape) == (200, 1)
    nv = torch.tensor(0.08, dtype=torch.float64)
    for M in (20, 50):
        spec = _snelson_spec(M)
        m, v = ot.predict(spec, X)
        assert float(m.abs().max()) == 0.0
        assert_close(v, torch.full_like(v, 0.7), tol=1e-9, what="fvar == Knn at q_sqrt = I")
        ve = ot.gaussian_variational_expectations(m, v, Y, nv).sum()
        assert_close(ve, torch.tensor(-1840.588157487725, dtype=torch.float64), tol=1e-11, what="sum VE at init")
        assert float(ot.prior_kl(spec)) == 0.0
    spec = _snelson_spec(20)
    spec["q_mu"] = torch.sin(torch.arange(20, dtype=torch.float64))[:, None]
    spec["q_sqrt"] = (0.1 * torch.tril(torch.cos(torch.arange(400, dtype=torch.float64)).reshape(20, 20)) + torch.eye(20, dtype=torch.float64))[None]
    m, v = ot.predict(spec, X)
    ve = ot.gaussian_variational_expectations(m, v, Y, nv).sum()
    kl = ot.prior_kl(spec)
    assert_close(ve, torch.tensor(-2861.4394568303373, dtype=torch.float64), tol=1e-11, what="sum VE")
    assert_close(kl, torch.tensor(5.312033758504997, dtype=torch.float64), tol=1e-12, what="KL")
    assert_close(ve - kl, torch.tensor(-2866.7514905888424, dtype=torch.float64), tol=1e-11, what="ELBO")


def test_lognormal_prior_anchor():
    """log-prior LogNormal(1, 0.5) at lengthscale 0.6 (test_svgp_equivalence.py:64-66)."""
    x, mu, sigma = 0.6, 1.0, 0.5
    lp = -math.log(x * sigma * math.sqrt(2 * math.pi)) - (math.log(x) - mu) ** 2 / (2 * sigma**2)
    assert abs(lp - (-4.2801538597345266)) < 1e-13


def test_kl_invariants():
    """reference tests/gpflux/layers/test_gp_layer.py:67-85."""
    rng = np.random.default_rng(0)
    spec = ot.make_spec(rng, 30, 2, 3)
    spec["q_mu"] = torch.zeros_like(spec["q_mu"])
    spec["q_sqrt"] = torch.eye(30, dtype=torch.float64).repeat(3, 1, 1)
    assert float(ot.prior_kl(spec)) == 0.0
    spec["q_mu"] = torch.ones_like(spec["q_mu"])
    assert float(ot.prior_kl(spec)) > 0.0


def test_nonwhite_kl_matches_dense_closed_form():
    rng = np.random.default_rng(1)
    spec = ot.make_spec(rng, 25, 2, 2, K=2, whiten=False)
    assert_close(np.asarray(on.prior_kl(spec)), ot.prior_kl(spec), tol=1e-9, what="non-white KL")


def test_compute_A_inv_b():
    """reference tests/gpflux/test_math.py:30-38 (decimal=3 there)."""
    rng = np.random.default_rng(2)
    X = torch.from_numpy(rng.standard_normal((100, 2)))
    A = ot.kernel_matrix("se", X, X, torch.ones(2, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64)) + 1e-6 * torch.eye(100, dtype=torch.float64)
    b = torch.from_numpy(rng.standard_normal((100, 3)))
    np.testing.assert_array_almost_equal((A @ ot.compute_A_inv_b(A, b)).numpy(), b.numpy(), decimal=3)


@pytest.mark.parametrize("kind", ["se", "matern12", "matern32", "matern52"])
def test_rff_inner_product_approximates_kernel(kind):
    """reference tests/.../fourier_features/test_random.py:116-138: Phi(x) Phi(y)^T ~ k(x, y), atol 5e-2."""
    rng = np.random.default_rng(3)
    D, n = 3, 40000
    ls = torch.from_numpy(rng.uniform(0.5, 1.5, D))
    var = torch.tensor(1.3, dtype=torch.float64)
    if kind == "se":
        W = torch.from_numpy(rng.standard_normal((n, D)))
    else:
        p = {"matern12": 0, "matern32": 1, "matern52": 2}[kind]
        nu = 2.0 * p + 1.0
        W = torch.from_numpy(rng.standard_normal((n, D)) / np.sqrt(rng.gamma(nu / 2, 2.0 / nu, (n, 1))))
    b = torch.from_numpy(rng.uniform(0, 2 * math.pi, (1, n)))
    X = torch.from_numpy(rng.standard_normal((20, D)))
    Yp = torch.from_numpy(rng.standard_normal((15, D)))
    approx = ot.rff_cosine(X, W, b, ls, var) @ ot.rff_cosine(Yp, W, b, ls, var).T
    exact = ot.kernel_matrix(kind, X, Yp, ls, var)
    assert float((approx - exact).abs().max()) < 5e-2
    approx2 = ot.rff_concat(X, W[: n // 2], ls, var) @ ot.rff_concat(Yp, W[: n // 2], ls, var).T
    assert float((approx2 - exact).abs().max()) < 5e-2


@pytest.mark.parametrize("whiten", [True, False])
def test_wilson_sample_is_a_consistent_function(whiten):
    """reference tests/gpflux/sampling/test_sample.py:78-100: same X twice -> identical values; and
    the sample interpolates towards the u-draw at Z."""
    rng = np.random.default_rng(4)
    M, D, P, L = 20, 2, 2, 100
    spec = ot.make_spec(rng, M, D, P, whiten=whiten)
    W = torch.from_numpy(rng.standard_normal((L, D)))
    b = torch.from_numpy(rng.uniform(0, 2 * math.pi, (1, L)))
    coeffs = torch.ones(L, 1, dtype=torch.float64)
    eps_w = torch.from_numpy(rng.standard_normal((L, P)))
    eps_u = torch.from_numpy(rng.standard_normal((P, M, 1)))
    pw, v = ot.wilson_setup(spec, W, b, coeffs, eps_w, eps_u)
    X = torch.from_numpy(rng.standard_normal((50, D)))
    f1 = ot.wilson_eval(spec, W, b, pw, v, X)
    f2 = ot.wilson_eval(spec, W, b, pw, v, X)
    np.testing.assert_array_almost_equal(f1.numpy(), f2.numpy(), decimal=7)
    assert tuple(f1.shape) == (50, P)


def test_deep_gp_loss_numpy_vs_torch():
    rng = np.random.default_rng(5)
    specs = [ot.make_spec(rng, 20, 3, 3, mean="identity"), ot.make_spec(rng, 20, 3, 1)]
    X = torch.from_numpy(rng.standard_normal((64, 3)))
    Y = torch.from_numpy(rng.standard_normal((64, 1)))
    eps = [torch.from_numpy(rng.standard_normal((64, 3))), None]
    nv = torch.tensor(0.1, dtype=torch.float64)
    lt, _ = ot.deep_gp_loss(specs, X, Y, eps, nv, 1000)
    ln = on.deep_gp_loss(specs, X, Y, eps, 0.1, 1000)
    assert_close(np.asarray(ln), lt, tol=1e-11, what="deep gp loss")


def test_oracle_matches_golden_next_cases():
    """tests/golden/next_cases.npz (natural gradient, full-covariance conditional, conditional-Gaussian
    sampler, latent layer) regenerates bit-for-bit-close from the oracle."""
    import os

    from oracle import natgrad_torch as ng

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_cases.npz"))
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    mu1, sq1 = ng.natgrad_step(t("natgrad/q_mu"), t("natgrad/q_sqrt"), t("natgrad/dq_mu"), t("natgrad/dq_sqrt"), float(g["natgrad/gamma"]))
    np.testing.assert_allclose(mu1.numpy(), g["natgrad/q_mu_new"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(sq1.numpy(), g["natgrad/q_sqrt_new"], rtol=1e-12, atol=1e-13)
    spec = {k: t(f"fullcov/{k}") for k in ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")}
    spec.update(kernel="se", mean="zero", whiten=True, jitter=1e-6)
    m, c = ot.conditional_full_cov(spec, t("fullcov/X"))
    np.testing.assert_allclose(m.numpy(), g["fullcov/mean"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(c.numpy(), g["fullcov/cov"], rtol=1e-11, atol=1e-12)
    o, kl = ot.latent_layer(t("latent/X"), t("latent/mean"), t("latent/std"), t("latent/prior_mean"), t("latent/prior_std"), t("latent/eps"))
    np.testing.assert_allclose(o.numpy(), g["latent/out"], rtol=1e-13)
    np.testing.assert_allclose(kl.numpy(), g["latent/kl"], rtol=1e-13)
