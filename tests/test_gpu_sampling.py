"""GPU parity of the random-Fourier-feature basis and the fused Wilson-sample evaluation
 against the oracle, plus the invariants of the reference's tests."""
import math

import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu

KINDS = ["se", "matern12", "matern32", "matern52"]


@pytest.mark.parametrize("N,D,L,K", [(50, 1, 100, 1), (333, 3, 257, 1), (128, 8, 1000, 2), (65, 20, 64, 3)])
def test_rff_features_match_oracle(N, D, L, K):
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(10)
    X = torch.from_numpy(rng.standard_normal((N, D)))
    W = torch.from_numpy(rng.standard_normal((K, L, D)))
    b = torch.from_numpy(rng.uniform(0, 2 * math.pi, (K, L)))
    ls = torch.from_numpy(rng.uniform(0.5, 2.0, (K, D)))
    var = torch.from_numpy(rng.uniform(0.5, 2.0, (K,)))
    got = ops.rff_features(X.cuda(), W.cuda(), b.cuda(), ls.cuda(), var.cuda())
    ref = torch.stack([ot.rff_cosine(X, W[k], b[k][None], ls[k], var[k]) for k in range(K)])
    assert_close(got, ref, scale=float(ref.abs().max()), tol=1e-9, what="cosine features")
    got = ops.rff_features(X.cuda(), W.cuda(), None, ls.cuda(), var.cuda(), concat=True)
    ref = torch.stack([ot.rff_concat(X, W[k], ls[k], var[k]) for k in range(K)])
    assert tuple(got.shape) == (K, N, 2 * L)  # reference test_random.py:231-247
    assert_close(got, ref, scale=float(ref.abs().max()), tol=1e-9, what="sin/cos features")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("S,N,D,P,L,M,shared", [(1, 100, 1, 1, 100, 10, True), (3, 257, 2, 3, 300, 40, True),
                                                (2, 130, 8, 8, 500, 64, False), (2, 70, 12, 2, 129, 33, False)])
def test_wilson_eval_matches_oracle(kind, S, N, D, P, L, M, shared):
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(11)
    spec = ot.make_spec(rng, M, D, P, kernel=kind)
    X = torch.from_numpy(rng.standard_normal((N, D) if shared else (S, N, D)))
    W = torch.from_numpy(rng.standard_normal((L, D)))
    b = torch.from_numpy(rng.uniform(0, 2 * math.pi, (1, L)))
    w = torch.from_numpy(rng.standard_normal((S, L, P)))
    v = torch.from_numpy(rng.standard_normal((S, M, P)))
    got = ops.wilson_eval(kind, X.cuda(), W.cuda(), b.cuda(), spec["lengthscales"][0].cuda(), spec["variance"].cuda(),
                          spec["Z"][0].cuda(), w.cuda(), v.cuda())
    ref = torch.stack([ot.wilson_eval(spec, W, b, w[s], v[s], X if shared else X[s]) for s in range(S)])
    assert_close(got, ref, scale=float(ref.abs().max()), tol=1e-9 if kind != "matern12" else 1e-7, what="wilson eval")


@pytest.mark.parametrize("whiten", [True, False])
def test_efficient_sample_matches_oracle_and_is_consistent(whiten):
    """reference tests/gpflux/sampling/test_sample.py:78-100 (consistency to 7 decimals) + parity of the
    whole Matheron set-up with the same N(0,1) draws."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers.basis_functions.fourier_features import RandomFourierFeaturesCosine
    from gpflux_b200.sampling import KernelWithFeatureDecomposition, efficient_sample
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(12)
    M, D, P, L, N = 20, 2, 2, 100, 60
    spec = ot.make_spec(rng, M, D, P, whiten=whiten, kernel="matern52")
    base = gf.Matern52(variance=float(spec["variance"][0]), lengthscales=spec["lengthscales"][0].numpy())
    feats = RandomFourierFeaturesCosine(base, L, input_dim=D)
    kernel = KernelWithFeatureDecomposition(base, feats, np.ones((L, 1)))
    iv = gf.InducingPoints(spec["Z"][0].numpy())
    eps_w = torch.from_numpy(rng.standard_normal((1, L, P)))
    eps_u = torch.from_numpy(rng.standard_normal((1, P, M, 1)))
    sample = efficient_sample(iv, kernel.cuda(), spec["q_mu"].cuda(), q_sqrt=spec["q_sqrt"].cuda(), whiten=whiten,
                              eps_w=eps_w.cuda(), eps_u=eps_u.cuda())
    X = torch.from_numpy(rng.standard_normal((N, D)))
    f1 = sample(X.cuda())
    f2 = sample(X.cuda())
    np.testing.assert_array_almost_equal(f1.cpu().numpy(), f2.cpu().numpy(), decimal=7)
    spec["mean"] = "zero"
    Wc, bc = feats.W.cpu(), feats.b.cpu()
    pw, v = ot.wilson_setup(spec, Wc, bc, torch.ones(L, 1, dtype=torch.float64), eps_w[0], eps_u[0])
    ref = ot.wilson_eval(spec, Wc, bc, pw, v, X)
    assert_close(f1, ref, scale=float(ref.abs().max()), tol=1e-8, what="efficient_sample")


def test_layer_sample_and_sample_dgp_shapes():
    """gp_layer.py:370-384 / deep_gp.py:295-307: layer.sample() adds the mean function; sample_dgp chains."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.layers.basis_functions.fourier_features import RandomFourierFeaturesCosine
    from gpflux_b200.models import DeepGP, sample_dgp
    from gpflux_b200.sampling import KernelWithFeatureDecomposition

    rng = np.random.default_rng(13)
    D, M, L = 2, 16, 200
    layers = []
    for i in range(2):
        base = gf.SquaredExponential(lengthscales=np.ones(D))
        feats = RandomFourierFeaturesCosine(base, L, input_dim=D)
        kern = KernelWithFeatureDecomposition(base, feats, np.ones((L, 1)))
        with pytest.warns(UserWarning):
            layers.append(GPLayer(kern, gf.InducingPoints(rng.standard_normal((M, D))), 100, gf.Identity(), num_latent_gps=D))
    model = DeepGP(layers, gf.Gaussian(0.1)).cuda()
    X = torch.from_numpy(rng.standard_normal((77, D))).cuda()
    # the conditional path still works with the wrapped kernel
    mean, var = layers[0].predict(X)
    assert tuple(mean.shape) == (77, D)
    draw = sample_dgp(model)
    f1, f2 = draw(X), draw(X)
    assert tuple(f1.shape) == (77, D)
    assert torch.equal(f1, f2)


def test_rff_layer_approximates_kernel_and_rejects_unsupported():
    """reference test_random.py:116-138 (atol 5e-2) and :96-113 (AssertionError)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200 import ops
    from gpflux_b200.layers.basis_functions.fourier_features import RandomFourierFeatures, RandomFourierFeaturesCosine

    torch.manual_seed(0)
    rng = np.random.default_rng(14)
    D, n = 3, 40000
    X = torch.from_numpy(rng.standard_normal((20, D))).cuda()
    Y = torch.from_numpy(rng.standard_normal((15, D))).cuda()
    for kern in (gf.SquaredExponential(variance=1.3, lengthscales=rng.uniform(0.5, 1.5, D)),
                 gf.Matern32(variance=0.7, lengthscales=rng.uniform(0.5, 1.5, D))):
        for cls in (RandomFourierFeatures, RandomFourierFeaturesCosine):
            layer = cls(kern, n, input_dim=D)
            u, v = layer(X), layer(Y)
            approx = u @ v.T
            ls = kern.lengthscales.value().detach().cuda()[None]
            var = kern.variance.value().detach().cuda().reshape(1)
            exact = ops.kernel_matrix(kern.kind, X[None], Y, ls, var)[0]
            assert float((approx - exact).abs().max()) < 5e-2
    with pytest.raises(AssertionError, match="Unsupported Kernel"):
        RandomFourierFeatures(gf.Kernel(), 10)
    mo = gf.SharedIndependent(gf.SquaredExponential(lengthscales=np.ones(D)), 3)
    layer = RandomFourierFeaturesCosine(mo, 50, input_dim=D)
    assert tuple(layer(X).shape) == (3, 20, 50)


def test_orthogonal_and_quadrature_features():
    """reference test_random.py:163-183 (ORF, atol 5e-2) and test_quadrature.py:69-98 (QFF with
    n = 128 nodes reproduces the SE kernel for lengthscales >= 0.75, rtol 1e-7 there)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200 import ops
    from gpflux_b200.layers.basis_functions.fourier_features import OrthogonalRandomFeatures, QuadratureFourierFeatures

    torch.manual_seed(1)
    rng = np.random.default_rng(15)
    D = 3
    kern = gf.SquaredExponential(variance=1.7, lengthscales=rng.uniform(0.8, 1.5, D))
    X = torch.from_numpy(rng.standard_normal((20, D))).cuda()
    Y = torch.from_numpy(rng.standard_normal((15, D))).cuda()
    ls = kern.lengthscales.value().detach().cuda()[None]
    var = kern.variance.value().detach().cuda().reshape(1)
    exact = ops.kernel_matrix("se", X[None], Y, ls, var)[0]
    orf = OrthogonalRandomFeatures(kern, 20000, input_dim=D)
    assert float((orf(X) @ orf(Y).T - exact).abs().max()) < 5e-2
    with pytest.raises(AssertionError):
        OrthogonalRandomFeatures(gf.Matern32(), 10)
    # quadrature features: 1-D, 128 nodes
    k1 = gf.SquaredExponential(variance=0.9, lengthscales=[0.8])
    x1 = torch.from_numpy(rng.standard_normal((30, 1))).cuda()
    y1 = torch.from_numpy(rng.standard_normal((25, 1))).cuda()
    qff = QuadratureFourierFeatures(k1, 128, input_dim=1)
    u, v = qff(x1), qff(y1)
    assert tuple(u.shape) == (30, 256)
    ex1 = ops.kernel_matrix("se", x1[None], y1, k1.lengthscales.value().detach().cuda()[None],
                            k1.variance.value().detach().cuda().reshape(1))[0]
    assert_close(u @ v.T, ex1, scale=0.9, tol=1e-7, what="QFF inner product == SE kernel")


@pytest.mark.parametrize("whiten", [True, False])
def test_conditional_gaussian_sampler_matches_oracle_and_is_consistent(whiten):
    """The O(N^3) fallback sampler for plain kernels (reference sample.py:85-134): with pinned N(0,1)
    draws it reproduces the oracle's sequence of conditioned evaluations; re-evaluating at the same
    points returns the same values to 2 decimals (reference tests/gpflux/sampling/test_sample.py:53-75)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.sampling.sample import efficient_sample
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(31)
    M, D, P = 12, 1, 2
    kern = gf.SquaredExponential(variance=1.3, lengthscales=0.7)
    Z = rng.uniform(-1, 1, (M, D))
    iv = gf.InducingPoints(Z)
    Kzz = ot.kernel_matrix("se", torch.from_numpy(Z), torch.from_numpy(Z), torch.tensor([0.7], dtype=torch.float64), torch.tensor(1.3, dtype=torch.float64))
    Lz = torch.linalg.cholesky(Kzz + 1e-6 * torch.eye(M, dtype=torch.float64))
    q_mu = Lz @ torch.from_numpy(rng.standard_normal((M, P)))
    q_sqrt = (1e-3 * Lz)[None].repeat(P, 1, 1)
    spec = {"kernel": "se", "Z": torch.from_numpy(Z)[None], "lengthscales": torch.tensor([[0.7]], dtype=torch.float64),
            "variance": torch.tensor([1.3], dtype=torch.float64), "q_mu": q_mu, "q_sqrt": q_sqrt, "mean": "zero",
            "whiten": whiten, "jitter": 1e-6}
    Xs = [torch.from_numpy(rng.uniform(-1, 1, (n, D))) for n in (40, 25, 1)]
    epss = [torch.from_numpy(rng.standard_normal((P, x.shape[0]))) for x in Xs]
    ref = ot.conditional_sample_sequence(spec, Xs, epss)
    sample = efficient_sample(iv.cuda(), kern.cuda(), q_mu.cuda(), q_sqrt=q_sqrt.cuda(), whiten=whiten)
    for x, e, r in zip(Xs, epss, ref):
        got = sample(x.cuda(), eps=e.cuda())
        assert tuple(got.shape) == (x.shape[0], P)
        # chol of near-singular covariances (1e-6 jitter) on both sides: conditioning-limited
        assert_close(got, r, scale=1.0, tol=1e-6, what="conditional sample")
    X = torch.from_numpy(np.linspace(-1, 1, 100)[:, None]).cuda()
    sample2 = efficient_sample(iv.cuda(), kern.cuda(), q_mu.cuda(), q_sqrt=q_sqrt.cuda(), whiten=whiten)
    a, b = sample2(X), sample2(X)
    np.testing.assert_array_almost_equal(a.cpu().numpy(), b.cpu().numpy(), decimal=2)


def test_adding_samples_and_mean_function():
    """Reference tests/gpflux/sampling/test_sample.py:111-130."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.sampling.sample import Sample

    class SampleMock(Sample):
        def __init__(self, a):
            self.a = a

        def __call__(self, X):
            return self.a * X

    X = torch.randn(100, 2, dtype=torch.float64)
    s1, s2 = SampleMock(1.0), SampleMock(2.0)
    assert torch.allclose((s1 + s2)(X), s1(X) + s2(X))
    s3 = SampleMock(3.0) + gf.Identity()
    assert torch.allclose(s3(X), 3.0 * X + X)


@pytest.mark.parametrize("concat", [False, True])
def test_rff_backward_matches_oracle_autograd(concat):
    """gpfx_rff_backward: gradients of the feature map w.r.t. X, lengthscales, variance."""
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(77)
    N, D, L, K = 150, 5, 200, 2
    X = torch.from_numpy(rng.standard_normal((N, D)))
    W = torch.from_numpy(rng.standard_normal((K, L, D)))
    b = torch.from_numpy(rng.uniform(0, 2 * np.pi, (K, L)))
    ls = torch.from_numpy(rng.uniform(0.7, 1.8, (K, D)))
    var = torch.from_numpy(rng.uniform(0.5, 1.5, (K,)))
    G = torch.from_numpy(rng.standard_normal((K, N, 2 * L if concat else L)))

    Xo, lso, varo = X.clone().requires_grad_(True), ls.clone().requires_grad_(True), var.clone().requires_grad_(True)
    ref = torch.stack([ot.rff_concat(Xo, W[k], lso[k], varo[k]) if concat else ot.rff_cosine(Xo, W[k], b[k], lso[k], varo[k])
                       for k in range(K)])
    (ref * G).sum().backward()

    Xg, lsg, varg = (t.cuda().requires_grad_(True) for t in (X, ls, var))
    out = ops.rff_features(Xg, W.cuda(), None if concat else b.cuda(), lsg, varg, concat=concat)
    (out * G.cuda()).sum().backward()
    assert_close(out, ref, what="features")
    assert_close(Xg.grad, Xo.grad, what="dX")
    assert_close(lsg.grad, lso.grad, what="dlengthscales")
    assert_close(varg.grad, varo.grad, what="dvariance")
