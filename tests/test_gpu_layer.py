"""GPU parity of the fused GPLayer forward/backward (through the C-ABI) against the CPU oracle:
outputs and every gradient, on the same seeded inputs and the same eps."""
import numpy as np
import pytest
import torch

from _util import assert_close, cond_tol, spec_to_cuda

pytestmark = pytest.mark.gpu

PARAMS = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")


def _run_oracle(spec, X, eps, w_f, w_m, w_v, w_kl):
    from oracle import svgp_torch as ot

    s = ot.spec_requires_grad(spec)
    Xr = X.clone().requires_grad_(True)
    f, m, v = ot.layer_forward(s, Xr, eps)
    kl = ot.prior_kl(s)
    loss = (f * w_f).sum() + (m * w_m).sum() + (v * w_v).sum() + w_kl * kl
    loss.backward()
    grads = {k: s[k].grad for k in PARAMS}
    grads["X"] = Xr.grad
    # the oracle's q_sqrt gradient lives on the lower triangle only (tril in base_conditional)
    return f.detach(), m.detach(), v.detach(), kl.detach(), grads


def _run_gpu(spec, X, eps, w_f, w_m, w_v, w_kl, solver="gemm"):
    from gpflux_b200 import ops

    s = spec_to_cuda(spec)
    p = {k: s[k].clone().requires_grad_(True) for k in PARAMS}
    Xg = X.cuda().requires_grad_(True)
    mean_kind = spec.get("mean", "zero")
    if mean_kind == "identity":
        mean_add = Xg
    elif mean_kind == "linear":
        mean_add = Xg @ s["mean_A"] + s["mean_b"]
    else:
        mean_add = None
    f, m, v, kl = ops.svgp_layer(Xg, p["Z"], p["lengthscales"], p["variance"], p["q_mu"], p["q_sqrt"],
                                 mean_add=mean_add, eps=None if eps is None else eps.cuda(),
                                 kernel=spec["kernel"], whiten=spec.get("whiten", True), jitter=spec["jitter"],
                                 solver=solver)
    loss = (f * w_f.cuda()).sum() + (m * w_m.cuda()).sum() + (v * w_v.cuda()).sum() + w_kl * kl
    loss.backward()
    grads = {k: p[k].grad for k in PARAMS}
    grads["X"] = Xg.grad
    return f.detach(), m.detach(), v.detach(), kl.detach(), grads


CASES = [
    # N, M, D, P, K, kernel, mean
    (200, 20, 1, 1, 1, "se", "zero"),        # cfg-1 shape (snelson SVGP, reference test M=20)
    (200, 50, 1, 1, 1, "se", "zero"),        # cfg-1 shape (BASELINE M=50)
    (333, 70, 5, 5, 1, "se", "identity"),    # ragged sizes, shared kernel, identity mean
    (256, 64, 4, 3, 3, "se", "linear"),      # separate kernels, linear mean
    (1000, 256, 8, 8, 1, "se", "identity"),  # cfg-2 hidden layer shape (reduced N)
    (515, 130, 3, 2, 2, "matern32", "zero"),
    (300, 100, 2, 2, 1, "matern52", "zero"),
    (260, 65, 2, 2, 1, "matern12", "zero"),
    (2048, 300, 6, 4, 4, "se", "zero"),
    (300, 64, 20, 3, 3, "se", "zero"),       # D > 15: GEMM-form Kuf backward with R = -dK K / 2 from the saved Kuf
    (260, 70, 18, 2, 1, "matern32", "zero"), # D > 15, not SE: GEMM-form backward with the recomputed R pass
    (1, 8, 2, 1, 1, "se", "zero"),           # a single row
    (129, 1, 1, 1, 1, "se", "zero"),         # a single inducing point
    (64, 64, 40, 2, 2, "se", "zero"),        # D = 40 > one staging chunk of the kernel-matrix kernel
    (300, 100, 3, 32, 32, "se", "zero"),     # c3-like: 32 separate kernels
    (127, 63, 5, 5, 1, "matern52", "identity"),  # odd sizes everywhere
    (255, 65, 15, 3, 1, "se", "zero"),       # D = 15: widest fused kernel-matrix backward (W = 16)
    (1030, 129, 16, 2, 2, "se", "zero"),     # D = 16: first GEMM-form shape
]


@pytest.mark.parametrize("N,M,D,P,K,kernel,mean", CASES)
def test_layer_forward_backward_parity(N, M, D, P, K, kernel, mean):
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(100 + N + M)
    spec = ot.make_spec(rng, M, D, P, K=K, kernel=kernel, mean=mean)
    X = torch.from_numpy(rng.standard_normal((N, D)))
    eps = torch.from_numpy(rng.standard_normal((N, P)))
    w_f = torch.from_numpy(rng.standard_normal((N, P)))
    w_m = torch.from_numpy(rng.standard_normal((N, P)))
    w_v = torch.from_numpy(rng.standard_normal((N, P)))
    w_kl = 0.37
    ref = _run_oracle(spec, X, eps, w_f, w_m, w_v, w_kl)
    got = _run_gpu(spec, X, eps, w_f, w_m, w_v, w_kl)
    var_scale = float(spec["variance"].max())
    tol = cond_tol(spec, 1e-9 if kernel != "matern12" else 1e-7)
    assert_close(got[1], ref[1], scale=1.0, tol=tol, what="fmean")
    assert_close(got[2], ref[2], scale=var_scale, tol=tol, what="fvar")
    assert_close(got[0], ref[0], scale=1.0, tol=tol, what="fsample")
    assert_close(got[3], ref[3], tol=tol, what="kl")
    for k in PARAMS + ("X",):
        assert_close(got[4][k], ref[4][k], tol=tol, what=f"grad {k}")


NONWHITE_CASES = [
    (200, 20, 1, 1, 1, "se", "zero"),
    (333, 70, 5, 5, 1, "se", "identity"),
    (256, 64, 4, 3, 3, "se", "linear"),
    (515, 130, 3, 2, 2, "matern32", "zero"),
    (1000, 256, 8, 8, 1, "se", "identity"),
]


@pytest.mark.parametrize("N,M,D,P,K,kernel,mean", NONWHITE_CASES)
def test_nonwhitened_layer_parity(N, M, D, P, K, kernel, mean):
    """whiten=False (SURVEY §8a row a9'): q(u) = N(q_mu, q_sqrt q_sqrt^T), KL against N(0, Kuu)."""
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(300 + N + M)
    spec = ot.make_spec(rng, M, D, P, K=K, kernel=kernel, mean=mean, whiten=False)
    X = torch.from_numpy(rng.standard_normal((N, D)))
    eps = torch.from_numpy(rng.standard_normal((N, P)))
    w_f, w_m, w_v = (torch.from_numpy(rng.standard_normal((N, P))) for _ in range(3))
    ref = _run_oracle(spec, X, eps, w_f, w_m, w_v, 0.37)
    got = _run_gpu(spec, X, eps, w_f, w_m, w_v, 0.37)
    tol = cond_tol(spec, 1e-9)
    assert_close(got[1], ref[1], scale=1.0, tol=tol, what="fmean")
    assert_close(got[2], ref[2], scale=float(spec["variance"].max()), tol=tol, what="fvar")
    assert_close(got[0], ref[0], scale=1.0, tol=tol, what="fsample")
    assert_close(got[3], ref[3], tol=tol, what="kl")
    for k in PARAMS + ("X",):
        assert_close(got[4][k], ref[4][k], tol=tol, what=f"grad {k}")


def test_layer_no_eps_returns_mean():
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(7)
    spec = ot.make_spec(rng, 40, 2, 2)
    X = torch.from_numpy(rng.standard_normal((90, 2)))
    s = spec_to_cuda(spec)
    f, m, v, kl = ops.svgp_layer(X.cuda(), s["Z"], s["lengthscales"], s["variance"], s["q_mu"], s["q_sqrt"])
    assert torch.equal(f, m)
    mr, vr = ot.predict(spec, X)
    assert_close(m, mr, scale=1.0, what="mean")
    assert_close(v, vr, scale=1.0, what="var")


def test_layer_init_state_matches_reference_invariants():
    """reference tests/gpflux/layers/test_gp_layer.py:67-85 and SURVEY §8c anchors: at
    q_mu = 0, q_sqrt = I the whitened conditional is the prior: fmean == 0, fvar == Knn, KL == 0."""
    from gpflux_b200 import ops

    M, N = 20, 200
    Z = torch.linspace(0, 6, M, dtype=torch.float64)[None, :, None].cuda()
    X = torch.linspace(0.05, 5.9, N, dtype=torch.float64)[:, None].cuda()
    ls = torch.tensor([[0.6]], dtype=torch.float64).cuda()
    var = torch.tensor([0.7], dtype=torch.float64).cuda()
    q_mu = torch.zeros(M, 1, dtype=torch.float64).cuda()
    q_sqrt = torch.eye(M, dtype=torch.float64)[None].cuda()
    f, m, v, kl = ops.svgp_layer(X, Z, ls, var, q_mu, q_sqrt)
    assert float(kl.item()) == 0.0
    assert float(m.abs().max().item()) == 0.0
    assert_close(v, torch.full((N, 1), 0.7, dtype=torch.float64), scale=0.7, what="fvar == Knn")


def test_backward_twice_raises():
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(8)
    spec = spec_to_cuda(ot.make_spec(rng, 30, 2, 2))
    X = torch.from_numpy(rng.standard_normal((50, 2))).cuda()
    q_mu = spec["q_mu"].clone().requires_grad_(True)
    f, m, v, kl = ops.svgp_layer(X, spec["Z"], spec["lengthscales"], spec["variance"], q_mu, spec["q_sqrt"])
    m.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        m.sum().backward()


@pytest.mark.parametrize("name", ["se_shared", "se_separate_identity", "m32_shared_linear", "m52_separate", "m12_shared"])
def test_layer_matches_committed_golden_vectors(name):
    """CUDA path vs tests/golden/layer_cases.npz (generated by tests/golden/make_golden.py)."""
    from test_cpu_oracle import load_case

    z, spec = load_case(name)
    X = torch.from_numpy(z[f"{name}/in/X"])
    eps = torch.from_numpy(z[f"{name}/in/eps"])
    w = torch.from_numpy(z[f"{name}/in/w"])
    got = _run_gpu(spec, X, eps, w[0], w[1], w[2], 0.37)
    tol = cond_tol(spec, 1e-9 if spec["kernel"] != "matern12" else 1e-7)
    assert_close(got[0], z[f"{name}/out/fsample"], scale=1.0, tol=tol, what="fsample")
    assert_close(got[1], z[f"{name}/out/fmean"], scale=1.0, tol=tol, what="fmean")
    assert_close(got[2], z[f"{name}/out/fvar"], scale=1.0, tol=tol, what="fvar")
    assert_close(got[3], z[f"{name}/out/kl"], tol=tol, what="kl")
    for k in PARAMS + ("X",):
        assert_close(got[4][k], z[f"{name}/grad/{k}"], tol=tol, what=f"grad {k}")


@pytest.mark.parametrize("K,whiten", [(1, True), (3, True), (1, False)])
def test_full_cov_predict_and_sample(K, whiten):
    """predict(full_cov=True) -> [Q,N,N] against the oracle's base_conditional full_cov branch, its
    diagonal against the marginal path, the [N,Q,N,Q] / [N,Q,Q] layouts of the reference's table
    (gp_layer.py:238-246), and the correlated sample loc + chol(cov + jitter I) eps."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(11)
    N, M, D, P = 150, 30, 2, 3
    spec = ot.make_spec(rng, M, D, P, K=K, whiten=whiten, q_sqrt_diag=0.3)
    kerns = [gf.SquaredExponential(variance=float(spec["variance"][k]), lengthscales=spec["lengthscales"][k].numpy()) for k in range(K)]
    ivs = [gf.InducingPoints(spec["Z"][k].numpy()) for k in range(K)]
    if K == 1:
        kernel = gf.SharedIndependent(kerns[0], P)
        iv = gf.SharedIndependentInducingVariables(ivs[0])
    else:
        kernel = gf.SeparateIndependent(kerns)
        iv = gf.SeparateIndependentInducingVariables(ivs)
    layer = GPLayer(kernel, iv, num_data=N, num_latent_gps=P, mean_function=gf.Zero(), whiten=whiten, full_cov=True)
    layer.q_mu.assign(spec["q_mu"].numpy())
    layer.q_sqrt.assign(spec["q_sqrt"].numpy())
    layer = layer.cuda()
    X = torch.from_numpy(rng.standard_normal((N, D)))
    m_ref, c_ref = ot.conditional_full_cov(spec, X)
    mean, cov = layer.predict(X.cuda(), full_cov=True)
    assert tuple(cov.shape) == (P, N, N)
    # not whitened: L^-T L^-1 Kuf is only defined to cond(Kuu) * eps (DESIGN.md "Tolerance and conditioning")
    tol = 1e-9 if whiten else cond_tol(spec)
    assert_close(mean, m_ref, tol=tol, what="full_cov mean")
    # (not whitened, S = q_sqrt q_sqrt^T is in u-space: A^T Kuu^-1 S Kuu^-1 A has entries >> 1)
    assert_close(cov, c_ref, scale=max(1.0, float(c_ref.abs().max())), tol=tol, what="full_cov cov")
    m2, v2 = layer.predict(X.cuda())
    assert_close(torch.diagonal(cov, dim1=1, dim2=2).T, v2, scale=max(1.0, float(c_ref.abs().max())), tol=tol,
                 what="diag(full cov) vs marginal")
    _, c4 = layer.predict(X.cuda(), full_cov=True, full_output_cov=True)
    assert tuple(c4.shape) == (N, P, N, P)
    assert torch.equal(c4[:, 1, :, 1], cov[1]) and float(c4[:, 0, :, 1].abs().max()) == 0.0
    _, c3 = layer.predict(X.cuda(), full_output_cov=True)
    assert tuple(c3.shape) == (N, P, P)
    assert torch.equal(torch.diagonal(c3, dim1=1, dim2=2), v2)
    # correlated sample (gp_layer.py:329-333, 358-368)
    eps = torch.from_numpy(rng.standard_normal((P, N)))
    dist = layer(X.cuda(), eps=eps.cuda())
    Lc = torch.linalg.cholesky(c_ref + 1e-6 * torch.eye(N, dtype=torch.float64))
    s_ref = (m_ref.T + (Lc @ eps[:, :, None])[:, :, 0]).T
    got = layer._convert_to_tensor_fn(dist)
    assert tuple(got.shape) == (N, P)
    assert tuple(dist.sample().shape) == (P, N)  # the TFP distribution itself is over [Q, N]
    assert_close(got, s_ref, tol=max(1e-7, 100 * tol), what="full_cov sample")  # chol of a near-singular N x N matrix
    layer.num_samples = 4
    assert tuple(layer(X.cuda())._tensor().shape) == (4, N, P)
    layer.full_output_cov = True
    with pytest.raises(NotImplementedError):
        layer(X.cuda())


@pytest.mark.parametrize("K,whiten", [(1, True), (3, True), (1, False), (3, False)])
def test_full_cov_gradients_match_oracle(K, whiten):
    """Gradients THROUGH the full-covariance branch (reference gp_layer.py:329-338 is differentiated by TF):
    d/d{X, Z, lengthscales, variance, q_mu, q_sqrt} of a random linear functional of (mean, cov), and of the
    correlated sample mean + chol(cov + jitter I) eps, against the oracle's autograd."""
    from gpflux_b200.conditionals import full_cov_conditional
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(21 + K)
    N, M, D, P = 40, 24, 2, 3
    spec = ot.make_spec(rng, M, D, P, K=K, whiten=whiten, q_sqrt_diag=0.3)
    X = torch.from_numpy(1.5 * rng.standard_normal((N, D)))
    Wm = torch.from_numpy(rng.standard_normal((N, P)))
    Wc = torch.from_numpy(rng.standard_normal((P, N, N)))
    eps = torch.from_numpy(rng.standard_normal((P, N)))
    names = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")

    so = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in spec.items()}
    Xo = X.clone().requires_grad_(True)
    m_ref, c_ref = ot.conditional_full_cov(so, Xo)
    Lc = torch.linalg.cholesky(c_ref + 1e-6 * torch.eye(N, dtype=torch.float64))
    s_ref = m_ref.T + (Lc @ eps[:, :, None])[:, :, 0]
    ((m_ref * Wm).sum() + (c_ref * Wc).sum() + (s_ref * Wm.T).sum()).backward()

    sg = {k: spec[k].cuda().requires_grad_(True) for k in names}
    Xg = X.cuda().requires_grad_(True)
    mean, cov = full_cov_conditional(Xg, spec["kernel"], sg["Z"], sg["lengthscales"], sg["variance"], sg["q_mu"], sg["q_sqrt"],
                                     whiten=whiten, jitter=spec["jitter"])
    scale, _, _ = ops.cholesky_inverse(cov + 1e-6 * torch.eye(N, dtype=torch.float64, device="cuda"))
    samp = mean.T + ops.gemm(scale, eps.cuda()[:, :, None].contiguous(), tri=1)[:, :, 0]
    ((mean * Wm.cuda()).sum() + (cov * Wc.cuda()).sum() + (samp * Wm.T.cuda()).sum()).backward()

    tol = max(1e-7, 100 * cond_tol(spec))  # the N x N Cholesky of a smooth kernel matrix amplifies rounding
    assert_close(samp, s_ref, tol=tol, what="sample")
    for k in names:
        assert_close(sg[k].grad, so[k].grad, tol=tol, what=f"grad {k}")
    assert_close(Xg.grad, Xo.grad, tol=tol, what="grad X")


def test_full_cov_layer_trains():
    """A full_cov layer in training mode: the data term reaches the layer's parameters (it used to be refused)."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer

    rng = np.random.default_rng(5)
    N, M, D, P = 30, 16, 2, 2
    kernel = gf.SharedIndependent(gf.SquaredExponential(lengthscales=np.ones(D)), P)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(rng.standard_normal((M, D))))
    layer = GPLayer(kernel, iv, num_data=N, num_latent_gps=P, mean_function=gf.Zero(), full_cov=True).cuda()
    X = torch.from_numpy(1.5 * rng.standard_normal((N, D))).cuda()
    dist = layer(X, training=True, eps=torch.from_numpy(rng.standard_normal((P, N))).cuda())
    f = layer._convert_to_tensor_fn(dist)
    loss = (f ** 2).sum() + layer.losses[0]
    loss.backward()
    for p in layer.parameters():
        if p.requires_grad:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all())
    assert float(layer.q_sqrt.unconstrained_variable.grad.abs().max()) > 0
    assert float(layer.inducing_variable.inducing_variable.Z.unconstrained_variable.grad.abs().max()) > 0
