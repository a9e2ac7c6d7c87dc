"""BASELINE.json's FULL layer sizes on the GPU, checked through size-independent properties (the CPU
oracle needs tens of seconds for one c3 layer, so it is used only at the c2 size here):

* c2 hidden layer (N=8192, M=256, D=8, P=8): outputs and every gradient against the oracle;
* rows are independent given the parameters (gp_layer.py:339-341): evaluating two row chunks
  separately reproduces the full-batch outputs, and the parameter gradients of the chunks ADD UP to the
  full-batch gradients (linearity of the backward pass in the row-wise cotangents);
* doubling the cotangents doubles every gradient;
* the reference's init-state invariants (tests/gpflux/layers/test_gp_layer.py:67-85): q_mu = 0,
  q_sqrt = I -> KL == 0 exactly, fmean == mean function, fvar == kernel variance;
* c3 layer shard (N=8192, M=1024, D=32, P=K=32 separate kernels): the same properties."""
import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu

PARAMS = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")


def _spec(rng, M, D, P, K, q_off=0.02, ls=(1.0, 2.0)):
    from oracle import svgp_torch as ot

    scale = np.sqrt(D) / 2
    return ot.make_spec(rng, M, D, P, K=K, mean="identity" if D == P else "zero", ls_lo=ls[0] * scale, ls_hi=ls[1] * scale,
                        q_sqrt_diag=0.3, q_sqrt_off=q_off)


def _gpu_layer(spec, X, eps, w_f, w_m, w_v, w_kl=1.0):
    from gpflux_b200 import ops

    p = {k: spec[k].detach().cuda().clone().requires_grad_(True) for k in PARAMS}
    Xg = X.clone().requires_grad_(True)
    mean_add = Xg if spec["mean"] == "identity" else None
    f, m, v, kl = ops.svgp_layer(Xg, p["Z"], p["lengthscales"], p["variance"], p["q_mu"], p["q_sqrt"],
                                 mean_add=mean_add, eps=eps, kernel=spec["kernel"], whiten=True, jitter=spec["jitter"])
    loss = (f * w_f).sum() + (m * w_m).sum() + (v * w_v).sum() + w_kl * kl
    loss.backward()
    g = {k: p[k].grad for k in PARAMS}
    g["X"] = Xg.grad
    return f.detach(), m.detach(), v.detach(), kl.detach(), g


def _properties(spec, N, D, P, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
    eps = torch.randn(N, P, dtype=torch.float64, device="cuda", generator=g)
    w_f = torch.randn(N, P, dtype=torch.float64, device="cuda", generator=g)
    w_m = torch.randn(N, P, dtype=torch.float64, device="cuda", generator=g)
    w_v = torch.randn(N, P, dtype=torch.float64, device="cuda", generator=g)
    f, m, v, kl, gr = _gpu_layer(spec, X, eps, w_f, w_m, w_v)
    assert bool(torch.isfinite(f).all()) and bool((v > 0).all())
    # rows are independent: two ragged chunks reproduce the full batch and their gradients add up
    cut = N // 2 + 37
    parts = []
    for sl, wkl in ((slice(0, cut), 1.0), (slice(cut, N), 0.0)):  # the KL term is counted once
        parts.append(_gpu_layer(spec, X[sl], eps[sl], w_f[sl], w_m[sl], w_v[sl], w_kl=wkl))
    for i, name in enumerate(("fsample", "fmean", "fvar")):
        assert_close(torch.cat([parts[0][i], parts[1][i]]), (f, m, v)[i], scale=1.0, tol=1e-12, what=f"chunked {name}")
    for k in PARAMS:
        assert_close(parts[0][4][k] + parts[1][4][k], gr[k], what=f"chunk-additive d{k}")
    assert_close(torch.cat([parts[0][4]["X"], parts[1][4]["X"]]), gr["X"], tol=1e-12, what="chunked dX")
    # linearity in the cotangents
    _, _, _, _, gr2 = _gpu_layer(spec, X, eps, 2 * w_f, 2 * w_m, 2 * w_v, w_kl=2.0)
    for k in PARAMS + ("X",):
        assert_close(gr2[k], 2 * gr[k], tol=1e-13, what=f"linearity d{k}")
    return X


def _init_invariants(spec, X, P):
    from gpflux_b200 import ops

    M = spec["q_mu"].shape[0]
    s = {k: spec[k].detach().cuda() for k in PARAMS}
    q_mu0 = torch.zeros_like(s["q_mu"])
    q_sqrt0 = torch.eye(M, dtype=torch.float64, device="cuda").repeat(P, 1, 1)
    f, m, v, kl = ops.svgp_layer(X, s["Z"], s["lengthscales"], s["variance"], q_mu0, q_sqrt0, mean_add=None, eps=None,
                                 kernel=spec["kernel"], whiten=True, jitter=spec["jitter"])
    assert float(kl) == 0.0
    assert float(m.abs().max()) == 0.0
    var = s["variance"] if s["variance"].numel() == P else s["variance"].expand(P)
    assert_close(v, var[None, :].expand_as(v), tol=1e-9, what="fvar == Knn at q_sqrt = I")


def test_c2_hidden_layer_full_size_against_oracle():
    from test_gpu_layer import _run_gpu, _run_oracle

    rng = np.random.default_rng(21)
    N, M, D, P = 8192, 256, 8, 8
    spec = _spec(rng, M, D, P, 1)
    X = torch.from_numpy(rng.standard_normal((N, D)))
    eps, w_f, w_m, w_v = (torch.from_numpy(rng.standard_normal((N, P))) for _ in range(4))
    ref = _run_oracle(spec, X, eps, w_f, w_m, w_v, 1.0)
    got = _run_gpu(spec, X, eps, w_f, w_m, w_v, 1.0)
    for i, name in enumerate(("fsample", "fmean", "fvar")):
        assert_close(got[i], ref[i], scale=1.0, what=name)
    assert_close(got[3], ref[3], what="kl")
    for k in PARAMS + ("X",):
        assert_close(got[4][k], ref[4][k], what=f"d{k}")


def test_c2_hidden_layer_full_size_properties():
    rng = np.random.default_rng(22)
    spec = _spec(rng, 256, 8, 8, 1)
    X = _properties(spec, 8192, 8, 8, seed=1)
    _init_invariants(spec, X, 8)


def test_c3_layer_shard_full_size_properties():
    rng = np.random.default_rng(23)
    spec = _spec(rng, 1024, 32, 32, 32, q_off=0.005, ls=(1.0, 3.0))
    X = _properties(spec, 8192, 32, 32, seed=2)
    _init_invariants(spec, X, 32)
    torch.cuda.empty_cache()


def test_empty_minibatch():
    """Zero rows: the reference returns empty [0, Q] tensors; the C-ABI rejects N = 0, so the host
    mirror answers without a launch."""
    from gpflux_b200.helpers import construct_gp_layer

    Z = np.random.default_rng(1).standard_normal((16, 2))
    layer = construct_gp_layer(100, 16, 2, 3, z_init=Z).cuda()
    X = torch.zeros((0, 2), dtype=torch.float64, device="cuda")
    mean, var = layer.predict(X)
    assert tuple(mean.shape) == (0, 3) and tuple(var.shape) == (0, 3)
    dist = layer(X, training=True)
    assert tuple(dist.sample().shape) == (0, 3)
    assert float(layer.losses[0]) == float(layer.prior_kl()) / layer.num_data


def test_million_row_prediction_spot_checked_against_oracle():
    """Maximum-size edge: one call with 2^20 query rows (BASELINE config 5 evaluates 1M points).  The
    first / last 64 rows are compared with the oracle; all outputs must be finite and fvar > 0."""
    from gpflux_b200 import ops
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(40)
    N, M, D, P = 1 << 20, 64, 2, 2
    spec = ot.make_spec(rng, M, D, P, K=1)
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
    s = {k: spec[k].cuda() for k in PARAMS}
    with torch.no_grad():
        f, m, v, kl = ops.svgp_layer(X, s["Z"], s["lengthscales"], s["variance"], s["q_mu"], s["q_sqrt"], mean_add=None,
                                     eps=None, kernel="se", whiten=True, jitter=spec["jitter"])
    assert bool(torch.isfinite(m).all()) and bool((v > 0).all())
    idx = torch.cat([torch.arange(64), torch.arange(N - 64, N)])
    m_ref, v_ref = ot.predict(spec, X[idx.cuda()].cpu())
    assert_close(m[idx.cuda()], m_ref, what="mean (spot check)")
    assert_close(v[idx.cuda()], v_ref, scale=1.0, what="var (spot check)")
    del f, m, v
    torch.cuda.empty_cache()
