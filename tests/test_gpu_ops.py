"""GPU parity of the C-ABI building blocks against the CPU oracle (same seeded inputs)."""
import math

import numpy as np
import pytest
import torch

from _util import assert_close

pytestmark = pytest.mark.gpu

KINDS = ["se", "matern12", "matern32", "matern52"]


def _ops():
    from gpflux_b200 import ops

    return ops


def _rand(rng, *shape):
    return torch.from_numpy(rng.standard_normal(shape))


@pytest.mark.parametrize(
    "b,M,N,K,tA,tB",
    [
        (1, 128, 128, 64, False, False),
        (2, 200, 150, 70, False, False),
        (1, 131, 257, 33, True, False),
        (3, 64, 64, 64, False, True),
        (1, 257, 129, 515, True, True),
        (2, 20, 200, 20, False, False),
        (1, 1, 7, 3, False, False),
        (1, 300, 1, 45, False, True),
        (1, 512, 1024, 512, False, False),
    ],
)
def test_gemm_plain(b, M, N, K, tA, tB):
    ops = _ops()
    rng = np.random.default_rng(1)
    A = _rand(rng, b, *((K, M) if tA else (M, K)))
    B = _rand(rng, b, *((N, K) if tB else (K, N)))
    C0 = _rand(rng, b, M, N)
    opA = A.transpose(1, 2) if tA else A
    opB = B.transpose(1, 2) if tB else B
    ref = 0.7 * opA @ opB - 1.3 * C0
    got = ops.gemm(A.cuda(), B.cuda(), transA=tA, transB=tB, alpha=0.7, beta=-1.3, C_in=C0.cuda())
    assert_close(got, ref, scale=math.sqrt(K), tol=1e-13, what="gemm")


@pytest.mark.parametrize("M,N", [(64, 100), (256, 300), (200, 513), (1024, 256)])
def test_gemm_triangular_modes(M, N):
    from gpflux_b200._cabi import TRI_LOWER_A, TRI_LOWER_B, TRI_UPPER_A

    ops = _ops()
    rng = np.random.default_rng(2)
    Lw = torch.tril(_rand(rng, 2, M, M))
    X = _rand(rng, 2, M, N)
    # lower-triangular A
    got = ops.gemm(Lw.cuda(), X.cuda(), tri=TRI_LOWER_A)
    assert_close(got, Lw @ X, scale=math.sqrt(M), tol=1e-13, what="lower_A")
    # upper-triangular op(A) = L^T stored as L
    got = ops.gemm(Lw.cuda(), X.cuda(), transA=True, tri=TRI_UPPER_A)
    assert_close(got, Lw.transpose(1, 2) @ X, scale=math.sqrt(M), tol=1e-13, what="upper_A")
    # lower-triangular B, and both
    L2 = torch.tril(_rand(rng, 2, M, M))
    got = ops.gemm(Lw.cuda(), L2.cuda(), tri=TRI_LOWER_A | TRI_LOWER_B)
    assert_close(got, Lw @ L2, scale=math.sqrt(M), tol=1e-13, what="lower_A*lower_B")
    # tril output of A B^T (contraction over N)
    Y = _rand(rng, 2, M, N)
    got = ops.gemm(X.cuda(), Y.cuda(), transB=True, tril_out=True, alpha=-1.0)
    assert_close(got, -torch.tril(X @ Y.transpose(1, 2)), scale=math.sqrt(N), tol=1e-13, what="tril_out")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("K,nu,nv,D,shared", [(1, 20, 200, 1, True), (3, 70, 130, 5, True), (2, 64, 64, 33, False), (1, 256, 1000, 8, True)])
def test_kernel_matrix_forward(kind, K, nu, nv, D, shared):
    from oracle import svgp_torch as ot

    ops = _ops()
    rng = np.random.default_rng(3)
    U = _rand(rng, K, nu, D)
    V = _rand(rng, nv, D) if shared else _rand(rng, K, nv, D)
    ls = torch.from_numpy(rng.uniform(0.7, 2.0, (K, D)))
    var = torch.from_numpy(rng.uniform(0.5, 2.0, (K,)))
    ref = torch.stack([ot.kernel_matrix(kind, U[k], V if shared else V[k], ls[k], var[k]) for k in range(K)])
    got = ops.kernel_matrix(kind, U.cuda(), V.cuda(), ls.cuda(), var.cuda())
    # Matern-1/2 has a sqrt at r -> 0 that amplifies rounding of the expansion form; compare on k scale
    assert_close(got, ref, scale=float(var.max()), tol=1e-9 if kind == "se" else 2e-8 if kind == "matern12" else 1e-9, what=f"K {kind}")


def test_kuu_jitter_and_symmetry():
    from oracle import svgp_torch as ot

    ops = _ops()
    rng = np.random.default_rng(4)
    spec = ot.make_spec(rng, 50, 3, 2, K=2)
    ref = torch.stack([ot.Kuu(spec, k) for k in range(2)])
    Z = spec["Z"].cuda()
    got = ops.kernel_matrix("se", Z, Z, spec["lengthscales"].cuda(), spec["variance"].cuda(), jitter=1e-6)
    assert_close(got, ref, scale=1.0, what="Kuu")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("K,nu,nv,D,shared", [(1, 20, 200, 1, True), (3, 70, 130, 5, True), (2, 40, 64, 33, False), (1, 64, 500, 8, True), (2, 33, 65, 40, True),
                                                (2, 100, 300, 3, False), (2, 130, 257, 15, True)])
def test_kernel_matrix_backward(kind, K, nu, nv, D, shared):
    from oracle import svgp_torch as ot

    ops = _ops()
    rng = np.random.default_rng(5)
    U = _rand(rng, K, nu, D).requires_grad_(True)
    V = (_rand(rng, nv, D) if shared else _rand(rng, K, nv, D)).requires_grad_(True)
    ls = torch.from_numpy(rng.uniform(0.9, 2.0, (K, D))).requires_grad_(True)
    var = torch.from_numpy(rng.uniform(0.5, 2.0, (K,))).requires_grad_(True)
    dK = _rand(rng, K, nu, nv)
    Kmat = torch.stack([ot.kernel_matrix(kind, U[k], V if shared else V[k], ls[k], var[k]) for k in range(K)])
    (Kmat * dK).sum().backward()
    dU, dV, dls, dvar = ops.kernel_matrix_bwd(kind, U.detach().cuda(), V.detach().cuda(), ls.detach().cuda(), var.detach().cuda(), dK.cuda())
    tol = 1e-9 if kind != "matern12" else 1e-7  # d/dr2 of exp(-r) ~ 1/r: ill-conditioned near r=0 in both forms
    for got, ref, name in [(dU, U.grad, "dU"), (dV, V.grad, "dV"), (dls, ls.grad, "dls"), (dvar, var.grad, "dvar")]:
        assert_close(got, ref, tol=tol, what=f"{kind} {name}")


@pytest.mark.parametrize("K,M", [(1, 20), (2, 50), (1, 64), (3, 100), (2, 256), (1, 300), (1, 1024)])
def test_cholesky_inverse(K, M):
    ops = _ops()
    rng = np.random.default_rng(6)
    G = _rand(rng, K, M, M)
    A = G @ G.transpose(1, 2) / M + 0.5 * torch.eye(M, dtype=torch.float64)
    L, Linv, info = ops.cholesky_inverse(A.cuda())
    assert int(info.abs().sum().item()) == 0
    Lref = torch.linalg.cholesky(A)
    assert_close(L, Lref, scale=1.0, tol=1e-12, what="L")
    eye = torch.eye(M, dtype=torch.float64).expand(K, M, M)
    assert_close((Linv.cpu() @ Lref), eye, scale=1.0, tol=1e-11, what="Linv L = I")
    assert float(torch.triu(L.cpu(), 1).abs().max()) == 0.0
    assert float(torch.triu(Linv.cpu(), 1).abs().max()) == 0.0


def test_cholesky_flags_non_pd():
    ops = _ops()
    A = torch.eye(70, dtype=torch.float64)[None].clone()
    A[0, 66, 66] = -1.0
    _, _, info = ops.cholesky_inverse(A.cuda())
    assert int(info[0].item()) == 67


# >= 2^20 elements take the TMA-staged (cp.async.bulk + mbarrier) kernel; odd / ragged tails included
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 65536 * 8 + 3, (1 << 20), (1 << 20) + 1, 5 * (1 << 20) + 2048 * 3 + 77])
def test_reparam_sample(n):
    ops = _ops()
    rng = np.random.default_rng(7)
    m, v, e = _rand(rng, n), torch.from_numpy(rng.uniform(0.1, 2.0, n)), _rand(rng, n)
    got = ops.reparam_sample(m.cuda(), v.cuda(), e.cuda())
    ref = m + torch.sqrt(v) * e
    # one sqrt + one (possibly fused) multiply-add: at most an ulp from the unfused CPU result
    assert_close(got, ref, tol=4e-16, what="sample")


@pytest.mark.parametrize("M,P", [(20, 1), (50, 3), (256, 8), (1024, 2)])
def test_kl_white(M, P):
    from oracle import svgp_torch as ot

    ops = _ops()
    rng = np.random.default_rng(8)
    spec = ot.make_spec(rng, M, 2, P)
    # garbage above the diagonal must be ignored (tf.linalg.band_part in gauss_kl)
    q_sqrt = spec["q_sqrt"] + torch.triu(torch.ones(M, M, dtype=torch.float64), 1) * 3.0
    got = ops.kl_white(spec["q_mu"].cuda(), q_sqrt.cuda())
    assert_close(got, ot.prior_kl(spec), what="kl")
    # reference tests/gpflux/layers/test_gp_layer.py:67-85: KL == 0.0 at q_mu = 0, q_sqrt = I
    z = ops.kl_white(torch.zeros(M, P, dtype=torch.float64).cuda(), torch.eye(M, dtype=torch.float64).repeat(P, 1, 1).cuda())
    assert float(z.item()) == 0.0


def test_gaussian_ve():
    from oracle import svgp_torch as ot

    ops = _ops()
    rng = np.random.default_rng(9)
    N, Q = 5000, 3
    Fmu = _rand(rng, N, Q).requires_grad_(True)
    Fvar = torch.from_numpy(rng.uniform(0.1, 1.0, (N, Q))).requires_grad_(True)
    Y = _rand(rng, N, Q)
    nv = torch.tensor(0.08, dtype=torch.float64, requires_grad=True)
    ref = ot.gaussian_variational_expectations(Fmu, Fvar, Y, nv).sum()
    (-ref / N).backward()
    gFmu = Fmu.detach().cuda().requires_grad_(True)
    gFvar = Fvar.detach().cuda().requires_grad_(True)
    gnv = nv.detach().cuda().requires_grad_(True)
    got = ops.gaussian_variational_expectations_sum(gFmu, gFvar, Y.cuda(), gnv)
    (-got / N).backward()
    assert_close(got, ref, what="ve")
    assert_close(gFmu.grad, Fmu.grad, what="dFmu")
    assert_close(gFvar.grad, Fvar.grad, what="dFvar")
    assert_close(gnv.grad, nv.grad, what="dnoise")


def test_rejects_cpu_and_float32():
    ops = _ops()
    x = torch.zeros(4, dtype=torch.float64)
    with pytest.raises(ValueError):
        ops.reparam_sample(x, x, x)
    y = torch.zeros(4, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        ops.reparam_sample(y, y, y)
