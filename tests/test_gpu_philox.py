"""GPU checks of the Philox variant of row a8 (gpfx_philox_raw / gpfx_randn_philox /
gpfx_reparam_sample_philox) against oracle/philox_numpy.py.

Scope of the claim: the generator's words are compared BIT FOR BIT with the numpy restatement
(itself pinned to three Philox4x32-10 known-answer vectors, see tests/test_cpu_philox.py and the
oracle's header for their provenance); the normals agree to 2e-14 absolute (sincospi / log on the
device vs numpy's cos / sin / log of the rounded angle: a few ulp of a value <= ~8.6).  This is an
added throughput path, not parity with GPflux's TFP noise stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NORMAL_ATOL = 2e-14


def _ops():
    from gpflux_b200 import ops

    return ops


@pytest.mark.parametrize(
    "seed,offset,n_pairs",
    [(0, 0, 1), (0, 0, 1000), (12345, 7, 4097), ((0xDEADBEEF << 32) | 0x01234567, (0xFFFFFFFF << 32) | 0x89ABCDEF, 333),
     ((1 << 64) - 1, (1 << 64) - 1, 65)],
)
def test_raw_words_bit_exact(seed, offset, n_pairs):
    from oracle import philox_numpy as px

    got = _ops().philox_raw(seed, offset, n_pairs).cpu().numpy().astype(np.uint32)
    np.testing.assert_array_equal(got, px.raw_words(seed, offset, n_pairs))


def test_raw_words_known_answer_on_device():
    from oracle import philox_numpy as px

    # seed 0 / offset 0 / pair 0 is the all-zero counter and key of the first known-answer vector
    got = _ops().philox_raw(0, 0, 1).cpu().numpy()[0]
    assert tuple(int(w) for w in got) == px.KAT[0][2]


@pytest.mark.parametrize("n", [0, 1, 2, 7, 1000, 65537, (1 << 20) + 3])
def test_randn_matches_oracle(n):
    from oracle import philox_numpy as px

    got = _ops().randn_philox((n,), seed=99, offset=5).cpu().numpy()
    want = px.randn(99, 5, n)
    assert got.shape == want.shape
    if n:
        assert np.max(np.abs(got - want)) <= NORMAL_ATOL


def test_randn_is_a_pure_function_of_its_arguments():
    ops = _ops()
    a = ops.randn_philox((513, 3), seed=1, offset=2)
    b = ops.randn_philox((513, 3), seed=1, offset=2)
    assert torch.equal(a, b)  # regenerated, e.g. by a backward pass: identical bits
    assert not torch.equal(a, ops.randn_philox((513, 3), seed=1, offset=3))
    assert not torch.equal(a, ops.randn_philox((513, 3), seed=2, offset=2))


def test_randn_moments():
    z = _ops().randn_philox((1 << 24,), seed=2026, offset=0)
    n = z.numel()
    assert bool(torch.isfinite(z).all())
    assert abs(float(z.mean())) < 5.0 / np.sqrt(n)
    assert abs(float(z.var()) - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(float((z**3).mean())) < 5.0 * np.sqrt(15.0 / n)
    assert abs(float((z**4).mean()) - 3.0) < 5.0 * np.sqrt(96.0 / n)


@pytest.mark.parametrize("shape,shift", [((33, 5), 0), ((8192, 8), 0), ((1001,), 1), ((64, 3), 1)])
def test_reparam_sample_philox(shape, shift):
    """shift = 1 offsets every buffer by one double (8-byte aligned only): the scalar path."""
    from oracle import philox_numpy as px

    ops = _ops()
    rng = np.random.default_rng(3)
    m, v = rng.standard_normal(shape), rng.uniform(0.1, 2.0, shape)
    n = m.size

    def dev(a):
        buf = torch.empty(n + shift, dtype=torch.float64, device="cuda")
        buf[shift:] = torch.from_numpy(a.reshape(-1))
        return buf[shift:].view(shape)

    f, eps = ops.reparam_sample_philox(dev(m), dev(v), seed=11, offset=4, return_eps=True)
    f_only = ops.reparam_sample_philox(dev(m), dev(v), seed=11, offset=4)
    want_f, want_eps = px.reparam_sample(m, v, 11, 4)
    assert np.max(np.abs(eps.cpu().numpy() - want_eps)) <= NORMAL_ATOL
    assert np.max(np.abs(f.cpu().numpy() - want_f)) <= NORMAL_ATOL * np.sqrt(2.0) + 4e-16 * np.max(np.abs(want_f))
    assert torch.equal(f, f_only)
    # and against the eps-in kernel on the device's own eps: the same map, at most an ulp apart
    f_in = ops.reparam_sample(dev(m), dev(v), eps.contiguous())
    assert float((f - f_in).abs().max()) <= 4e-16 * float(f.abs().max())


def test_gp_layer_seeded_noise():
    """GPLayer.seed_noise: call c uses stream offset c; fsample = fmean + sqrt(fvar) * that eps."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.layers import GPLayer
    from oracle import philox_numpy as px

    rng = np.random.default_rng(0)
    N, M, D, P = 257, 16, 3, 2
    layer = GPLayer(gf.SharedIndependent(gf.SquaredExponential(lengthscales=np.ones(D)), P),
                    gf.SharedIndependentInducingVariables(gf.InducingPoints(rng.standard_normal((M, D)))),
                    num_data=1000, mean_function=gf.Zero()).cuda()
    layer.q_mu.assign(0.3 * rng.standard_normal((M, P)))
    X = torch.from_numpy(rng.standard_normal((N, D))).cuda()
    layer.seed_noise(42)
    for call in range(2):
        f = layer(X)._tensor()
        mean, var = layer.predict(X)
        want = mean.detach().cpu().numpy() + np.sqrt(var.detach().cpu().numpy()) * px.randn(42, call, N * P).reshape(N, P)
        assert np.max(np.abs(f.detach().cpu().numpy() - want)) <= 1e-12
    layer.seed_noise(None)
    assert layer._noise_seed is None
