"""CPU checks of oracle/philox_numpy.py, the checker for the Philox variant of row a8.

What is pinned and how: `philox_numpy.KAT` holds three Philox4x32-10 known-answer vectors recalled
from Random123's published kat_vectors file (not available offline, see the oracle's header).  The
test asserts the restatement reproduces them bit for bit; everything else here is a statistical or
structural property.  None of this is parity with GPflux, which has no counter-based generator."""
import numpy as np

from oracle import philox_numpy as px


def test_known_answer_vectors():
    for ctr, key, want in px.KAT:
        got = tuple(int(w) for w in px.philox4x32_10(ctr, key))
        assert got == want, (ctr, key, [hex(g) for g in got])


def test_counter_layout_and_streams():
    w = px.raw_words(seed=0, offset=0, n_pairs=3)
    assert tuple(int(v) for v in w[0]) == px.KAT[0][2]  # seed 0, offset 0, pair 0 is the all-zero KAT
    # pair i of one stream equals pair 0 of a hand-built counter
    got = px.philox4x32_10((2, 0, 7, 0), (5, 1))
    assert tuple(int(v) for v in px.raw_words(seed=(1 << 32) | 5, offset=7, n_pairs=3)[2]) == tuple(int(v) for v in got)
    a = px.randn(1, 0, 1001)
    assert a.shape == (1001,)
    np.testing.assert_array_equal(a[:1000], px.randn(1, 0, 1000))  # odd n only truncates
    assert not np.array_equal(a[:10], px.randn(1, 1, 10))  # offset selects a different stream
    assert not np.array_equal(a[:10], px.randn(2, 0, 10))


def test_normal_moments():
    z = px.randn(seed=1234, offset=0, n=1 << 20)
    n = z.size
    assert np.all(np.isfinite(z))
    assert abs(z.mean()) < 5.0 / np.sqrt(n)
    assert abs(z.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(z**3)) < 5.0 * np.sqrt(15.0 / n)
    assert abs(np.mean(z**4) - 3.0) < 5.0 * np.sqrt(96.0 / n)
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5.0 / np.sqrt(n / 2)  # the two Box-Muller outputs


def test_reparam_sample():
    rng = np.random.default_rng(0)
    m, v = rng.standard_normal((33, 5)), rng.uniform(0.1, 2.0, (33, 5))
    f, eps = px.reparam_sample(m, v, seed=9, offset=3)
    np.testing.assert_array_equal(eps.reshape(-1), px.randn(9, 3, 165))
    np.testing.assert_allclose(f, m + np.sqrt(v) * eps, rtol=0, atol=0)
