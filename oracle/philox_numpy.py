"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the counter-based noise
generator behind gpfx_randn_philox / gpfx_reparam_sample_philox.  Only tests/, smoke() and
bench.py's cpu_baseline leg may import this.

This is NOT a restatement of anything in GPflux: the reference draws its reparameterisation
noise through TFP (`MultivariateNormalDiag(...).sample()`, gpflux/layers/gp_layer.py:339-363),
whose stream cannot be reproduced here.  The Philox variant of row a8 is an addition
that is equal in distribution only; parity with the reference is carried by the eps-in entry
points, never by this one.

Algorithm: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2,
3", SC'11; Random123 library).  Pinning: `KAT` below holds three known-answer vectors of the
philox4x32_10 rows of Random123's `kat_vectors` file.  That file is not available offline; the
vectors were written down from memory of the published file BEFORE this restatement was run, and
the restatement (built from the paper's round function and constants) reproduces all three.  Two
independent recollections agreeing on 384 bits is the evidence; it is not a check against a file
on disk.

Layout used by the CUDA kernels (and mirrored here):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (i & 0xffffffff, i >> 32, offset & 0xffffffff, offset >> 32)   for pair index i
    u1 = (((x1 << 32 | x0) >> 11) + 0.5) * 2^-53,  u2 likewise from (x3, x2)       in (0, 1)
    r = sqrt(-2 ln u1);  out[2i] = r cos(2 pi u2);  out[2i+1] = r sin(2 pi u2)     (Box-Muller)
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

# (counter, key, expected) - provenance in the module docstring
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox4x32_10(counter, key):
    """counter: 4 arrays (or ints) of 32-bit words, key: 2 ints -> 4 uint64 arrays holding 32-bit words."""
    c = [np.asarray(x, dtype=np.uint64) for x in counter]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        c = [(p1 >> _S32) ^ c[1] ^ np.uint64(k0), p1 & _LO, (p0 >> _S32) ^ c[3] ^ np.uint64(k1), p0 & _LO]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c


def raw_words(seed: int, offset: int, n_pairs: int) -> np.ndarray:
    """[n_pairs, 4] uint32: the four output words of pair index 0..n_pairs-1."""
    i = np.arange(n_pairs, dtype=np.uint64)
    ctr = (i & _LO, i >> _S32, np.uint64(offset & 0xFFFFFFFF), np.uint64((offset >> 32) & 0xFFFFFFFF))
    x = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return np.stack([np.broadcast_to(w, i.shape) for w in x], axis=1).astype(np.uint32)


def randn(seed: int, offset: int, n: int) -> np.ndarray:
    """n float64 standard normals, element 2i / 2i+1 from pair i."""
    n_pairs = (n + 1) // 2
    w = raw_words(seed, offset, n_pairs).astype(np.uint64)
    u1 = ((((w[:, 1] << _S32) | w[:, 0]) >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0**-53
    u2 = ((((w[:, 3] << _S32) | w[:, 2]) >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0**-53
    r = np.sqrt(-2.0 * np.log(u1))
    out = np.empty(2 * n_pairs)
    out[0::2] = r * np.cos(2.0 * np.pi * u2)
    out[1::2] = r * np.sin(2.0 * np.pi * u2)
    return out[:n]


def reparam_sample(mean: np.ndarray, var: np.ndarray, seed: int, offset: int):
    """f = mean + sqrt(var) * eps with eps = randn(seed, offset, mean.size) in C order; returns (f, eps)."""
    eps = randn(seed, offset, mean.size).reshape(mean.shape)
    return mean + np.sqrt(var) * eps, eps
