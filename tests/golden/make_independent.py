"""Independent pin for the oracle (Next #4): a THIRD derivation of the SVGP layer that
shares no code and no algorithmic structure with ``oracle/`` or the CUDA path, evaluated with mpmath
at 50 significant digits and stored as ``tests/golden/independent_cases.npz``.

What is different from the oracle (which follows GPflow's ``base_conditional`` / ``gauss_kl``):

* squared distances in the DIRECT form  r2 = sum_d ((x_d - z_d) / l_d)^2  (GPflow: expansion form);
* the conditional in the dense closed form of the sparse variational posterior
      q(u) = N(m, S):   mu = Kfu Kuu^-1 m,   var = kff - diag(Kfu Kuu^-1 Kuf) + diag(Kfu Kuu^-1 S Kuu^-1 Kuf)
  through an LU-based inverse of Kuu (GPflow: Cholesky + triangular solves, whitened algebra);
  a whitened layer is mapped to q(u) by u = L v (m = L q_mu, S = L Sq L^T, L = chol(Kuu) from mpmath);
* the KL in the dense closed form  1/2 [tr(Kp^-1 S) + m^T Kp^-1 m - M + log det Kp - log det S];
* gradients by central FINITE DIFFERENCES of that closed form in 50-digit arithmetic (h = 1e-15:
  truncation error ~1e-30, rounding error ~1e-35), not by automatic differentiation.

It is still not the reference (GPflow/TF cannot run here): reports keep saying "parity unpinned
(reference not runnable)".  What it rules out is the golden files being the oracle's own echo.

Run:  python tests/golden/make_independent.py   (about a minute; mpmath only)
"""
from __future__ import annotations

import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))
JITTER = mp.mpf("1e-6")


def k_of_r2(kind, r2, variance):
    if kind == "se":
        return variance * mp.exp(-r2 / 2)
    r = mp.sqrt(r2)
    if kind == "matern12":
        return variance * mp.exp(-r)
    if kind == "matern32":
        s = mp.sqrt(3) * r
        return variance * (1 + s) * mp.exp(-s)
    if kind == "matern52":
        s = mp.sqrt(5) * r
        return variance * (1 + s + s * s / 3) * mp.exp(-s)
    raise ValueError(kind)


def kmat(kind, A, B, ls, variance):
    out = mp.matrix(len(A), len(B))
    for i, a in enumerate(A):
        for j, b in enumerate(B):
            r2 = mp.fsum(((a[d] - b[d]) / ls[d]) ** 2 for d in range(len(ls)))
            out[i, j] = k_of_r2(kind, r2, variance)
    return out


def layer(case, params, X):
    """-> fmean [N][P], fvar [N][P], kl with every quantity an mpf."""
    kind, white = case["kernel"], case["whiten"]
    Z, ls, var, q_mu, q_sqrt = params["Z"], params["lengthscales"], params["variance"], params["q_mu"], params["q_sqrt"]
    K = len(Z)
    M, P = len(q_mu), len(q_mu[0])
    N = len(X)
    fmean = [[None] * P for _ in range(N)]
    fvar = [[None] * P for _ in range(N)]
    kl = mp.mpf(0)
    cache = {}
    for p in range(P):
        k = p if K > 1 else 0
        if k not in cache:
            Kuu = kmat(kind, Z[k], Z[k], ls[k], var[k]) + JITTER * mp.eye(M)
            Kuf = kmat(kind, Z[k], X, ls[k], var[k])
            Kinv = mp.inverse(Kuu)  # LU-based, 50 digits
            KinvKuf = Kinv * Kuf  # [M, N]
            cache[k] = (Kuu, Kuf, KinvKuf, mp.cholesky(Kuu) if white else None, Kinv)
        Kuu, Kuf, KinvKuf, L, Kinv = cache[k]
        Lq = mp.matrix(M, M)
        for i in range(M):
            for j in range(i + 1):
                Lq[i, j] = q_sqrt[p][i][j]
        Sq = Lq * Lq.T
        m = mp.matrix([q_mu[i][p] for i in range(M)])
        if white:
            m_u, S_u = L * m, L * Sq * L.T
        else:
            m_u, S_u = m, Sq
        for n in range(N):
            a = KinvKuf[:, n]
            fmean[n][p] = (a.T * m_u)[0]
            fvar[n][p] = var[k] - (Kuf[:, n].T * a)[0] + (a.T * S_u * a)[0]
        # KL[q || prior] in the parameterisation the layer optimises
        if white:
            Kp_inv_S_trace = mp.fsum(Sq[i, i] for i in range(M))
            maha = (m.T * m)[0]
            logdet_p = mp.mpf(0)
        else:
            KS = Kinv * Sq
            Kp_inv_S_trace = mp.fsum(KS[i, i] for i in range(M))
            maha = (m.T * Kinv * m)[0]
            logdet_p = mp.log(mp.det(Kuu))
        kl += (Kp_inv_S_trace + maha - M + logdet_p - mp.log(mp.det(Sq))) / 2
    return fmean, fvar, kl


def to_mp(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        return mp.mpf(float(a))
    return [to_mp(x) for x in a]


def scalar_loss(case, params, X, w_m, w_v, w_kl):
    fm, fv, kl = layer(case, params, X)
    N, P = len(fm), len(fm[0])
    return mp.fsum(w_m[n][p] * fm[n][p] + w_v[n][p] * fv[n][p] for n in range(N) for p in range(P)) + w_kl * kl


def fd_gradient(fn, leaf_list, index_iter, h=mp.mpf("1e-15")):
    """Central difference of fn() w.r.t. nested-list entries."""
    grads = []
    for idx in index_iter:
        ref = leaf_list
        for i in idx[:-1]:
            ref = ref[i]
        x0 = ref[idx[-1]]
        ref[idx[-1]] = x0 + h
        fp = fn()
        ref[idx[-1]] = x0 - h
        fm = fn()
        ref[idx[-1]] = x0
        grads.append((fp - fm) / (2 * h))
    return grads


def indices(shape, lower_only=False):
    out = []
    for idx in np.ndindex(*shape):
        if lower_only and idx[-1] > idx[-2]:
            continue
        out.append(tuple(int(i) for i in idx))
    return out


CASES = [
    # name, kernel, whiten, N, M, D, P, K
    ("se_shared_white", "se", True, 5, 5, 2, 2, 1),
    ("se_separate_white", "se", True, 4, 4, 3, 2, 2),
    ("matern32_shared_white", "matern32", True, 4, 5, 2, 1, 1),
    ("se_shared_nonwhite", "se", False, 4, 4, 2, 2, 1),
    ("matern52_separate_nonwhite", "matern52", False, 3, 4, 2, 2, 2),
    ("se_separate_white_m12", "se", True, 9, 12, 3, 3, 3),
    ("matern12_shared_white", "matern12", True, 4, 5, 2, 2, 1),
]


def main():
    out = {}
    names = []
    for ci, (name, kind, white, N, M, D, P, K) in enumerate(CASES):
        rng = np.random.default_rng(1000 + ci)
        f64 = dict(
            Z=rng.standard_normal((K, M, D)) * 1.5,
            lengthscales=rng.uniform(0.8, 1.6, (K, D)),
            variance=rng.uniform(0.6, 1.4, (K,)),
            q_mu=0.4 * rng.standard_normal((M, P)),
            q_sqrt=np.stack([np.diag(rng.uniform(0.4, 0.9, M)) + 0.15 * np.tril(rng.standard_normal((M, M)), -1)
                             for _ in range(P)]),
        )
        X64 = rng.standard_normal((N, D)) * 1.5
        w_m64, w_v64 = rng.standard_normal((N, P)), rng.standard_normal((N, P))
        w_kl = 0.37
        case = {"kernel": kind, "whiten": white}
        params = {k: to_mp(v) for k, v in f64.items()}
        X = to_mp(X64)
        w_m, w_v = to_mp(w_m64), to_mp(w_v64)
        fm, fv, kl = layer(case, params, X)
        loss_fn = lambda: scalar_loss(case, params, X, w_m, w_v, mp.mpf(w_kl))  # noqa: E731
        pre = f"{name}/"
        names.append(name)
        out[pre + "meta"] = np.array([kind, str(int(white))])
        for k, v in f64.items():
            out[pre + k] = v
        out[pre + "X"] = X64
        out[pre + "w_m"], out[pre + "w_v"], out[pre + "w_kl"] = w_m64, w_v64, np.float64(w_kl)
        out[pre + "fmean"] = np.array([[float(x) for x in row] for row in fm])
        out[pre + "fvar"] = np.array([[float(x) for x in row] for row in fv])
        out[pre + "kl"] = np.float64(float(kl))
        out[pre + "loss"] = np.float64(float(loss_fn()))
        for key, leaf, shape, lower in (("Z", params["Z"], f64["Z"].shape, False),
                                        ("lengthscales", params["lengthscales"], f64["lengthscales"].shape, False),
                                        ("variance", params["variance"], f64["variance"].shape, False),
                                        ("q_mu", params["q_mu"], f64["q_mu"].shape, False),
                                        ("q_sqrt", params["q_sqrt"], f64["q_sqrt"].shape, True),
                                        ("X", X, X64.shape, False)):
            idx = indices(shape, lower_only=lower)
            g = fd_gradient(loss_fn, leaf, idx)
            G = np.zeros(shape)
            for i, gi in zip(idx, g):
                G[i] = float(gi)
            out[pre + "grad_" + key] = G
        print(name, "loss", mp.nstr(loss_fn(), 20), "kl", mp.nstr(kl, 20))
    out["names"] = np.array(names)
    np.savez(os.path.join(HERE, "independent_cases.npz"), **out)


if __name__ == "__main__":
    main()
