"""numpy/scipy float64 restatement (forward only) of the GPflow algorithm GPflux's hot path calls.

TEST INFRASTRUCTURE ONLY - PARITY UNPINNED (see ``oracle/__init__.py``).  Written independently
of ``oracle.svgp_torch`` (scipy.linalg solvers, explicit per-output loops) so that the two can be
cross-checked; the layer-spec dict is the same, with numpy arrays.  Reference call sites cited per
function are relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cholesky, solve_triangular

DEFAULT_JITTER = 1e-6


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def spec_to_numpy(spec):
    return {k: (_np(v) if hasattr(v, "shape") else v) for k, v in spec.items()}


def kernel_matrix(kind, X, X2, ls, var):
    """gpflow stationary kernels through square_distance's expansion."""
    Xs, X2s = X / ls, X2 / ls
    r2 = (Xs**2).sum(1)[:, None] + (X2s**2).sum(1)[None, :] - 2.0 * Xs @ X2s.T
    if kind == "se":
        return var * np.exp(-0.5 * r2)
    r = np.sqrt(np.maximum(r2, 1e-36))
    if kind == "matern12":
        return var * np.exp(-r)
    if kind == "matern32":
        return var * (1 + math.sqrt(3) * r) * np.exp(-math.sqrt(3) * r)
    if kind == "matern52":
        return var * (1 + math.sqrt(5) * r + 5.0 / 3.0 * r * r) * np.exp(-math.sqrt(5) * r)
    raise ValueError(kind)


def conditional(spec, X):
    """base_conditional per latent GP (gp_layer.py:257-266) -> fmean [N,P], fvar [N,P]."""
    s = spec_to_numpy(spec)
    X = _np(X)
    K, M, _ = s["Z"].shape
    P = s["q_mu"].shape[1]
    N = X.shape[0]
    fmean = np.zeros((N, P))
    fvar = np.zeros((N, P))
    for p in range(P):
        k = p if K > 1 else 0
        Z, ls, var = s["Z"][k], s["lengthscales"][k], s["variance"][k]
        Kmm = kernel_matrix(s["kernel"], Z, Z, ls, var) + s.get("jitter", DEFAULT_JITTER) * np.eye(M)
        Kmn = kernel_matrix(s["kernel"], Z, X, ls, var)
        Lm = cholesky(Kmm, lower=True)
        A = solve_triangular(Lm, Kmn, lower=True)
        v = var - np.einsum("mn,mn->n", A, A)
        if not s.get("whiten", True):
            A = solve_triangular(Lm.T, A, lower=False)
        Lq = np.tril(s["q_sqrt"][p])
        LTA = Lq.T @ A
        fmean[:, p] = A.T @ s["q_mu"][:, p]
        fvar[:, p] = v + np.einsum("mn,mn->n", LTA, LTA)
    return fmean, fvar


def mean_function(spec, X):
    s = spec_to_numpy(spec)
    X = _np(X)
    kind = s.get("mean", "zero")
    if kind == "identity":
        return X
    if kind == "zero":
        return np.zeros((X.shape[0], s["q_mu"].shape[1]))
    return X @ s["mean_A"] + s["mean_b"]


def predict(spec, X):
    m, v = conditional(spec, X)
    return m + mean_function(spec, X), v


def prior_kl(spec):
    """gauss_kl (gp_layer.py:303-311), dense closed form per output:
    KL[N(m,S)||N(0,Kp)] = 1/2 [ m^T Kp^-1 m - M + tr(Kp^-1 S) - logdet S + logdet Kp ]."""
    s = spec_to_numpy(spec)
    K, M, _ = s["Z"].shape
    P = s["q_mu"].shape[1]
    total = 0.0
    for p in range(P):
        k = p if K > 1 else 0
        Lq = np.tril(s["q_sqrt"][p])
        S = Lq @ Lq.T
        m = s["q_mu"][:, p]
        if s.get("whiten", True):
            Kp = np.eye(M)
        else:
            Z, ls, var = s["Z"][k], s["lengthscales"][k], s["variance"][k]
            Kp = kernel_matrix(s["kernel"], Z, Z, ls, var) + s.get("jitter", DEFAULT_JITTER) * np.eye(M)
        Kpi = np.linalg.inv(Kp)
        _, logdetS = np.linalg.slogdet(S)
        _, logdetK = np.linalg.slogdet(Kp)
        total += 0.5 * (m @ Kpi @ m - M + np.trace(Kpi @ S) - logdetS + logdetK)
    return total


def gaussian_variational_expectations(Fmu, Fvar, Y, noise_variance):
    Fmu, Fvar, Y = _np(Fmu), _np(Fvar), _np(Y)
    nv = float(noise_variance)
    return (-0.5 * math.log(2 * math.pi) - 0.5 * math.log(nv) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / nv).sum(-1)


def deep_gp_loss(specs, X, Y, eps, noise_variance, num_data):
    """deep_gp.py:151-231 training loss, forward only."""
    h = _np(X)
    kl = 0.0
    for i, spec in enumerate(specs):
        m, v = predict(spec, h)
        kl += prior_kl(spec)
        if i < len(specs) - 1:
            h = m + np.sqrt(v) * _np(eps[i])
    ve = gaussian_variational_expectations(m, v, Y, noise_variance)
    return kl / num_data - ve.mean()


def rff_cosine(X, W, b, ls, var):
    X, W, b = _np(X), _np(W), _np(b)
    return math.sqrt(2.0 * float(var) / W.shape[0]) * np.cos((X / _np(ls)) @ W.T + b)
