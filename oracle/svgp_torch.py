"""torch-CPU float64 restatement (with autograd) of the GPflow algorithm GPflux's hot path calls.

TEST INFRASTRUCTURE ONLY - PARITY UNPINNED (see ``oracle/__init__.py``).

Each function cites the reference call site it stands in for (paths relative to /root/reference)
and the GPflow 2.9.x symbol restated (GPflow is not vendored in the reference).

A *layer spec* is a plain dict of tensors:

    kernel        "se" | "matern12" | "matern32" | "matern52"
    Z             [K, M, D]   inducing points, K = 1 (SharedIndependent) or P (SeparateIndependent)
    lengthscales  [K, D]      (ARD; constrained/positive values)
    variance      [K]
    q_mu          [M, P]
    q_sqrt        [P, M, M]   (lower triangle is used)
    mean          "identity" | "zero" | "linear"   (+ mean_A [D, P], mean_b [P] for "linear")
    whiten        bool
    jitter        float (gpflow.config.default_jitter() = 1e-6)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter(), read at gpflux/math.py:36, sample.py:169


# ----------------------------------------------------------------------------------------------
# kernels: gpflow.kernels.stationaries (square_distance expansion form), SURVEY §2a row 4
# ----------------------------------------------------------------------------------------------
def square_distance(X: torch.Tensor, X2: torch.Tensor) -> torch.Tensor:
    """gpflow.utilities.ops.square_distance: ||x||^2 + ||z||^2 - 2 x.z (the expansion form)."""
    Xs = (X * X).sum(-1)
    X2s = (X2 * X2).sum(-1)
    return Xs[:, None] + X2s[None, :] - 2.0 * X @ X2.T


def k_r2(kind: str, r2: torch.Tensor, variance: torch.Tensor) -> torch.Tensor:
    """gpflow.kernels.{SquaredExponential,Matern12,Matern32,Matern52}.K_r2 / K_r."""
    if kind == "se":
        return variance * torch.exp(-0.5 * r2)
    r = torch.sqrt(torch.clamp(r2, min=1e-36))  # gpflow IsotropicStationary.scaled_difference
    if kind == "matern12":
        return variance * torch.exp(-r)
    if kind == "matern32":
        s3 = math.sqrt(3.0)
        return variance * (1.0 + s3 * r) * torch.exp(-s3 * r)
    if kind == "matern52":
        s5 = math.sqrt(5.0)
        return variance * (1.0 + s5 * r + 5.0 / 3.0 * r * r) * torch.exp(-s5 * r)
    raise ValueError(kind)


def kernel_matrix(kind, X, X2, lengthscales, variance):
    """k(X, X2) for one latent kernel.  X [n, D], X2 [m, D], lengthscales [D], variance []."""
    return k_r2(kind, square_distance(X / lengthscales, X2 / lengthscales), variance)


def Kuu(spec: Dict, k: int) -> torch.Tensor:
    """gpflow.covariances.Kuu(InducingPoints, kernel, jitter) -> [M, M] (gp_layer.py:257-266)."""
    Z = spec["Z"][k]
    K = kernel_matrix(spec["kernel"], Z, Z, spec["lengthscales"][k], spec["variance"][k])
    return K + spec.get("jitter", DEFAULT_JITTER) * torch.eye(Z.shape[0], dtype=Z.dtype)


def Kuf(spec: Dict, k: int, X: torch.Tensor) -> torch.Tensor:
    """gpflow.covariances.Kuf(InducingPoints, kernel, X) -> [M, N]."""
    return kernel_matrix(spec["kernel"], spec["Z"][k], X, spec["lengthscales"][k], spec["variance"][k])


# ----------------------------------------------------------------------------------------------
# conditional: gpflow.conditionals.util.base_conditional (full_cov=False), gp_layer.py:257-266
# ----------------------------------------------------------------------------------------------
def mean_function(spec: Dict, X: torch.Tensor) -> torch.Tensor:
    """gpflow.mean_functions.{Identity,Zero,Linear} (gp_layer.py:256)."""
    kind = spec.get("mean", "zero")
    P = spec["q_mu"].shape[1]
    if kind == "identity":
        return X
    if kind == "zero":
        return torch.zeros(X.shape[0], P, dtype=X.dtype)
    if kind == "linear":
        return X @ spec["mean_A"] + spec["mean_b"]
    raise ValueError(kind)


def conditional(spec: Dict, X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fmean [N, P], fvar [N, P] of q(f(X)); the reference runs base_conditional per latent GP
    for SeparateIndependent (serial tf.map_fn) and once for SharedIndependent."""
    Kn = spec["Z"].shape[0]
    P = spec["q_mu"].shape[1]
    white = spec.get("whiten", True)
    means, variances = [], []
    cache = {}
    for p in range(P):
        k = p if Kn > 1 else 0
        if k not in cache:
            Kmm = Kuu(spec, k)
            Kmn = Kuf(spec, k, X)
            Lm = torch.linalg.cholesky(Kmm)
            A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)  # [M, N]
            fvar0 = spec["variance"][k] - (A * A).sum(0)  # Knn - sum A^2
            if not white:
                A = torch.linalg.solve_triangular(Lm.T, A, upper=True)
            cache[k] = (A, fvar0)
        A, fvar0 = cache[k]
        Lq = torch.tril(spec["q_sqrt"][p])  # tf.linalg.band_part(q_sqrt, -1, 0)
        LTA = Lq.T @ A  # [M, N]
        means.append(A.T @ spec["q_mu"][:, p])
        variances.append(fvar0 + (LTA * LTA).sum(0))
    return torch.stack(means, 1), torch.stack(variances, 1)


def conditional_full_cov(spec: Dict, X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fmean [N, P], fcov [P, N, N]: base_conditional's ``full_cov=True`` branch
    (fvar = Knn - A^T A + LTA^T LTA), reached from GPLayer.predict(full_cov=True)
    (gp_layer.py:257-266) and sampled through math._cholesky_with_jitter (gp_layer.py:329-333)."""
    Kn = spec["Z"].shape[0]
    P = spec["q_mu"].shape[1]
    white = spec.get("whiten", True)
    means, covs = [], []
    for p in range(P):
        k = p if Kn > 1 else 0
        Lm = torch.linalg.cholesky(Kuu(spec, k))
        A = torch.linalg.solve_triangular(Lm, Kuf(spec, k, X), upper=False)
        Knn = kernel_matrix(spec["kernel"], X, X, spec["lengthscales"][k], spec["variance"][k])
        fcov = Knn - A.T @ A
        if not white:
            A = torch.linalg.solve_triangular(Lm.T, A, upper=True)
        LTA = torch.tril(spec["q_sqrt"][p]).T @ A
        means.append(A.T @ spec["q_mu"][:, p])
        covs.append(fcov + LTA.T @ LTA)
    return torch.stack(means, 1) , torch.stack(covs, 0)


def predict(spec: Dict, X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """GPLayer.predict (gp_layer.py:226-268): conditional + mean function."""
    m, v = conditional(spec, X)
    return m + mean_function(spec, X), v


# ----------------------------------------------------------------------------------------------
# KL: gpflow.kullback_leiblers.gauss_kl via prior_kl (gp_layer.py:303-311)
# ----------------------------------------------------------------------------------------------
def prior_kl(spec: Dict) -> torch.Tensor:
    q_mu, q_sqrt = spec["q_mu"], spec["q_sqrt"]
    M, P = q_mu.shape
    Lq = torch.tril(q_sqrt)
    Lq_diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
    white = spec.get("whiten", True)
    if white:
        mahalanobis = (q_mu * q_mu).sum()
        trace = (Lq * Lq).sum()
        logdet_p = q_mu.new_zeros(())
    else:
        Kn = spec["Z"].shape[0]
        mahalanobis = q_mu.new_zeros(())
        trace = q_mu.new_zeros(())
        logdet_p = q_mu.new_zeros(())
        for p in range(P):
            k = p if Kn > 1 else 0
            Lp = torch.linalg.cholesky(Kuu(spec, k))
            alpha = torch.linalg.solve_triangular(Lp, q_mu[:, p : p + 1], upper=False)
            mahalanobis = mahalanobis + (alpha * alpha).sum()
            LpiLq = torch.linalg.solve_triangular(Lp, Lq[p], upper=False)
            trace = trace + (LpiLq * LpiLq).sum()
            logdet_p = logdet_p + torch.log(torch.diagonal(Lp) ** 2).sum()
    twoKL = mahalanobis - M * P - torch.log(Lq_diag * Lq_diag).sum() + trace + logdet_p
    return 0.5 * twoKL


# ----------------------------------------------------------------------------------------------
# sample: MultivariateNormalDiag(loc, sqrt(cov)).sample() with explicit eps (gp_layer.py:339-363)
# ----------------------------------------------------------------------------------------------
def layer_forward(spec: Dict, X: torch.Tensor, eps: Optional[torch.Tensor]):
    """-> (f_sample or mean, mean, var).  eps [N, P] supplied by the caller (TF's RNG stream is
    not reproducible outside TF; 'same seeds' parity means 'same eps', SURVEY §7 hard part 4)."""
    mean, var = predict(spec, X)
    if eps is None:
        return mean, mean, var
    return mean + torch.sqrt(var) * eps, mean, var


# ----------------------------------------------------------------------------------------------
# Gaussian likelihood: gpflow.likelihoods.Gaussian (likelihood_layer.py:84-93)
# ----------------------------------------------------------------------------------------------
def gaussian_variational_expectations(Fmu, Fvar, Y, noise_variance) -> torch.Tensor:
    """-> [N]: sum over output dims of the closed-form E_q[log N(y | f, s2)]."""
    ve = (
        -0.5 * math.log(2.0 * math.pi)
        - 0.5 * torch.log(noise_variance)
        - 0.5 * ((Y - Fmu) ** 2 + Fvar) / noise_variance
    )
    return ve.sum(-1)


# ----------------------------------------------------------------------------------------------
# DeepGP: deep_gp.py:151-231
# ----------------------------------------------------------------------------------------------
def deep_gp_loss(
    specs: Sequence[Dict],
    X: torch.Tensor,
    Y: torch.Tensor,
    eps: Sequence[Optional[torch.Tensor]],
    noise_variance: torch.Tensor,
    num_data: int,
) -> Tuple[torch.Tensor, Dict]:
    """The Keras training loss: sum_l KL_l/num_data + mean_n(-VE_n)   (gp_layer.py:286-289,
    likelihood_layer.py:87-89).  Hidden layers are sampled with eps[l] (deep_gp.py:176-181); the
    last layer's (mean, var) feed the likelihood (likelihood_layer.py:84-90).
    ELBO = -loss * num_data (deep_gp.py:226-231)."""
    h = X
    kls = []
    fm = fv = None
    for i, spec in enumerate(specs):
        last = i == len(specs) - 1
        f, fm, fv = layer_forward(spec, h, None if last else eps[i])
        kls.append(prior_kl(spec))
        h = f
    ve = gaussian_variational_expectations(fm, fv, Y, noise_variance)
    loss = sum(kls) / num_data + (-ve).mean()
    return loss, {"kl": kls, "fmean": fm, "fvar": fv, "ve_sum": ve.sum()}


# ----------------------------------------------------------------------------------------------
# Fourier features: basis_functions/fourier_features/base.py:59-76, utils.py:37-54
# ----------------------------------------------------------------------------------------------
def rff_cosine(X, W, b, lengthscales, variance) -> torch.Tensor:
    """RandomFourierFeaturesCosine.call: sqrt(2 s2 / n) cos((X/l) W^T + b) -> [N, n]
    (random/base.py:147-152, 274-295)."""
    n = W.shape[0]
    return torch.sqrt(2.0 * variance / n) * torch.cos((X / lengthscales) @ W.T + b)


def rff_concat(X, W, lengthscales, variance) -> torch.Tensor:
    """RandomFourierFeatures.call: sqrt(2 s2 / 2n) [sin, cos]((X/l) W^T) -> [N, 2n]
    (random/base.py:155-213)."""
    n = W.shape[0]
    proj = (X / lengthscales) @ W.T
    return torch.sqrt(2.0 * variance / (2 * n)) * torch.cat([torch.sin(proj), torch.cos(proj)], -1)


# ----------------------------------------------------------------------------------------------
# Matheron-rule sampling: sampling/sample.py:137-200, math.py:42-62
# ----------------------------------------------------------------------------------------------
def compute_A_inv_b(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """gpflux/math.py:42-62: Cholesky + two triangular solves."""
    L = torch.linalg.cholesky(A)
    Linv_b = torch.linalg.solve_triangular(L, b, upper=False)
    return torch.linalg.solve_triangular(L.T, Linv_b, upper=True)


def wilson_setup(spec: Dict, W, b, coeffs, eps_w, eps_u):
    """sample.py:158-181.  Single (kernel, Z) pair (Kmm asserted [M, M], sample.py:170).
    W [L, D], b [1, L] define the cosine features; coeffs [L, 1] the feature coefficients;
    eps_w [L, P], eps_u [P, M, 1] are the N(0,1) draws the reference takes from tf.random.normal.
    -> prior_weights [L, P], v [M, P]."""
    q_mu, q_sqrt = spec["q_mu"], spec["q_sqrt"]
    prior_weights = torch.sqrt(coeffs) * eps_w  # [L, P]
    u_noise = q_sqrt @ eps_u  # [P, M, 1]  (NB: reference multiplies the *full* q_sqrt)
    Kmm = Kuu(spec, 0)
    u = q_mu + u_noise[..., 0].T
    if spec.get("whiten", True):
        u = torch.linalg.cholesky(Kmm) @ u
    phi_Z = rff_cosine(spec["Z"][0], W, b, spec["lengthscales"][0], spec["variance"][0])
    v = compute_A_inv_b(Kmm, u - phi_Z @ prior_weights)
    return prior_weights, v


def wilson_eval(spec: Dict, W, b, prior_weights, v, X) -> torch.Tensor:
    """WilsonSample.__call__ (sample.py:184-198) + mean function (gp_layer.py:383, sample.py:72-82)."""
    phi_X = rff_cosine(X, W, b, spec["lengthscales"][0], spec["variance"][0])
    Knm = Kuf(spec, 0, X).T
    return phi_X @ prior_weights + Knm @ v + mean_function(spec, X)


def draw_conditional_sample(mean, cov, f_old, eps, jitter: float = DEFAULT_JITTER) -> torch.Tensor:
    """gpflux/sampling/utils.py:27-78 with gpflow's sample_mvn(full_cov=True) and
    math._cholesky_with_jitter: a draw of f_new | f_old for the joint N(mean [P,N+M], cov [P,N+M,N+M]).
    ``eps`` [P, M] is the N(0,1) draw."""
    P, N = f_old.shape
    M = mean.shape[-1] - N
    eye = lambda n: jitter * torch.eye(n, dtype=cov.dtype)  # noqa: E731
    cov_new = cov[:, N:, N:]
    mean_new = mean[:, N:]
    if N > 0:
        L_old = torch.linalg.cholesky(cov[:, :N, :N] + eye(N))
        A = torch.linalg.solve_triangular(L_old, cov[:, :N, N:], upper=False)
        cov_new = cov_new - A.transpose(-1, -2) @ A
        AM = torch.linalg.solve_triangular(L_old, (f_old - mean[:, :N])[:, :, None], upper=False)
        mean_new = mean_new + (A.transpose(-1, -2) @ AM)[:, :, 0]
    chol = torch.linalg.cholesky(cov_new + eye(M))
    return mean_new + (chol @ eps[:, :, None])[:, :, 0]


def conditional_sample_sequence(spec: Dict, Xs: Sequence[torch.Tensor], epss: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """_efficient_sample_conditional_gaussian (sample.py:85-134): successive evaluations of ONE
    function draw at the point sets ``Xs`` (each conditioned on everything evaluated before)."""
    P = spec["q_mu"].shape[1]
    X_all = None
    f = torch.zeros((0, P), dtype=torch.float64)
    out = []
    for X_new, eps in zip(Xs, epss):
        X_all = X_new if X_all is None else torch.cat([X_all, X_new], 0)
        mean, cov = conditional_full_cov(spec, X_all)
        f_new = draw_conditional_sample(mean.T, cov, f.T, eps, spec.get("jitter", DEFAULT_JITTER)).T
        f = torch.cat([f, f_new], 0)
        out.append(f_new)
    return out


# ----------------------------------------------------------------------------------------------
# helpers used by tests / bench to build deterministic specs
# ----------------------------------------------------------------------------------------------
def make_spec(
    rng,
    M: int,
    D: int,
    P: int,
    K: int = 1,
    kernel: str = "se",
    mean: str = "zero",
    whiten: bool = True,
    variance: float = 1.0,
    ls_lo: float = 1.0,
    ls_hi: float = 2.0,
    q_sqrt_diag: float = 0.5,
    q_sqrt_off: float = 0.05,
) -> Dict:
    """Seeded random layer spec (numpy Generator ``rng``)."""
    import numpy as np

    spec = {
        "kernel": kernel,
        "Z": torch.from_numpy(rng.standard_normal((K, M, D))),
        "lengthscales": torch.from_numpy(rng.uniform(ls_lo, ls_hi, (K, D))),
        "variance": torch.from_numpy(variance * rng.uniform(0.8, 1.2, (K,))),
        "q_mu": torch.from_numpy(0.3 * rng.standard_normal((M, P))),
        "q_sqrt": torch.from_numpy(
            np.stack(
                [
                    q_sqrt_diag * np.eye(M) * rng.uniform(0.5, 1.5, (M,))[None, :]
                    + q_sqrt_off * np.tril(rng.standard_normal((M, M)), -1)
                    for _ in range(P)
                ]
            )
        ),
        "mean": mean,
        "whiten": whiten,
        "jitter": DEFAULT_JITTER,
    }
    if mean == "linear":
        spec["mean_A"] = torch.from_numpy(rng.standard_normal((D, P)) / math.sqrt(D))
        spec["mean_b"] = torch.from_numpy(0.1 * rng.standard_normal((P,)))
    return spec


PARAM_KEYS = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")


def spec_requires_grad(spec: Dict, keys=PARAM_KEYS) -> Dict:
    out = dict(spec)
    for k in keys:
        out[k] = spec[k].detach().clone().requires_grad_(True)
    return out


# ----------------------------------------------------------------------------------------------
# LatentVariableLayer: gpflux/layers/latent_variable_layer.py:140-228 (diagonal Gaussians)
# ----------------------------------------------------------------------------------------------
def latent_layer(X, mean, std, prior_mean, prior_std, eps):
    """-> (concat(X, mean + std * eps), sum_n KL[N(mean, std^2) || N(prior_mean, prior_std^2)]);
    the KL is tfp's kl_divergence of diagonal normals (:221-228)."""
    w = mean + std * eps
    kl = (torch.log(prior_std / std) + (std**2 + (mean - prior_mean) ** 2) / (2.0 * prior_std**2) - 0.5).sum()
    return torch.cat([X, w], dim=-1), kl
