"""The synthetic workloads of BASELINE.json / SURVEY §8(d), built once here so that ``bench.py``,
the ``tools/`` scripts and the parity tests evaluate exactly the same models.

Host-side construction only (numpy + the reference-shaped builders); no arithmetic of the hot path.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

N_DATA_C2 = 1_000_000
N_DATA_C3 = 1_000_000

C2 = dict(name="c2", layers=2, M=256, D=8, P=8, batch=8192)
C3 = dict(name="c3", layers=5, M=1024, D=32, P=32, batch=8192)

C2_WORKLOAD = ("c2: 2-layer DeepGP (build_constant_input_dim_deep_gp), D=8, N=1M synthetic, "
               "minibatch 8192/GPU, M=256, float64")
C3_WORKLOAD = ("c3: 5-layer DeepGP, M=1024, SeparateIndependent P=32 (32 SE kernels + 32 inducing sets per hidden "
               "layer, P=1 last layer), D=32 synthetic, minibatch 8192 rows/GPU (65536 at 8 GPUs), float64")


def f_fwd(N, M, D, P, K):
    """algorithmic flops of one layer forward (triangular work counted once)."""
    return (2 * K * M * M * D + 2 * K * M * N * D + K * M**3 / 3 + K * M * M * N + 2 * M * N * P
            + 2 * K * M * N + P * M * M * N + 2 * P * M * N + 2 * P * (M * M + 2 * M))


def layer_shapes(cfg: dict, batch: int):
    """[(N, M, D, P, K)] per layer of a configuration."""
    L, M, D, P = cfg["layers"], cfg["M"], cfg["D"], cfg["P"]
    if cfg["name"] == "c2":
        return [(batch, M, D, P, 1)] * (L - 1) + [(batch, M, D, 1, 1)]
    return [(batch, M, D, P, P)] * (L - 1) + [(batch, M, D, 1, 1)]


def algorithmic_flops_per_step(cfg: dict, batch: int) -> float:
    """fwd + bwd = 3 x sum of F_fwd (the SURVEY §8(d) convention)."""
    return 3.0 * sum(f_fwd(*s) for s in layer_shapes(cfg, batch))


# ------------------------------------------------------------------------------------------- c2
def make_data_c2(n: int = N_DATA_C2) -> Tuple[np.ndarray, np.ndarray]:
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, C2["D"]))
    Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((n, 1))
    return X, Y


def build_c2(X: np.ndarray, layers: int = C2["layers"], M: int = C2["M"]):
    """BASELINE configs[1]: exactly ``build_constant_input_dim_deep_gp`` with the values of the
    reference's benchmarking/main.py:78-83, Z = first M rows (k-means skipped)."""
    import torch

    from .architectures import Config, build_constant_input_dim_deep_gp

    torch.manual_seed(0)
    cfg = Config(num_inducing=M, inner_layer_qsqrt_factor=1e-5, likelihood_noise_variance=1e-2)
    model = build_constant_input_dim_deep_gp(X, layers, cfg, z_init=X[:M])
    # move q(u) away from the exact initial point so that no term of the step is trivially zero
    rng = np.random.default_rng(1)
    for layer in model.f_layers:
        Ml, P = layer.q_mu.shape
        layer.q_mu.assign(0.1 * rng.standard_normal((Ml, P)))
        q = layer.q_sqrt.value().detach().cpu().numpy()
        scale = float(np.abs(np.diagonal(q, axis1=1, axis2=2)).mean())
        layer.q_sqrt.assign(q + 0.05 * scale * np.tril(rng.standard_normal(q.shape), -1))
    return model


# ------------------------------------------------------------------------------------------- c3
def make_data_c3(n: int, D: int = C3["D"], seed: int = 100) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic UCI-shaped rows: X ~ N(0,1) [n, D], Y = sin(mean(X)) * 2 + 0.1 noise."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D))
    Y = 2.0 * np.sin(X.sum(1, keepdims=True) / np.sqrt(D)) + 0.1 * rng.standard_normal((n, 1))
    return X, Y


def build_c3(layers: int = C3["layers"], M: int = C3["M"], P: int = C3["P"], D: int = C3["D"],
             num_data: int = N_DATA_C3, seed: int = 0):
    """BASELINE configs[2] as SURVEY §8(d) spells it out: ``layers - 1`` hidden GPLayers with
    ``SeparateIndependent`` of P SquaredExponential kernels (ARD lengthscales ~ U(1,3) * sqrt(D)/2 so that
    ||x - z|| / l = O(1), variance 1e-2) and ``SeparateIndependentInducingVariables`` (P sets of
    Z ~ N(0,1) [M, D]), Identity mean; a last P=1 layer (variance 1, Zero mean); q_mu ~ 0.1 N(0,1),
    q_sqrt = 1e-2 I + 1e-3 tril(N(0,1)); Gaussian likelihood, noise variance 1e-2."""
    from . import gpflow_compat as gf
    from .layers import GPLayer, LikelihoodLayer
    from .models import DeepGP

    if P != D:
        raise ValueError("hidden layers use the Identity mean function: P must equal D")
    rng = np.random.default_rng(seed)
    f_layers = []
    for i in range(layers):
        last = i == layers - 1
        Pl = 1 if last else P
        kernels = [gf.SquaredExponential(variance=1.0 if last else 1e-2,
                                         lengthscales=rng.uniform(1.0, 3.0, D) * np.sqrt(D) / 2)
                   for _ in range(Pl)]
        ivs = [gf.InducingPoints(rng.standard_normal((M, D))) for _ in range(Pl)]
        kern = gf.SeparateIndependent(kernels) if Pl > 1 else gf.SharedIndependent(kernels[0], 1)
        iv = (gf.SeparateIndependentInducingVariables(ivs) if Pl > 1
              else gf.SharedIndependentInducingVariables(ivs[0]))
        layer = GPLayer(kern, iv, num_data=num_data, mean_function=gf.Zero() if last else gf.Identity(),
                        name=f"gp_{i}")
        layer.q_mu.assign(0.1 * rng.standard_normal((M, Pl)))
        q = np.stack([1e-2 * np.eye(M) + 1e-3 * np.tril(rng.standard_normal((M, M)), -1) for _ in range(Pl)])
        layer.q_sqrt.assign(q)
        f_layers.append(layer)
    return DeepGP(f_layers, LikelihoodLayer(gf.Gaussian(0.01)))
