"""Generates tests/golden/next_cases.npz: seeded inputs and the oracle's outputs for the SURVEY §8f
rows - natural-gradient step, full-covariance conditional, conditional-Gaussian sampler, latent
variable layer.  Like layer_cases.npz these pin the ORACLE ("parity unpinned": the reference cannot
be imported here), and are what the CUDA path is checked against on the GPU box.

Run from the repo root:  python tests/golden/make_golden_next.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import natgrad_torch as ng  # noqa: E402
from oracle import svgp_torch as ot  # noqa: E402


def main():
    out = {}
    # natural-gradient step (XiNat)
    rng = np.random.default_rng(2000)
    M, P, gamma = 33, 2, 0.4
    q_mu = rng.standard_normal((M, P))
    q_sqrt = np.stack([np.tril(0.1 * rng.standard_normal((M, M))) / np.sqrt(M) + 0.6 * np.eye(M) for _ in range(P)])
    gm = rng.standard_normal((M, P))
    G = np.stack([np.tril(0.1 * rng.standard_normal((M, M))) / np.sqrt(M) for _ in range(P)])
    mu1, sq1 = ng.natgrad_step(*(torch.from_numpy(a) for a in (q_mu, q_sqrt, gm, G)), gamma)
    out.update({"natgrad/q_mu": q_mu, "natgrad/q_sqrt": q_sqrt, "natgrad/dq_mu": gm, "natgrad/dq_sqrt": G,
                "natgrad/gamma": np.array(gamma), "natgrad/q_mu_new": mu1.numpy(), "natgrad/q_sqrt_new": sq1.numpy()})
    # full-covariance conditional (whitened, separate kernels) and conditional-Gaussian sampler (shared)
    rng = np.random.default_rng(2001)
    spec = ot.make_spec(rng, 14, 2, 2, K=2, q_sqrt_diag=0.3)
    X = torch.from_numpy(rng.standard_normal((40, 2)))
    m, c = ot.conditional_full_cov(spec, X)
    for k in ("Z", "lengthscales", "variance", "q_mu", "q_sqrt"):
        out[f"fullcov/{k}"] = spec[k].numpy()
    out.update({"fullcov/X": X.numpy(), "fullcov/mean": m.numpy(), "fullcov/cov": c.numpy()})
    rng = np.random.default_rng(2002)
    spec = ot.make_spec(rng, 10, 1, 2, K=1, q_sqrt_diag=0.05, q_sqrt_off=0.0)
    Xs = [torch.from_numpy(rng.uniform(-1, 1, (n, 1))) for n in (20, 9)]
    epss = [torch.from_numpy(rng.standard_normal((2, x.shape[0]))) for x in Xs]
    fs = ot.conditional_sample_sequence(spec, Xs, epss)
    for k in ("Z", "lengthscales", "variance", "q_mu", "q_sqrt"):
        out[f"condsample/{k}"] = spec[k].numpy()
    for i in range(2):
        out[f"condsample/X{i}"] = Xs[i].numpy()
        out[f"condsample/eps{i}"] = epss[i].numpy()
        out[f"condsample/f{i}"] = fs[i].numpy()
    # latent variable layer
    rng = np.random.default_rng(2003)
    N, D, W = 50, 1, 2
    Xl = torch.from_numpy(rng.standard_normal((N, D)))
    mean = torch.from_numpy(rng.standard_normal((N, W)))
    std = torch.from_numpy(rng.uniform(0.2, 1.5, (N, W)))
    pm, ps = torch.from_numpy(rng.standard_normal(W)), torch.from_numpy(rng.uniform(0.5, 1.5, W))
    eps = torch.from_numpy(rng.standard_normal((N, W)))
    o, kl = ot.latent_layer(Xl, mean, std, pm, ps, eps)
    out.update({"latent/X": Xl.numpy(), "latent/mean": mean.numpy(), "latent/std": std.numpy(), "latent/prior_mean": pm.numpy(),
                "latent/prior_std": ps.numpy(), "latent/eps": eps.numpy(), "latent/out": o.numpy(), "latent/kl": kl.numpy()})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "next_cases.npz"), **out)
    print("wrote next_cases.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
