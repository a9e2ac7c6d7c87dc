"""Generates tests/golden/layer_cases.npz: seeded inputs and the oracle's outputs/gradients for a
few small GPLayer cases.  The reference (GPflux -> GPflow -> TensorFlow) cannot be imported here
(SURVEY.md fact 2), so these vectors pin the ORACLE (regression + cross-implementation check), not
the reference: "parity unpinned".  snelson1d.npz is the reference's own data fixture
(tests/snelson1d.npz), copied unchanged.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import svgp_torch as ot  # noqa: E402

CASES = [
    # name, N, M, D, P, K, kernel, mean
    ("se_shared", 64, 24, 3, 2, 1, "se", "zero"),
    ("se_separate_identity", 50, 17, 2, 2, 2, "se", "identity"),
    ("m32_shared_linear", 40, 20, 4, 3, 1, "matern32", "linear"),
    ("m52_separate", 33, 16, 1, 2, 2, "matern52", "zero"),
    ("m12_shared", 30, 12, 2, 1, 1, "matern12", "zero"),
]


def main():
    out = {}
    for idx, (name, N, M, D, P, K, kernel, mean) in enumerate(CASES):
        rng = np.random.default_rng(1000 + idx)
        spec = ot.make_spec(rng, M, D, P, K=K, kernel=kernel, mean=mean)
        X = torch.from_numpy(rng.standard_normal((N, D)))
        eps = torch.from_numpy(rng.standard_normal((N, P)))
        w = torch.from_numpy(rng.standard_normal((3, N, P)))
        s = ot.spec_requires_grad(spec)
        Xr = X.clone().requires_grad_(True)
        f, m, v = ot.layer_forward(s, Xr, eps)
        kl = ot.prior_kl(s)
        loss = (f * w[0]).sum() + (m * w[1]).sum() + (v * w[2]).sum() + 0.37 * kl
        loss.backward()
        for k in ("Z", "lengthscales", "variance", "q_mu", "q_sqrt"):
            out[f"{name}/in/{k}"] = spec[k].numpy()
            out[f"{name}/grad/{k}"] = s[k].grad.numpy()
        if mean == "linear":
            out[f"{name}/in/mean_A"] = spec["mean_A"].numpy()
            out[f"{name}/in/mean_b"] = spec["mean_b"].numpy()
        out[f"{name}/in/X"] = X.numpy()
        out[f"{name}/in/eps"] = eps.numpy()
        out[f"{name}/in/w"] = w.numpy()
        out[f"{name}/out/fsample"] = f.detach().numpy()
        out[f"{name}/out/fmean"] = m.detach().numpy()
        out[f"{name}/out/fvar"] = v.detach().numpy()
        out[f"{name}/out/kl"] = kl.detach().numpy()
        out[f"{name}/grad/X"] = Xr.grad.numpy()
        out[f"{name}/meta"] = np.array([N, M, D, P, K])
        out[f"{name}/kernel"] = np.array(kernel)
        out[f"{name}/mean"] = np.array(mean)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "layer_cases.npz"), **out)
    print("wrote layer_cases.npz with", len(CASES), "cases")


if __name__ == "__main__":
    main()
