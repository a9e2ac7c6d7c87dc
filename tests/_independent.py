"""Loader for tests/golden/independent_cases.npz (made by tests/golden/make_independent.py: mpmath,
50 digits, dense closed form + finite-difference gradients - shares no code with oracle/)."""
import os

import numpy as np
import torch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "independent_cases.npz")
PARAMS = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")


def load_cases():
    d = np.load(PATH)
    cases = []
    for name in d["names"]:
        name = str(name)
        g = lambda k: d[f"{name}/{k}"]  # noqa: E731
        kind, white = g("meta")
        spec = {k: torch.from_numpy(np.array(g(k))) for k in PARAMS}
        spec.update(kernel=str(kind), whiten=bool(int(white)), mean="zero", jitter=1e-6)
        case = dict(name=name, spec=spec, X=torch.from_numpy(np.array(g("X"))),
                    w_m=torch.from_numpy(np.array(g("w_m"))), w_v=torch.from_numpy(np.array(g("w_v"))),
                    w_kl=float(g("w_kl")),
                    fmean=torch.from_numpy(np.array(g("fmean"))), fvar=torch.from_numpy(np.array(g("fvar"))),
                    kl=torch.tensor(float(g("kl")), dtype=torch.float64),
                    loss=torch.tensor(float(g("loss")), dtype=torch.float64),
                    grads={k: torch.from_numpy(np.array(g("grad_" + k))) for k in PARAMS + ("X",)})
        cases.append(case)
    return cases


def case_names():
    return [str(n) for n in np.load(PATH)["names"]]
