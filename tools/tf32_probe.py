"""Diagnostic: error of the float32 opt-in against the float64 oracle per output / gradient, for the
contraction variants (GPFX_FP32_TF32=0: float64 contraction; GPFX_TF32_PRODUCTS=mask: which products
take the 3xTF32 kernel).  python tools/tf32_probe.py [f64]"""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from gpflux_b200 import _cabi, gpflow_compat as gf, ops  # noqa: E402
from oracle import svgp_torch as ot  # noqa: E402

use32 = "f64" not in sys.argv
if use32:
    gf.set_default_float(np.float32)
rng = np.random.default_rng(9)
N, M, D, P = 1024, 128, 4, 4
spec = ot.make_spec(rng, M, D, P, K=P, kernel="se", mean="identity", whiten=True)
spec32 = {k: (v.float().double() if isinstance(v, torch.Tensor) and v.is_floating_point() else v) for k, v in spec.items()}
X = torch.from_numpy(rng.standard_normal((N, D))).float()
eps = torch.from_numpy(rng.standard_normal((N, P))).float()
w_f, w_m, w_v = (torch.from_numpy(rng.standard_normal((N, P))).float() for _ in range(3))
names = ("Z", "lengthscales", "variance", "q_mu", "q_sqrt")
so = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in spec32.items()}
Xo = X.double().clone().requires_grad_(True)
f_ref, m_ref, v_ref = ot.layer_forward(so, Xo, eps.double())
kl_ref = ot.prior_kl(so)
((f_ref * w_f.double()).sum() + (m_ref * w_m.double()).sum() + (v_ref * w_v.double()).sum() + 0.37 * kl_ref).backward()
cast = (lambda t: t.float()) if use32 else (lambda t: t.double())
p = {k: cast(spec32[k]).cuda().requires_grad_(True) for k in names}
Xg = cast(X).cuda().requires_grad_(True)
f, m, v, kl = ops.svgp_layer(Xg, p["Z"], p["lengthscales"], p["variance"], p["q_mu"], p["q_sqrt"], mean_add=Xg,
                             eps=cast(eps).cuda(), kernel="se", whiten=True, jitter=spec["jitter"])
((f * cast(w_f).cuda()).sum() + (m * cast(w_m).cuda()).sum() + (v * cast(w_v).cuda()).sum() + 0.37 * kl).backward()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


print("float32" if use32 else "float64", "contraction mode", _cabi.load().gpfx_get_contraction(),
      " ".join(f"{n}={rel(x, y):.1e}" for n, x, y in [("m", m, m_ref), ("v", v, v_ref), ("kl", kl, kl_ref)]
               + [(k, p[k].grad, so[k].grad) for k in names] + [("X", Xg.grad, Xo.grad)]),
      "cond(Kuu)=%.1e" % max(float(torch.linalg.cond(ot.Kuu(spec32, k))) for k in range(P)))
