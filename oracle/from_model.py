"""Builds oracle layer specs (CPU float64 tensors) from a gpflux_b200 DeepGP / GPLayer, so that the
checker evaluates exactly the parameters the CUDA path was given.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

from typing import Dict, List

import torch


def spec_from_layer(layer) -> Dict:
    from gpflux_b200 import gpflow_compat as gc

    ivs = layer._inducing_points()
    D = int(ivs[0].Z.shape[1])
    kind, Z, ls, var = layer._packed(D)
    mf = layer.mean_function
    spec = {
        "kernel": kind,
        "Z": Z.detach().cpu().clone(),
        "lengthscales": ls.detach().cpu().clone(),
        "variance": var.detach().cpu().clone(),
        "q_mu": layer.q_mu.value().detach().cpu().clone(),
        "q_sqrt": layer.q_sqrt.value().detach().cpu().clone(),
        "whiten": layer.whiten,
        "jitter": gc.default_jitter(),
    }
    if isinstance(mf, gc.Identity):
        spec["mean"] = "identity"
    elif isinstance(mf, gc.Zero):
        spec["mean"] = "zero"
    elif isinstance(mf, gc.Linear):
        spec["mean"] = "linear"
        spec["mean_A"] = mf.A.value().detach().cpu().clone()
        spec["mean_b"] = mf.b.value().detach().cpu().clone()
    else:
        raise NotImplementedError(type(mf))
    return spec


def specs_from_model(model) -> List[Dict]:
    return [spec_from_layer(layer) for layer in model.f_layers]


def noise_variance_from_model(model) -> torch.Tensor:
    return model.likelihood_layer.likelihood.variance.value().detach().cpu().clone()
