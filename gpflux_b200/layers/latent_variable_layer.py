"""``LatentVariableLayer``: mirror of gpflux/layers/latent_variable_layer.py:32-228,
the step before the GP path in conditional-density DGPs (BASELINE config 4).

The arithmetic - posterior reparameterised sample, diagonal-Gaussian KL[q || p], concatenation with
the layer input, and their backward - is one fused kernel pair behind ``gpfx_latent_forward`` /
``gpfx_latent_backward``.  The encoder is the host's business (any module mapping
``concat([X, Y])`` to ``(means, stds)``); ``DirectlyParameterizedNormalDiag`` mirrors
gpflux/encoders/directly_parameterized_encoder.py:28-114."""
from __future__ import annotations

import abc
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops
from ..gpflow_compat import Parameter, default_float


class MultivariateNormalDiagPrior:
    """Stand-in for the ``tfp.distributions.MultivariateNormalDiag(loc, scale_diag)`` prior the
    reference passes to the layer (only diagonal Gaussians are supported on the B200 path)."""

    def __init__(self, loc, scale_diag):
        self.loc = torch.as_tensor(np.asarray(loc, dtype=np.float64)).reshape(-1)
        self.scale_diag = torch.as_tensor(np.asarray(scale_diag, dtype=np.float64)).reshape(-1)
        if self.loc.shape != self.scale_diag.shape:
            raise ValueError("loc and scale_diag must have the same shape [W]")

    @property
    def event_size(self) -> int:
        return int(self.loc.shape[0])


class LayerWithObservations(nn.Module, metaclass=abc.ABCMeta):
    """Mirror of latent_variable_layer.py:32-55: layers whose call takes ``observations=[X, Y]``."""

    needs_observations = True


class LatentVariableLayer(LayerWithObservations):
    def __init__(self, prior: MultivariateNormalDiagPrior, encoder: nn.Module, compositor=None,
                 name: Optional[str] = None):
        super().__init__()
        if compositor is not None:
            raise NotImplementedError("only the default Concatenate(axis=-1) compositor is implemented")
        self.prior = prior
        self.encoder = encoder
        self.name = name
        self.losses = []
        self.metrics = {}

    def _prior(self, device):
        return self.prior.loc.to(device), self.prior.scale_diag.to(device)

    def call(self, layer_inputs, observations=None, training: Optional[bool] = None, seed: Optional[int] = None,
             eps: Optional[torch.Tensor] = None):
        """Reference latent_variable_layer.py:115-158.  Training: sample the encoder's posterior and
        add mean_n KL[q_n || p] as the layer loss; otherwise sample the prior.  ``eps`` [N, W] pins the
        N(0,1) draw (parity runs)."""
        if hasattr(layer_inputs, "_tensor"):
            layer_inputs = layer_inputs._tensor()
        N = layer_inputs.shape[0]
        W = self.prior.event_size
        dev = layer_inputs.device
        pm, ps = self._prior(dev)
        if eps is None:
            gen = None
            if seed is not None:
                gen = torch.Generator(device=dev)
                gen.manual_seed(seed)
            eps = torch.randn((N, W), dtype=default_float(), device=dev, generator=gen)
        if training:
            if observations is None:
                raise ValueError("LatentVariableLayer requires observations when training")
            means, stds = self._inference_posteriors(observations, training=True)
            out, kl_sum = ops.latent_sample_and_kl(layer_inputs, means, stds, pm, ps, eps)
            loss_per_datapoint = kl_sum / N
        else:
            out, _ = ops.latent_sample_and_kl(layer_inputs, pm, ps, pm, ps, eps)
            loss_per_datapoint = torch.zeros((), dtype=default_float(), device=dev)
        self.losses = [loss_per_datapoint]
        name = f"{self.name}_local_kl" if self.name else "local_kl"
        self.metrics = {name: loss_per_datapoint.detach()}
        return out

    forward = call

    def _inference_posteriors(self, observations, training: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Reference :160-184: the encoder is called on concat([inputs, targets])."""
        encoder_inputs = torch.cat(list(observations), dim=-1)
        means, stds = self.encoder(encoder_inputs)
        return means, stds


class DirectlyParameterizedNormalDiag(nn.Module):
    """Mirror of gpflux/encoders/directly_parameterized_encoder.py: one (mean, std) per datapoint -
    "not compatible with mini-batching" (:42-44), so the batch must be the whole dataset."""

    def __init__(self, num_data: int, latent_dim: int, means: Optional[np.ndarray] = None):
        super().__init__()
        if means is None:
            means = 0.01 * np.random.randn(num_data, latent_dim)
        elif tuple(means.shape) != (num_data, latent_dim):
            raise ValueError(
                f"means must have shape [num_data, latent_dim] = [{num_data}, {latent_dim}]; got {means.shape} instead."
            )
        self.means = Parameter(means, name="w_means")
        self.stds = Parameter(1e-5 * np.ones_like(means), transform="positive", name="w_stds")

    def forward(self, inputs=None, *args, **kwargs):
        if inputs is not None and inputs.shape[0] != self.means.shape[0]:
            raise ValueError(
                f"DirectlyParameterizedNormalDiag holds {self.means.shape[0]} datapoints, got a batch of {inputs.shape[0]}"
            )
        return self.means.value(), self.stds.value()
