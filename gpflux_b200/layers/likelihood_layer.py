"""``LikelihoodLayer``: mirror of gpflux/layers/likelihood_layer.py:34-140.

The Gaussian closed form (variational expectations and their gradients) runs on the device through
``gpfx_gaussian_ve``; non-Gaussian likelihoods are out of scope."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from ..gpflow_compat import Gaussian, Likelihood, default_float


class LikelihoodOutputs:
    """Mirror of gpflux.layers.likelihood_layer.LikelihoodOutputs (:102-140): the mean and variance
    of the last GP layer's marginals and, when not training, of ``y``."""

    def __init__(self, f_mean, f_var, y_mean, y_var):
        self.f_mean = f_mean
        self.f_var = f_var
        self.y_mean = y_mean
        self.y_var = y_var

    def _tensor(self) -> torch.Tensor:
        return self.f_mean

    @property
    def shape(self):
        return self.f_mean.shape

    @property
    def dtype(self):
        return self.f_mean.dtype


class LikelihoodLayer(nn.Module):
    def __init__(self, likelihood: Likelihood):
        super().__init__()
        if not isinstance(likelihood, Gaussian):
            raise NotImplementedError("only gpflow.likelihoods.Gaussian is implemented on the B200 path")
        self.likelihood = likelihood
        self.losses = []

    def call(self, inputs, targets: Optional[torch.Tensor] = None, training: Optional[bool] = None) -> LikelihoodOutputs:
        """Training: adds mean(-variational_expectations) as the layer loss (:84-90); otherwise
        computes predict_mean_and_var (:91-93)."""
        F_mean = inputs.loc
        F_var = inputs.variance()
        if training:
            assert targets is not None
            ve_sum = self.likelihood.variational_expectations_sum(F_mean, F_var, targets)
            loss_per_datapoint = ve_sum / float(-F_mean.shape[0])  # = -mean(VE), one kernel
            Y_mean = Y_var = None
        else:
            loss_per_datapoint = torch.zeros((), dtype=default_float(), device=F_mean.device)
            Y_mean, Y_var = self.likelihood.predict_mean_and_var(F_mean, F_var)
        self.losses = [loss_per_datapoint]
        return LikelihoodOutputs(F_mean, F_var, Y_mean, Y_var)

    forward = call
