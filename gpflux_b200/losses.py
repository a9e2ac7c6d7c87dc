"""``LikelihoodLoss``: mirror of gpflux/losses.py:33-91 - the Keras-loss form of the
likelihood term, for models that end in a plain ``GPLayer`` instead of a ``LikelihoodLayer``.

``loss(y_true, f_prediction)`` returns what Keras' ``Loss.__call__`` returns for the reference class
(the batch mean of ``call``):  ``-mean_n E_q(f_n)[log p(y_n | f_n)]`` when the prediction is the
layer's ``MultivariateNormalDiag``, ``-mean_n log p(y_n | f_n)`` when it is a tensor of samples.  Both
run on the device through ``gpfx_gaussian_ve`` (the log-density is the variational expectation at
zero variance); only ``gpflow.likelihoods.Gaussian`` is implemented on the B200 path."""
from __future__ import annotations

import torch

from .gpflow_compat import Gaussian, Likelihood
from .layers.gp_layer import MultivariateNormalDiag


class LikelihoodLoss:
    def __init__(self, likelihood: Likelihood):
        if not isinstance(likelihood, Gaussian):
            raise NotImplementedError("only gpflow.likelihoods.Gaussian is implemented on the B200 path")
        self.likelihood = likelihood

    def __call__(self, y_true: torch.Tensor, f_prediction) -> torch.Tensor:
        if isinstance(f_prediction, MultivariateNormalDiag):
            F_mu, F_var = f_prediction.loc, f_prediction.variance()
        else:  # a tensor of samples: -log p(y | f)
            F_mu = f_prediction
            F_var = torch.zeros_like(F_mu)
        ve_sum = self.likelihood.variational_expectations_sum(F_mu, F_var, y_true)
        return ve_sum / float(-F_mu.shape[0])

    def call(self, y_true, f_prediction):
        """The reference returns the per-datapoint vector here and lets Keras average it; the device
        kernel reduces in the same pass, so only the averaged value (``__call__``) is exposed."""
        raise NotImplementedError("use LikelihoodLoss.__call__: the per-datapoint vector is reduced on the device")
