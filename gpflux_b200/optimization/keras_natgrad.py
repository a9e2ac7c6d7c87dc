"""``NatGradModel`` / ``NatGradWrapper``: mirror of gpflux/optimization/keras_natgrad.py:37-265
(SURVEY §8f-3).

The reference subclasses ``tf.keras.Model`` and overrides ``train_step`` so that the q(u)
parameters of the ``natgrad_layers`` are updated by ``gpflow.optimizers.NaturalGradient`` and
everything else by a regular Keras optimiser.  Here the Keras model is the ``TrainingModel``
stand-in of ``gpflux_b200.models.deep_gp``; the natural-gradient arithmetic runs on the device in
``gpfx_natgrad_step`` (no host maths)."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from ..gpflow_compat import NaturalGradient, Parameter
from ..layers.gp_layer import GPLayer
from ..models.deep_gp import TrainingModel


class NatGradModel(TrainingModel):
    """Drop-in for the training model that trains ``natgrad_layers`` with natural gradients.
    ``compile`` takes a list: one ``NaturalGradient`` per natgrad layer followed by one regular
    optimiser (a ``torch.optim.Optimizer`` or a factory ``params -> optimizer``) for all other
    parameters (keras_natgrad.py:48-53)."""

    def __init__(self, deep_gp):
        super().__init__(deep_gp)
        self._all_optimizers = None

    # -- natgrad_layers (keras_natgrad.py:55-86)
    @property
    def natgrad_layers(self) -> List[GPLayer]:
        if not hasattr(self, "_natgrad_layers"):
            raise AttributeError(f"natgrad_layers must be set before training {self.__class__.__name__}")
        return self._natgrad_layers

    @natgrad_layers.setter
    def natgrad_layers(self, layers: Union[List[GPLayer], bool]) -> None:
        if isinstance(layers, bool):
            self._natgrad_layers = [l for l in self.deep_gp.f_layers if isinstance(l, GPLayer)] if layers else []
        else:
            self._natgrad_layers = list(layers)

    @property
    def natgrad_optimizers(self) -> Optional[List[NaturalGradient]]:
        return None if self._all_optimizers is None else self._all_optimizers[:-1]

    @property
    def optimizer(self):
        """The last optimiser (the hack of keras_natgrad.py:98-112)."""
        return None if getattr(self, "_all_optimizers", None) is None else self._all_optimizers[-1]

    @optimizer.setter
    def optimizer(self, optimizers) -> None:
        if optimizers is None:
            self._all_optimizers = None
            return
        if optimizers is self.optimizer:
            return
        if not isinstance(optimizers, (tuple, list)):
            raise TypeError(
                "`optimizer` needs to be a list of NaturalGradient optimizers for each element of "
                f"`natgrad_layers` followed by one optimizer for all other parameters, but was {optimizers}"
            )
        if isinstance(optimizers[-1], NaturalGradient):
            raise TypeError(
                "The last element of the optimizer list applies to the non-variational parameters and "
                f"cannot be a NaturalGradient optimizer, but was {optimizers[-1]}"
            )
        if not all(isinstance(o, NaturalGradient) for o in optimizers[:-1]):
            raise TypeError(
                "The all-but-last elements of the optimizer list must be NaturalGradient instances, one for "
                f"each element of natgrad_layers, but were {optimizers[:-1]}"
            )
        self._all_optimizers = list(optimizers)

    def compile(self, optimizer=None, **_):
        self.optimizer = optimizer

    def _split_natgrad_params_and_other_vars(self) -> Tuple[List[Tuple[Parameter, Parameter]], List[torch.nn.Parameter]]:
        """keras_natgrad.py:147-163."""
        variational_params = [(layer.q_mu, layer.q_sqrt) for layer in self.natgrad_layers]
        variational_ids = {id(p.unconstrained_variable) for vp in variational_params for p in vp}
        other_vars = [v for v in self.deep_gp.parameters() if v.requires_grad and id(v) not in variational_ids]
        return variational_params, other_vars

    def _apply_backwards_pass(self, loss: torch.Tensor) -> None:
        """keras_natgrad.py:165-193."""
        num_natgrad_layers = len(self.natgrad_layers)
        num_natgrad_opt = len(self.natgrad_optimizers)
        if num_natgrad_opt != num_natgrad_layers:
            raise ValueError(
                f"Model has {num_natgrad_opt} NaturalGradient optimizers, but {num_natgrad_layers} "
                "variational distributions in natgrad_layers"
            )
        variational_params, other_vars = self._split_natgrad_params_and_other_vars()
        variational_vars = [p.unconstrained_variable for vp in variational_params for p in vp]
        grads = torch.autograd.grad(loss, variational_vars + other_vars, allow_unused=True)
        vgrads, other_grads = grads[: len(variational_vars)], grads[len(variational_vars):]
        for i, (opt, (q_mu, q_sqrt)) in enumerate(zip(self.natgrad_optimizers, variational_params)):
            opt._natgrad_apply_gradients(vgrads[2 * i], vgrads[2 * i + 1], q_mu, q_sqrt)
        last = self._all_optimizers[-1]
        if callable(last) and not isinstance(last, torch.optim.Optimizer):
            last = self._all_optimizers[-1] = last(other_vars)
        for v, g in zip(other_vars, other_grads):
            v.grad = g
        last.step()

    def train_step(self, inputs, targets) -> torch.Tensor:
        """keras_natgrad.py:195-216."""
        loss = self.deep_gp.loss(inputs, targets)
        self._apply_backwards_pass(loss)
        return loss.detach()

    def check_status(self) -> None:
        """Non-PD Kuu in any layer or a failed factorisation inside the last natural-gradient step -> raise
        (the reference raises from tf.linalg.cholesky).  Reads a few int32 back: call it where the loss is
        read (e.g. once per logging interval), not inside the step."""
        self.deep_gp.check_status()
        for opt in self.natgrad_optimizers:
            opt.check_status()


class NatGradWrapper(NatGradModel):
    """Wraps the model returned by ``DeepGP.as_training_model`` (keras_natgrad.py:219-265)."""

    def __init__(self, base_model: TrainingModel):
        super().__init__(base_model.deep_gp)
        self.base_model = base_model
