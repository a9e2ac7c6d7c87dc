"""``DeepGP``: mirror of gpflux/models/deep_gp.py:32-307 - the caller/driver of the hot path
(SURVEY §8a row a15).  Contains no arithmetic: it chains ``GPLayer`` calls (each one fused
forward on the device), evaluates the likelihood layer and sums the layer losses."""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from ..gpflow_compat import Likelihood, default_float
from ..layers.likelihood_layer import LikelihoodLayer


class DeepGP(nn.Module):
    def __init__(
        self,
        f_layers: List[nn.Module],
        likelihood: Union[LikelihoodLayer, Likelihood],
        *,
        input_dim: Optional[int] = None,
        target_dim: Optional[int] = None,
        default_model_class=None,
        num_data: Optional[int] = None,
    ):
        super().__init__()
        self.input_dim = input_dim
        self.target_dim = target_dim
        self.f_layers = nn.ModuleList(f_layers)
        if isinstance(likelihood, Likelihood):
            self.likelihood_layer = LikelihoodLayer(likelihood)
        else:
            self.likelihood_layer = likelihood
        self.default_model_class = default_model_class
        self.num_data = self._validate_num_data(f_layers, num_data)
        # B200 addition: the parameter-only half of every GPLayer (Kuu, Cholesky, L^-1, KL) does not
        # depend on the minibatch, so it is launched for all layers up front - layer 0 on the current
        # stream, the others on side streams - and overlaps the data path (also as parallel branches
        # when the step is captured in a CUDA graph).
        self.overlap_factorization = True
        self._side_streams = {}

    @staticmethod
    def _validate_num_data(f_layers, num_data: Optional[int] = None) -> int:
        """Reference deep_gp.py:113-135."""
        for i, layer in enumerate(f_layers):
            layer_num_data = getattr(layer, "num_data", None)
            if num_data is None:
                num_data = layer_num_data
            elif layer_num_data is not None and num_data != layer_num_data:
                raise ValueError(f"f_layers[{i}].num_data is inconsistent with num_data={num_data}")
        if num_data is None:
            raise ValueError("Could not determine num_data; please provide explicitly")
        return num_data

    @staticmethod
    def _validate_dtype(x) -> None:
        """Reference deep_gp.py:137-149."""
        if x.dtype != default_float():
            raise ValueError(
                f"x needs to have dtype {default_float()} (according to "
                f"gpflow.default_float()), however got x with {x.dtype} dtype."
            )

    def _evaluate_deep_gp(self, inputs, targets, training: Optional[bool] = None, eps: Optional[Sequence] = None):
        """f(x) = f_n(... f_2(f_1(x))) (reference deep_gp.py:151-181).  A fresh reparameterised
        sample is drawn between every pair of layers (gp_layer.py:358-363); ``eps[i]`` pins it."""
        features = inputs
        observations = [inputs, targets] if targets is not None else None
        self._prefactor(inputs)
        for i, layer in enumerate(self.f_layers):
            kw = {}
            if eps is not None and eps[i] is not None:
                kw["eps"] = eps[i]
            if getattr(layer, "needs_observations", False):
                kw["observations"] = observations
            features = layer(features, training=training, **kw)
        return features

    def _prefactor(self, inputs) -> None:
        if not self.overlap_factorization or len(self.f_layers) < 2 or not inputs.is_cuda:
            return
        dev = inputs.device
        for i, layer in enumerate(self.f_layers):
            if not hasattr(layer, "factorize"):
                continue
            if i == 0:
                # nothing to overlap with: the first layer keeps the fused single-node path (no
                # gradient-accumulation kernels between a factor and a cond node)
                continue
            key = (dev.index, i)
            if key not in self._side_streams:
                # high priority: the factor chains are short kernels that slot in between the CTAs of the
                # data path's large GEMMs
                self._side_streams[key] = torch.cuda.Stream(device=dev, priority=-1)
            layer.factorize(stream=self._side_streams[key])

    def check_status(self) -> None:
        """Raise if any GPLayer met a non-positive-definite Kuu in its last evaluation (one small
        device->host read per layer; see ``GPLayer.check_status``)."""
        for layer in self.f_layers:
            if hasattr(layer, "check_status"):
                layer.check_status()

    def _evaluate_likelihood(self, f_outputs, targets, training: Optional[bool] = None):
        return self.likelihood_layer(f_outputs, targets=targets, training=training)

    @staticmethod
    def _ingest(x, what):
        """torch tensors pass through; any other DLPack producer is validated and viewed zero-copy."""
        if x is None or isinstance(x, torch.Tensor):
            return x
        if hasattr(x, "__dlpack__") or type(x).__name__ == "PyCapsule":
            from ..dlpack import from_dlpack

            return from_dlpack(x, ndim=2, what=what)
        return x

    def call(self, inputs, targets=None, training: Optional[bool] = None, eps: Optional[Sequence] = None):
        inputs, targets = self._ingest(inputs, "inputs"), self._ingest(targets, "targets")
        self._validate_dtype(inputs)
        if targets is not None:
            self._validate_dtype(targets)
        f_outputs = self._evaluate_deep_gp(inputs, targets=targets, training=training, eps=eps)
        return self._evaluate_likelihood(f_outputs, targets=targets, training=training)

    forward = call

    def predict_f(self, inputs) -> Tuple[torch.Tensor, torch.Tensor]:
        """Mean and variance of f (reference deep_gp.py:208-218)."""
        inputs = self._ingest(inputs, "inputs")
        self._validate_dtype(inputs)
        f_distribution = self._evaluate_deep_gp(inputs, targets=None)
        return f_distribution.loc, f_distribution.variance()

    @property
    def losses(self) -> List[torch.Tensor]:
        return [
            loss
            for layer in itertools.chain(self.f_layers, [self.likelihood_layer])
            for loss in getattr(layer, "losses", [])
        ]

    def loss(self, inputs, targets, eps: Optional[Sequence] = None) -> torch.Tensor:
        """The Keras training loss: sum of the layer losses after a training-mode call."""
        self.call(inputs, targets, training=True, eps=eps)
        losses = self.losses
        total = losses[0]
        for term in losses[1:]:  # a handful of scalars: adds beat a stack + reduce
            total = total + term
        return total

    def loss_and_grad_chunked(self, inputs, targets, rows_per_chunk: int, eps: Optional[Sequence] = None) -> torch.Tensor:
        """The training loss of a minibatch that does not fit the device in one piece (the saved Kuf / A / B
        of a layer are 24 M P bytes per row: 45 GB for 8192 rows of BASELINE config 3), evaluated in row chunks.
        Rows are independent given the parameters, and
            loss(X, Y) = sum_c (N_c / N) loss(X_c, Y_c)
        exactly (every chunk loss carries the full KL / num_data term and the weights sum to one), so each
        chunk is evaluated and back-propagated on its own and the gradients accumulate in ``.grad``.
        Returns the detached total loss.  B200 addition: the reference evaluates the whole minibatch."""
        inputs, targets = self._ingest(inputs, "inputs"), self._ingest(targets, "targets")
        n = int(inputs.shape[0])
        if rows_per_chunk <= 0:
            raise ValueError("rows_per_chunk must be positive")
        total = None
        for s0 in range(0, n, rows_per_chunk):
            s1 = min(n, s0 + rows_per_chunk)
            eps_c = None if eps is None else [None if e is None else e[s0:s1] for e in eps]
            part = self.loss(inputs[s0:s1], targets[s0:s1], eps=eps_c) * ((s1 - s0) / n)
            part.backward()
            total = part.detach() if total is None else total + part.detach()
        return total

    def elbo(self, data, eps: Optional[Sequence] = None) -> torch.Tensor:
        """The ELBO, not the per-datapoint loss (reference deep_gp.py:220-231)."""
        X, Y = data
        return -self.loss(X, Y, eps=eps) * self.num_data

    def as_training_model(self, model_class=None):
        """Reference deep_gp.py:239-270.  Keras' functional Model is replaced by a thin trainer
        with the same ``compile(optimizer)`` / ``fit({"inputs": X, "targets": Y}, ...)`` calls."""
        return TrainingModel(self)

    def as_prediction_model(self, model_class=None):
        """Reference deep_gp.py:272-292."""
        return PredictionModel(self)


class TrainingModel:
    """Stand-in for the ``tf.keras.Model`` returned by ``DeepGP.as_training_model``."""

    def __init__(self, deep_gp: DeepGP):
        self.deep_gp = deep_gp
        self.optimizer = None

    def __call__(self, data, training: bool = True):
        return self.deep_gp(data["inputs"], data["targets"], training=training)

    def compile(self, optimizer=None, **_):
        self.optimizer = optimizer or torch.optim.Adam(self.deep_gp.parameters(), lr=0.01)

    def train_step(self, inputs, targets) -> torch.Tensor:
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.deep_gp.loss(inputs, targets)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def fit(self, data, epochs: int = 1, batch_size: Optional[int] = None, verbose: int = 0, shuffle: bool = True, **_):
        if self.optimizer is None:
            self.compile()
        X, Y = data["inputs"], data["targets"]
        n = X.shape[0]
        bs = batch_size or n
        history = {"loss": []}
        for epoch in range(epochs):
            perm = torch.randperm(n, device=X.device) if shuffle else torch.arange(n, device=X.device)
            total, count = 0.0, 0
            for s in range(0, n, bs):
                idx = perm[s : s + bs]
                step_loss = float(self.train_step(X[idx], Y[idx]))  # the step's one device->host read
                if step_loss != step_loss or step_loss in (float("inf"), float("-inf")):
                    self.deep_gp.check_status()  # raises GpfxError(E_NOT_PD) where TF raises from cholesky
                total += step_loss
                count += 1
            history["loss"].append(total / max(count, 1))
            if verbose:
                print(f"epoch {epoch + 1}/{epochs} loss {history['loss'][-1]:.6f}")
        return history


class PredictionModel:
    """Stand-in for the ``tf.keras.Model`` returned by ``DeepGP.as_prediction_model``."""

    def __init__(self, deep_gp: DeepGP):
        self.deep_gp = deep_gp

    def __call__(self, inputs):
        with torch.no_grad():
            return self.deep_gp(inputs)

    predict = __call__


def sample_dgp(model: DeepGP):
    """Chained efficient posterior samples (reference deep_gp.py:295-307)."""
    from ..sampling.sample import Sample

    function_draws = [layer.sample() for layer in model.f_layers]

    class ChainedSample(Sample):
        def __call__(self, X):
            for f in function_draws:
                X = f(X)
            return X

    return ChainedSample()
