"""Minimal stand-ins for the GPflow types GPflux's public API takes as arguments.

GPflow/TensorFlow are not installable in this environment, so the objects a
user passes to ``GPLayer`` / ``DeepGP`` / ``helpers`` - kernels, inducing variables, mean
functions, the Gaussian likelihood, ``Parameter`` - are restated here as thin ``torch.nn.Module``
holders of parameters.  They contain NO hot-path arithmetic: ``GPLayer`` packs their parameters
into the ``[K,M,D]`` / ``[K,D]`` / ``[K]`` arrays the C-ABI takes.  Names, constructor arguments
and attribute names follow GPflow 2.9 (``gpflow.kernels``, ``gpflow.inducing_variables``,
``gpflow.mean_functions``, ``gpflow.likelihoods``, ``gpflow.Parameter``).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn


_CONFIG = {"jitter": 1e-6, "float": torch.float64}


def default_float() -> torch.dtype:
    """gpflow.config.default_float() (read at reference gp_layer.py:163)."""
    return _CONFIG["float"]


def set_default_float(value) -> None:
    """gpflow.config.set_default_float: float64 (default) or the float32 opt-in.

    float32 is a STORAGE / interface precision here: parameters, inputs, outputs and gradients are
    float32, the device kernels still contract and accumulate in float64 on the DMMA pipe (values are
    widened on the way in, rounded once on the way out), so results sit within float32 rounding of the
    float64 path - inside the 1e-4 the north star allows for the opt-in, at the float64 path's speed.
    Parameters created before the switch keep their dtype."""
    dt = {np.float32: torch.float32, np.float64: torch.float64, "float32": torch.float32,
          "float64": torch.float64}.get(value, value)
    if isinstance(dt, np.dtype):
        dt = {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64}.get(dt, dt)
    if dt not in (torch.float32, torch.float64):
        raise ValueError(f"default float must be float32 or float64, got {value!r}")
    _CONFIG["float"] = dt


def default_jitter() -> float:
    """gpflow.config.default_jitter() (read at reference math.py:36, sample.py:169)."""
    return _CONFIG["jitter"]


def set_default_jitter(value: float) -> None:
    """gpflow.config.set_default_jitter."""
    _CONFIG["jitter"] = float(value)


def _as_tensor(value) -> torch.Tensor:
    if isinstance(value, torch.Tensor):
        return value.detach().to(default_float()).clone()
    return torch.as_tensor(np.asarray(value, dtype=np.float64)).to(default_float())


def _inv_softplus(y: torch.Tensor) -> torch.Tensor:
    # log(exp(y) - 1), stable for large y
    return torch.where(y > 30.0, y, torch.log(torch.expm1(y)))


class Parameter(nn.Module):
    """gpflow.Parameter: a constrained value backed by an unconstrained trainable variable.

    ``transform``: None (identity), "positive" (softplus, GPflow's ``positive()`` with a zero
    lower bound) or "triangular".  GPflow's ``triangular()`` packs the lower triangle with TFP's
    FillTriangular; here the unconstrained variable is the full ``[..., M, M]`` array whose strict
    upper triangle receives zero gradient and is ignored (``tril`` applied on read) - the same
    optimisation problem for any element-wise optimiser, without the gather/scatter.
    """

    def __init__(self, value, transform: Optional[str] = None, trainable: bool = True, name: Optional[str] = None,
                 prior=None):
        super().__init__()
        v = _as_tensor(value)
        if transform == "positive":
            if bool((v <= 0).any()):
                raise ValueError(f"Parameter {name or ''}: positive transform needs positive initial values")
            u = _inv_softplus(v)
        elif transform in (None, "triangular"):
            u = v
        else:
            raise ValueError(f"unknown transform {transform!r}")
        self.transform = transform
        self.param_name = name
        self.prior = prior
        self.unconstrained_variable = nn.Parameter(u, requires_grad=trainable)

    @property
    def trainable(self) -> bool:
        return self.unconstrained_variable.requires_grad

    def value(self) -> torch.Tensor:
        u = self.unconstrained_variable
        if self.transform == "positive":
            return torch.nn.functional.softplus(u)
        if self.transform == "triangular":
            return torch.tril(u)
        return u

    def forward(self) -> torch.Tensor:
        return self.value()

    def assign(self, value) -> None:
        v = _as_tensor(value).to(self.unconstrained_variable.device)
        with torch.no_grad():
            self.unconstrained_variable.copy_(_inv_softplus(v) if self.transform == "positive" else v)

    def numpy(self) -> np.ndarray:
        return self.value().detach().cpu().numpy()

    @property
    def shape(self):
        return self.unconstrained_variable.shape

    def log_prior_density(self) -> torch.Tensor:
        """gpflow.Parameter.log_prior_density (reference gp_layer.py:287): 0 without a prior.
        With a prior (any object with ``log_prob``), GPflow evaluates it on the constrained value."""
        if self.prior is None:
            return torch.zeros((), dtype=default_float(), device=self.unconstrained_variable.device)
        return self.prior.log_prob(self.value()).sum()


class NaturalGradient:
    """gpflow.optimizers.NaturalGradient (default XiNat parameterisation): the optimiser
    ``NatGradModel`` applies to each layer's ``(q_mu, q_sqrt)`` (reference
    gpflux/optimization/keras_natgrad.py:187-191).  The step itself is one ``gpfx_natgrad_step``
    call on the device."""

    def __init__(self, gamma: float):
        self.gamma = float(gamma)

    def _natgrad_apply_gradients(self, q_mu_grad, q_sqrt_grad, q_mu: "Parameter", q_sqrt: "Parameter") -> None:
        from . import ops

        if q_sqrt.transform not in (None, "triangular") or q_mu.transform is not None:
            raise NotImplementedError("NaturalGradient expects an unconstrained q_mu and a triangular q_sqrt")
        with torch.no_grad():
            mu_new, sqrt_new, info = ops.natgrad_step(
                q_mu.value(), q_sqrt.value(), q_mu_grad, torch.tril(q_sqrt_grad), self.gamma)
            q_mu.unconstrained_variable.copy_(mu_new)
            q_sqrt.unconstrained_variable.copy_(sqrt_new)
        self.last_info = info  # [3,P] Cholesky status, read lazily (no sync here)

    def check_status(self) -> None:
        """Raise if the last step met a covariance / precision that is not positive definite (the three
        factorisations of the XiNat step; one device->host read - call it where the loss is read back)."""
        info = getattr(self, "last_info", None)
        if info is not None and int(info.abs().max()) != 0:
            from ._cabi import E_NOT_PD, GpfxError

            which = ("current covariance", "new precision", "new covariance")
            r, p = [int(i) for i in torch.nonzero(info)[0]]
            raise GpfxError(f"gpfx status {E_NOT_PD} (GPFX_E_NOT_PD): natural-gradient step, {which[r]} of latent GP {p} "
                            "is not positive definite (step size gamma too large?)")


def set_trainable(module: nn.Module, flag: bool) -> None:
    """gpflow.set_trainable."""
    for p in module.parameters():
        p.requires_grad_(flag)


# ------------------------------------------------------------------------------------------ kernels
class Kernel(nn.Module):
    """gpflow.kernels.Kernel"""


class Stationary(Kernel):
    """gpflow.kernels.Stationary: ``variance`` and (ARD) ``lengthscales``, both positive()."""

    kind: str = ""

    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        super().__init__()
        if kwargs.get("active_dims") is not None:
            raise NotImplementedError("active_dims is not supported by the B200 SVGP path")
        self.variance = Parameter(variance, transform="positive", name="variance")
        self.lengthscales = Parameter(lengthscales, transform="positive", name="lengthscales")

    @property
    def ard(self) -> bool:
        return self.lengthscales.shape != torch.Size([])

    def lengthscales_vector(self, input_dim: int) -> torch.Tensor:
        ls = self.lengthscales.value()
        if ls.dim() == 0:
            return ls.expand(input_dim)
        if ls.shape[0] != input_dim:
            raise ValueError(f"kernel has {ls.shape[0]} lengthscales, input has dim {input_dim}")
        return ls

    @property
    def trainable_parameters(self) -> List[Parameter]:
        return [p for p in (self.variance, self.lengthscales) if p.trainable]


class SquaredExponential(Stationary):
    kind = "se"


RBF = SquaredExponential


class Matern12(Stationary):
    kind = "matern12"


class Matern32(Stationary):
    kind = "matern32"


class Matern52(Stationary):
    kind = "matern52"


class MultioutputKernel(Kernel):
    """gpflow.kernels.MultioutputKernel"""

    @property
    def num_latent_gps(self) -> int:
        raise NotImplementedError

    @property
    def latent_kernels(self) -> Sequence[Kernel]:
        raise NotImplementedError

    @property
    def trainable_parameters(self) -> List[Parameter]:
        out = []
        for k in self.latent_kernels:
            out.extend(k.trainable_parameters)
        return out


class SharedIndependent(MultioutputKernel):
    """gpflow.kernels.SharedIndependent(kernel, output_dim)"""

    def __init__(self, kernel: Kernel, output_dim: int):
        super().__init__()
        self.kernel = kernel
        self.output_dim = output_dim

    @property
    def num_latent_gps(self) -> int:
        return self.output_dim

    @property
    def latent_kernels(self):
        return (self.kernel,)


class SeparateIndependent(MultioutputKernel):
    """gpflow.kernels.SeparateIndependent(kernels)"""

    def __init__(self, kernels: Sequence[Kernel]):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)

    @property
    def num_latent_gps(self) -> int:
        return len(self.kernels)

    @property
    def latent_kernels(self):
        return tuple(self.kernels)


# ------------------------------------------------------------------------------ inducing variables
class InducingVariables(nn.Module):
    """gpflow.inducing_variables.InducingVariables"""


class InducingPoints(InducingVariables):
    """gpflow.inducing_variables.InducingPoints(Z)"""

    def __init__(self, Z):
        super().__init__()
        self.Z = Parameter(Z, name="Z")

    @property
    def num_inducing(self) -> int:
        return int(self.Z.shape[0])


class MultioutputInducingVariables(InducingVariables):
    """gpflow.inducing_variables.MultioutputInducingVariables"""


class SharedIndependentInducingVariables(MultioutputInducingVariables):
    def __init__(self, inducing_variable: InducingPoints):
        super().__init__()
        self.inducing_variable = inducing_variable

    @property
    def num_inducing(self) -> int:
        return self.inducing_variable.num_inducing

    @property
    def inducing_variables(self):
        return (self.inducing_variable,)


class SeparateIndependentInducingVariables(MultioutputInducingVariables):
    def __init__(self, inducing_variable_list: Sequence[InducingPoints]):
        super().__init__()
        self.inducing_variable_list = nn.ModuleList(inducing_variable_list)

    @property
    def num_inducing(self) -> int:
        # "currently the same for each dim" (reference runtime_checks.py:73)
        return self.inducing_variable_list[0].num_inducing

    @property
    def inducing_variables(self):
        return tuple(self.inducing_variable_list)


FallbackSeparateIndependentInducingVariables = SeparateIndependentInducingVariables
FallbackSharedIndependentInducingVariables = SharedIndependentInducingVariables


# ---------------------------------------------------------------------------------- mean functions
class MeanFunction(nn.Module):
    """gpflow.mean_functions.MeanFunction"""


class Identity(MeanFunction):
    def forward(self, X):
        return X


class Zero(MeanFunction):
    def __init__(self, output_dim: int = 1):
        super().__init__()
        self.output_dim = output_dim

    def forward(self, X):
        return None  # GPLayer treats None as "add nothing" (zeros of the right shape in GPflow)


class Linear(MeanFunction):
    """gpflow.mean_functions.Linear(A, b): X @ A + b"""

    def __init__(self, A=None, b=None):
        super().__init__()
        A = np.ones((1, 1)) if A is None else A
        b = np.zeros(1) if b is None else b
        self.A = Parameter(A, name="A")
        self.b = Parameter(b, name="b")

    def forward(self, X):
        return X @ self.A.value() + self.b.value()


# ------------------------------------------------------------------------------------ likelihoods
class Likelihood(nn.Module):
    """gpflow.likelihoods.Likelihood"""


class Gaussian(Likelihood):
    """gpflow.likelihoods.Gaussian(variance): closed-form variational expectations (computed on
    the device by gpfx_gaussian_ve) and predict_mean_and_var."""

    def __init__(self, variance: float = 1.0):
        super().__init__()
        self.variance = Parameter(variance, transform="positive", name="likelihood_variance")

    def variational_expectations_sum(self, Fmu, Fvar, Y):
        from . import ops

        return ops.gaussian_variational_expectations_sum(Fmu, Fvar, Y, self.variance.value())

    def predict_mean_and_var(self, Fmu, Fvar):
        return Fmu, Fvar + self.variance.value()
