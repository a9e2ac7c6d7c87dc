"""Mirrors gpflux/runtime_checks.py:29-74 (host-side shape discovery only; no GPU work)."""
from __future__ import annotations

from typing import Optional, Tuple

from .exceptions import GPLayerIncompatibilityException
from .gpflow_compat import (
    MeanFunction,
    MultioutputInducingVariables,
    MultioutputKernel,
    SeparateIndependentInducingVariables,
)


def verify_compatibility(kernel, mean_function, inducing_variable) -> Tuple[int, int]:
    """Checks that the arguments are compatible for use in a ``GPLayer``.

    :returns: number of inducing variables and number of latent GPs
    :raises GPLayerIncompatibilityException: if an incompatibility is detected
    """
    if not isinstance(inducing_variable, MultioutputInducingVariables):
        raise GPLayerIncompatibilityException(
            "`inducing_variable` must be a `gpflow.inducing_variables.MultioutputInducingVariables`"
        )
    if not isinstance(kernel, MultioutputKernel):
        raise GPLayerIncompatibilityException("`kernel` must be a `gpflow.kernels.MultioutputKernel`")
    if not isinstance(mean_function, MeanFunction):
        raise GPLayerIncompatibilityException("`kernel` must be a `gpflow.mean_functions.MeanFunction`")

    latent_inducing_points: Optional[int] = None
    if isinstance(inducing_variable, SeparateIndependentInducingVariables):
        latent_inducing_points = len(inducing_variable.inducing_variable_list)

    num_latent_gps = kernel.num_latent_gps
    if latent_inducing_points is not None and latent_inducing_points != num_latent_gps:
        raise GPLayerIncompatibilityException(
            f"The number of latent GPs ({num_latent_gps}) does not match "
            f"the number of separate independent inducing_variables ({latent_inducing_points})"
        )
    return inducing_variable.num_inducing, num_latent_gps
