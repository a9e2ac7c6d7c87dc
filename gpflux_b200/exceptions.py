"""Mirrors gpflux/exceptions.py."""


class GPLayerIncompatibilityException(Exception):
    r"""Raised when :class:`~gpflux_b200.layers.GPLayer` is misconfigured
    (reference gpflux/exceptions.py, gpflux/runtime_checks.py:47-71)."""


class InvalidPlottingDimension(Exception):
    """Kept for name compatibility with gpflux.exceptions."""
