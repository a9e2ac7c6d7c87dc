"""Mirror of gpflux/optimization/__init__.py."""
from .keras_natgrad import NatGradModel, NatGradWrapper

__all__ = ["NatGradModel", "NatGradWrapper"]
