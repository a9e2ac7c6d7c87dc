from .deep_gp import DeepGP, sample_dgp

__all__ = ["DeepGP", "sample_dgp"]
