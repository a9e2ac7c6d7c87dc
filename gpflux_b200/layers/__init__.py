from .gp_layer import GPLayer, MultivariateNormalDiag
from .likelihood_layer import LikelihoodLayer, LikelihoodOutputs

__all__ = ["GPLayer", "LikelihoodLayer", "LikelihoodOutputs", "MultivariateNormalDiag"]
