from .gp_layer import GPLayer, MultivariateNormalDiag
from .latent_variable_layer import LatentVariableLayer, LayerWithObservations
from .likelihood_layer import LikelihoodLayer, LikelihoodOutputs

__all__ = ["GPLayer", "LatentVariableLayer", "LayerWithObservations", "LikelihoodLayer", "LikelihoodOutputs",
           "MultivariateNormalDiag"]
