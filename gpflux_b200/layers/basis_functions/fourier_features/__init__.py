from .random import (
    OrthogonalRandomFeatures,
    QuadratureFourierFeatures,
    RandomFourierFeatures,
    RandomFourierFeaturesCosine,
)

__all__ = ["OrthogonalRandomFeatures", "QuadratureFourierFeatures", "RandomFourierFeatures", "RandomFourierFeaturesCosine"]
