from .random import RandomFourierFeatures, RandomFourierFeaturesCosine

__all__ = ["RandomFourierFeatures", "RandomFourierFeaturesCosine"]
