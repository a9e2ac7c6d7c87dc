from .kernel_with_feature_decomposition import KernelWithFeatureDecomposition
from .sample import Sample, efficient_sample

__all__ = ["KernelWithFeatureDecomposition", "Sample", "efficient_sample"]
