from .constant_input_dim_deep_gp import Config, build_constant_input_dim_deep_gp

__all__ = ["Config", "build_constant_input_dim_deep_gp"]
