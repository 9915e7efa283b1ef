"""Network building blocks (mirror of `rsl_rl.networks`)."""
from .memory import Memory
from .mlp import MLP
from .normalization import EmpiricalDiscountedVariationNormalization, EmpiricalNormalization

__all__ = ["Memory", "MLP", "EmpiricalNormalization", "EmpiricalDiscountedVariationNormalization"]
