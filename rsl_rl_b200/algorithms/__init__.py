"""RL algorithms (mirror of `rsl_rl.algorithms`)."""
from .distillation import Distillation
from .ppo import PPO

__all__ = ["PPO", "Distillation"]
