"""RL algorithms (mirror of `rsl_rl.algorithms`; Distillation is out of scope)."""
from .ppo import PPO

__all__ = ["PPO"]
