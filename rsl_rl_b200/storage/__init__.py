"""Transition storage (mirror of `rsl_rl.storage`)."""
from .rollout_storage import RolloutStorage

__all__ = ["RolloutStorage"]
