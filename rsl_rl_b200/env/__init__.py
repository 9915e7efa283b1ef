"""Environment definitions (mirror of `rsl_rl.env`)."""
from .synthetic import PinnedHostEnv, ReplayVecEnv, SynthVecEnv
from .vec_env import VecEnv

__all__ = ["VecEnv", "SynthVecEnv", "ReplayVecEnv", "PinnedHostEnv"]
