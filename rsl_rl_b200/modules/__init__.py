"""Policy modules (mirror of `rsl_rl.modules`; the distillation-only StudentTeacher* are out of scope)."""
from .actor_critic import ActorCritic
from .actor_critic_recurrent import ActorCriticRecurrent
from .rnd import RandomNetworkDistillation, resolve_rnd_config
from .symmetry import resolve_symmetry_config

__all__ = ["ActorCritic", "ActorCriticRecurrent", "RandomNetworkDistillation", "resolve_rnd_config",
           "resolve_symmetry_config"]
