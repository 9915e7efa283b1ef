"""Runners (mirror of `rsl_rl.runners`)."""
from .on_policy_runner import OnPolicyRunner  # isort:skip
from .distillation_runner import DistillationRunner

__all__ = ["OnPolicyRunner", "DistillationRunner"]
