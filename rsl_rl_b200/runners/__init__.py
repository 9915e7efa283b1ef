"""Runners (mirror of `rsl_rl.runners`; DistillationRunner is out of scope)."""
from .on_policy_runner import OnPolicyRunner

__all__ = ["OnPolicyRunner"]
