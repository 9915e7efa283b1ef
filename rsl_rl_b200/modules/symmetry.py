"""Symmetry configuration hook (mirror of `rsl_rl/modules/symmetry.py:9-24`)."""
from __future__ import annotations


def resolve_symmetry_config(alg_cfg, env):
    """The user's augmentation function receives the env, so it is injected into the config."""
    if "symmetry_cfg" in alg_cfg and alg_cfg["symmetry_cfg"] is not None:
        alg_cfg["symmetry_cfg"]["_env"] = env
    return alg_cfg
