"""Vectorised-environment contract (mirror of `rsl_rl/env/vec_env.py:13-98` of the reference).

Implementations are supplied by the simulation framework; rsl_rl_b200 only consumes the interface:

attributes  num_envs, num_actions, max_episode_length, episode_length_buf, device, cfg
            (+ `unwrapped.step_dt` when RND is enabled, reference modules/rnd.py:208)
methods     get_observations() -> TensorDict{group: [num_envs, d_group]}
            step(actions[num_envs, num_actions]) -> (TensorDict, rewards[num_envs] fp32,
                 dones[num_envs] any int/bool dtype, extras)
extras      "time_outs": [num_envs] flags of terminations caused by the time limit (bootstrapped),
            "log" / "episode": dict of scalars or tensors averaged by the runner's logger.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch


class VecEnv(ABC):
    num_envs: int
    num_actions: int
    max_episode_length: "int | torch.Tensor"
    episode_length_buf: torch.Tensor
    device: "torch.device | str"
    cfg: "dict | object"

    @abstractmethod
    def get_observations(self):
        """Current observations as a TensorDict-like mapping of observation groups."""
        raise NotImplementedError

    @abstractmethod
    def step(self, actions: torch.Tensor):
        """Applies `actions` and returns (observations, rewards, dones, extras)."""
        raise NotImplementedError
