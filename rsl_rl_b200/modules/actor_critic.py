"""Gaussian MLP actor + MLP critic (mirror of `rsl_rl/modules/actor_critic.py` of the reference).

Same constructor, sub-module names (`actor`, `critic`, `actor_obs_normalizer`, `critic_obs_normalizer`),
parameters (`std` / `log_std`) and methods, therefore the same `state_dict`.  Every dense layer runs in
the library's GEMM kernels (networks/mlp.py).  The methods below are the general, autograd-capable
surface; the PPO rollout/update of a feed-forward policy drives the same layers through
`algorithms/_engine.py` without autograd.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.distributions import Normal

from ..networks import MLP, EmpiricalNormalization

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
_ENTROPY_CONST = 0.5 + 0.5 * math.log(2 * math.pi)


class GaussianHeadMixin:
    """Distribution bookkeeping shared by the feed-forward and the recurrent actor-critic."""

    noise_std_type: str
    state_dependent_std: bool

    def _init_noise(self, num_actions, init_noise_std):
        if self.state_dependent_std:
            torch.nn.init.zeros_(self.actor[-2].weight[num_actions:])
            if self.noise_std_type == "scalar":
                torch.nn.init.constant_(self.actor[-2].bias[num_actions:], init_noise_std)
            elif self.noise_std_type == "log":
                torch.nn.init.constant_(self.actor[-2].bias[num_actions:], math.log(init_noise_std + 1e-7))
            else:
                raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")
        elif self.noise_std_type == "scalar":
            self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        elif self.noise_std_type == "log":
            self.log_std = nn.Parameter(torch.log(init_noise_std * torch.ones(num_actions)))
        else:
            raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")
        self._mean = None
        self._std = None
        self._noise_fn = None  # test hook: callable(like) -> standard-normal tensor
        Normal.set_default_validate_args(False)  # reference actor_critic.py:98 (global side effect kept)

    # -- distribution surface (reference actor_critic.py:106-146,170-171)
    @property
    def distribution(self):
        return None if self._mean is None else Normal(self._mean, self._std)

    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        if self._std is None and not self.state_dependent_std:  # before the first forward: the parameter itself
            return self.std.detach() if self.noise_std_type == "scalar" else torch.exp(self.log_std.detach())
        return self._std

    @property
    def entropy(self):
        return (_ENTROPY_CONST + torch.log(self._std)).sum(dim=-1)

    def update_distribution(self, obs):
        """`obs`: the (normalised, concatenated) network input -- the reference's parameter name (actor_critic.py:119)."""
        if self.state_dependent_std:
            mean_and_std = self.actor(obs)
            if self.noise_std_type == "scalar":
                mean, std = torch.unbind(mean_and_std, dim=-2)
            elif self.noise_std_type == "log":
                mean, log_std = torch.unbind(mean_and_std, dim=-2)
                std = torch.exp(log_std)
            else:
                raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")
        else:
            mean = self.actor(obs)
            if self.noise_std_type == "scalar":
                std = self.std.expand_as(mean)
            elif self.noise_std_type == "log":
                std = torch.exp(self.log_std).expand_as(mean)
            else:
                raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")
        self._mean, self._std = mean, std

    def draw_noise(self, like: torch.Tensor) -> torch.Tensor:
        """z ~ N(0,1) with the shape of `like`, drawn exactly like `torch.normal(mean, std)` draws it
        (normal_(0,1) on an empty tensor of that shape), so the stream matches the reference's."""
        if self._noise_fn is not None:
            return self._noise_fn(like)
        return torch.empty_like(like).normal_()

    def sample(self):
        with torch.no_grad():
            z = self.draw_noise(self._mean)
            return z.mul_(self._std).add_(self._mean)  # two roundings, like Normal.sample()

    def get_actions_log_prob(self, actions):
        var = self._std ** 2
        return (-((actions - self._mean) ** 2) / (2 * var) - torch.log(self._std) - _LOG_SQRT_2PI).sum(dim=-1)

    def set_gemm_precision(self, precision: str):
        from ..networks.memory import Memory
        for m in self.modules():
            if isinstance(m, (MLP, Memory)):
                # (the autograd path of an MLP / LSTM has no fp16-storage variant: tensor cores there mean tf32)
                m.precision = "tf32" if precision == "fp16" else precision


class ActorCritic(GaussianHeadMixin, nn.Module):
    is_recurrent = False

    def __init__(self, obs, obs_groups, num_actions, actor_obs_normalization=False, critic_obs_normalization=False,
                 actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256], activation="elu",
                 init_noise_std=1.0, noise_std_type: str = "scalar", state_dependent_std=False, **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        super().__init__()
        self.obs_groups = obs_groups
        self.num_actor_obs = self._count_obs(obs, obs_groups["policy"])
        self.num_critic_obs = self._count_obs(obs, obs_groups["critic"])
        self.state_dependent_std = state_dependent_std
        self.noise_std_type = noise_std_type

        out = [2, num_actions] if state_dependent_std else num_actions
        self.actor = MLP(self.num_actor_obs, out, actor_hidden_dims, activation)
        self.actor_obs_normalization = actor_obs_normalization
        self.actor_obs_normalizer = (EmpiricalNormalization(self.num_actor_obs) if actor_obs_normalization
                                     else torch.nn.Identity())
        print(f"Actor MLP: {self.actor}")
        self.critic = MLP(self.num_critic_obs, 1, critic_hidden_dims, activation)
        self.critic_obs_normalization = critic_obs_normalization
        self.critic_obs_normalizer = (EmpiricalNormalization(self.num_critic_obs) if critic_obs_normalization
                                      else torch.nn.Identity())
        print(f"Critic MLP: {self.critic}")
        self._init_noise(num_actions, init_noise_std)

    @staticmethod
    def _count_obs(obs, groups):
        n = 0
        for g in groups:
            assert len(obs[g].shape) == 2, "The ActorCritic module only supports 1D observations."
            n += obs[g].shape[-1]
        return n

    @property
    def critic_shares_actor_input(self) -> bool:
        return (list(self.obs_groups["policy"]) == list(self.obs_groups["critic"])
                and not self.actor_obs_normalization and not self.critic_obs_normalization)

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    def act(self, obs, **kwargs):
        self.update_distribution(self.actor_obs_normalizer(self.get_actor_obs(obs)))
        return self.sample()

    def act_inference(self, obs):
        return self.actor(self.actor_obs_normalizer(self.get_actor_obs(obs)))

    def evaluate(self, obs, **kwargs):
        return self.critic(self.critic_obs_normalizer(self.get_critic_obs(obs)))

    def _cat(self, obs, groups):
        if len(groups) == 1:
            return obs[groups[0]]
        return torch.cat([obs[g] for g in groups], dim=-1)

    def get_actor_obs(self, obs):
        return self._cat(obs, self.obs_groups["policy"])

    def get_critic_obs(self, obs):
        return self._cat(obs, self.obs_groups["critic"])

    def update_normalization(self, obs):
        if self.actor_obs_normalization:
            self.actor_obs_normalizer.update(self.get_actor_obs(obs))
        if self.critic_obs_normalization:
            self.critic_obs_normalizer.update(self.get_critic_obs(obs))

    def load_state_dict(self, state_dict, strict=True):
        """Returns True: loading a PPO checkpoint always resumes training (reference actor_critic.py:181-195)."""
        super().load_state_dict(state_dict, strict=strict)
        return True
