"""Recurrent actor-critic: an LSTM/GRU memory in front of each MLP (mirror of
`rsl_rl/modules/actor_critic_recurrent.py`; sub-modules `memory_a`, `memory_c`, `actor`, `critic`)."""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from ..networks import MLP, EmpiricalNormalization, Memory
from .actor_critic import ActorCritic, GaussianHeadMixin


class ActorCriticRecurrent(GaussianHeadMixin, nn.Module):
    is_recurrent = True

    def __init__(self, obs, obs_groups, num_actions, actor_obs_normalization=False, critic_obs_normalization=False,
                 actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256], activation="elu",
                 init_noise_std=1.0, noise_std_type: str = "scalar", state_dependent_std=False, rnn_type="lstm",
                 rnn_hidden_dim=256, rnn_num_layers=1, **kwargs):
        if "rnn_hidden_size" in kwargs:
            warnings.warn("The argument `rnn_hidden_size` is deprecated and will be removed in a future version. "
                          "Please use `rnn_hidden_dim` instead.", DeprecationWarning)
            if rnn_hidden_dim == 256:
                rnn_hidden_dim = kwargs.pop("rnn_hidden_size")
        if kwargs:
            print("ActorCriticRecurrent.__init__ got unexpected arguments, which will be ignored: "
                  + str(kwargs.keys()))
        super().__init__()
        self.obs_groups = obs_groups
        self.num_actor_obs = ActorCritic._count_obs(obs, obs_groups["policy"])
        self.num_critic_obs = ActorCritic._count_obs(obs, obs_groups["critic"])
        self.state_dependent_std = state_dependent_std
        self.noise_std_type = noise_std_type

        self.memory_a = Memory(self.num_actor_obs, type=rnn_type, num_layers=rnn_num_layers,
                               hidden_size=rnn_hidden_dim)
        out = [2, num_actions] if state_dependent_std else num_actions
        self.actor = MLP(rnn_hidden_dim, out, actor_hidden_dims, activation)
        self.actor_obs_normalization = actor_obs_normalization
        self.actor_obs_normalizer = (EmpiricalNormalization(self.num_actor_obs) if actor_obs_normalization
                                     else torch.nn.Identity())
        print(f"Actor RNN: {self.memory_a}")
        print(f"Actor MLP: {self.actor}")
        self.memory_c = Memory(self.num_critic_obs, type=rnn_type, num_layers=rnn_num_layers,
                               hidden_size=rnn_hidden_dim)
        self.critic = MLP(rnn_hidden_dim, 1, critic_hidden_dims, activation)
        self.critic_obs_normalization = critic_obs_normalization
        self.critic_obs_normalizer = (EmpiricalNormalization(self.num_critic_obs) if critic_obs_normalization
                                      else torch.nn.Identity())
        print(f"Critic RNN: {self.memory_c}")
        print(f"Critic MLP: {self.critic}")
        self._init_noise(num_actions, init_noise_std)

    def reset(self, dones=None):
        self.memory_a.reset(dones)
        self.memory_c.reset(dones)

    def forward(self):
        raise NotImplementedError

    _cat = ActorCritic._cat
    get_actor_obs = ActorCritic.get_actor_obs
    get_critic_obs = ActorCritic.get_critic_obs
    update_normalization = ActorCritic.update_normalization

    def act(self, obs, masks=None, hidden_states=None):
        x = self.actor_obs_normalizer(self.get_actor_obs(obs))
        out_mem = self.memory_a(x, masks, hidden_states).squeeze(0)
        self.update_distribution(out_mem)
        return self.sample()

    def act_inference(self, obs):
        x = self.actor_obs_normalizer(self.get_actor_obs(obs))
        return self.actor(self.memory_a(x).squeeze(0))

    def evaluate(self, obs, masks=None, hidden_states=None):
        x = self.critic_obs_normalizer(self.get_critic_obs(obs))
        return self.critic(self.memory_c(x, masks, hidden_states).squeeze(0))

    def act_evaluate(self, obs, masks=None, hidden_states=(None, None)):
        """`act(obs, masks, hidden_states[0])` followed by `evaluate(obs, masks, hidden_states[1])` with the two memories
        advanced together (grouped LSTM launches): returns (actions, values).  Same results, same RNG consumption."""
        x_a = self.actor_obs_normalizer(self.get_actor_obs(obs))
        x_c = self.critic_obs_normalizer(self.get_critic_obs(obs))
        out_a, out_c = Memory.forward_pair(self.memory_a, self.memory_c, x_a, x_c, masks, hidden_states[0],
                                           hidden_states[1])
        self.update_distribution(out_a.squeeze(0))
        actions = self.sample()
        return actions, self.critic(out_c.squeeze(0))

    def get_hidden_states(self):
        return self.memory_a.hidden_states, self.memory_c.hidden_states

    def load_state_dict(self, state_dict, strict=True):
        super().load_state_dict(state_dict, strict=strict)
        return True
