"""Seeded synthetic VecEnv used by the tests, the smoke test and bench.py.

obs ~ N(0,1) per group, rewards ~ N(0,1), dones ~ Bernoulli(p_done) as int64, half of the dones are
flagged as time-outs; actions are ignored.  Random numbers come from the env's own generator so the
policy's RNG stream is untouched.  `ReplayVecEnv` replays a recorded stream (golden fixtures).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..utils.tensordict_lite import TensorDict
from .vec_env import VecEnv


class SynthVecEnv(VecEnv):
    """obs ~ N(0,1), rewards ~ N(0,1), dones ~ Bernoulli(p_done) (int64), time_outs = dones & Bernoulli(p_timeout).

    The random stream is drawn `block` env steps at a time (one launch per quantity) and handed out one step
    per `step()` call, so the simulator stand-in costs a few microseconds per step instead of eight launches."""

    def __init__(self, num_envs: int, group_dims: Optional[Dict[str, int]] = None, num_actions: int = 12,
                 device="cuda:0", seed: int = 0, p_done: float = 0.02, p_timeout: float = 0.5, block: int = 24):
        self.num_envs, self.num_actions = num_envs, num_actions
        self.group_dims = dict(group_dims or {"policy": 48})
        self.device = device
        self.max_episode_length = 1000
        self.episode_length_buf = torch.zeros(num_envs, dtype=torch.long, device=device)
        self.cfg = {}
        self.step_dt = 0.02
        self.unwrapped = self
        self.p_done, self.p_timeout = p_done, p_timeout
        self.block = max(1, int(block))
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self._i = self.block  # forces a refill on first use
        self._refill()
        self._obs = self._next_obs()

    def _refill(self):
        n, dev, b, g = self.num_envs, self.device, self.block, self.gen
        self._obs_blk = {k: torch.randn(b + 1, n, d, generator=g, device=dev) for k, d in self.group_dims.items()}
        self._rew_blk = torch.randn(b, n, generator=g, device=dev)
        done = torch.rand(b, n, generator=g, device=dev) < self.p_done
        self._done_blk = done.long()
        self._to_blk = (torch.rand(b, n, generator=g, device=dev) < self.p_timeout) & done
        # per-step views made once per block: step() then costs no tensor operations at all
        self._obs_rows = [TensorDict({k: v[j] for k, v in self._obs_blk.items()}, batch_size=[self.num_envs], device=dev)
                          for j in range(b + 1)]
        self._rew_rows, self._done_rows, self._to_rows = (list(t.unbind(0)) for t in
                                                          (self._rew_blk, self._done_blk, self._to_blk))
        self._i = 0
        self._j = 0

    def _next_obs(self):
        obs = self._obs_rows[self._j]
        self._j += 1
        return obs

    def get_observations(self):
        return self._obs

    def step(self, actions):
        if self._i >= self.block:
            self._refill()
        i = self._i
        self._i += 1
        self._obs = self._next_obs()
        return self._obs, self._rew_rows[i], self._done_rows[i], {"time_outs": self._to_rows[i]}


class ReplayVecEnv(VecEnv):
    """Replays recorded tensors: obs[k]: [S+1, N, d], rewards/dones/time_outs: [S, N]."""

    def __init__(self, obs: Dict[str, torch.Tensor], rewards, dones, time_outs, num_actions: int, device="cuda:0"):
        self._obs = {k: v.to(device) for k, v in obs.items()}
        self._rew, self._done = rewards.to(device), dones.to(device)
        self._to = None if time_outs is None else time_outs.to(device)
        first = next(iter(self._obs.values()))
        self.num_envs, self.num_actions = first.shape[1], num_actions
        self.device = device
        self.max_episode_length = 1000
        self.episode_length_buf = torch.zeros(self.num_envs, dtype=torch.long, device=device)
        self.cfg = {}
        self.step_dt = 0.02
        self.unwrapped = self
        self.t = 0
        self.actions_seen = []

    def _cur(self):
        return TensorDict({k: v[self.t] for k, v in self._obs.items()}, batch_size=[self.num_envs], device=self.device)

    def get_observations(self):
        return self._cur()

    def step(self, actions):
        self.actions_seen.append(actions.clone())
        r, d = self._rew[self.t], self._done[self.t]
        extras = {} if self._to is None else {"time_outs": self._to[self.t]}
        self.t += 1
        return self._cur(), r, d, extras


class PinnedHostEnv(VecEnv):
    """Host-resident environment for end-to-end timing: `pool` pre-generated steps of observations /
    rewards / dones / time-outs live in PINNED host memory and are cycled; every step() hands the runner
    host tensors (so the runner's `.to(device)` is a real host->device copy) and receives the actions
    on the host (a real device->host copy).  No host-side random number generation inside the loop."""

    def __init__(self, num_envs: int, group_dims=None, num_actions: int = 12, pool: int = 32, seed: int = 0,
                 p_done: float = 0.02, p_timeout: float = 0.5):
        self.num_envs, self.num_actions = num_envs, num_actions
        self.group_dims = dict(group_dims or {"policy": 48})
        self.device = "cpu"
        self.max_episode_length = 1000
        self.episode_length_buf = torch.zeros(num_envs, dtype=torch.long)
        self.cfg = {}
        self.step_dt = 0.02
        self.unwrapped = self
        g = torch.Generator().manual_seed(seed)
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        self._obs = {k: pin(torch.randn(pool, num_envs, d, generator=g)) for k, d in self.group_dims.items()}
        self._rew = pin(torch.randn(pool, num_envs, generator=g))
        dones = torch.rand(pool, num_envs, generator=g) < p_done
        self._done = pin(dones.long())
        self._to = pin((torch.rand(pool, num_envs, generator=g) < p_timeout) & dones)
        self.pool, self.t = pool, 0
        self.h2d_bytes_per_step = (sum(self.group_dims.values()) * 4 + 4 + 8 + 1) * num_envs
        self.d2h_bytes_per_step = num_actions * 4 * num_envs

    def _cur(self):
        i = self.t % self.pool
        return TensorDict({k: v[i] for k, v in self._obs.items()}, batch_size=[self.num_envs], device="cpu")

    def get_observations(self):
        return self._cur()

    def step(self, actions):
        i = self.t % self.pool
        r, d, to = self._rew[i], self._done[i], self._to[i]
        self.t += 1
        return self._cur(), r, d, {"time_outs": to}
