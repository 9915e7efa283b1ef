"""Rollout buffer: structure of arrays in HBM, written / scanned / gathered by sm_100a kernels.

Mirror of `rsl_rl/storage/rollout_storage.py` of the reference: same constructor, same tensor
attributes (`observations, rewards, actions, dones, values, actions_log_prob, mu, sigma, returns,
advantages, saved_hidden_states_a/c, step`), same methods and the same 10-tuple minibatch protocol.

What changes underneath
  add_transitions          one launch (`rslrl_store_transition`) instead of 8+ `copy_`s
  compute_returns          one cooperative kernel (`rslrl_gae_norm`) instead of ~400 launches
  mini_batch_generator     `torch.randperm` (kept verbatim so the index stream is bit-identical to the
                           reference on the same device) + ONE gather launch per minibatch for all fields
  recurrent_mini_batch_generator   trajectory index computed once per call by a segmented scan; padded
                           observations / masks / hidden states are gathers from it
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _native as nat
from ..utils.tensordict_lite import TensorDict, is_tensordict_like
from ..utils.utils import TrajectoryIndex


_FLAGS = nat._FLAG_DTYPES
_store_step_fn = None


def _store_step():
    global _store_step_fn
    if _store_step_fn is None:
        _store_step_fn = nat.lib().rslrl_store_step
    return _store_step_fn


class RolloutStorage:
    class Transition:
        """One env step of data (reference rollout_storage.py:15-29)."""

        def __init__(self):
            self.observations = None
            self.actions = None
            self.privileged_actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None

        def clear(self):
            self.__init__()

    def __init__(self, training_type, num_envs, num_transitions_per_env, obs, actions_shape, device="cpu"):
        nat.require_cuda_device(device, "rsl_rl_b200.RolloutStorage (it lives in GPU memory)")
        if training_type not in ("rl", "distillation"):
            raise ValueError(f"Unknown training_type {training_type!r}: expected 'rl' or 'distillation'.")
        nat.lib()  # fail loudly now if the CUDA library is missing
        self.training_type = training_type
        self.device = device
        self.num_transitions_per_env = T = num_transitions_per_env
        self.num_envs = N = num_envs
        self.actions_shape = actions_shape
        A = int(torch.Size(actions_shape).numel())
        self._A = A

        f32 = dict(dtype=torch.float32, device=device)
        self.observations = TensorDict(
            {k: torch.zeros(T, *v.shape, **f32) for k, v in obs.items()}, batch_size=[T, N], device=device)
        self._obs_keys: List[str] = list(self.observations.keys())
        self.rewards = torch.zeros(T, N, 1, **f32)
        self.actions = torch.zeros(T, N, *actions_shape, **f32)
        self.dones = torch.zeros(T, N, 1, dtype=torch.uint8, device=device)
        if training_type == "distillation":   # reference rollout_storage.py:57-59
            self.privileged_actions = torch.zeros(T, N, *actions_shape, **f32)
        if training_type == "rl":             # reference rollout_storage.py:61-68
            self.values = torch.zeros(T, N, 1, **f32)
            self.actions_log_prob = torch.zeros(T, N, 1, **f32)
            self.mu = torch.zeros(T, N, *actions_shape, **f32)
            self.sigma = torch.zeros(T, N, *actions_shape, **f32)
            self.returns = torch.zeros(T, N, 1, **f32)
            self.advantages = torch.zeros(T, N, 1, **f32)

        self.saved_hidden_states_a = None
        self.saved_hidden_states_c = None
        self.step = 0

        self._store_plans = {}
        self._gae_ws = (torch.zeros(int(nat.lib().rslrl_gae_workspace_bytes(N)), dtype=torch.uint8, device=device)
                        if training_type == "rl" else None)
        # hooks used by the parity tests to inject a recorded permutation
        self._randperm = lambda n: torch.randperm(n, requires_grad=False, device=self.device)

    # ---------------------------------------------------------------------------------------------
    # writes
    # ---------------------------------------------------------------------------------------------
    def add_transitions(self, transition: "RolloutStorage.Transition"):
        """Reference rollout_storage.py:77-103."""
        if self.training_type == "distillation":
            self._store_distillation(transition)
        else:
            self._store(transition)

    def _store_distillation(self, tr):
        """Row `step` of observations, actions, privileged (teacher) actions, rewards and dones in ONE launch of the
        same kernel (`rslrl_store_transition`; the teacher's actions travel in the segment PPO uses for the action mean,
        the PPO-only fields are absent)."""
        if self.step >= self.num_transitions_per_env:
            raise OverflowError("Rollout buffer overflow! You should call clear() before adding new transitions.")
        s, N, A = self.step, self.num_envs, self._A
        srcs, dsts, dims = [], [], []
        for k in self._obs_keys:
            src = nat.f32c(tr.observations[k])
            dst = self.observations[k][s]
            srcs.append(src)
            dsts.append(dst.data_ptr())
            dims.append(int(dst[0].numel()))
        actions, privileged = nat.f32c(tr.actions), nat.f32c(tr.privileged_actions)
        rewards = nat.f32c(tr.rewards)
        dones, dcode = nat.flag_tensor(tr.dones)
        nat.require_cuda(*srcs, actions, privileged, rewards, dones)
        if actions.numel() != N * A or privileged.numel() != N * A or rewards.numel() != N or dones.numel() != N:
            raise ValueError("add_transitions: transition fields do not match the storage shape")
        nat.check(nat.lib().rslrl_store_transition(
            len(srcs), nat.void_array([t.data_ptr() for t in srcs]), nat.void_array(dsts), nat.int_array(dims),
            actions.data_ptr(), privileged.data_ptr(), None, 0, None, None, rewards.data_ptr(), dones.data_ptr(), dcode,
            None, 0, None, 0.0, self.actions[s].data_ptr(), self.privileged_actions[s].data_ptr(), None, None, None,
            self.rewards[s].data_ptr(), self.dones[s].data_ptr(), N, A, nat.stream_ptr()))
        self._save_hidden_states(tr.hidden_states)
        self.step += 1

    def _store(self, tr, time_outs: Optional[torch.Tensor] = None, intrinsic: Optional[torch.Tensor] = None,
               gamma: float = 0.0):
        """add_transitions fused with the reward bootstrap of PPO.process_env_step (ppo.py:150-164):
        stored reward = rewards (+ intrinsic) + gamma * values * time_outs."""
        if self.step >= self.num_transitions_per_env:
            raise OverflowError("Rollout buffer overflow! You should call clear() before adding new transitions.")
        s, N, A = self.step, self.num_envs, self._A
        srcs, dsts, dims = [], [], []
        for k in self._obs_keys:
            src = nat.f32c(tr.observations[k])
            nat.require_cuda(src)
            dst = self.observations[k][s]
            srcs.append(src)
            dsts.append(dst.data_ptr())
            dims.append(int(dst[0].numel()))
        actions, mu = nat.f32c(tr.actions), nat.f32c(tr.action_mean)
        sigma = tr.action_sigma
        if sigma.dim() == 2 and sigma.stride(0) == 0:  # std parameter expanded over the batch (stride-0 view)
            sigma, sigma_stride = sigma[0], 0
        elif sigma.dim() == 1:
            sigma_stride = 0
        else:
            sigma, sigma_stride = nat.f32c(sigma), A
        sigma = nat.f32c(sigma)
        values, logp = nat.f32c(tr.values), nat.f32c(tr.actions_log_prob)
        rewards = nat.f32c(tr.rewards)
        dones, dcode = nat.flag_tensor(tr.dones)
        nat.require_cuda(actions, mu, sigma, values, logp, rewards, dones)
        to_ptr, tcode = None, 0
        if time_outs is not None:
            time_outs, tcode = nat.flag_tensor(time_outs.to(self.device))
            to_ptr = time_outs.data_ptr()
        intr = None if intrinsic is None else nat.f32c(intrinsic)
        # fields that PPO.act already wrote in place (views of row `step`) need no copy
        skip = lambda src, dst: None if src.data_ptr() == dst.data_ptr() else src.data_ptr()
        nat.check(nat.lib().rslrl_store_transition(
            len(srcs), nat.void_array([t.data_ptr() for t in srcs]), nat.void_array(dsts), nat.int_array(dims),
            skip(actions, self.actions[s]), skip(mu, self.mu[s]), skip(sigma, self.sigma[s]), sigma_stride,
            values.data_ptr(), skip(logp, self.actions_log_prob[s]),
            rewards.data_ptr(), dones.data_ptr(), dcode, to_ptr, tcode, nat.ptr(intr), float(gamma),
            self.actions[s].data_ptr(), self.mu[s].data_ptr(), self.sigma[s].data_ptr(), self.values[s].data_ptr(),
            self.actions_log_prob[s].data_ptr(), self.rewards[s].data_ptr(), self.dones[s].data_ptr(), N, A,
            nat.stream_ptr()))
        self._save_hidden_states(tr.hidden_states)
        self.step += 1

    def _store_in_place(self, obs, rewards, dones, time_outs=None, intrinsic=None, gamma: float = 0.0):
        """Fast path of `_store` for PPO's fused rollout: actions / mean / sigma / value / log-prob of this step
        were already written into row `step` by `PPO.act`, so only observations, rewards (with the bootstrap)
        and dones move.  Everything fixed for a step index sits in a C-side plan (`rslrl_store_plan_t`) built once; per
        env step only the source pointers cross the ctypes boundary."""
        s = self.step
        if s >= self.num_transitions_per_env:
            raise OverflowError("Rollout buffer overflow! You should call clear() before adding new transitions.")
        plan = self._store_plans.get(s)
        if plan is None or plan[2] != gamma:
            sp = nat.StorePlan()
            for i, k in enumerate(self._obs_keys):
                d = self.observations[k][s]
                sp.obs_dst[i], sp.obs_dim[i] = d.data_ptr(), int(d[0].numel())
            sp.values_row, sp.rewards_row, sp.dones_row = self.values[s].data_ptr(), self.rewards[s].data_ptr(), self.dones[s].data_ptr()
            sp.gamma, sp.n_groups, sp.N, sp.A = float(gamma), len(self._obs_keys), self.num_envs, self._A
            import ctypes
            plan = (sp, nat.void_array([0] * len(self._obs_keys)), gamma, ctypes.byref(sp))
            self._store_plans[s] = plan
        src_arr, plan_ref = plan[1], plan[3]
        f32 = torch.float32
        keep = []   # converted copies must outlive the launch (their blocks could otherwise be recycled by the
        for i, k in enumerate(self._obs_keys):   # conversions below, which are enqueued before the kernel)
            src = obs[k]
            if src.dtype is not f32 or not src.is_contiguous():
                src = nat.f32c(src)
                keep.append(src)
            if not src.is_cuda:
                nat.require_cuda(src)
            src_arr[i] = src.data_ptr()
        if rewards.dtype is not f32 or not rewards.is_contiguous():
            rewards = nat.f32c(rewards)
        dcode = _FLAGS.get(dones.dtype)
        if dcode is None or not dones.is_contiguous():
            dones, dcode = nat.flag_tensor(dones)
        to_ptr, tcode = None, 0
        if time_outs is not None:
            if not time_outs.is_cuda and not time_outs.is_pinned():   # pinned host flags are read in place by the kernel
                time_outs = time_outs.to(self.device)
            tcode = _FLAGS.get(time_outs.dtype)
            if tcode is None or not time_outs.is_contiguous():
                time_outs, tcode = nat.flag_tensor(time_outs)
            to_ptr = time_outs.data_ptr()
        intr_t = None
        if intrinsic is not None:
            intr_t = nat.f32c(intrinsic)
            nat.require_cuda(intr_t)
        if not (rewards.is_cuda and dones.is_cuda and (time_outs is None or time_outs.is_cuda)):
            nat.require_device_readable(rewards, dones, time_outs)   # device, or pinned host memory read in place
        rc = _store_step()(plan_ref, src_arr, rewards.data_ptr(), dones.data_ptr(), dcode, to_ptr, tcode,
                           None if intr_t is None else intr_t.data_ptr(), nat.stream_ptr())
        if rc:
            nat.check(rc)
        del keep, intr_t
        self.step = s + 1

    def _save_hidden_states(self, hidden_states):
        """Reference rollout_storage.py:105-122 (LSTM: (h, c) tuples, GRU: single tensors)."""
        if hidden_states is None or hidden_states == (None, None):
            return
        hid_a = hidden_states[0] if isinstance(hidden_states[0], tuple) else (hidden_states[0],)
        hid_c = hidden_states[1] if isinstance(hidden_states[1], tuple) else (hidden_states[1],)
        T = self.num_transitions_per_env
        if self.saved_hidden_states_a is None:
            self.saved_hidden_states_a = [torch.zeros(T, *h.shape, device=self.device) for h in hid_a]
            self.saved_hidden_states_c = [torch.zeros(T, *h.shape, device=self.device) for h in hid_c]
        for i in range(len(hid_a)):
            self.saved_hidden_states_a[i][self.step].copy_(hid_a[i])
            self.saved_hidden_states_c[i][self.step].copy_(hid_c[i])

    def clear(self):
        self.step = 0

    # ---------------------------------------------------------------------------------------------
    # GAE
    # ---------------------------------------------------------------------------------------------
    def compute_returns(self, last_values, gamma, lam, normalize_advantage: bool = True):
        """Reference rollout_storage.py:127-149, as one cooperative kernel."""
        lv = nat.f32c(last_values.detach())
        nat.require_cuda(lv)
        nat.check(nat.lib().rslrl_gae_norm(
            self.rewards.data_ptr(), self.values.data_ptr(), self.dones.data_ptr(), lv.data_ptr(),
            self.returns.data_ptr(), self.advantages.data_ptr(), self._gae_ws.data_ptr(),
            self.num_transitions_per_env, self.num_envs, float(gamma), float(lam), int(bool(normalize_advantage)),
            nat.stream_ptr()))

    # ---------------------------------------------------------------------------------------------
    # minibatches (feed-forward)
    # ---------------------------------------------------------------------------------------------
    def generator(self):
        """Reference rollout_storage.py:152-157: the stored steps in order, as views of row t."""
        if self.training_type != "distillation":
            raise ValueError("This function is only available for distillation training.")
        for i in range(self.num_transitions_per_env):
            yield self.observations[i], self.actions[i], self.privileged_actions[i], self.dones[i]

    def _flat_fields(self):
        """(source tensors [T*N, d], widths) in the order obs groups..., actions, values, advantages, returns,
        old log-prob, old mu, old sigma."""
        fields = [self.observations[k] for k in self._obs_keys]
        fields += [self.actions, self.values, self.advantages, self.returns, self.actions_log_prob, self.mu, self.sigma]
        widths = [int(f[0, 0].numel()) for f in fields]
        return fields, widths

    def _gather(self, fields, widths, idx: torch.Tensor, outs, strides=None, cols=None):
        n = len(fields)
        strides = widths if strides is None else strides
        cols = [0] * n if cols is None else cols
        for start in range(0, n, nat.MAX_SEGMENTS):
            sl = slice(start, start + nat.MAX_SEGMENTS)
            nat.check(nat.lib().rslrl_permute_gather(
                len(fields[sl]), nat.void_array([f.data_ptr() for f in fields[sl]]),
                nat.void_array([o.data_ptr() for o in outs[sl]]), nat.int_array(widths[sl]), nat.int_array(strides[sl]),
                nat.int_array(cols[sl]), idx.data_ptr(), idx.numel(), nat.stream_ptr()))

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """Reference rollout_storage.py:160-203: one permutation per call, reused by every epoch; yields
        (obs, actions, values, advantages, returns, old_log_prob, old_mu, old_sigma, (None, None), None)."""
        if self.training_type != "rl":
            raise ValueError("This function is only available for reinforcement learning training.")
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = self._randperm(num_mini_batches * mini_batch_size)
        if indices.dtype != torch.int64:
            indices = indices.long()
        fields, widths = self._flat_fields()
        n_obs = len(self._obs_keys)
        f32 = dict(dtype=torch.float32, device=self.device)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mini_batch_size:(i + 1) * mini_batch_size]
                outs = [torch.empty(mini_batch_size, *f.shape[2:], **f32) for f in fields]
                self._gather(fields, widths, idx, outs)
                obs_batch = TensorDict(dict(zip(self._obs_keys, outs[:n_obs])), batch_size=[mini_batch_size],
                                       device=self.device)
                actions, values, advantages, returns, old_logp, old_mu, old_sigma = outs[n_obs:]
                yield obs_batch, actions, values, advantages, returns, old_logp, old_mu, old_sigma, (None, None), None

    # ---------------------------------------------------------------------------------------------
    # minibatches (recurrent)
    # ---------------------------------------------------------------------------------------------
    def recurrent_mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """Reference rollout_storage.py:206-260: no shuffling, a minibatch is a contiguous block of env
        columns; observations are split at dones and zero padded to T; hidden states are taken at the
        trajectory starts.  The trajectory index is built once (segmented scan, one host read of M+1
        offsets) and every epoch reuses the gathered tensors (they are epoch invariant)."""
        if self.training_type != "rl":
            raise ValueError("This function is only available for reinforcement learning training.")
        index = TrajectoryIndex(self.dones)
        offs = index.offsets_host()
        mini_batch_size = self.num_envs // num_mini_batches
        cached = []
        for i in range(num_mini_batches):
            start, stop = i * mini_batch_size, (i + 1) * mini_batch_size
            first_traj, last_traj = offs[start], offs[stop]
            obs_batch = TensorDict(
                {k: index.split_and_pad(self.observations[k], first_traj, last_traj) for k in self._obs_keys},
                batch_size=[self.num_transitions_per_env, last_traj - first_traj], device=self.device)
            masks_batch = index.masks(first_traj, last_traj)
            hid_a = hid_c = None
            if self.saved_hidden_states_a is not None:
                hid_a = [index.hidden_at_starts(s, first_traj, last_traj) for s in self.saved_hidden_states_a]
                hid_c = [index.hidden_at_starts(s, first_traj, last_traj) for s in self.saved_hidden_states_c]
                hid_a = hid_a[0] if len(hid_a) == 1 else hid_a
                hid_c = hid_c[0] if len(hid_c) == 1 else hid_c
            cached.append((start, stop, obs_batch, masks_batch, hid_a, hid_c))
        self._last_trajectory_index = index
        for _ in range(num_epochs):
            for start, stop, obs_batch, masks_batch, hid_a, hid_c in cached:
                yield (obs_batch, self.actions[:, start:stop], self.values[:, start:stop],
                       self.advantages[:, start:stop], self.returns[:, start:stop],
                       self.actions_log_prob[:, start:stop], self.mu[:, start:stop], self.sigma[:, start:stop],
                       (hid_a, hid_c), masks_batch)
