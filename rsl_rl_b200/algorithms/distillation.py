"""Teacher-student distillation (mirror of `rsl_rl/algorithms/distillation.py` of the reference).

Same constructor and methods (`init_storage, act, process_env_step, update, broadcast_parameters, reduce_parameters`), same
attributes (`policy, optimizer, storage, transition, last_hidden_states, num_updates, learning_rate, ...`).  Underneath:

* rollout: student and teacher forward on the library's GEMM / LSTM kernels, the step goes into the rollout buffer with
  one launch (`RolloutStorage._store_distillation`);
* update: the behaviour-cloning loss over `gradient_length` consecutive steps is differentiated through the hand-written
  backward of the MLP / LSTM kernels (networks/mlp.py, networks/lstm.py); gradients land in the flat gradient buffer,
  clip + Adam is the fused optimiser kernel (`FusedAdam.fused_step`);
* no host synchronisation inside the loop: the reference reads `behavior_loss.item()` every step
  (distillation.py:122), here the sum stays on the device (fp64, same order of additions) and is read once.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _native as nat
from ..storage import RolloutStorage
from ..utils import resolve_optimizer
from ..networks.mlp import linear_forward
from .flat_params import FlatParamGroup, FusedAdam
from .ppo import DEFAULT_GEMM_PRECISION


def _detached(hidden):
    """`hidden` (None, a tensor, or nested tuples of them) with every tensor detached."""
    if hidden is None:
        return None
    if isinstance(hidden, torch.Tensor):
        return hidden.detach()
    return tuple(_detached(h) for h in hidden)


class Distillation:
    """Distillation algorithm for training a student model to mimic a teacher model."""

    def __init__(self, policy, num_learning_epochs=1, gradient_length=15, learning_rate=1e-3, max_grad_norm=None,
                 loss_type="mse", optimizer="adam", device="cpu", multi_gpu_cfg: dict | None = None,
                 # --- rsl_rl_b200 extension (optional) ---
                 gemm_precision: str | None = None):   # see PPO: "tf32x3" (default) | "fp32" | "tf32" | "fp16" (-> tf32)
        nat.require_cuda_device(device, "rsl_rl_b200.Distillation")
        if gemm_precision is None:
            gemm_precision = os.environ.get("RSLRL_GEMM_PRECISION", DEFAULT_GEMM_PRECISION)
        if gemm_precision not in nat.PREC_CODES:
            raise ValueError(f"gemm_precision must be one of {sorted(nat.PREC_CODES)}, got {gemm_precision!r}")
        loss_fns = {"mse": nn.functional.mse_loss, "huber": nn.functional.huber_loss}
        if loss_type not in loss_fns:
            raise ValueError(f"Unknown loss type: {loss_type}. Supported types are: {list(loss_fns.keys())}")
        optimizer_cls = resolve_optimizer(optimizer)   # ValueError for unknown names, like the reference
        nat.lib()
        self.device = device
        self.is_multi_gpu = multi_gpu_cfg is not None
        if multi_gpu_cfg is not None:
            self.gpu_global_rank = multi_gpu_cfg["global_rank"]
            self.gpu_world_size = multi_gpu_cfg["world_size"]
        else:
            self.gpu_global_rank = 0
            self.gpu_world_size = 1

        self.policy = policy
        self.policy.to(self.device)
        self.policy.set_gemm_precision(gemm_precision)
        self.gemm_precision = gemm_precision
        self.storage: RolloutStorage = None  # type: ignore
        self.rnd = None   # (the runner's logging / checkpoint code asks every algorithm)

        # Every parameter of the policy, in `policy.parameters()` order like the reference's optimiser (so optimiser
        # state dicts index the same tensors), lives in one flat buffer.  Teacher and noise parameters never receive a
        # gradient: their flat gradient stays exactly zero and Adam leaves them unchanged.
        self._group = FlatParamGroup(self.policy.parameters(), self.device)
        if optimizer_cls is torch.optim.Adam:
            self.optimizer = FusedAdam(self._group, lr=learning_rate)
        else:   # AdamW / SGD / RMSprop: torch's own CUDA optimiser on the flat-buffer views
            self.optimizer = optimizer_cls(self._group.params, lr=learning_rate)
        self._fused_adam = isinstance(self.optimizer, FusedAdam)
        self._student_slice = self._flat_slice(self.policy.student)
        # parameters update() never differentiates (teacher side, noise): the reference's optimiser sees `grad is None`
        # for them and skips them -- which matters for optimisers that act on a parameter without a gradient (AdamW's
        # weight decay); their flat gradient views are hidden from a stock optimiser's step for the same effect
        learnable = {id(p) for name, mod in self.policy.named_children() if name in ("student", "memory_s")
                     for p in mod.parameters()}
        self._never_trained = [p for p in self._group.params if id(p) not in learnable]

        self.transition = RolloutStorage.Transition()
        self.last_hidden_states = None

        self.num_learning_epochs = num_learning_epochs
        self.gradient_length = gradient_length
        self.learning_rate = learning_rate
        self.max_grad_norm = max_grad_norm
        self.loss_fn = loss_fns[loss_type]
        self.num_updates = 0
        self._loss_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._mb_hook = None   # test hook: called after every optimiser step
        # the stock feed-forward pair: grouped student / teacher launches in the rollout, batched windows in update()
        from ..modules.student_teacher import StudentTeacher
        plain = type(policy) is StudentTeacher and os.environ.get("RSLRL_DISTILL_BATCHED", "1") != "0"
        self._batched_update = plain
        # a student with a native LSTM memory: every run of steps between two optimiser steps as ONE multi-step LSTM call
        # over the trajectories the dones cut it into (RSLRL_DISTILL_WINDOWED=0: the step-by-step restatement of the
        # reference's loop, which GRU students always use)
        from ..modules.student_teacher_recurrent import StudentTeacherRecurrent
        self._windowed_recurrent = (type(policy) is StudentTeacherRecurrent and policy.memory_s.native
                                    and os.environ.get("RSLRL_DISTILL_WINDOWED", "1") != "0")
        self._grouped_rollout = plain and (len(policy.student.linear_layers()) == len(policy.teacher.linear_layers())
                                           and not any(isinstance(m, nn.Unflatten) for net in (policy.student, policy.teacher)
                                                       for m in net))
        self.last_update_info = {}

    def _flat_slice(self, module: nn.Module) -> slice:
        """The contiguous span of the flat buffers that holds `module`'s parameters (padding between tensors is zero)."""
        ids = {id(p) for p in module.parameters()}
        idx = [i for i, p in enumerate(self._group.params) if id(p) in ids]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise RuntimeError("the student's parameters are not contiguous in policy.parameters()")
        g = self._group
        return slice(g.offsets[idx[0]], g.offsets[idx[-1]] + g.params[idx[-1]].numel())

    # ------------------------------------------------------------------------------------------
    def init_storage(self, training_type, num_envs, num_transitions_per_env, obs, actions_shape):
        self.storage = RolloutStorage(training_type, num_envs, num_transitions_per_env, obs, actions_shape, self.device)

    def act(self, obs):
        """Reference distillation.py:84-90."""
        tr = self.transition
        if self._grouped_rollout:
            tr.actions, tr.privileged_actions = self._act_grouped(obs)
        else:
            tr.actions = self.policy.act(obs).detach()
            tr.privileged_actions = self.policy.evaluate(obs).detach()
        tr.observations = obs
        return tr.actions

    def _act_grouped(self, obs):
        """`policy.act(obs)` and `policy.evaluate(obs)` of a feed-forward StudentTeacher with layer i of the student and
        layer i of the teacher in ONE grouped GEMM launch (no autograd: both results are detached by the caller anyway).
        Same kernels, same operands, same results as the two module calls."""
        pol = self.policy
        hs = nat.f32c(pol.student_obs_normalizer(pol.get_student_obs(obs)))
        ht = nat.f32c(pol.teacher_obs_normalizer(pol.get_teacher_obs(obs)))
        nat.require_cuda(hs, ht)
        prec = nat.PREC_CODES[pol.student.precision]
        s_layers, t_layers = pol.student.linear_layers(), pol.teacher.linear_layers()
        s_acts, t_acts = pol.student.layer_activation_codes(), pol.teacher.layer_activation_codes()
        f32 = dict(dtype=torch.float32, device=hs.device)
        with torch.no_grad():
            for ls, lt, a_s, a_t in zip(s_layers, t_layers, s_acts, t_acts):
                ys = torch.empty(hs.shape[0], ls.weight.shape[0], **f32)
                yt = torch.empty(ht.shape[0], lt.weight.shape[0], **f32)
                linear_forward([(hs, ls.weight, ls.bias, ys, a_s), (ht, lt.weight, lt.bias, yt, a_t)], prec)
                hs, ht = ys, yt
            pol.set_distribution(hs)
            return pol._sample(), ht

    def process_env_step(self, obs, rewards, dones, extras):
        """Reference distillation.py:92-102."""
        self.policy.update_normalization(obs)
        tr = self.transition
        tr.rewards = rewards
        tr.dones = dones
        self.storage.add_transitions(tr)
        tr.clear()
        self.policy.reset(dones)

    # ------------------------------------------------------------------------------------------
    def update(self):
        """Reference distillation.py:104-150.  The loss of `gradient_length` consecutive steps is summed (across epoch
        boundaries too: `cnt` and the running sum are not reset between epochs), differentiated once, then one optimiser
        step; hidden states are carried through the whole rollout and cut at the optimiser steps and at dones."""
        self.num_updates += 1
        pol, st = self.policy, self.storage
        self._group.ensure_bound()
        if self._fused_adam:
            self.optimizer.dev_state.set_lr(self.learning_rate)
        self._group.zero_grad()
        self._loss_sum.zero_()
        if self._batched_update:
            return self._update_feedforward()
        if self._windowed_recurrent:
            return self._update_recurrent_windows()
        loss = 0
        cnt = 0
        for _ in range(self.num_learning_epochs):
            pol.reset(hidden_states=self.last_hidden_states)
            pol.detach_hidden_states()
            for obs, _, privileged_actions, dones in st.generator():
                actions = pol.act_inference(obs)
                behavior_loss = self.loss_fn(actions, privileged_actions)
                loss = loss + behavior_loss
                self._loss_sum += behavior_loss.detach()
                cnt += 1
                if cnt % self.gradient_length == 0:
                    loss.backward()   # accumulates into the (zeroed) flat gradient views
                    self._optimizer_step()
                    pol.detach_hidden_states()
                    loss = 0
                    if self._mb_hook is not None:
                        self._mb_hook()
                pol.reset(dones.view(-1))
                pol.detach_hidden_states(dones.view(-1))
        mean_behavior_loss = float(self._loss_sum.item()) / cnt   # the one device->host read of update()
        st.clear()
        # (the reference keeps the attached tensors here and detaches them at the start of the next update; the values are
        #  the same, detaching now releases the autograd graph of the last `cnt % gradient_length` steps)
        self.last_hidden_states = _detached(pol.get_hidden_states())
        pol.detach_hidden_states()
        if self._fused_adam:
            self.optimizer.sync_host_state(self.learning_rate)
        return {"behavior": mean_behavior_loss}

    def _update_feedforward(self):
        """update() of a feed-forward student: a student without memory sees every stored step independently and the
        weights only change at the optimiser steps, so the `gradient_length` steps between two optimiser steps are ONE
        batch of gradient_length * num_envs rows -- one forward, one backward per optimiser step instead of
        gradient_length of each, on GEMMs large enough for the tensor cores.  The objective is the same sum of per-step
        mean losses (so is its gradient, up to the order of fp32 additions); windows that run across an epoch boundary
        (the reference does not reset its step counter between epochs) are two contiguous row ranges of the buffer."""
        pol, st = self.policy, self.storage
        T, gl = st.num_transitions_per_env, self.gradient_length
        total = self.num_learning_epochs * T
        n_steps = 0

        def window_loss(first: int, count: int):
            """Sum over steps [first, first + count) (modulo T) of the per-step mean loss, as a device scalar."""
            out = 0
            k = first
            while k < first + count:
                a = k % T
                b = min(a + (first + count - k), T)
                obs = st.observations[a:b].flatten(0, 1)
                target = st.privileged_actions[a:b]
                actions = pol.act_inference(obs).view(target.shape)
                per_step = self.loss_fn(actions, target, reduction="none").flatten(1).mean(dim=1)
                self._loss_sum += per_step.detach().sum(dtype=torch.float64)
                out = out + per_step.sum()
                k += b - a
            return out

        while n_steps + gl <= total:
            window_loss(n_steps, gl).backward()   # accumulates into the (zeroed) flat gradient views
            self._optimizer_step()
            n_steps += gl
            if self._mb_hook is not None:
                self._mb_hook()
        if n_steps < total:   # the trailing steps only enter the reported mean (the reference drops their gradient)
            with torch.no_grad():
                window_loss(n_steps, total - n_steps)
        mean_behavior_loss = float(self._loss_sum.item()) / total   # the one device->host read of update()
        st.clear()
        self.last_hidden_states = pol.get_hidden_states()
        if self._fused_adam:
            self.optimizer.sync_host_state(self.learning_rate)
        return {"behavior": mean_behavior_loss}

    def _update_recurrent_windows(self):
        """update() of a student with a native LSTM memory, one multi-step LSTM call per window instead of one call per
        stored step.  Gradients never leave a window (hidden states are detached after every optimiser step) and never cross
        a done (the state is zeroed there), so a run of consecutive steps is a set of independent trajectories -- exactly
        what the recurrent PPO update is made of: the run's dones are scanned into a trajectory index, the normalised
        observations are split and padded, the LSTM runs over all of them at once (hoisted input projection, time loop in C,
        backward through time over the padded grid), the outputs are un-padded into [steps, envs] for the student MLP.
        A trajectory that starts at the first step of the run continues from the carried state, every other one from
        zero; the state carried out of the run is every environment's state at the end of its last trajectory (read from
        the LSTM's state buffers), zeroed where that step was done.  A window that crosses the epoch boundary is two runs:
        the reference re-installs the start-of-rollout state at every epoch start while its loss sum and step counter
        carry on."""
        from ..networks.lstm import lstm_forward, tap_states
        from ..utils.utils import TrajectoryIndex, unpad_trajectories
        pol, st = self.policy, self.storage
        mem = pol.memory_s
        T, gl, N = st.num_transitions_per_env, self.gradient_length, st.num_envs
        L, H = mem.rnn.num_layers, mem.rnn.hidden_size

        def run_loss(t0: int, count: int):
            """Sum of the per-step mean losses of stored steps [t0, t0 + count); advances `mem.hidden_states`."""
            obs = st.observations[t0:t0 + count].flatten(0, 1)
            x = pol.student_obs_normalizer(pol.get_student_obs(obs))
            x = x.reshape(count, N, x.shape[-1])
            dones = st.dones[t0:t0 + count]
            index = TrajectoryIndex(dones)
            n_traj = index.n_traj                                    # the one host read of the run
            padded = index.split_and_pad(x, 0, n_traj)               # [count, n_traj, D], zero padded
            masks = index.masks(0, n_traj)
            envs = index.traj_env[:n_traj].long()
            prev = mem.hidden_states
            if prev is None:
                h0 = c0 = torch.zeros(L, n_traj, H, dtype=torch.float32, device=padded.device)
            else:
                continues = (index.traj_start[:n_traj] == 0).view(1, n_traj, 1)
                h0, c0 = (torch.where(continues, s.detach()[:, envs], torch.zeros((), dtype=s.dtype, device=s.device))
                          for s in prev)
            with tap_states() as layers:
                out, _ = lstm_forward(mem.rnn, padded, (h0, c0), mem.precision)
            feats = unpad_trajectories(out, masks)                   # [count, N, H]
            target = st.privileged_actions[t0:t0 + count]
            actions = pol.student(feats.reshape(count * N, H)).view(target.shape)
            per_step = self.loss_fn(actions, target, reduction="none").flatten(1).mean(dim=1)
            self._loss_sum += per_step.detach().sum(dtype=torch.float64)
            # every environment's state after the run: the end of its last trajectory, zero if that step was done
            last = index.env_traj_offset[1:N + 1].long() - 1
            length = index.traj_len[:n_traj].long()[last]
            alive = (dones[count - 1].view(1, N, 1) != 1)
            h_n = torch.stack([h_all[length, last] for (h_all, _), in layers]) * alive
            c_n = torch.stack([c_all[length, last] for (_, c_all), in layers]) * alive
            mem.hidden_states = (h_n, c_n)
            if pol.teacher_recurrent and pol.memory_t.hidden_states is not None:   # (only zeroed at dones, never advanced)
                for t in range(count):
                    pol.memory_t.reset(dones[t].view(-1))
            return per_step.sum()

        total, done_steps, in_window, loss = self.num_learning_epochs * T, 0, 0, 0
        trainable = (total // gl) * gl        # steps after the last full window only enter the reported mean
        for _ in range(self.num_learning_epochs):
            pol.reset(hidden_states=self.last_hidden_states)
            pol.detach_hidden_states()
            t = 0
            while t < T:
                if done_steps >= trainable:
                    with torch.no_grad():
                        run_loss(t, T - t)
                    done_steps, t = done_steps + (T - t), T
                    continue
                count = min(gl - in_window, T - t)
                loss = loss + run_loss(t, count)
                in_window, t, done_steps = in_window + count, t + count, done_steps + count
                if in_window == gl:
                    loss.backward()   # accumulates into the (zeroed) flat gradient views
                    self._optimizer_step()
                    pol.detach_hidden_states()
                    loss, in_window = 0, 0
                    if self._mb_hook is not None:
                        self._mb_hook()
        mean_behavior_loss = float(self._loss_sum.item()) / total
        st.clear()
        self.last_hidden_states = _detached(pol.get_hidden_states())
        pol.detach_hidden_states()
        if self._fused_adam:
            self.optimizer.sync_host_state(self.learning_rate)
        return {"behavior": mean_behavior_loss}

    def _optimizer_step(self):
        """zero_grad / backward happened; now (multi-GPU) average, clip the STUDENT MLP's gradients (reference
        distillation.py:131-132 clips `policy.student.parameters()` only -- a recurrent student's memory is not
        clipped), step, and leave the gradient buffer zeroed for the next accumulation."""
        g = self._group
        if self.is_multi_gpu:
            self.reduce_parameters()
        if self.max_grad_norm:
            sl = g.grads[self._student_slice]
            total_norm = torch.linalg.vector_norm(sl)
            sl.mul_(torch.clamp(self.max_grad_norm / (total_norm + 1e-6), max=1.0))
        if self._fused_adam:
            self.optimizer.fused_step(world_size=1, max_grad_norm=0.0, adaptive=False, use_kl=False, zero_grad=True)
        else:
            hidden = [(p, p.grad) for p in self._never_trained]
            for p, _ in hidden:
                p.grad = None
            try:
                self.optimizer.step()
            finally:
                for p, grad in hidden:
                    p.grad = grad
            g.zero_grad()

    # ------------------------------------------------------------------------------------------
    # multi-GPU helpers (reference distillation.py:156-185)
    # ------------------------------------------------------------------------------------------
    def broadcast_parameters(self):
        """Rank 0's parameters and buffers to every rank: one flat broadcast (the reference pickles the state dict)."""
        self._group.ensure_bound()
        torch.distributed.broadcast(self._group.values, src=0)
        for b in self.policy.buffers():
            torch.distributed.broadcast(b, src=0)

    def reduce_parameters(self):
        """Averages the gradients over all ranks in place: one all-reduce of the flat gradient buffer."""
        self._group.ensure_bound()
        torch.distributed.all_reduce(self._group.grads, op=torch.distributed.ReduceOp.SUM)
        self._group.grads /= self.gpu_world_size
