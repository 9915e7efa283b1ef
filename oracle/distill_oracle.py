"""TEST INFRASTRUCTURE -- numpy restatement of the reference's teacher-student distillation for a FEED-FORWARD student
(SURVEY 8 f4).  Only `tests/` may import it; it is pinned to golden traces recorded from the unmodified reference
(`tests/golden/distill_mlp_small.npz`, `distill_mlp_adamw.npz`; `tests/test_oracle_vs_golden.py`).

What it follows, step by step:
  * `StudentTeacher.act / evaluate`                (rsl_rl/modules/student_teacher.py:123-138)
  * `EmpiricalNormalization.forward / update`       (rsl_rl/networks/normalization.py:44-66)
  * `Distillation.act / process_env_step`           (rsl_rl/algorithms/distillation.py:84-102)
  * `RolloutStorage` in "distillation" mode         (rsl_rl/storage/rollout_storage.py:57-59,88-90,152-157)
  * `Distillation.update`                           (rsl_rl/algorithms/distillation.py:104-150): the loss of
    `gradient_length` consecutive steps is summed and differentiated once; the step counter and the running sum carry
    across epochs; steps after the last full window only enter the reported mean; clip on the student's parameters only.
Recurrent students are pinned by the golden traces alone (`distill_lstm_small`, `distill_lstm2_cross`, `distill_gru_small`).
Dense layers, Adam and the norm clip are the PPO oracle's (`oracle/ppo_oracle.py`).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import ppo_oracle as O

F32 = np.float32


class Normalizer:
    """Running mean / variance with the reference's batched (Chan) update in its fp32 operation order."""

    def __init__(self, mean, var, std, count, eps=1e-2):
        self.mean, self.var, self.std = (np.array(a, dtype=F32).reshape(1, -1) for a in (mean, var, std))
        self.count, self.eps = int(count), F32(eps)

    def __call__(self, x):
        return ((x - self.mean) / (self.std + self.eps)).astype(F32)

    def update(self, x):
        n = x.shape[0]
        self.count += n
        rate = F32(n / self.count)
        var_x = x.astype(F32).var(axis=0, keepdims=True, dtype=F32)
        mean_x = x.astype(F32).mean(axis=0, keepdims=True, dtype=F32)
        delta = mean_x - self.mean
        self.mean = (self.mean + rate * delta).astype(F32)
        self.var = (self.var + rate * (var_x - self.var + delta * (mean_x - self.mean))).astype(F32)
        self.std = np.sqrt(self.var)


def _normalizer(state: dict, prefix: str):
    if prefix + "._mean" not in state:
        return lambda x: x
    return Normalizer(state[prefix + "._mean"], state[prefix + "._var"], state[prefix + "._std"], state[prefix + ".count"])


def _loss_and_grad(kind: str, actions, target):
    """mean-reduced MSE / Huber (delta 1) and its gradient w.r.t. `actions` (torch.nn.functional.mse_loss / huber_loss)."""
    d = (actions - target).astype(F32)
    n = F32(d.size)
    if kind == "mse":
        return float((d.astype(np.float64) ** 2).mean()), (F32(2.0) * d / n).astype(F32)
    if kind == "huber":
        a = np.abs(d)
        val = np.where(a < 1, 0.5 * d.astype(np.float64) ** 2, a.astype(np.float64) - 0.5).mean()
        return float(val), (np.where(a < 1, d, np.sign(d)) / n).astype(F32)
    raise ValueError(kind)


def rollout(state: dict, obs_student, obs_teacher, z, rewards, dones, cfg):
    """T steps of `Distillation.act` + `process_env_step` on a recorded stream.  obs_*: [T+1, N, d] (row t is what the
    policy sees at step t; the student's normaliser is updated with row t+1, distillation.py:94).  Returns the storage
    dict and mutates the student normaliser statistics inside `state`."""
    act = cfg.get("activation", "elu")
    T = rewards.shape[0]
    s_norm, t_norm = _normalizer(state, "student_obs_normalizer"), _normalizer(state, "teacher_obs_normalizer")
    student, teacher = O.mlp_layers(state, "student"), O.mlp_layers(state, "teacher")
    sigma = state["std"] if "std" in state else np.exp(state["log_std"])
    st = dict(obs_student=obs_student[:T].astype(F32), actions=[], privileged_actions=[],
              rewards=rewards.astype(F32)[..., None], dones=dones.astype(np.uint8)[..., None])
    for t in range(T):
        mean, _ = O.mlp_forward(student, s_norm(obs_student[t]), act)
        st["actions"].append(O.gaussian_sample(mean, np.broadcast_to(sigma, mean.shape).astype(F32), z[t]))
        st["privileged_actions"].append(O.mlp_forward(teacher, t_norm(obs_teacher[t]), act)[0])
        if isinstance(s_norm, Normalizer):
            s_norm.update(obs_student[t + 1])
    st["actions"], st["privileged_actions"] = np.stack(st["actions"]), np.stack(st["privileged_actions"])
    if isinstance(s_norm, Normalizer):
        state["student_obs_normalizer._mean"], state["student_obs_normalizer._var"] = s_norm.mean, s_norm.var
        state["student_obs_normalizer._std"], state["student_obs_normalizer.count"] = s_norm.std, np.array(s_norm.count)
    return st


def update(state: dict, adam_state, storage, cfg, optimizer_step=None):
    """`Distillation.update` for a feed-forward student.  Mutates the `student.*` entries of `state` in place; returns
    (mean behaviour loss, [flat student parameters after every optimiser step]).  `optimizer_step(params, grads)`
    replaces Adam (e.g. AdamW) when given."""
    act, kind = cfg.get("activation", "elu"), cfg.get("loss_type", "mse")
    gl, E, lr, max_norm = cfg["gradient_length"], cfg["num_learning_epochs"], cfg["learning_rate"], cfg.get("max_grad_norm")
    s_norm = _normalizer(state, "student_obs_normalizer")
    names = [k for k in state if k.startswith("student.")]
    params = OrderedDict((k, state[k]) for k in names)
    T = storage["obs_student"].shape[0]
    acc = OrderedDict((k, np.zeros_like(v)) for k, v in params.items())
    loss_sum, cnt, trace = 0.0, 0, []
    for _ in range(E):
        for t in range(T):
            layers = O.mlp_layers(state, "student")
            actions, cache = O.mlp_forward(layers, s_norm(storage["obs_student"][t]), act)
            value, dact = _loss_and_grad(kind, actions, storage["privileged_actions"][t])
            loss_sum += value
            cnt += 1
            grads, _ = O.mlp_backward(layers, cache, dact, act)      # d(sum of step losses): accumulate step by step
            for li, (gw, gb) in enumerate(grads):
                acc[f"student.{2 * li}.weight"] += gw
                acc[f"student.{2 * li}.bias"] += gb
            if cnt % gl == 0:
                g = OrderedDict((k, v.copy()) for k, v in acc.items())
                if max_norm:
                    O.clip_grad_norm(g, max_norm)                      # student MLP only (distillation.py:131-132)
                if optimizer_step is None:
                    O.adam_step(params, g, adam_state, lr)
                else:
                    optimizer_step(params, g)
                for v in acc.values():
                    v[...] = 0
                trace.append(np.concatenate([v.reshape(-1) for v in params.values()]))
        # (the reference keeps `loss` and `cnt` across the epoch boundary: nothing is reset here)
    return loss_sum / cnt, trace


def adamw_step(params, grads, state, lr, weight_decay=1e-2, **adam_kw):
    """torch.optim.AdamW defaults: decoupled decay `p *= 1 - lr * weight_decay`, then the Adam update; in place."""
    for p in params.values():
        p *= F32(1.0 - lr * weight_decay)
    O.adam_step(params, grads, state, lr, **adam_kw)
