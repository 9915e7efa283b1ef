#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- regenerate tests/golden/*.npz by running the UNMODIFIED reference.

Usage (build container only; the GPU box has no /root/reference):

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

The reference is imported from `--ref` with `oracle/_shim` on sys.path (stand-in for the missing
`tensordict` package).  Nothing is patched inside the reference; the script only wraps *torch* entry
points (torch.normal, torch.randperm, clip_grad_norm_, Optimizer.step) to record the random draws and
per-minibatch intermediates that the reference does not expose.  Everything runs on CPU, fp32,
torch defaults (no TF32), single thread count as configured by torch.
"""

from __future__ import annotations

import argparse
import os
import sys
import tempfile
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference(ref):
    sys.path.insert(0, os.path.join(HERE, "_shim"))
    sys.path.insert(0, ref)
    import torch  # noqa: F401
    import rsl_rl  # noqa: F401

    assert os.path.realpath(rsl_rl.__file__).startswith(os.path.realpath(ref)), rsl_rl.__file__


# ---------------------------------------------------------------------------------------------
# recording helpers
# ---------------------------------------------------------------------------------------------

class Recorder:
    """Wraps torch entry points; `with Recorder() as rec:` restores them on exit."""

    def __init__(self):
        self.normal_z = []       # list of z arrays, one per torch.normal(mean, std) call
        self.perms = []          # list of randperm outputs
        self.clip_calls = []     # (flat raw grads, total_norm)
        self.step_calls = []     # (lr, flat params after step)  for the policy optimiser only
        self.policy_optimizer = None

    def __enter__(self):
        import torch

        self._torch = torch
        self._normal = torch.normal
        self._randperm = torch.randperm
        self._clip = torch.nn.utils.clip_grad_norm_
        rec = self

        def normal(mean, std, *a, **k):
            state = torch.get_rng_state()
            out = rec._normal(mean, std, *a, **k)
            after = torch.get_rng_state()
            torch.set_rng_state(state)
            z = torch.empty_like(out).normal_()
            assert torch.equal(torch.get_rng_state(), after), "normal_(0,1) does not reproduce torch.normal's stream"
            assert torch.equal(z * std + mean, out), "torch.normal != normal_(0,1)*std+mean"
            rec.normal_z.append(z.numpy().copy())
            return out

        def randperm(n, *a, **k):
            out = rec._randperm(n, *a, **k)
            rec.perms.append(out.numpy().copy())
            return out

        def clip(parameters, max_norm, *a, **k):
            params = list(parameters)
            raw = torch.cat([p.grad.reshape(-1) for p in params]).numpy().copy()
            total = rec._clip(params, max_norm, *a, **k)
            rec.clip_calls.append((raw, float(total)))
            return total

        torch.normal = normal
        torch.randperm = randperm
        torch.nn.utils.clip_grad_norm_ = clip
        return self

    def watch_optimizer(self, optimizer, module):
        torch = self._torch
        rec = self
        orig_step = optimizer.step

        def step(*a, **k):
            r = orig_step(*a, **k)
            flat = torch.cat([p.detach().reshape(-1) for p in module.parameters()]).numpy().copy()
            rec.step_calls.append((float(optimizer.param_groups[0]["lr"]), flat))
            return r

        optimizer.step = step

    def __exit__(self, *exc):
        torch = self._torch
        torch.normal = self._normal
        torch.randperm = self._randperm
        torch.nn.utils.clip_grad_norm_ = self._clip
        return False


def make_env_class():
    import torch
    from tensordict import TensorDict
    from rsl_rl.env import VecEnv

    class RecordingEnv(VecEnv):
        """Seeded synthetic env; remembers everything it emitted."""

        def __init__(self, num_envs, group_dims, num_actions, seed=0, p_done=0.02, p_timeout=0.5):
            self.num_envs, self.num_actions = num_envs, num_actions
            self.group_dims = dict(group_dims)
            self.max_episode_length = 1000
            self.episode_length_buf = torch.zeros(num_envs, dtype=torch.long)
            self.device = "cpu"
            self.cfg = {}
            self.step_dt = 0.02
            self.unwrapped = self
            self.p_done, self.p_timeout = p_done, p_timeout
            self.gen = torch.Generator().manual_seed(seed)
            self.obs_log, self.rew_log, self.done_log, self.to_log, self.act_log = [], [], [], [], []
            self._obs = self._draw_obs()

        def _draw_obs(self):
            obs = {k: torch.randn(self.num_envs, d, generator=self.gen) for k, d in self.group_dims.items()}
            self.obs_log.append({k: v.numpy().copy() for k, v in obs.items()})
            return TensorDict(obs, batch_size=[self.num_envs], device="cpu")

        def get_observations(self):
            return self._obs

        def step(self, actions):
            self.act_log.append(actions.numpy().copy())
            rew = torch.randn(self.num_envs, generator=self.gen)
            dones = (torch.rand(self.num_envs, generator=self.gen) < self.p_done).long()
            time_outs = (torch.rand(self.num_envs, generator=self.gen) < self.p_timeout) & (dones > 0)
            self._obs = self._draw_obs()
            self.rew_log.append(rew.numpy().copy())
            self.done_log.append(dones.numpy().copy())
            self.to_log.append(time_outs.numpy().copy())
            return self._obs, rew, dones, {"time_outs": time_outs}

    return RecordingEnv


def _state_arrays(module):
    return OrderedDict((k, v.detach().numpy().copy()) for k, v in module.state_dict().items())


def _storage_arrays(storage):
    out = {f"obs.{k}": v.numpy().copy() for k, v in storage.observations.items()}
    for name in ("rewards", "actions", "dones", "values", "actions_log_prob", "mu", "sigma", "returns", "advantages"):
        out[name] = getattr(storage, name).numpy().copy()
    return out


# ---------------------------------------------------------------------------------------------
# scenarios
# ---------------------------------------------------------------------------------------------

def scenario_ppo_mlp(out_dir, name, *, num_envs, num_steps, group_dims, obs_groups, num_actions, hidden,
                     activation, noise_std_type, alg_overrides, iterations, seed=0, p_done=0.1,
                     policy_class="ActorCritic", policy_extra=None, rnd_cfg=None, symmetry=False, checkpoint=False):
    import torch
    from rsl_rl.runners import OnPolicyRunner

    RecordingEnv = make_env_class()
    torch.manual_seed(seed)
    env = RecordingEnv(num_envs, group_dims, num_actions, seed=seed + 100, p_done=p_done)
    alg = dict(class_name="PPO", learning_rate=1e-3, num_learning_epochs=2, num_mini_batches=4,
               schedule="adaptive", value_loss_coef=1.0, clip_param=0.2, use_clipped_value_loss=True,
               desired_kl=0.01, entropy_coef=0.01, gamma=0.99, lam=0.95, max_grad_norm=1.0,
               normalize_advantage_per_mini_batch=False)
    alg.update(alg_overrides)
    alg_run = dict(alg)
    if rnd_cfg is not None:
        alg_run["rnd_cfg"] = dict(rnd_cfg)
    symmetry_flags = dict(use_data_augmentation=True, use_mirror_loss=True, mirror_loss_coeff=0.5)
    if isinstance(symmetry, dict):   # other combinations of the two switches (ppo.py:226-244, 318-348)
        symmetry_flags.update(symmetry)
    if symmetry:
        sys.path.insert(0, os.path.join(HERE, ".."))
        from oracle.symmetry_example import mirror_augmentation
        alg_run["symmetry_cfg"] = dict(symmetry_flags, data_augmentation_func=mirror_augmentation)
    cfg = dict(num_steps_per_env=num_steps, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in obs_groups.items()},
               policy=dict(class_name=policy_class, actor_hidden_dims=list(hidden), critic_hidden_dims=list(hidden),
                           activation=activation, init_noise_std=1.0, noise_std_type=noise_std_type,
                           **(policy_extra or {})),
               algorithm=alg_run)
    data = {}
    meta = dict(num_envs=num_envs, num_steps=num_steps, num_actions=num_actions, hidden=list(hidden),
                activation=activation, noise_std_type=noise_std_type, iterations=iterations,
                obs_groups={k: list(v) for k, v in obs_groups.items()}, group_dims=dict(group_dims),
                alg={k: v for k, v in alg.items() if k != "class_name"}, policy_class=policy_class,
                policy_extra=dict(policy_extra or {}), rnd_cfg=None if rnd_cfg is None else dict(rnd_cfg),
                symmetry=bool(symmetry), symmetry_cfg=dict(symmetry_flags))
    with Recorder() as rec:
        runner = OnPolicyRunner(env, cfg, log_dir=tempfile.mkdtemp(prefix="golden_"), device="cpu")
        rec.watch_optimizer(runner.alg.optimizer, runner.alg.policy)
        logged = []                       # what the reference's log() sees: episode reward / length buffers (f1)
        orig_log = runner.log

        def log_and_record(locs, *a, **k):
            logged.append(dict(rew=list(locs["rewbuffer"]), length=list(locs["lenbuffer"]),
                               loss={kk: float(vv) for kk, vv in locs["loss_dict"].items()}))
            return orig_log(locs, *a, **k)

        runner.log = log_and_record
        for k, v in _state_arrays(runner.alg.policy).items():
            data[f"init.{k}"] = v
        meta["param_names"] = [k for k, _ in runner.alg.policy.named_parameters()]
        if runner.alg.rnd:
            for k, v in _state_arrays(runner.alg.rnd).items():
                data[f"init_rnd.{k}"] = v
        for it in range(iterations):
            n_norm, n_perm, n_clip, n_step = len(rec.normal_z), len(rec.perms), len(rec.clip_calls), len(rec.step_calls)
            first_obs = len(env.obs_log) - 1
            first_step = len(env.rew_log)
            runner.learn(1)
            p = f"it{it}."
            T = num_steps
            # env stream seen by this iteration: obs rows first_obs .. first_obs+T
            for g in group_dims:
                data[p + f"stream.obs.{g}"] = np.stack([env.obs_log[first_obs + t][g] for t in range(T + 1)])
            data[p + "stream.rewards"] = np.stack(env.rew_log[first_step:first_step + T])
            data[p + "stream.dones"] = np.stack(env.done_log[first_step:first_step + T])
            data[p + "stream.time_outs"] = np.stack(env.to_log[first_step:first_step + T])
            zs = rec.normal_z[n_norm:]
            data[p + "z_rollout"] = np.stack(zs[:T])                 # [T,N,A]
            data[p + "z_update_count"] = np.array(len(zs[T:]))        # discarded draws (ppo.py:249)
            if policy_class == "ActorCritic":
                assert len(rec.perms) == n_perm + 1
                data[p + "perm"] = rec.perms[-1]
            for k, v in _storage_arrays(runner.alg.storage).items():
                data[p + "storage." + k] = v
            clips = rec.clip_calls[n_clip:]
            steps = rec.step_calls[n_step:]
            data[p + "mb.raw_grads"] = np.stack([c[0] for c in clips])
            data[p + "mb.total_norm"] = np.array([c[1] for c in clips])
            data[p + "mb.lr"] = np.array([s[0] for s in steps])
            data[p + "mb.params_after"] = np.stack([s[1] for s in steps])
            for k, v in _state_arrays(runner.alg.policy).items():
                data[p + f"final.{k}"] = v
            if runner.alg.rnd:
                for k, v in _state_arrays(runner.alg.rnd).items():
                    data[p + f"final_rnd.{k}"] = v
            if runner.alg.storage.saved_hidden_states_a is not None:
                for i, h in enumerate(runner.alg.storage.saved_hidden_states_a):
                    data[p + f"storage.hidden_a{i}"] = h.numpy().copy()
                for i, h in enumerate(runner.alg.storage.saved_hidden_states_c):
                    data[p + f"storage.hidden_c{i}"] = h.numpy().copy()
            ost = runner.alg.optimizer.state_dict()["state"]
            data[p + "adam.exp_avg"] = np.concatenate([ost[i]["exp_avg"].numpy().reshape(-1) for i in range(len(ost))])
            data[p + "adam.exp_avg_sq"] = np.concatenate([ost[i]["exp_avg_sq"].numpy().reshape(-1) for i in range(len(ost))])
            data[p + "adam.step"] = np.array(float(ost[0]["step"]))
            data[p + "learning_rate"] = np.array(runner.alg.learning_rate, dtype=np.float64)
            if logged:
                data[p + "log.rewbuffer"] = np.array(logged[-1]["rew"], dtype=np.float64)
                data[p + "log.lenbuffer"] = np.array(logged[-1]["length"], dtype=np.float64)
                for kk, vv in logged[-1]["loss"].items():
                    data[p + "log.loss." + kk] = np.array(vv, dtype=np.float64)
        if checkpoint:   # a reference-written checkpoint (on_policy_runner.py:289-305) for the round-trip test (f3)
            ckpt = os.path.join(out_dir, name + "_ref_checkpoint.pt")
            runner.save(ckpt)
            print("wrote", ckpt, f"{os.path.getsize(ckpt) / 1024:.0f} KiB")
    data["meta"] = np.array(repr(meta))
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **data)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def scenario_distillation(out_dir, name, *, num_envs, num_steps, group_dims, obs_groups, teacher_groups, num_actions,
                          hidden, activation, policy_class, policy_extra, teacher_class, teacher_extra, alg_overrides,
                          iterations, seed=0, p_done=0.1, checkpoint=False):
    """Teacher-student distillation: `DistillationRunner.learn` of the unmodified reference
    (runners/distillation_runner.py:56-150, algorithms/distillation.py:84-150) with the teacher taken from a reference
    PPO policy's state dict (student_teacher.py:168-206: `actor.*` -> `teacher.*`).  Recorded per iteration: the env
    stream, the sampling noise, the rollout buffer, the parameters after every optimiser step, the loss, the carried
    hidden states."""
    import torch
    import rsl_rl.modules as ref_modules
    from rsl_rl.runners import DistillationRunner

    RecordingEnv = make_env_class()
    torch.manual_seed(seed)
    env = RecordingEnv(num_envs, group_dims, num_actions, seed=seed + 100, p_done=p_done)
    # the "RL checkpoint" the teacher comes from: a reference actor-critic whose policy set is the teacher's observations
    obs0 = env.get_observations()
    teacher_policy = getattr(ref_modules, teacher_class)(
        obs0, {"policy": list(teacher_groups), "critic": list(teacher_groups)}, num_actions,
        actor_hidden_dims=list(hidden), critic_hidden_dims=list(hidden), activation=activation, **teacher_extra)
    if teacher_extra.get("actor_obs_normalization"):
        g = torch.Generator().manual_seed(seed + 5)
        for _ in range(3):   # non-trivial running statistics in the checkpoint
            fake = {k: 1.5 * torch.randn(num_envs, d, generator=g) + 0.3 for k, d in group_dims.items()}
            teacher_policy.update_normalization(fake)
    teacher_sd = teacher_policy.state_dict()
    alg = dict(class_name="Distillation", num_learning_epochs=1, gradient_length=15, learning_rate=1e-3,
               max_grad_norm=None, loss_type="mse", optimizer="adam")
    alg.update(alg_overrides)
    cfg = dict(num_steps_per_env=num_steps, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in obs_groups.items()},
               policy=dict(class_name=policy_class, student_hidden_dims=list(hidden), teacher_hidden_dims=list(hidden),
                           activation=activation, init_noise_std=0.1, **policy_extra),
               algorithm=dict(alg))
    meta = dict(num_envs=num_envs, num_steps=num_steps, num_actions=num_actions, hidden=list(hidden),
                activation=activation, iterations=iterations, obs_groups={k: list(v) for k, v in obs_groups.items()},
                group_dims=dict(group_dims), alg={k: v for k, v in alg.items() if k != "class_name"},
                policy_class=policy_class, policy_extra=dict(policy_extra), teacher_class=teacher_class,
                teacher_groups=list(teacher_groups), teacher_extra=dict(teacher_extra))
    data = {}
    for k, v in teacher_sd.items():
        data[f"teacher_ckpt.{k}"] = v.detach().numpy().copy()
    with Recorder() as rec:
        runner = DistillationRunner(env, cfg, log_dir=tempfile.mkdtemp(prefix="golden_"), device="cpu")
        resumed = runner.alg.policy.load_state_dict(teacher_sd)
        assert resumed is False and runner.alg.policy.loaded_teacher
        rec.watch_optimizer(runner.alg.optimizer, runner.alg.policy)
        for k, v in _state_arrays(runner.alg.policy).items():
            data[f"init.{k}"] = v
        meta["param_names"] = [k for k, _ in runner.alg.policy.named_parameters()]
        meta["student_param_names"] = ["student." + k for k, _ in runner.alg.policy.student.named_parameters()]
        losses = []
        orig_update = runner.alg.update
        snap = {}

        def update_and_record():
            st = runner.alg.storage
            for kk, vv in st.observations.items():
                snap[f"obs.{kk}"] = vv.numpy().copy()
            for kk in ("actions", "privileged_actions", "rewards", "dones"):
                snap[kk] = getattr(st, kk).numpy().copy()
            out = orig_update()
            losses.append(out)
            return out

        runner.alg.update = update_and_record
        for it in range(iterations + int(checkpoint)):
            if it == iterations:
                # a checkpoint WRITTEN BY THE REFERENCE (on_policy_runner.py:289-305) after the last regular iteration, then
                # one more iteration of the same runner: what a run resumed from that checkpoint has to reproduce (optimiser
                # moments and step count, normaliser statistics, iteration counter)
                ckpt = os.path.join(out_dir, name + "_ref_checkpoint.pt")
                runner.save(ckpt)
                print("wrote", ckpt, f"{os.path.getsize(ckpt) / 1024:.0f} KiB")
                data["resume.iter"] = np.array(runner.current_learning_iteration)
            n_norm, n_clip, n_step = len(rec.normal_z), len(rec.clip_calls), len(rec.step_calls)
            first_obs, first_step = len(env.obs_log) - 1, len(env.rew_log)
            runner.learn(1)
            p = f"it{it}." if it < iterations else "resume."
            T = num_steps
            for g in group_dims:
                data[p + f"stream.obs.{g}"] = np.stack([env.obs_log[first_obs + t][g] for t in range(T + 1)])
            data[p + "stream.rewards"] = np.stack(env.rew_log[first_step:first_step + T])
            data[p + "stream.dones"] = np.stack(env.done_log[first_step:first_step + T])
            zs = rec.normal_z[n_norm:]
            assert len(zs) == T, len(zs)
            data[p + "z_rollout"] = np.stack(zs)
            for kk, vv in snap.items():
                data[p + "storage." + kk] = vv
            clips, steps = rec.clip_calls[n_clip:], rec.step_calls[n_step:]
            if clips:
                data[p + "mb.raw_grads"] = np.stack([c[0] for c in clips])     # student MLP gradients only
                data[p + "mb.total_norm"] = np.array([c[1] for c in clips])
            data[p + "mb.lr"] = np.array([s[0] for s in steps])
            n_params = sum(q.numel() for q in runner.alg.policy.parameters())
            data[p + "mb.params_after"] = (np.stack([s[1] for s in steps]) if steps   # (no optimiser step when the
                                           else np.zeros((0, n_params), np.float32))   #  window is longer than the update)
            data[p + "loss.behavior"] = np.array(losses[-1]["behavior"], dtype=np.float64)
            for k, v in _state_arrays(runner.alg.policy).items():
                data[p + f"final.{k}"] = v
            hs = runner.alg.last_hidden_states
            if hs is not None:
                for tag, h in zip(("s", "t"), hs):
                    if h is None:
                        continue
                    for i, x in enumerate(h if isinstance(h, tuple) else (h,)):
                        data[p + f"hidden_{tag}{i}"] = x.detach().numpy().copy()
            data[p + "num_updates"] = np.array(runner.alg.num_updates)
    data["meta"] = np.array(repr(meta))
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **data)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def scenario_loss_terms(out_dir):
    """Per-minibatch loss values + autograd gradients w.r.t. (mu, sigma, value) from the reference's
    own expressions, evaluated through a real PPO.update() with the optimiser step recorded."""
    import torch
    from tensordict import TensorDict
    from rsl_rl.algorithms import PPO
    from rsl_rl.modules import ActorCritic

    torch.manual_seed(3)
    N, T, O, A = 32, 6, 10, 5
    obs = TensorDict({"policy": torch.randn(N, O)}, batch_size=[N])
    data = {}
    for tag, clipped in (("clipped", True), ("unclipped", False)):
        policy = ActorCritic(obs, {"policy": ["policy"], "critic": ["policy"]}, A, actor_hidden_dims=[16, 8],
                             critic_hidden_dims=[16, 8], activation="elu")
        alg = PPO(policy, num_learning_epochs=1, num_mini_batches=2, use_clipped_value_loss=clipped, device="cpu")
        alg.init_storage("rl", N, T, obs, [A])
        g = torch.Generator().manual_seed(11)
        with torch.inference_mode():
            o = obs
            for t in range(T):
                alg.act(o)
                o = TensorDict({"policy": torch.randn(N, O, generator=g)}, batch_size=[N])
                dones = (torch.rand(N, generator=g) < 0.2).long()
                alg.process_env_step(o, torch.randn(N, generator=g), dones, {"time_outs": (dones > 0)})
            alg.compute_returns(o)
        # perturb the policy so that ratio != 1 and both clip branches are exercised
        with torch.no_grad():
            for p in policy.parameters():
                p.add_(0.15 * torch.randn(p.shape, generator=g))
            policy.std.clamp_(min=0.3)
        # replicate one minibatch by hand with autograd on the reference's modules
        st = alg.storage
        idx = torch.arange(N * T)
        ob = st.observations.flatten(0, 1)[idx]
        policy.act(ob)
        mu = policy.action_mean.detach().clone().requires_grad_(True)
        sigma = policy.action_std.detach().clone().requires_grad_(True)
        value = policy.evaluate(ob).detach().clone().requires_grad_(True)
        from torch.distributions import Normal

        dist = Normal(mu, sigma)
        actions = st.actions.flatten(0, 1)
        logp = dist.log_prob(actions).sum(-1)
        old_logp = st.actions_log_prob.flatten(0, 1)
        adv = st.advantages.flatten(0, 1)
        ret = st.returns.flatten(0, 1)
        tv = st.values.flatten(0, 1)
        ent = dist.entropy().sum(-1)
        # --- the reference's expressions (ppo.py:297-315), evaluated verbatim on these leaves
        ratio = torch.exp(logp - torch.squeeze(old_logp))
        surrogate = -torch.squeeze(adv) * ratio
        surrogate_clipped = -torch.squeeze(adv) * torch.clamp(ratio, 1.0 - alg.clip_param, 1.0 + alg.clip_param)
        surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
        if clipped:
            value_clipped = tv + (value - tv).clamp(-alg.clip_param, alg.clip_param)
            value_loss = torch.max((value - ret).pow(2), (value_clipped - ret).pow(2)).mean()
        else:
            value_loss = (ret - value).pow(2).mean()
        loss = surrogate_loss + alg.value_loss_coef * value_loss - alg.entropy_coef * ent.mean()
        loss.backward()
        with torch.inference_mode():
            kl = torch.sum(torch.log(sigma / st.sigma.flatten(0, 1) + 1.0e-5)
                           + (torch.square(st.sigma.flatten(0, 1)) + torch.square(st.mu.flatten(0, 1) - mu))
                           / (2.0 * torch.square(sigma)) - 0.5, axis=-1).mean()
        for k, v in dict(mu=mu, sigma=sigma, value=value, actions=actions, old_logp=old_logp, advantages=adv,
                         returns=ret, old_values=tv, old_mu=st.mu.flatten(0, 1), old_sigma=st.sigma.flatten(0, 1),
                         dmu=mu.grad, dsigma=sigma.grad, dvalue=value.grad, surrogate=surrogate_loss,
                         value_loss=value_loss, entropy=ent.mean(), loss=loss, kl=kl, logp=logp).items():
            data[f"{tag}.{k}"] = v.detach().numpy().copy()
    path = os.path.join(out_dir, "loss_terms.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def scenario_gae(out_dir):
    import torch
    from tensordict import TensorDict
    from rsl_rl.storage import RolloutStorage

    rng = np.random.default_rng(5)
    data = {}
    cases = {}
    T, N = 24, 96
    base = (rng.random((T, N)) < 0.05)
    cases["random"] = base
    cases["none"] = np.zeros((T, N), bool)
    cases["all"] = np.ones((T, N), bool)
    first = np.zeros((T, N), bool); first[0] = True
    cases["first_row"] = first
    last = np.zeros((T, N), bool); last[-1] = True
    cases["last_row"] = last
    cases["dense"] = (rng.random((T, N)) < 0.5)
    for name, dones in cases.items():
        for norm in (True, False):
            obs = TensorDict({"policy": torch.zeros(N, 3)}, batch_size=[N])
            st = RolloutStorage("rl", N, T, obs, [2], "cpu")
            rewards = rng.standard_normal((T, N, 1)).astype(np.float32)
            values = rng.standard_normal((T, N, 1)).astype(np.float32) * 2
            last_values = rng.standard_normal((N, 1)).astype(np.float32)
            st.rewards.copy_(torch.from_numpy(rewards))
            st.values.copy_(torch.from_numpy(values))
            st.dones.copy_(torch.from_numpy(dones[..., None].astype(np.uint8)))
            st.compute_returns(torch.from_numpy(last_values), 0.99, 0.95, normalize_advantage=norm)
            p = f"{name}.{'norm' if norm else 'raw'}."
            data[p + "rewards"], data[p + "values"], data[p + "dones"] = rewards, values, dones[..., None].astype(np.uint8)
            data[p + "last_values"] = last_values
            data[p + "returns"] = st.returns.numpy().copy()
            data[p + "advantages"] = st.advantages.numpy().copy()
    # one odd-shaped case (T=5, N=7, different gamma/lam)
    T, N = 5, 7
    obs = TensorDict({"policy": torch.zeros(N, 3)}, batch_size=[N])
    st = RolloutStorage("rl", N, T, obs, [2], "cpu")
    rewards = rng.standard_normal((T, N, 1)).astype(np.float32)
    values = rng.standard_normal((T, N, 1)).astype(np.float32)
    dones = (rng.random((T, N, 1)) < 0.3).astype(np.uint8)
    last_values = rng.standard_normal((N, 1)).astype(np.float32)
    st.rewards.copy_(torch.from_numpy(rewards)); st.values.copy_(torch.from_numpy(values)); st.dones.copy_(torch.from_numpy(dones))
    st.compute_returns(torch.from_numpy(last_values), 0.9, 0.8, normalize_advantage=True)
    p = "odd.norm."
    data[p + "rewards"], data[p + "values"], data[p + "dones"], data[p + "last_values"] = rewards, values, dones, last_values
    data[p + "returns"], data[p + "advantages"] = st.returns.numpy().copy(), st.advantages.numpy().copy()
    data["odd.gamma_lam"] = np.array([0.9, 0.8])
    path = os.path.join(out_dir, "gae_cases.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def scenario_trajectories(out_dir):
    import torch
    from rsl_rl.utils import split_and_pad_trajectories, unpad_trajectories

    rng = np.random.default_rng(9)
    data = {}
    for name, (T, N, D, p) in dict(small=(8, 16, 3, 0.2), none=(6, 5, 2, 0.0), all=(4, 6, 2, 1.0),
                                   c4like=(24, 64, 5, 0.02)).items():
        x = rng.standard_normal((T, N, D)).astype(np.float32)
        dones = (rng.random((T, N, 1)) < p).astype(np.uint8)
        padded, masks = split_and_pad_trajectories(torch.from_numpy(x), torch.from_numpy(dones))
        unp = unpad_trajectories(padded, masks)
        # hidden-state extraction exactly as rollout_storage.py:218-244 does it, for all envs at once
        L, H = 2, 4
        saved = rng.standard_normal((T, L, N, H)).astype(np.float32)
        d = torch.from_numpy(dones).squeeze(-1)
        lwd = torch.zeros_like(d, dtype=torch.bool)
        lwd[1:] = d[:-1]
        lwd[0] = True
        hid = torch.from_numpy(saved).permute(2, 0, 1, 3)[lwd.permute(1, 0)].transpose(1, 0).contiguous()
        data[f"{name}.x"], data[f"{name}.dones"] = x, dones
        data[f"{name}.padded"], data[f"{name}.masks"] = padded.numpy().copy(), masks.numpy().copy()
        data[f"{name}.unpadded"] = unp.numpy().copy()
        data[f"{name}.saved_hidden"], data[f"{name}.hidden_at_starts"] = saved, hid.numpy().copy()
        data[f"{name}.last_was_done"] = lwd.numpy().copy()
    path = os.path.join(out_dir, "trajectory_cases.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    _import_reference(args.ref)
    os.makedirs(args.out, exist_ok=True)
    todo = {
        "gae_cases": lambda: scenario_gae(args.out),
        "trajectory_cases": lambda: scenario_trajectories(args.out),
        "loss_terms": lambda: scenario_loss_terms(args.out),
        "ppo_mlp_small": lambda: scenario_ppo_mlp(
            args.out, "ppo_mlp_small", num_envs=64, num_steps=8, group_dims={"policy": 12},
            obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=4, hidden=[32, 16, 8],
            activation="elu", noise_std_type="scalar", alg_overrides={}, iterations=2),
        "ppo_mlp_groups": lambda: scenario_ppo_mlp(
            args.out, "ppo_mlp_groups", num_envs=48, num_steps=6, group_dims={"proprio": 9, "scan": 7, "priv": 5},
            obs_groups={"policy": ["proprio", "scan"], "critic": ["proprio", "priv"]}, num_actions=3,
            hidden=[24, 12], activation="relu", noise_std_type="log",
            alg_overrides=dict(use_clipped_value_loss=False, num_learning_epochs=3, num_mini_batches=2,
                               learning_rate=3e-4, max_grad_norm=0.5), iterations=1, seed=4),
    }
    todo["ppo_lstm_small"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_lstm_small", num_envs=32, num_steps=8, group_dims={"policy": 10},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=3, hidden=[24, 12],
        activation="elu", noise_std_type="scalar", alg_overrides=dict(num_learning_epochs=2), iterations=2, seed=7,
        p_done=0.15, policy_class="ActorCriticRecurrent",
        policy_extra=dict(rnn_type="lstm", rnn_hidden_dim=16, rnn_num_layers=1))
    todo["ppo_rnd_sym_small"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_rnd_sym_small", num_envs=48, num_steps=6, group_dims={"policy": 11},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=4, hidden=[24, 12],
        activation="elu", noise_std_type="scalar", alg_overrides=dict(num_learning_epochs=2), iterations=2, seed=8,
        rnd_cfg=dict(weight=1.0, weight_schedule=None, reward_normalization=False, state_normalization=False,
                     learning_rate=1e-3, num_outputs=1, predictor_hidden_dims=[-1], target_hidden_dims=[-1]),
        symmetry=True)
    # observation normalisers on (EmpiricalNormalization, networks/normalization.py:44-66); 4 iterations so that the CUDA
    # path is checked eager -> graph capture -> graph replay with the running statistics changing every env step
    todo["ppo_mlp_norm"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_mlp_norm", num_envs=64, num_steps=8, group_dims={"policy": 12, "priv": 6},
        obs_groups={"policy": ["policy"], "critic": ["policy", "priv"]}, num_actions=4, hidden=[32, 16, 8],
        activation="elu", noise_std_type="scalar", alg_overrides={}, iterations=4, seed=12,
        policy_extra=dict(actor_obs_normalization=True, critic_obs_normalization=True), checkpoint=True)
    # RND with state and reward normalisation (EmpiricalDiscountedVariationNormalization, normalization.py:75-130)
    todo["ppo_rnd_norm_small"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_rnd_norm_small", num_envs=48, num_steps=6, group_dims={"policy": 11},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=4, hidden=[24, 12],
        activation="elu", noise_std_type="scalar", alg_overrides=dict(num_learning_epochs=2), iterations=3, seed=18,
        rnd_cfg=dict(weight=1.0, weight_schedule=None, reward_normalization=True, state_normalization=True,
                     learning_rate=1e-3, num_outputs=1, predictor_hidden_dims=[-1], target_hidden_dims=[-1]))
    # state-dependent standard deviation (second actor head, actor_critic.py:119-128), GRU memories
    # (networks/memory.py:23) and teacher-student distillation
    todo["ppo_sdstd_log_small"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_sdstd_log_small", num_envs=48, num_steps=6, group_dims={"policy": 11},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=4, hidden=[24, 12],
        activation="elu", noise_std_type="log", alg_overrides=dict(num_learning_epochs=2), iterations=2, seed=21,
        policy_extra=dict(state_dependent_std=True))
    todo["ppo_sdstd_scalar_small"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_sdstd_scalar_small", num_envs=48, num_steps=6, group_dims={"policy": 11},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=4, hidden=[24, 12],
        activation="tanh", noise_std_type="scalar", alg_overrides=dict(num_learning_epochs=2), iterations=2, seed=22,
        policy_extra=dict(state_dependent_std=True))
    todo["ppo_gru_small"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_gru_small", num_envs=32, num_steps=8, group_dims={"policy": 10},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=3, hidden=[24, 12],
        activation="elu", noise_std_type="scalar", alg_overrides=dict(num_learning_epochs=2), iterations=2, seed=23,
        p_done=0.15, policy_class="ActorCriticRecurrent",
        policy_extra=dict(rnn_type="gru", rnn_hidden_dim=16, rnn_num_layers=1))
    todo["distill_mlp_small"] = lambda: scenario_distillation(
        args.out, "distill_mlp_small", num_envs=48, num_steps=10, group_dims={"policy": 9, "priv": 5},
        obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]}, teacher_groups=["policy", "priv"],
        num_actions=3, hidden=[24, 12], activation="elu", policy_class="StudentTeacher",
        policy_extra=dict(student_obs_normalization=True, teacher_obs_normalization=True, noise_std_type="scalar"),
        teacher_class="ActorCritic", teacher_extra=dict(actor_obs_normalization=True),
        alg_overrides=dict(num_learning_epochs=2, gradient_length=4, max_grad_norm=1.0, loss_type="mse"),
        iterations=2, seed=31, checkpoint=True)
    # a stock optimiser with weight decay (utils.py:52-75 resolve_optimizer): parameters without a gradient must not move
    todo["distill_mlp_adamw"] = lambda: scenario_distillation(
        args.out, "distill_mlp_adamw", num_envs=32, num_steps=6, group_dims={"policy": 8, "priv": 4},
        obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]}, teacher_groups=["policy", "priv"],
        num_actions=3, hidden=[16, 8], activation="tanh", policy_class="StudentTeacher",
        policy_extra=dict(noise_std_type="log"), teacher_class="ActorCritic", teacher_extra={},
        alg_overrides=dict(num_learning_epochs=1, gradient_length=2, max_grad_norm=None, loss_type="huber",
                           optimizer="adamw", learning_rate=3e-3),
        iterations=1, seed=33)
    todo["distill_lstm_small"] = lambda: scenario_distillation(
        args.out, "distill_lstm_small", num_envs=32, num_steps=8, group_dims={"policy": 10, "priv": 4},
        obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]}, teacher_groups=["policy", "priv"],
        num_actions=3, hidden=[24, 12], activation="elu", policy_class="StudentTeacherRecurrent",
        policy_extra=dict(noise_std_type="log", rnn_type="lstm", rnn_hidden_dim=16, rnn_num_layers=1,
                          teacher_recurrent=True),
        teacher_class="ActorCriticRecurrent", teacher_extra=dict(rnn_type="lstm", rnn_hidden_dim=16, rnn_num_layers=1),
        alg_overrides=dict(num_learning_epochs=1, gradient_length=3, max_grad_norm=0.5, loss_type="huber"),
        iterations=2, seed=32, p_done=0.15)
    # a shape the one-launch fused forward and the 16-bit storage chain cover (3 hidden layers, widths multiples of 64,
    # 4 actions, 16-byte observation rows): drives the tf32 / fp16 host paths above the emulated ABI on the CPU
    todo["ppo_mlp_fused_shape"] = lambda: scenario_ppo_mlp(
        args.out, "ppo_mlp_fused_shape", num_envs=48, num_steps=6, group_dims={"policy": 12},
        obs_groups={"policy": ["policy"], "critic": ["policy"]}, num_actions=4, hidden=[128, 128, 64],
        activation="elu", noise_std_type="scalar", alg_overrides=dict(num_learning_epochs=2, num_mini_batches=2),
        iterations=1, seed=41)
    # two LSTM layers, a feed-forward teacher, windows of 3 over two epochs of 8 steps: the third window crosses the epoch
    # boundary (where the reference re-installs the start-of-rollout hidden state), one trailing step
    todo["distill_lstm2_cross"] = lambda: scenario_distillation(
        args.out, "distill_lstm2_cross", num_envs=24, num_steps=8, group_dims={"policy": 6, "priv": 3},
        obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]}, teacher_groups=["policy", "priv"],
        num_actions=2, hidden=[16, 8], activation="elu", policy_class="StudentTeacherRecurrent",
        policy_extra=dict(student_obs_normalization=True, rnn_type="lstm", rnn_hidden_dim=8, rnn_num_layers=2,
                          teacher_recurrent=False),
        teacher_class="ActorCritic", teacher_extra={},
        alg_overrides=dict(num_learning_epochs=2, gradient_length=3, max_grad_norm=None, loss_type="mse"),
        iterations=2, seed=34, p_done=0.2)
    # GRU student and teacher (single-tensor hidden states), teacher from a recurrent PPO checkpoint
    todo["distill_gru_small"] = lambda: scenario_distillation(
        args.out, "distill_gru_small", num_envs=24, num_steps=7, group_dims={"policy": 6, "priv": 3},
        obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]}, teacher_groups=["policy", "priv"],
        num_actions=2, hidden=[16, 8], activation="elu", policy_class="StudentTeacherRecurrent",
        policy_extra=dict(rnn_type="gru", rnn_hidden_dim=8, rnn_num_layers=1, teacher_recurrent=True),
        teacher_class="ActorCriticRecurrent", teacher_extra=dict(rnn_type="gru", rnn_hidden_dim=8, rnn_num_layers=1),
        alg_overrides=dict(num_learning_epochs=2, gradient_length=3, max_grad_norm=1.0, loss_type="mse"),
        iterations=2, seed=35, p_done=0.2)
    for name, fn in todo.items():
        if args.only in (None, name):
            fn()


if __name__ == "__main__":
    main()
