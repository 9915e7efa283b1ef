"""On-policy training loop (mirror of `rsl_rl/runners/on_policy_runner.py` of the reference).

Same constructor, `learn / save / load / get_inference_policy / train_mode / eval_mode /
add_git_repo_to_log`, same config dictionary schema, same checkpoint keys, same log scalars.
Differences that matter for the hot path:

* no per-step host synchronisation: episode statistics (reference :113-139 does `nonzero` +
  `.cpu().tolist()` every env step) are accumulated in device buffers and flushed ONCE per iteration,
  producing the same deque contents in the same order;
* `log_dir=None` is a supported benchmarking mode (the reference crashes in `store_code_state`);
* optional per-phase CUDA-event timing (`phase_times` after `learn`).
"""
from __future__ import annotations

import os
import statistics
import time
import warnings
from collections import deque

import torch

import rsl_rl_b200
from rsl_rl_b200.algorithms import PPO
from rsl_rl_b200.env import VecEnv
from rsl_rl_b200.modules import ActorCritic, ActorCriticRecurrent, resolve_rnd_config, resolve_symmetry_config
from rsl_rl_b200.utils import resolve_obs_groups, store_code_state

_CLASSES = {"PPO": PPO, "ActorCritic": ActorCritic, "ActorCriticRecurrent": ActorCriticRecurrent}


class _EpisodeStats:
    """Device-side episode bookkeeping, flushed to the host once per iteration."""

    def __init__(self, num_envs, num_steps, device, with_rnd):
        z = lambda *s: torch.zeros(*s, dtype=torch.float, device=device)
        self.names = ["reward", "length"] + (["ereward", "ireward"] if with_rnd else [])
        self.cur = {n: z(num_envs) for n in self.names}
        self.snap = {n: z(num_steps, num_envs) for n in self.names}
        self.done = torch.zeros(num_steps, num_envs, dtype=torch.bool, device=device)
        self.t = 0

    def step(self, rewards, dones, intrinsic):
        c = self.cur
        if intrinsic is not None:
            c["ereward"] += rewards
            c["ireward"] += intrinsic
            c["reward"] += rewards + intrinsic
        else:
            c["reward"] += rewards
        c["length"] += 1
        d = dones > 0
        self.done[self.t] = d
        for n in self.names:
            self.snap[n][self.t] = c[n]
            c[n].masked_fill_(d, 0)
        self.t += 1

    def flush(self, buffers):
        """buffers: name -> deque.  Appends finished-episode values in (step, env) order."""
        t = self.t
        self.t = 0
        if t == 0:
            return
        done = self.done[:t]
        idx = done.nonzero(as_tuple=True)
        vals = torch.stack([self.snap[n][:t][idx] for n in self.names]).cpu()  # single device->host read
        for i, n in enumerate(self.names):
            buffers[n].extend(vals[i].tolist())


class OnPolicyRunner:
    """On-policy runner for training and evaluation of actor-critic methods."""

    def __init__(self, env: VecEnv, train_cfg: dict, log_dir: "str | None" = None, device="cpu"):
        self.cfg = train_cfg
        self.alg_cfg = train_cfg["algorithm"]
        self.policy_cfg = train_cfg["policy"]
        self.device = device
        self.env = env
        self._configure_multi_gpu()
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg["save_interval"]

        obs = self.env.get_observations()
        self.cfg["obs_groups"] = resolve_obs_groups(obs, self.cfg["obs_groups"], self._default_obs_sets())
        self.alg = self._construct_algorithm(obs)

        self.disable_logs = self.is_distributed and self.gpu_global_rank != 0
        self.log_dir = log_dir
        self.writer = None
        self.logger_type = self.cfg.get("logger", "tensorboard").lower()
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.git_status_repos = [rsl_rl_b200.__file__]
        self.profile_phases = False
        self.phase_times = {}
        self._needs_returns = hasattr(self.alg, "compute_returns")   # (behaviour cloning has no returns / advantages)
        self._env_device, self._action_ring, self._action_slot = None, None, 0

    def _default_obs_sets(self) -> list:
        """Observation sets (besides 'policy') the algorithm reads; missing ones fall back to same-named groups or to
        the 'policy' set (reference on_policy_runner.py:44-47)."""
        sets = ["critic"]
        if "rnd_cfg" in self.alg_cfg and self.alg_cfg["rnd_cfg"] is not None:
            sets.append("rnd_state")
        return sets

    # ------------------------------------------------------------------------------------------
    def _check_ready(self):
        """Hook: conditions a subclass wants checked before training starts (DistillationRunner: a loaded teacher)."""

    def learn(self, num_learning_iterations: int, init_at_random_ep_len: bool = False):
        self._prepare_logging_writer()
        self._check_ready()
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf,
                                                             high=int(self.env.max_episode_length))
        obs = self.env.get_observations().to(self.device)
        self.train_mode()

        logging = self.log_dir is not None
        ep_infos = []
        buffers = {n: deque(maxlen=100) for n in ("reward", "length", "ereward", "ireward")}
        stats = _EpisodeStats(self.env.num_envs, self.num_steps_per_env, self.device, bool(self.alg.rnd)) if logging else None

        if self.is_distributed:
            print(f"Synchronizing parameters for rank {self.gpu_global_rank}...")
            self.alg.broadcast_parameters()

        # Host-resident env + fused rollout + no logging: the small per-step outputs (rewards, dones, time-outs) stay in
        # the env's PINNED buffers and the store kernel reads them in place (one launch instead of three copies + one).
        zero_copy = (getattr(self.alg, "rollout_fused", False) and not logging and not getattr(self.alg, "rnd", None)
                     and os.environ.get("RSLRL_ZERO_COPY_STEP", "1") != "0")
        prof = self.profile_phases
        if prof:
            ev = lambda: torch.cuda.Event(enable_timing=True)
            acc = {"collection": 0.0, "returns": 0.0, "update": 0.0}
        start_iter = self.current_learning_iteration
        tot_iter = start_iter + num_learning_iterations
        for it in range(start_iter, tot_iter):
            start = time.time()
            if prof:
                e0, e1, e2, e3 = ev(), ev(), ev(), ev()
                e0.record()
            with torch.inference_mode():
                for _ in range(self.num_steps_per_env):
                    actions = self.alg.act(obs)
                    obs, rewards, dones, extras = self.env.step(self._actions_for_env(actions))
                    # host-resident envs: pinned buffers are copied asynchronously (stream ordered; the blocking
                    # device->host copy of the next actions is the synchronisation point the env relies on)
                    obs = self._to_device(obs)
                    if not (zero_copy and not rewards.is_cuda and rewards.is_pinned() and dones.is_pinned()):
                        rewards, dones = self._to_device(rewards), self._to_device(dones)
                    self.alg.process_env_step(obs, rewards, dones, extras)
                    if logging:
                        if "episode" in extras:
                            ep_infos.append(extras["episode"])
                        elif "log" in extras:
                            ep_infos.append(extras["log"])
                        stats.step(rewards, dones, self.alg.intrinsic_rewards if self.alg.rnd else None)
                stop = time.time()
                collection_time = stop - start
                start = stop
                if prof:
                    e1.record()
                if self._needs_returns:
                    self.alg.compute_returns(obs)
                if prof:
                    e2.record()
            loss_dict = self.alg.update()
            if prof:
                e3.record()
                e3.synchronize()
                acc["collection"] += e0.elapsed_time(e1)
                acc["returns"] += e1.elapsed_time(e2)
                acc["update"] += e2.elapsed_time(e3)
            stop = time.time()
            learn_time = stop - start
            self.current_learning_iteration = it
            if logging:
                stats.flush(buffers)
            if logging and not self.disable_logs:
                rewbuffer, lenbuffer = buffers["reward"], buffers["length"]
                erewbuffer, irewbuffer = buffers["ereward"], buffers["ireward"]
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
            ep_infos.clear()
            if it == start_iter and logging and not self.disable_logs:
                git_file_paths = store_code_state(self.log_dir, self.git_status_repos)
                if self.logger_type in ["wandb", "neptune"] and git_file_paths:
                    for path in git_file_paths:
                        self.writer.save_file(path)
        if prof and num_learning_iterations > 0:
            self.phase_times = {k: v / num_learning_iterations for k, v in acc.items()}
        if logging and not self.disable_logs:
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))

    def _actions_for_env(self, actions):
        """`actions.to(self.env.device)` of the reference (on_policy_runner.py:104).  For a host-resident env the copy
        lands in a small ring of PINNED buffers (asynchronous copy + one stream synchronisation) instead of a pageable
        staging copy: this read-back is the one blocking point of every env step."""
        if not actions.is_cuda:
            return actions.to(self.env.device)
        env_dev = self._env_device
        if env_dev is None:
            env_dev = self._env_device = torch.device(self.env.device)
        if env_dev.type != "cpu":
            return actions if env_dev == actions.device else actions.to(env_dev)
        host = getattr(self.alg, "last_host_actions", None)
        if host is not None and host.shape == actions.shape:
            # the rollout kernel has already written this step's actions into pinned host memory: wait for it, no copy
            torch.cuda.current_stream(actions.device).synchronize()
            return host
        ring = self._action_ring
        if ring is None or ring[0].shape != actions.shape or ring[0].dtype != actions.dtype:
            ring = self._action_ring = [torch.empty(actions.shape, dtype=actions.dtype, pin_memory=True) for _ in range(2)]
            self._action_slot = 0
            if (getattr(self.alg, "rollout_fused", False) and actions.dim() == 2 and actions.dtype == torch.float32
                    and os.environ.get("RSLRL_HOST_ACTIONS", "0") == "1"):
                # opt-in: measured SLOWER than the copy engine on this box (e2e 7.44 vs 6.16 ms per iteration): the head
                # epilogue owns one row per lane, so the 4-byte stores reach PCIe uncoalesced
                self.alg.enable_host_actions(torch.empty(2, *actions.shape, dtype=torch.float32, pin_memory=True))
        self._action_slot ^= 1
        host = ring[self._action_slot]
        host.copy_(actions, non_blocking=True)
        torch.cuda.current_stream(actions.device).synchronize()
        return host

    def _to_device(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(self.device, non_blocking=(not x.is_cuda) and x.is_pinned())
        try:
            return x.to(self.device, non_blocking=True)
        except TypeError:
            return x.to(self.device)

    # ------------------------------------------------------------------------------------------
    def log(self, locs: dict, width: int = 80, pad: int = 35):
        """Same scalar groups and console table as the reference (on_policy_runner.py:177-287)."""
        collection_size = self.num_steps_per_env * self.env.num_envs * self.gpu_world_size
        iteration_time = locs["collection_time"] + locs["learn_time"]
        self.tot_timesteps += collection_size
        self.tot_time += iteration_time
        it = locs["it"]
        w = self.writer

        ep_string = ""
        if locs["ep_infos"]:
            for key in locs["ep_infos"][0]:
                parts = []
                for info in locs["ep_infos"]:
                    if key not in info:
                        continue
                    v = info[key]
                    if not isinstance(v, torch.Tensor):
                        v = torch.Tensor([v])
                    if len(v.shape) == 0:
                        v = v.unsqueeze(0)
                    parts.append(v.to(self.device))
                value = torch.mean(torch.cat(parts)) if parts else torch.tensor(float("nan"))
                if "/" in key:
                    w.add_scalar(key, value, it)
                    ep_string += f"""{f'{key}:':>{pad}} {value:.4f}\n"""
                else:
                    w.add_scalar("Episode/" + key, value, it)
                    ep_string += f"""{f'Mean episode {key}:':>{pad}} {value:.4f}\n"""

        mean_std = float(self.alg.policy.action_std.mean()) if self.alg.policy.action_std is not None else float("nan")
        fps = int(collection_size / iteration_time)
        for key, value in locs["loss_dict"].items():
            w.add_scalar(f"Loss/{key}", value, it)
        w.add_scalar("Loss/learning_rate", self.alg.learning_rate, it)
        w.add_scalar("Policy/mean_noise_std", mean_std, it)
        w.add_scalar("Perf/total_fps", fps, it)
        w.add_scalar("Perf/collection time", locs["collection_time"], it)
        w.add_scalar("Perf/learning_time", locs["learn_time"], it)

        have_eps = len(locs["rewbuffer"]) > 0
        rnd = hasattr(self.alg, "rnd") and self.alg.rnd
        if have_eps:
            if rnd:
                w.add_scalar("Rnd/mean_extrinsic_reward", statistics.mean(locs["erewbuffer"]), it)
                w.add_scalar("Rnd/mean_intrinsic_reward", statistics.mean(locs["irewbuffer"]), it)
                w.add_scalar("Rnd/weight", self.alg.rnd.weight, it)
            w.add_scalar("Train/mean_reward", statistics.mean(locs["rewbuffer"]), it)
            w.add_scalar("Train/mean_episode_length", statistics.mean(locs["lenbuffer"]), it)
            if self.logger_type != "wandb":
                w.add_scalar("Train/mean_reward/time", statistics.mean(locs["rewbuffer"]), self.tot_time)
                w.add_scalar("Train/mean_episode_length/time", statistics.mean(locs["lenbuffer"]), self.tot_time)

        title = f" \033[1m Learning iteration {it}/{locs['tot_iter']} \033[0m "
        out = (f"""{'#' * width}\n{title.center(width, ' ')}\n\n"""
               f"""{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, """
               f"""learning {locs['learn_time']:.3f}s)\n"""
               f"""{'Mean action noise std:':>{pad}} {mean_std:.2f}\n""")
        for key, value in locs["loss_dict"].items():
            label = f"Mean {key} loss:" if have_eps else f"{key}:"
            out += f"""{label:>{pad}} {value:.4f}\n"""
        if have_eps:
            if rnd:
                out += (f"""{'Mean extrinsic reward:':>{pad}} {statistics.mean(locs['erewbuffer']):.2f}\n"""
                        f"""{'Mean intrinsic reward:':>{pad}} {statistics.mean(locs['irewbuffer']):.2f}\n""")
            out += f"""{'Mean reward:':>{pad}} {statistics.mean(locs['rewbuffer']):.2f}\n"""
            out += f"""{'Mean episode length:':>{pad}} {statistics.mean(locs['lenbuffer']):.2f}\n"""
        out += ep_string
        done_its = it - locs["start_iter"] + 1
        remaining = locs["start_iter"] + locs["num_learning_iterations"] - it
        eta = self.tot_time / done_its * remaining
        out += (f"""{'-' * width}\n"""
                f"""{'Total timesteps:':>{pad}} {self.tot_timesteps}\n"""
                f"""{'Iteration time:':>{pad}} {iteration_time:.2f}s\n"""
                f"""{'Time elapsed:':>{pad}} {time.strftime("%H:%M:%S", time.gmtime(self.tot_time))}\n"""
                f"""{'ETA:':>{pad}} {time.strftime("%H:%M:%S", time.gmtime(eta))}\n""")
        print(out)

    # ------------------------------------------------------------------------------------------
    def save(self, path: str, infos=None):
        """Checkpoint with the reference's keys (on_policy_runner.py:289-305)."""
        saved = {
            "model_state_dict": self.alg.policy.state_dict(),
            "optimizer_state_dict": self.alg.optimizer.state_dict(),
            "iter": self.current_learning_iteration,
            "infos": infos,
        }
        if hasattr(self.alg, "rnd") and self.alg.rnd:
            saved["rnd_state_dict"] = self.alg.rnd.state_dict()
            saved["rnd_optimizer_state_dict"] = self.alg.rnd_optimizer.state_dict()
        torch.save(saved, path)
        if self.logger_type in ["neptune", "wandb"] and not self.disable_logs and self.writer is not None:
            self.writer.save_model(path, self.current_learning_iteration)

    def load(self, path: str, load_optimizer: bool = True, map_location: "str | None" = None):
        loaded = torch.load(path, weights_only=False, map_location=map_location)
        resumed = self.alg.policy.load_state_dict(loaded["model_state_dict"])
        if hasattr(self.alg, "rnd") and self.alg.rnd:
            self.alg.rnd.load_state_dict(loaded["rnd_state_dict"])
        if load_optimizer and resumed:
            self.alg.optimizer.load_state_dict(loaded["optimizer_state_dict"])
            self.alg.learning_rate = self.alg.optimizer.param_groups[0]["lr"]
            if hasattr(self.alg, "rnd") and self.alg.rnd:
                self.alg.rnd_optimizer.load_state_dict(loaded["rnd_optimizer_state_dict"])
        if resumed:
            self.current_learning_iteration = loaded["iter"]
        return loaded["infos"]

    def get_inference_policy(self, device=None):
        self.eval_mode()
        if device is not None:
            self.alg.policy.to(device)
        return self.alg.policy.act_inference

    def train_mode(self):
        self.alg.policy.train()
        if hasattr(self.alg, "rnd") and self.alg.rnd:
            self.alg.rnd.train()

    def eval_mode(self):
        self.alg.policy.eval()
        if hasattr(self.alg, "rnd") and self.alg.rnd:
            self.alg.rnd.eval()

    def add_git_repo_to_log(self, repo_file_path):
        self.git_status_repos.append(repo_file_path)

    # ------------------------------------------------------------------------------------------
    def _configure_multi_gpu(self):
        """torchrun-style env vars -> process group (reference on_policy_runner.py:353-395)."""
        self.gpu_world_size = int(os.getenv("WORLD_SIZE", "1"))
        self.is_distributed = self.gpu_world_size > 1
        if not self.is_distributed:
            self.gpu_local_rank = 0
            self.gpu_global_rank = 0
            self.multi_gpu_cfg = None
            return
        self.gpu_local_rank = int(os.getenv("LOCAL_RANK", "0"))
        self.gpu_global_rank = int(os.getenv("RANK", "0"))
        self.multi_gpu_cfg = {"global_rank": self.gpu_global_rank, "local_rank": self.gpu_local_rank,
                              "world_size": self.gpu_world_size}
        if self.device != f"cuda:{self.gpu_local_rank}":
            raise ValueError(
                f"Device '{self.device}' does not match expected device for local rank '{self.gpu_local_rank}'.")
        if self.gpu_local_rank >= self.gpu_world_size:
            raise ValueError(
                f"Local rank '{self.gpu_local_rank}' is greater than or equal to world size '{self.gpu_world_size}'.")
        if self.gpu_global_rank >= self.gpu_world_size:
            raise ValueError(
                f"Global rank '{self.gpu_global_rank}' is greater than or equal to world size '{self.gpu_world_size}'.")
        torch.cuda.set_device(self.gpu_local_rank)
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group(backend="nccl", rank=self.gpu_global_rank,
                                                 world_size=self.gpu_world_size)

    def _construct_algorithm(self, obs) -> PPO:
        self.alg_cfg = resolve_rnd_config(self.alg_cfg, obs, self.cfg["obs_groups"], self.env)
        self.alg_cfg = resolve_symmetry_config(self.alg_cfg, self.env)
        if self.cfg.get("empirical_normalization") is not None:
            warnings.warn("The `empirical_normalization` parameter is deprecated. Please set `actor_obs_normalization` "
                          "and `critic_obs_normalization` as part of the `policy` configuration instead.",
                          DeprecationWarning)
            for key in ("actor_obs_normalization", "critic_obs_normalization"):
                if self.policy_cfg.get(key) is None:
                    self.policy_cfg[key] = self.cfg["empirical_normalization"]
        policy_cls = self._resolve_class(self.policy_cfg.pop("class_name"))
        policy = policy_cls(obs, self.cfg["obs_groups"], self.env.num_actions, **self.policy_cfg).to(self.device)
        alg_cls = self._resolve_class(self.alg_cfg.pop("class_name"))
        alg: PPO = alg_cls(policy, device=self.device, **self.alg_cfg, multi_gpu_cfg=self.multi_gpu_cfg)
        alg.init_storage("rl", self.env.num_envs, self.num_steps_per_env, obs, [self.env.num_actions])
        return alg

    @staticmethod
    def _resolve_class(name):
        if isinstance(name, type):
            return name
        if name in _CLASSES:
            return _CLASSES[name]
        raise ValueError(f"Unknown class_name '{name}'. Available in rsl_rl_b200: {sorted(_CLASSES)}")

    def _prepare_logging_writer(self):
        if self.log_dir is None or self.writer is not None or self.disable_logs:
            return
        if self.logger_type == "neptune":
            raise NotImplementedError("The Neptune writer is outside the PPO hot path; use logger='tensorboard'.")
        if self.logger_type == "wandb":
            raise NotImplementedError("The W&B writer is outside the PPO hot path; use logger='tensorboard'.")
        if self.logger_type == "tensorboard":
            from torch.utils.tensorboard import SummaryWriter

            self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
        else:
            raise ValueError("Logger type not found. Please choose 'neptune', 'wandb' or 'tensorboard'.")
