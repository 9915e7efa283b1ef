"""Teacher-student training loop (mirror of `rsl_rl/runners/distillation_runner.py` of the reference).

Same constructor, config schema (`policy.class_name` in {StudentTeacher, StudentTeacherRecurrent}, `algorithm.class_name =
"Distillation"`), `learn` contract (refuses to start before a teacher was loaded) and checkpoints; logging, `save`,
`load`, `get_inference_policy` are inherited from `OnPolicyRunner`.  Like there, episode statistics are accumulated on the
device and flushed once per iteration instead of one host synchronisation per env step (reference :105-117).
"""
from __future__ import annotations

import os
import time
from collections import deque

import torch

from rsl_rl_b200.algorithms import Distillation
from rsl_rl_b200.env import VecEnv
from rsl_rl_b200.modules import StudentTeacher, StudentTeacherRecurrent
from rsl_rl_b200.utils import store_code_state

from .on_policy_runner import OnPolicyRunner, _EpisodeStats

_CLASSES = {"Distillation": Distillation, "StudentTeacher": StudentTeacher,
            "StudentTeacherRecurrent": StudentTeacherRecurrent}


class DistillationRunner(OnPolicyRunner):
    """On-policy runner for training and evaluation of teacher-student training."""

    def __init__(self, env: VecEnv, train_cfg: dict, log_dir: "str | None" = None, device="cpu"):
        super().__init__(env, train_cfg, log_dir, device)

    def _default_obs_sets(self) -> list:
        return ["teacher"]   # reference distillation_runner.py:40

    def learn(self, num_learning_iterations: int, init_at_random_ep_len: bool = False):
        self._prepare_logging_writer()
        if not self.alg.policy.loaded_teacher:
            raise ValueError("Teacher model parameters not loaded. Please load a teacher model to distill.")
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf,
                                                             high=int(self.env.max_episode_length))
        obs = self.env.get_observations().to(self.device)
        self.train_mode()

        logging = self.log_dir is not None
        ep_infos = []
        buffers = {n: deque(maxlen=100) for n in ("reward", "length")}
        stats = _EpisodeStats(self.env.num_envs, self.num_steps_per_env, self.device, False) if logging else None

        if self.is_distributed:
            print(f"Synchronizing parameters for rank {self.gpu_global_rank}...")
            self.alg.broadcast_parameters()

        start_iter = self.current_learning_iteration
        tot_iter = start_iter + num_learning_iterations
        for it in range(start_iter, tot_iter):
            start = time.time()
            with torch.inference_mode():
                for _ in range(self.num_steps_per_env):
                    actions = self.alg.act(obs)
                    obs, rewards, dones, extras = self.env.step(self._actions_for_env(actions))
                    obs, rewards, dones = self._to_device(obs), self._to_device(rewards), self._to_device(dones)
                    self.alg.process_env_step(obs, rewards, dones, extras)
                    if logging:
                        if "episode" in extras:
                            ep_infos.append(extras["episode"])
                        elif "log" in extras:
                            ep_infos.append(extras["log"])
                        stats.step(rewards, dones, None)
                stop = time.time()
                collection_time = stop - start
                start = stop
            loss_dict = self.alg.update()
            stop = time.time()
            learn_time = stop - start
            self.current_learning_iteration = it
            if logging:
                stats.flush(buffers)
            if logging and not self.disable_logs:
                rewbuffer, lenbuffer = buffers["reward"], buffers["length"]
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
            ep_infos.clear()
            if it == start_iter and logging and not self.disable_logs:
                git_file_paths = store_code_state(self.log_dir, self.git_status_repos)
                if self.logger_type in ["wandb", "neptune"] and git_file_paths:
                    for path in git_file_paths:
                        self.writer.save_file(path)
        if logging and not self.disable_logs:
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))

    def _construct_algorithm(self, obs) -> Distillation:
        policy_cls = self._resolve(self.policy_cfg.pop("class_name"))
        policy = policy_cls(obs, self.cfg["obs_groups"], self.env.num_actions, **self.policy_cfg).to(self.device)
        alg_cls = self._resolve(self.alg_cfg.pop("class_name"))
        alg: Distillation = alg_cls(policy, device=self.device, **self.alg_cfg, multi_gpu_cfg=self.multi_gpu_cfg)
        alg.init_storage("distillation", self.env.num_envs, self.num_steps_per_env, obs, [self.env.num_actions])
        return alg

    @staticmethod
    def _resolve(name):
        if isinstance(name, type):
            return name
        if name in _CLASSES:
            return _CLASSES[name]
        raise ValueError(f"Unknown class_name '{name}'. Available to the DistillationRunner: {sorted(_CLASSES)}")
