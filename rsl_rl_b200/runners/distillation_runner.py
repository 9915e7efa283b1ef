"""Teacher-student training loop (mirror of `rsl_rl/runners/distillation_runner.py` of the reference).

Same constructor, config schema (`policy.class_name` in {StudentTeacher, StudentTeacherRecurrent}, `algorithm.class_name =
"Distillation"`), `learn` contract (refuses to start before a teacher was loaded) and checkpoints.  The iteration loop, the
logging, `save`, `load` and `get_inference_policy` are `OnPolicyRunner`'s: the reference's second runner repeats its first
one's loop without the return computation (distillation_runner.py:56-150 against on_policy_runner.py:61-175); here the
loop exists once, with episode statistics accumulated on the device and flushed once per iteration instead of one host
synchronisation per env step (reference :105-117), and asks the algorithm whether it has returns to compute.
"""
from __future__ import annotations

from rsl_rl_b200.algorithms import Distillation
from rsl_rl_b200.env import VecEnv
from rsl_rl_b200.modules import StudentTeacher, StudentTeacherRecurrent

from .on_policy_runner import OnPolicyRunner

_CLASSES = {"Distillation": Distillation, "StudentTeacher": StudentTeacher,
            "StudentTeacherRecurrent": StudentTeacherRecurrent}


class DistillationRunner(OnPolicyRunner):
    """On-policy runner for training and evaluation of teacher-student training."""

    def __init__(self, env: VecEnv, train_cfg: dict, log_dir: "str | None" = None, device="cpu"):
        super().__init__(env, train_cfg, log_dir, device)

    def _default_obs_sets(self) -> list:
        return ["teacher"]   # reference distillation_runner.py:40

    def _check_ready(self):
        if not self.alg.policy.loaded_teacher:   # reference distillation_runner.py:59-61
            raise ValueError("Teacher model parameters not loaded. Please load a teacher model to distill.")

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
