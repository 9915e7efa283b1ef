"""Student / teacher pair for behaviour cloning (mirror of `rsl_rl/modules/student_teacher.py` of the reference).

Same constructor, sub-module names (`student`, `teacher`, `student_obs_normalizer`, `teacher_obs_normalizer`), noise
parameter (`std` / `log_std`) and methods, therefore the same `state_dict`; `load_state_dict` accepts a PPO checkpoint
(`actor.*` becomes the teacher, reference student_teacher.py:168-206) or a distillation checkpoint.  Both MLPs run on the
library's GEMM kernels (networks/mlp.py): the teacher without autograd, the student through `_MLPFunction` (hand-written
backward) when `Distillation.update` differentiates it.  `_StudentTeacherBase` holds what the feed-forward and the
recurrent variant (student_teacher_recurrent.py) share.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.distributions import Normal

from ..networks import MLP, EmpiricalNormalization


def count_obs(obs, groups) -> int:
    """Width of the concatenation of the 1-D observation groups `groups`."""
    n = 0
    for g in groups:
        assert len(obs[g].shape) == 2, "The StudentTeacher module only supports 1D observations."
        n += obs[g].shape[-1]
    return n


class _StudentTeacherBase(nn.Module):
    """Noise parameter, Gaussian action distribution, observation sets and checkpoint rules of both variants."""

    def _init_noise(self, num_actions: int, init_noise_std: float, noise_std_type: str):
        self.noise_std_type = noise_std_type
        if noise_std_type == "scalar":
            self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        elif noise_std_type == "log":
            self.log_std = nn.Parameter(torch.log(init_noise_std * torch.ones(num_actions)))
        else:
            raise ValueError(f"Unknown standard deviation type: {noise_std_type}. Should be 'scalar' or 'log'")
        self.distribution = None
        self._noise_fn = None   # test hook: callable(like) -> standard-normal tensor (replays recorded draws)
        Normal.set_default_validate_args(False)   # reference student_teacher.py:89 (global side effect kept)

    @staticmethod
    def _normalizer(enabled: bool, width: int) -> nn.Module:
        return EmpiricalNormalization(width) if enabled else torch.nn.Identity()

    # -- distribution surface (reference student_teacher.py:97-121)
    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        if self.distribution is None:   # before the first act(): the parameter itself (the runner's log reads it)
            return self.std.detach() if self.noise_std_type == "scalar" else torch.exp(self.log_std.detach())
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, obs):
        self.set_distribution(self.student(obs))

    def set_distribution(self, mean):
        """The action distribution around a student output `mean` computed elsewhere (Distillation's grouped rollout)."""
        if self.noise_std_type == "scalar":
            std = self.std.expand_as(mean)
        elif self.noise_std_type == "log":
            std = torch.exp(self.log_std).expand_as(mean)
        else:
            raise ValueError(f"Unknown standard deviation type: {self.noise_std_type}. Should be 'scalar' or 'log'")
        self.distribution = Normal(mean, std)

    def _sample(self):
        """`distribution.sample()`: normal_(0,1) on a tensor of the mean's shape, scaled and shifted with two roundings
        (what torch.normal(mean, std) does), so the random stream and the bits match the reference's."""
        d = self.distribution
        with torch.no_grad():
            z = self._noise_fn(d.mean) if self._noise_fn is not None else torch.empty_like(d.mean).normal_()
            return z.mul_(d.stddev).add_(d.mean)

    # -- observation sets
    def _cat(self, obs, groups):
        if len(groups) == 1:
            return obs[groups[0]]
        return torch.cat([obs[g] for g in groups], dim=-1)

    def get_student_obs(self, obs):
        return self._cat(obs, self.obs_groups["policy"])

    def get_teacher_obs(self, obs):
        return self._cat(obs, self.obs_groups["teacher"])

    def forward(self):
        raise NotImplementedError

    def train(self, mode=True):
        super().train(mode)
        self.teacher.eval()   # the teacher never trains (reference student_teacher.py:152-156)
        self.teacher_obs_normalizer.eval()
        return self

    def update_normalization(self, obs):
        if self.student_obs_normalization:
            self.student_obs_normalizer.update(self.get_student_obs(obs))

    def set_gemm_precision(self, precision: str):
        from ..networks.memory import Memory
        for m in self.modules():
            if isinstance(m, (MLP, Memory)):
                # (an MLP / LSTM called as a module has no fp16-storage variant: tensor cores there mean tf32)
                m.precision = "tf32" if precision == "fp16" else precision

    # -- checkpoints
    def _load_teacher_extras(self, state_dict, strict):
        """Hook of the recurrent variant: the teacher's memory from `memory_a.*`."""

    def load_state_dict(self, state_dict, strict=True):
        """Loads either a PPO checkpoint (keys with `actor`: only the teacher is filled, returns False = training does
        not resume) or a distillation checkpoint (keys with `student`: everything, returns True).  The runner's `load`
        uses the flag to decide whether optimiser state and iteration counter are restored."""
        keys = list(state_dict.keys())
        if any("actor" in k for k in keys):
            teacher = {k.replace("actor.", ""): v for k, v in state_dict.items() if "actor." in k}
            normalizer = {k.replace("actor_obs_normalizer.", ""): v for k, v in state_dict.items()
                          if "actor_obs_normalizer." in k}
            self.teacher.load_state_dict(teacher, strict=strict)
            self.teacher_obs_normalizer.load_state_dict(normalizer, strict=strict)
            self._load_teacher_extras(state_dict, strict)
            resumed = False
        elif any("student" in k for k in keys):
            super().load_state_dict(state_dict, strict=strict)
            resumed = True
        else:
            raise ValueError("state_dict does not contain student or teacher parameters")
        self.loaded_teacher = True
        self.teacher.eval()
        self.teacher_obs_normalizer.eval()
        return resumed


class StudentTeacher(_StudentTeacherBase):
    is_recurrent = False

    def __init__(self, obs, obs_groups, num_actions, student_obs_normalization=False, teacher_obs_normalization=False,
                 student_hidden_dims=[256, 256, 256], teacher_hidden_dims=[256, 256, 256], activation="elu",
                 init_noise_std=0.1, noise_std_type: str = "scalar", **kwargs):
        if kwargs:
            print("StudentTeacher.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        super().__init__()
        self.loaded_teacher = False
        self.obs_groups = obs_groups
        num_student_obs = count_obs(obs, obs_groups["policy"])
        num_teacher_obs = count_obs(obs, obs_groups["teacher"])

        self.student = MLP(num_student_obs, num_actions, student_hidden_dims, activation)
        self.student_obs_normalization = student_obs_normalization
        self.student_obs_normalizer = self._normalizer(student_obs_normalization, num_student_obs)
        print(f"Student MLP: {self.student}")

        self.teacher = MLP(num_teacher_obs, num_actions, teacher_hidden_dims, activation)
        self.teacher.eval()
        self.teacher_obs_normalization = teacher_obs_normalization
        self.teacher_obs_normalizer = self._normalizer(teacher_obs_normalization, num_teacher_obs)
        print(f"Teacher MLP: {self.teacher}")

        self._init_noise(num_actions, init_noise_std, noise_std_type)

    def reset(self, dones=None, hidden_states=None):
        pass

    def act(self, obs):
        self.update_distribution(self.student_obs_normalizer(self.get_student_obs(obs)))
        return self._sample()

    def act_inference(self, obs):
        return self.student(self.student_obs_normalizer(self.get_student_obs(obs)))

    def evaluate(self, obs):
        x = self.teacher_obs_normalizer(self.get_teacher_obs(obs))
        with torch.no_grad():
            return self.teacher(x)

    def get_hidden_states(self):
        return None

    def detach_hidden_states(self, dones=None):
        pass
