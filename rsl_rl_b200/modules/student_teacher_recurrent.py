"""Recurrent student / teacher pair (mirror of `rsl_rl/modules/student_teacher_recurrent.py` of the reference): a memory
(`memory_s`) in front of the student MLP and, with `teacher_recurrent=True`, one (`memory_t`) in front of the teacher.
LSTM memories run on the library's kernels (networks/lstm.py), including backward through time across the
`gradient_length` steps `Distillation.update` accumulates; the hidden state of step t-1 is an autograd input of step t.
"""
from __future__ import annotations

import warnings

import torch

from ..networks import MLP, Memory
from .student_teacher import _StudentTeacherBase, count_obs


class StudentTeacherRecurrent(_StudentTeacherBase):
    is_recurrent = True

    def __init__(self, obs, obs_groups, num_actions, student_obs_normalization=False, teacher_obs_normalization=False,
                 student_hidden_dims=[256, 256, 256], teacher_hidden_dims=[256, 256, 256], activation="elu",
                 init_noise_std=0.1, noise_std_type: str = "scalar", rnn_type="lstm", rnn_hidden_dim=256,
                 rnn_num_layers=1, teacher_recurrent=False, **kwargs):
        if "rnn_hidden_size" in kwargs:
            warnings.warn("The argument `rnn_hidden_size` is deprecated and will be removed in a future version. "
                          "Please use `rnn_hidden_dim` instead.", DeprecationWarning)
            if rnn_hidden_dim == 256:   # the old name only overrides an untouched default
                rnn_hidden_dim = kwargs.pop("rnn_hidden_size")
        if kwargs:
            print("StudentTeacherRecurrent.__init__ got unexpected arguments, which will be ignored: "
                  + str(kwargs.keys()))
        super().__init__()
        self.loaded_teacher = False
        self.teacher_recurrent = teacher_recurrent
        self.obs_groups = obs_groups
        num_student_obs = count_obs(obs, obs_groups["policy"])
        num_teacher_obs = count_obs(obs, obs_groups["teacher"])

        self.memory_s = Memory(num_student_obs, type=rnn_type, num_layers=rnn_num_layers, hidden_size=rnn_hidden_dim)
        self.student = MLP(rnn_hidden_dim, num_actions, student_hidden_dims, activation)
        self.student_obs_normalization = student_obs_normalization
        self.student_obs_normalizer = self._normalizer(student_obs_normalization, num_student_obs)
        print(f"Student RNN: {self.memory_s}")
        print(f"Student MLP: {self.student}")

        if teacher_recurrent:
            self.memory_t = Memory(num_teacher_obs, type=rnn_type, num_layers=rnn_num_layers,
                                   hidden_size=rnn_hidden_dim)
            num_teacher_obs = rnn_hidden_dim   # (the reference sizes the teacher's normaliser after this, :85-91)
        self.teacher = MLP(num_teacher_obs, num_actions, teacher_hidden_dims, activation)
        self.teacher_obs_normalization = teacher_obs_normalization
        self.teacher_obs_normalizer = self._normalizer(teacher_obs_normalization, num_teacher_obs)
        if teacher_recurrent:
            print(f"Teacher RNN: {self.memory_t}")
        print(f"Teacher MLP: {self.teacher}")

        self._init_noise(num_actions, init_noise_std, noise_std_type)

    def reset(self, dones=None, hidden_states=None):
        if hidden_states is None:
            hidden_states = (None, None)
        self.memory_s.reset(dones, hidden_states[0])
        if self.teacher_recurrent:
            self.memory_t.reset(dones, hidden_states[1])

    def _student_features(self, obs):
        x = self.student_obs_normalizer(self.get_student_obs(obs))
        return self.memory_s(x).squeeze(0)

    def act(self, obs):
        self.update_distribution(self._student_features(obs))
        return self._sample()

    def act_inference(self, obs):
        return self.student(self._student_features(obs))

    def evaluate(self, obs):
        x = self.teacher_obs_normalizer(self.get_teacher_obs(obs))
        with torch.no_grad():
            if self.teacher_recurrent:
                self.memory_t.eval()
                x = self.memory_t(x).squeeze(0)
            return self.teacher(x)

    def get_hidden_states(self):
        return self.memory_s.hidden_states, (self.memory_t.hidden_states if self.teacher_recurrent else None)

    def detach_hidden_states(self, dones=None):
        self.memory_s.detach_hidden_states(dones)
        if self.teacher_recurrent:
            self.memory_t.detach_hidden_states(dones)

    def _load_teacher_extras(self, state_dict, strict):
        if self.teacher_recurrent:   # a recurrent PPO policy's actor memory becomes the teacher's
            memory = {k.replace("memory_a.", ""): v for k, v in state_dict.items() if "memory_a." in k}
            self.memory_t.load_state_dict(memory, strict=strict)
