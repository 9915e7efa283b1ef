"""Running-statistics normalisers (mirror of `rsl_rl/networks/normalization.py`; buffers keep the
reference's names `_mean, _var, _std, count` so checkpoints are interchangeable).

These are off in every benchmark configuration; they are plain device-side
tensor algebra here and are the next fusion candidate (section 8f)."""
from __future__ import annotations

import torch
from torch import nn


class EmpiricalNormalization(nn.Module):
    """(x - mean) / (std + eps) with batched running mean / variance (reference normalization.py:14-72)."""

    def __init__(self, shape, eps=1e-2, until=None):
        super().__init__()
        self.eps = eps
        self.until = until
        self.register_buffer("_mean", torch.zeros(shape).unsqueeze(0))
        self.register_buffer("_var", torch.ones(shape).unsqueeze(0))
        self.register_buffer("_std", torch.ones(shape).unsqueeze(0))
        self.register_buffer("count", torch.tensor(0, dtype=torch.long))

    @property
    def mean(self):
        return self._mean.squeeze(0).clone()

    @property
    def std(self):
        return self._std.squeeze(0).clone()

    def forward(self, x):
        return (x - self._mean) / (self._std + self.eps)

    @torch.jit.unused
    def update(self, x):
        if not self.training:
            return
        if self.until is not None and self.count >= self.until:
            return
        n = x.shape[0]
        self.count += n
        rate = n / self.count
        var_x = torch.var(x, dim=0, unbiased=False, keepdim=True)
        mean_x = torch.mean(x, dim=0, keepdim=True)
        delta = mean_x - self._mean
        self._mean += rate * delta
        self._var += rate * (var_x - self._var + delta * (mean_x - self._mean))
        torch.sqrt(self._var, out=self._std)   # in place: CUDA graphs and act() plans hold this buffer's address

    @torch.jit.unused
    def inverse(self, y):
        return y * (self._std + self.eps) + self._mean


class _DiscountedAverage:
    """avg_t = gamma * avg_{t-1} + r_t (reference normalization.py:108-130)."""

    def __init__(self, gamma):
        self.avg = None
        self.gamma = gamma

    def update(self, rew: torch.Tensor) -> torch.Tensor:
        self.avg = rew if self.avg is None else self.avg * self.gamma + rew
        return self.avg


class EmpiricalDiscountedVariationNormalization(nn.Module):
    """Divides rewards by a running std of the discounted reward sum (reference normalization.py:75-101)."""

    def __init__(self, shape, eps=1e-2, gamma=0.99, until=None):
        super().__init__()
        self.emp_norm = EmpiricalNormalization(shape, eps, until)
        self.disc_avg = _DiscountedAverage(gamma)

    def forward(self, rew):
        if self.training:
            self.emp_norm.update(self.disc_avg.update(rew))
        if self.emp_norm._std > 0:
            return rew / self.emp_norm._std
        return rew
