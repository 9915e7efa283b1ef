"""Running-statistics normalisers (mirror of `rsl_rl/networks/normalization.py`; buffers keep the
reference's names `_mean, _var, _std, count` so checkpoints are interchangeable).

On CUDA tensors `update()` is ONE launch (`rslrl_empirical_norm_update`: batch statistics + Chan update, the `until`
test on the device, no host synchronisation -- the reference's `if self.count >= self.until` is a device->host sync per
env step) and `forward()` is one launch (`rslrl_empirical_normalize`, bit-exact `(x - mean) / (std + eps)`).  All four
buffers are updated IN PLACE: CUDA graphs and the rollout launch plans hold their addresses.  `forward()` keeps the plain
tensor expression for tensors that need autograd, for tracing / export and for host tensors (inference surface);
`update()` is a training-path operation and exists on CUDA only."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import _native as nat


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
        self._ws = None

    @property
    def mean(self):
        return self._mean.squeeze(0).clone()

    @property
    def std(self):
        return self._std.squeeze(0).clone()

    def _rows(self, x):
        """x viewed as [N, D] fp32 rows (D = number of statistics) -> (tensor, N, D, row stride) or None."""
        D = self._mean.numel()
        if not (nat.on_device(x) and x.dtype == torch.float32 and nat.on_device(self._mean)):
            return None
        if x.dim() == 0 or (D > 1 and x.shape[-1] != D):
            return None
        if x.dim() == 2 and D > 1 and x.stride(1) == 1:
            return x, x.shape[0], D, x.stride(0)
        x = x.contiguous()
        return x, x.numel() // D, D, D

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`out` (optional, [N, >= D] fp32 with unit column stride): write the result there (any row pitch)."""
        if torch.jit.is_scripting():   # export surface: the reference's expression (normalization.py:44-47)
            return (x - self._mean) / (self._std + self.eps)
        else:
            return self._forward_kernels(x, out)

    @torch.jit.unused
    def _forward_kernels(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        rows = None if (x.requires_grad and torch.is_grad_enabled()) or torch.jit.is_tracing() else self._rows(x)
        if rows is None:
            y = (x - self._mean) / (self._std + self.eps)
            return y if out is None else out.copy_(y)
        xr, N, D, ldx = rows
        if out is None:
            y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            ldy = D
        else:
            y, ldy = out, out.stride(0)
        nat.check(nat.lib().rslrl_empirical_normalize(xr.data_ptr(), ldx, N, D, self._mean.data_ptr(), self._std.data_ptr(),
                                                      float(self.eps), y.data_ptr(), ldy, nat.stream_ptr()))
        return y

    @torch.jit.unused
    def update(self, x):
        if not self.training:
            return
        rows = self._rows(x.detach())
        if rows is None:
            raise RuntimeError("rsl_rl_b200.EmpiricalNormalization.update runs on CUDA fp32 tensors only (got "
                               f"{x.dtype} on {x.device}); there is no CPU path")
        xr, N, D, ldx = rows
        if self._ws is None or self._ws.device != x.device:
            self._ws = torch.zeros(int(nat.lib().rslrl_empirical_norm_workspace_bytes(D)), dtype=torch.uint8,
                                   device=x.device)
        until = -1 if self.until is None else int(self.until)
        nat.check(nat.lib().rslrl_empirical_norm_update(xr.data_ptr(), ldx, N, D, self._mean.data_ptr(),
                                                        self._var.data_ptr(), self._std.data_ptr(), self.count.data_ptr(),
                                                        until, self._ws.data_ptr(), nat.stream_ptr()))

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
        # (the reference branches on `_std > 0` on the host, normalization.py:98-101: a device->host sync per env step;
        # the same selection is made elementwise on the device here)
        std = self.emp_norm._std
        return torch.where(std > 0, rew / std, rew)
