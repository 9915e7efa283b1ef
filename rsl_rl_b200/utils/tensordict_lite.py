"""Minimal observation container used when the optional `tensordict` package is not installed.

The reference types observations as `tensordict.TensorDict` (rsl_rl/env/vec_env.py:10,48-54) but the
PPO path only relies on a small duck-typed surface: `obs[key]`, `key in obs`, `.keys()/.items()/
.values()`, `.to(device)`, `.batch_size`/`.shape`, `obs[int | slice | LongTensor | tuple]`,
`.flatten(a, b)`, `.copy_()`, `.detach()`, `.clone()`.  Everything in rsl_rl_b200 is written against that
surface, so a real TensorDict works unchanged; this class provides it otherwise.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch


def _as_int(v):
    return None if v is None else int(v)


def _plain_index(index):
    if isinstance(index, slice):
        return slice(_as_int(index.start), _as_int(index.stop), _as_int(index.step))
    if isinstance(index, tuple):
        return tuple(_plain_index(i) for i in index)
    if isinstance(index, torch.Tensor) and index.dim() == 0 and index.dtype != torch.bool:
        return int(index)
    return index


class TensorDictLite:
    """Dict of tensors sharing leading `batch_size` dimensions."""

    def __init__(self, source: Optional[Dict[str, torch.Tensor]] = None, batch_size: Optional[Iterable[int]] = None,
                 device=None):
        self._t: Dict[str, torch.Tensor] = dict(source) if source is not None else {}
        if batch_size is None:
            first = next(iter(self._t.values()), None)
            batch_size = first.shape[:1] if first is not None else []
        self.batch_size = torch.Size(batch_size)
        self.device = device if device is not None else next((v.device for v in self._t.values()), None)

    # ---- mapping
    def keys(self):
        return self._t.keys()

    def values(self):
        return self._t.values()

    def items(self):
        return self._t.items()

    def __contains__(self, key) -> bool:
        return key in self._t

    def __iter__(self):
        return iter(self._t)

    def __len__(self) -> int:
        return int(self.batch_size[0]) if len(self.batch_size) else 0

    def __setitem__(self, key: str, value: torch.Tensor):
        self._t[key] = value

    def get(self, key, default=None):
        return self._t.get(key, default)

    @property
    def shape(self) -> torch.Size:
        return self.batch_size

    # ---- indexing
    def __getitem__(self, index):
        if isinstance(index, str):
            return self._t[index]
        index = _plain_index(index)
        out = {k: v[index] for k, v in self._t.items()}
        nb = len(self.batch_size)
        if out:
            k0 = next(iter(out))
            tail = self._t[k0].dim() - nb
            new_bs = out[k0].shape[: out[k0].dim() - tail]
        else:
            new_bs = torch.empty(tuple(self.batch_size), device="meta")[index].shape
        return TensorDictLite(out, batch_size=new_bs, device=self.device)

    # ---- tensor-like helpers
    def _map(self, fn, batch_size=None, device=None):
        return TensorDictLite({k: fn(v) for k, v in self._t.items()},
                              batch_size=self.batch_size if batch_size is None else batch_size,
                              device=self.device if device is None else device)

    def to(self, device, non_blocking: bool = False):
        """non_blocking is honoured only for pinned host sources (the only case where it is safe and useful)."""
        dev = torch.device(device)
        if all(v.device == dev for v in self._t.values()):
            return self
        return self._map(lambda v: v.to(dev, non_blocking=non_blocking and (v.is_cuda or v.is_pinned())), device=device)

    def detach(self):
        return self._map(lambda v: v.detach())

    def clone(self):
        return self._map(lambda v: v.clone())

    def contiguous(self):
        return self._map(lambda v: v.contiguous())

    def flatten(self, start_dim: int, end_dim: int):
        bs = torch.empty(tuple(self.batch_size), device="meta").flatten(start_dim, end_dim).shape
        return self._map(lambda v: v.flatten(start_dim, end_dim), batch_size=bs)

    def copy_(self, other):
        for k, v in self._t.items():
            v.copy_(other[k])
        return self

    def __repr__(self):
        body = ", ".join(f"{k}: {tuple(v.shape)}" for k, v in self._t.items())
        return f"TensorDictLite({{{body}}}, batch_size={list(self.batch_size)}, device={self.device})"


try:  # pragma: no cover - depends on the environment
    from tensordict import TensorDict as TensorDict  # type: ignore
except Exception:  # tensordict is optional
    TensorDict = TensorDictLite


def is_tensordict_like(x) -> bool:
    return hasattr(x, "keys") and hasattr(x, "items") and not isinstance(x, torch.Tensor)
