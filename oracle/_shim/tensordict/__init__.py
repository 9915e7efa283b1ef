"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `tensordict` package.

`tensordict` is a hard import of the reference (rsl_rl/storage/rollout_storage.py:9,
rsl_rl/utils/utils.py:14, rsl_rl/env/vec_env.py:10) but is not installed in this image and
cannot be installed (no network).  `oracle/make_golden.py` puts this directory on
`sys.path` so that the UNMODIFIED reference can be executed on CPU to produce the golden
vectors under `tests/golden/`.  It performs no arithmetic; it only mimics the container
semantics the PPO path touches.  Nothing in the product package imports it.
"""

from __future__ import annotations

import torch


class TensorDict:
    def __init__(self, source=None, batch_size=None, device=None):
        self._leaves = dict(source or {})
        self.batch_size = torch.Size(batch_size if batch_size is not None else [])
        self.device = device

    # -- mapping surface
    def keys(self):
        return self._leaves.keys()

    def values(self):
        return self._leaves.values()

    def items(self):
        return self._leaves.items()

    def __contains__(self, key):
        return key in self._leaves

    def __setitem__(self, key, value):
        self._leaves[key] = value

    def __len__(self):
        return self.batch_size[0] if len(self.batch_size) else 0

    @property
    def shape(self):
        return self.batch_size

    # -- indexing: str -> leaf, anything else -> indexed container
    def __getitem__(self, index):
        if isinstance(index, str):
            return self._leaves[index]
        leaves = {k: v[index] for k, v in self._leaves.items()}
        nb = len(self.batch_size)
        probe = torch.empty(tuple(self.batch_size), device="meta")
        if isinstance(index, torch.Tensor):
            if index.dtype == torch.bool:
                raise NotImplementedError("boolean indexing is not on the PPO path")
            index = index.to("meta")
        elif isinstance(index, tuple):
            index = tuple(
                (slice(int(i.start) if i.start is not None else None, int(i.stop) if i.stop is not None else None, i.step)
                 if isinstance(i, slice) else i)
                for i in index
            )
        elif isinstance(index, slice):
            index = slice(
                int(index.start) if index.start is not None else None,
                int(index.stop) if index.stop is not None else None,
                index.step,
            )
        new_bs = probe[index].shape
        return TensorDict(leaves, batch_size=new_bs, device=self.device)

    # -- tensor-like surface
    def copy_(self, other):
        for k, v in self._leaves.items():
            v.copy_(other[k])
        return self

    def _map(self, fn, batch_size=None):
        return TensorDict(
            {k: fn(v) for k, v in self._leaves.items()},
            batch_size=self.batch_size if batch_size is None else batch_size,
            device=self.device,
        )

    def flatten(self, start, end):
        probe = torch.empty(tuple(self.batch_size), device="meta").flatten(start, end)
        return self._map(lambda v: v.flatten(start, end), batch_size=probe.shape)

    def to(self, device):
        out = self._map(lambda v: v.to(device))
        out.device = device
        return out

    def detach(self):
        return self._map(lambda v: v.detach())

    def clone(self):
        return self._map(lambda v: v.clone())

    def __repr__(self):
        inner = ", ".join(f"{k}: {tuple(v.shape)}" for k, v in self._leaves.items())
        return f"TensorDict({{{inner}}}, batch_size={list(self.batch_size)})"
