"""TEST / BENCH INFRASTRUCTURE -- times the UNMODIFIED reference (`baseline/_ref/rsl_rl`, installed with
`pip install --no-deps --target baseline/_ref`) on the benchmark workload.

Only `bench.py`'s reference arm / `reference_cuda` leg and `tests/` may import this module.  Nothing inside the
reference is patched: the harness builds the reference's own `OnPolicyRunner` on a seeded synthetic `VecEnv`
(obs ~ N(0,1), rewards ~ N(0,1), dones ~ Bernoulli(0.02) int64, half of them time-outs),
calls its public `learn()` and reads the wall clock around it.  `oracle/_shim/tensordict` stands in for the
`tensordict` package (not installed in this image, no arithmetic inside).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> str | None:
    """None when the reference install is present, else the reason it is not."""
    if not os.path.exists(os.path.join(REF_DIR, "rsl_rl", "__init__.py")):
        return f"{REF_DIR}/rsl_rl is missing (pip install --no-deps --target baseline/_ref /root/reference)"
    return None


def import_reference():
    why = available()
    if why:
        raise RuntimeError(why)
    for p in (os.path.join(HERE, "_shim"), REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import rsl_rl  # noqa: F401

    assert os.path.realpath(rsl_rl.__file__).startswith(os.path.realpath(REF_DIR)), rsl_rl.__file__
    return rsl_rl


def make_env(num_envs, group_dims, num_actions, device, seed=0, p_done=0.02, p_timeout=0.5, block=24):
    """Seeded synthetic env on `device` as a subclass of the REFERENCE's VecEnv (same stream layout as
    rsl_rl_b200.env.SynthVecEnv: the random numbers of `block` steps are drawn at once)."""
    import torch
    from rsl_rl.env import VecEnv
    from tensordict import TensorDict

    class RefSynthEnv(VecEnv):
        def __init__(self):
            self.num_envs, self.num_actions = num_envs, num_actions
            self.group_dims = dict(group_dims)
            self.device = device
            self.max_episode_length = 1000
            self.episode_length_buf = torch.zeros(num_envs, dtype=torch.long, device=device)
            self.cfg = {}
            self.step_dt = 0.02
            self.unwrapped = self
            self.gen = torch.Generator(device=device)
            self.gen.manual_seed(seed)
            self._refill()
            self._obs = self._next_obs()

        def _refill(self):
            n, g = self.num_envs, self.gen
            self._obs_blk = {k: torch.randn(block + 1, n, d, generator=g, device=device)
                             for k, d in self.group_dims.items()}
            self._rew_blk = torch.randn(block, n, generator=g, device=device)
            done = torch.rand(block, n, generator=g, device=device) < p_done
            self._done_blk = done.long()
            self._to_blk = (torch.rand(block, n, generator=g, device=device) < p_timeout) & done
            self._i = 0
            self._j = 0

        def _next_obs(self):
            obs = TensorDict({k: v[self._j] for k, v in self._obs_blk.items()}, batch_size=[self.num_envs],
                             device=device)
            self._j += 1
            return obs

        def get_observations(self):
            return self._obs

        def step(self, actions):
            if self._i >= block:
                self._refill()
            i = self._i
            self._i += 1
            self._obs = self._next_obs()
            return self._obs, self._rew_blk[i], self._done_blk[i], {"time_outs": self._to_blk[i]}

    return RefSynthEnv()


def reference_cfg(train_cfg: dict) -> dict:
    """bench.py's train_cfg without the rsl_rl_b200-only keyword arguments."""
    import copy

    cfg = copy.deepcopy(train_cfg)
    for k in ("gemm_precision", "use_cuda_graph", "rng_compat"):
        cfg["algorithm"].pop(k, None)
    return cfg


class ReferenceRun:
    """The reference's OnPolicyRunner on the synthetic env; `iterate(k)` -> seconds for k learn() iterations."""

    def __init__(self, train_cfg, num_envs, group_dims, num_actions, device="cpu", seed=0):
        import torch

        import_reference()
        from rsl_rl.runners import OnPolicyRunner

        self.torch, self.device = torch, device
        torch.manual_seed(0)
        self.env = make_env(num_envs, group_dims, num_actions, device, seed=seed)
        with contextlib.redirect_stdout(io.StringIO()):
            self.runner = OnPolicyRunner(self.env, reference_cfg(train_cfg), log_dir=None, device=device)
        # No logging on either arm: log_dir=None (as bench.py's own arm) and the reference's own `disable_logs` switch
        # (what it sets on non-zero ranks), without which learn() insists on a log directory to store the git state.
        self.runner.disable_logs = True

    def iterate(self, iters: int) -> float:
        torch = self.torch
        if str(self.device).startswith("cuda"):
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            self.runner.learn(iters)
        if str(self.device).startswith("cuda"):
            torch.cuda.synchronize()
        return time.perf_counter() - t0

