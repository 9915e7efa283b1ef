"""Recurrent distillation, step-by-step update against the windowed one (RSLRL_DISTILL_WINDOWED=1: one multi-step LSTM call
per run of steps between two optimiser steps) on one GPU: same seeds, one learn() iteration each, largest relative
parameter difference; then the windowed path timed.  Config of profiles/experiments/distill_timing.py (4096 envs, 120 steps,
LSTM(256) + MLP [512, 256, 128], gradient_length 15, 2 epochs).  One JSON line per result, flushed as it goes."""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "profiles", "experiments"))
import distill_timing as DT  # noqa: E402


def build(windowed: bool):
    from rsl_rl_b200.env import SynthVecEnv
    from rsl_rl_b200.runners import DistillationRunner
    os.environ["RSLRL_DISTILL_WINDOWED"] = "1" if windowed else "0"
    torch.manual_seed(0)
    env = SynthVecEnv(DT.N, DT.GROUPS, DT.A, device=DT.DEV, seed=0, block=DT.T)
    with contextlib.redirect_stdout(io.StringIO()):
        runner = DistillationRunner(env, DT.cfg_for(True), log_dir=None, device=DT.DEV)
    assert runner.alg._windowed_recurrent == windowed
    pol = runner.alg.policy
    pol.load_state_dict(pol.state_dict())
    return runner


def learn(runner, iters):
    DT._sync()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        runner.learn(iters)
    DT._sync()
    return (time.perf_counter() - t0) / iters


def flat(runner):
    return torch.cat([p.detach().reshape(-1) for p in runner.alg.policy.parameters()])


def main():
    if len(sys.argv) > 1:   # dry run off the GPU: distill_windowed_check.py <device> <envs> <steps>
        DT.DEV, DT.N, DT.T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    a, b = build(False), build(True)
    p0 = flat(a).clone()
    torch.manual_seed(1)
    learn(a, 1)
    torch.manual_seed(1)
    learn(b, 1)
    pa, pb = flat(a), flat(b)
    print(json.dumps(dict(check="windowed vs step-by-step after one learn() iteration, same seeds",
                          max_abs_diff=float((pa - pb).abs().max()), max_abs_param=float(pa.abs().max()),
                          max_abs_update=float((pa - p0).abs().max()))), flush=True)
    for name, r in (("windowed", b), ("step_by_step", a)):
        s = learn(r, 2)
        print(json.dumps(dict(impl="rsl_rl_b200", update=name, ms_per_iter=round(1e3 * s, 3),
                              env_steps_per_s=round(DT.N * DT.T / s, 1), iters=2, warmup=1)), flush=True)


if __name__ == "__main__":
    main()
