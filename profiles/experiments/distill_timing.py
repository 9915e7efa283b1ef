"""Teacher-student distillation timed on one GPU: this library's DistillationRunner and the UNMODIFIED
reference's (baseline/_ref, eager PyTorch on the same GPU), same synthetic env, same config, one JSON line per arm.

Config: 4096 envs, student obs 48, teacher obs 48 + 24, MLPs [512, 256, 128], 120 steps per env, Distillation with
gradient_length 15, 2 learning epochs (the reference's Isaac Lab distillation defaults), recurrent variant = LSTM(256)
memories in front of both MLPs.  env-steps/s = T * N / seconds per learn() iteration (on_policy_runner.py:179,209's
definition).  Usage: python profiles/experiments/distill_timing.py [--iters 3] [--warmup 2] [--skip-reference]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N, T, A = 4096, 120, 12
DEV = "cuda:0"
GROUPS = {"policy": 48, "priv": 24}
OBS_GROUPS = {"policy": ["policy"], "teacher": ["policy", "priv"]}


def cfg_for(recurrent: bool) -> dict:
    policy = dict(class_name="StudentTeacherRecurrent" if recurrent else "StudentTeacher",
                  student_hidden_dims=[512, 256, 128], teacher_hidden_dims=[512, 256, 128], activation="elu",
                  init_noise_std=0.1)
    if recurrent:
        policy.update(rnn_type="lstm", rnn_hidden_dim=256, rnn_num_layers=1, teacher_recurrent=True)
    return dict(num_steps_per_env=T, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in OBS_GROUPS.items()},
                policy=policy,
                algorithm=dict(class_name="Distillation", num_learning_epochs=2, gradient_length=15, learning_rate=1e-3,
                               max_grad_norm=1.0, loss_type="mse"))


def _sync():
    if DEV.startswith("cuda"):
        torch.cuda.synchronize()


def time_runner(runner, warmup, iters):
    pol = runner.alg.policy
    pol.load_state_dict(pol.state_dict())   # "a distillation checkpoint": marks the (random) teacher as loaded
    with contextlib.redirect_stdout(io.StringIO()):
        runner.learn(warmup)
        _sync()
        t0 = time.perf_counter()
        runner.learn(iters)
        _sync()
    return (time.perf_counter() - t0) / iters


def ours(recurrent, warmup, iters):
    from rsl_rl_b200 import _native as nat
    from rsl_rl_b200.env import SynthVecEnv
    from rsl_rl_b200.runners import DistillationRunner
    torch.manual_seed(0)
    env = SynthVecEnv(N, GROUPS, A, device=DEV, seed=0, block=T)
    with contextlib.redirect_stdout(io.StringIO()):
        runner = DistillationRunner(env, cfg_for(recurrent), log_dir=None, device=DEV)
    before = nat.launch_count()
    s = time_runner(runner, warmup, iters)
    return s, (nat.launch_count() - before) // (warmup + iters)


def reference(recurrent, warmup, iters):
    from oracle import reference_arm
    reference_arm.import_reference()
    from rsl_rl.runners import DistillationRunner
    torch.manual_seed(0)
    env = reference_arm.make_env(N, GROUPS, A, DEV, seed=0, block=T)
    with contextlib.redirect_stdout(io.StringIO()):
        runner = DistillationRunner(env, cfg_for(recurrent), log_dir=None, device=DEV)
    runner.disable_logs = True
    return time_runner(runner, warmup, iters), None


def main():
    global N, T, DEV
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--envs", type=int, default=N)
    ap.add_argument("--steps", type=int, default=T)
    ap.add_argument("--device", default=DEV)   # (anything but a CUDA device only makes sense for the reference arm)
    args = ap.parse_args()
    N, T, DEV = args.envs, args.steps, args.device
    arms = [("rsl_rl_b200", ours, False), ("reference_cuda_eager", reference, False),
            ("rsl_rl_b200", ours, True), ("reference_cuda_eager", reference, True)]
    if args.skip_reference:
        arms = [a for a in arms if a[1] is ours]
    for impl, fn, recurrent in arms:
        try:
            s, launches = fn(recurrent, args.warmup, args.iters)
            line = dict(impl=impl, workload=f"distillation {'lstm256+' if recurrent else ''}mlp512x256x128 N={N} T={T} "
                        "gradient_length=15 epochs=2", ms_per_iter=round(1e3 * s, 3),
                        env_steps_per_s=round(N * T / s, 1), iters=args.iters, warmup=args.warmup)
            if launches is not None:
                line["library_launches_per_iter"] = int(launches)
        except Exception as exc:   # one arm failing must not lose the others
            line = dict(impl=impl, recurrent=recurrent, error=f"{type(exc).__name__}: {exc}")
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
