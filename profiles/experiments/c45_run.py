"""BASELINE configs C4 (ActorCriticRecurrent, LSTM 256) and C5 (RND intrinsic reward + symmetry augmentation with mirror
loss) at 4096 envs, 24 steps, on the seeded synthetic VecEnv: iteration time of the composite update path
(PPO._update_composite: native GEMM / storage / GAE / trajectory / optimiser kernels under autograd, cuDNN LSTM cell).
usage: python profiles/experiments/c45_run.py [c4|c5] [precision]"""
import os, sys, io, contextlib, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import torch
import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
dev, N = "cuda:0", 4096


_cache = {}


def _signed_perm(dim, device):   # built once per width: a function that uploads host data per call cannot be graph-captured
    if dim not in _cache:
        rng = np.random.default_rng(1000 + dim)
        _cache[dim] = (torch.as_tensor(rng.permutation(dim), device=device),
                       torch.as_tensor(rng.choice([-1.0, 1.0], size=dim), dtype=torch.float32, device=device))
    return _cache[dim]


def _mirror(x):
    perm, sign = _signed_perm(x.shape[-1], x.device)
    return x[..., perm] * sign


def mirror_augmentation(obs=None, actions=None, env=None):
    obs_out = None
    if obs is not None:
        n = obs.batch_size[0]
        obs_out = type(obs)({k: torch.cat([v, _mirror(v)], dim=0) for k, v in obs.items()}, batch_size=[2 * n])
    return obs_out, (None if actions is None else torch.cat([actions, _mirror(actions)], dim=0))


cfg = bench.train_cfg(prec)
if which == "c4":
    cfg["policy"].update(class_name="ActorCriticRecurrent", rnn_type="lstm", rnn_hidden_dim=256, rnn_num_layers=1)
else:
    cfg["obs_groups"]["rnd_state"] = ["policy"]
    cfg["algorithm"]["rnd_cfg"] = dict(weight=1.0, weight_schedule=None, reward_normalization=False,
                                       state_normalization=False, learning_rate=1e-3, num_outputs=1,
                                       predictor_hidden_dims=[-1], target_hidden_dims=[-1])
    cfg["algorithm"]["symmetry_cfg"] = dict(use_data_augmentation=True, use_mirror_loss=True, mirror_loss_coeff=0.5,
                                            data_augmentation_func=mirror_augmentation)
with contextlib.redirect_stdout(io.StringIO()):
    env = SynthVecEnv(N, {"policy": 48}, 12, device=dev, seed=0)
    runner = OnPolicyRunner(env, cfg, log_dir=None, device=dev)
    runner.learn(int(os.environ.get('RSLRL_C45_WARMUP', '4')))
torch.cuda.synchronize()
runner.profile_phases = True
K = int(os.environ.get('RSLRL_C45_ITERS', '5'))
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    runner.learn(K)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
p = torch.cat([q.detach().reshape(-1) for q in runner.alg.policy.parameters()])
print(f"{which.upper()} {prec} envs={N}: {dt*1e3:.1f} ms/iter, {N*24/dt/1e6:.2f} M env-steps/s, phases "
      f"{ {k: round(v, 2) for k, v in runner.phase_times.items()} }, lr {runner.alg.learning_rate:.3e}, "
      f"params finite {bool(torch.isfinite(p).all())}, fused path {runner.alg.fused}, composite graph {runner.alg._cgraph is not None}")
if os.environ.get("RSLRL_HOST_PROFILE") == "1":   # where the host time of one iteration goes
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    with contextlib.redirect_stdout(io.StringIO()):
        runner.learn(2)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
    print(s.getvalue()[:7000])
if os.environ.get("RSLRL_KERNEL_TABLE") == "1":   # device-time table of one iteration (CUPTI through torch.profiler)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        with contextlib.redirect_stdout(io.StringIO()):
            runner.learn(1)
        torch.cuda.synchronize()
    rows = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    rows.sort(key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"kernel table of one iteration: {tot / 1e3:.2f} ms of device time, {sum(e.count for e in rows)} launches")
    for e in rows[:40]:
        print(f"{100 * e.device_time_total / tot:6.2f}%  {e.device_time_total / 1e3:9.3f} ms  n={e.count:5d}  "
              f"mean {e.device_time_total / e.count:8.2f} us  {e.key[:110]}")
