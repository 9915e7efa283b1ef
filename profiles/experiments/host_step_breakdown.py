"""Host time per rollout step by call (perf_counter around each call, no device syncs, no profiler).
usage: python profiles/experiments/host_step_breakdown.py [device|host]"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import PinnedHostEnv, SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

mode = sys.argv[1] if len(sys.argv) > 1 else "device"
dev = "cuda:0"
with contextlib.redirect_stdout(io.StringIO()):
    env = (SynthVecEnv(4096, {"policy": 48}, 12, device=dev, seed=0) if mode == "device"
           else PinnedHostEnv(4096, {"policy": 48}, 12, pool=32, seed=0))
    runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=dev)
    runner.learn(5)
alg = runner.alg
pc = time.perf_counter
acc = dict(act=0.0, actions_to_env=0.0, env_step=0.0, to_device=0.0, process=0.0, returns=0.0, update_issue=0.0, total=0.0)
iters = 20
obs = env.get_observations().to(dev)
for it in range(iters):
    torch.cuda.synchronize()
    t_start = pc()
    with torch.inference_mode():
        for _ in range(24):
            t0 = pc(); actions = alg.act(obs)
            t1 = pc(); a = runner._actions_for_env(actions)
            t2 = pc(); obs, rewards, dones, extras = env.step(a)
            t3 = pc(); obs, rewards, dones = runner._to_device(obs), runner._to_device(rewards), runner._to_device(dones)
            t4 = pc(); alg.process_env_step(obs, rewards, dones, extras)
            t5 = pc()
            acc["act"] += t1 - t0; acc["actions_to_env"] += t2 - t1; acc["env_step"] += t3 - t2
            acc["to_device"] += t4 - t3; acc["process"] += t5 - t4
        t6 = pc(); alg.compute_returns(obs); acc["returns"] += pc() - t6
    t7 = pc(); alg.update(); acc["update_issue"] += pc() - t7
    torch.cuda.synchronize()
    acc["total"] += pc() - t_start
print(mode, {k: (f"{v / iters / 24 * 1e6:.1f} us/step" if k in ("act", "actions_to_env", "env_step", "to_device", "process")
                 else f"{v / iters * 1e3:.3f} ms/iter") for k, v in acc.items()})
