"""Host-side time per rollout step, split by phase (no device syncs inside the loop)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import io, contextlib
import torch
import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

dev = "cuda:0"
with contextlib.redirect_stdout(io.StringIO()):
    env = SynthVecEnv(4096, {"policy": 48}, 12, device=dev, seed=0)
    runner = OnPolicyRunner(env, bench.train_cfg("tf32"), log_dir=None, device=dev)
    runner.learn(3)
alg = runner.alg
obs = env.get_observations()
acc = dict(act=0.0, env=0.0, to=0.0, proc=0.0)
iters = 10
torch.cuda.synchronize()
t_all = time.perf_counter()
with torch.inference_mode():
    for it in range(iters):
        for _ in range(24):
            t0 = time.perf_counter(); a = alg.act(obs)
            t1 = time.perf_counter(); obs, r, d, ex = env.step(a)
            t2 = time.perf_counter(); obs, r, d = obs.to(dev), r.to(dev), d.to(dev)
            t3 = time.perf_counter(); alg.process_env_step(obs, r, d, ex)
            t4 = time.perf_counter()
            acc["act"] += t1 - t0; acc["env"] += t2 - t1; acc["to"] += t3 - t2; acc["proc"] += t4 - t3
        alg.storage.clear()
torch.cuda.synchronize()
tot = time.perf_counter() - t_all
n = iters * 24
print({k: f"{v / n * 1e6:.1f} us" for k, v in acc.items()}, f"total/step {tot / n * 1e6:.1f} us (incl. device drain)")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with torch.inference_mode():
    for _ in range(24 * 3):
        a = alg.act(obs); obs, r, d, ex = env.step(a); alg.process_env_step(obs, r, d, ex)
        if alg.storage.step == 24: alg.storage.clear()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
