"""Host-side profile of the learn() loop (cProfile): where the Python time of a rollout step goes.
usage: python profiles/experiments/host_profile.py [device|host] [iters]"""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import PinnedHostEnv, SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

mode = sys.argv[1] if len(sys.argv) > 1 else "device"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda:0"
with contextlib.redirect_stdout(io.StringIO()):
    env = (SynthVecEnv(4096, {"policy": 48}, 12, device=dev, seed=0) if mode == "device"
           else PinnedHostEnv(4096, {"policy": 48}, 12, pool=32, seed=0))
    runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=dev)
    runner.learn(5)
torch.cuda.synchronize()
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    runner.learn(iters)
torch.cuda.synchronize()
print(f"{mode}: {(time.perf_counter() - t0) / iters * 1e3:.3f} ms/iter unprofiled")
pr = cProfile.Profile()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    runner.learn(iters)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue()[:9000])
