"""cProfile of OnPolicyRunner.learn with the env on the host (the e2e leg of bench.py): where the per-step host time goes."""
import os, sys, cProfile, pstats, io, contextlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import PinnedHostEnv
from rsl_rl_b200.runners import OnPolicyRunner
dev = "cuda:0"
torch.manual_seed(0)
henv = PinnedHostEnv(bench.NUM_ENVS, {"policy": bench.OBS_DIM}, bench.NUM_ACTIONS, pool=32, seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    runner = OnPolicyRunner(henv, bench.train_cfg(sys.argv[1] if len(sys.argv) > 1 else "fp16"), log_dir=None, device=dev)
    runner.learn(4)
torch.cuda.synchronize()
pr = cProfile.Profile()
with contextlib.redirect_stdout(io.StringIO()):
    pr.enable(); runner.learn(20); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
