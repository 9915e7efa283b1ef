"""Host cost of replaying the per-step act() graphs: 24 distinct graph execs in rotation vs one exec repeatedly."""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner
dev = "cuda:0"
env = SynthVecEnv(bench.NUM_ENVS, {"policy": bench.OBS_DIM}, bench.NUM_ACTIONS, device=dev, seed=0)
with contextlib.redirect_stdout(io.StringIO()):
    runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=dev)
    runner.learn(4)
plans = runner.alg._act_plans
graphs = [plans[s]["graph"] for s in range(24)]
def host_cost(seq, rounds=50):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(rounds):
        for g in seq:
            g.replay()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = rounds * len(seq)
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6
print("24 graphs in rotation: host %.1f us/replay, incl. GPU drain %.1f us" % host_cost(graphs))
print("one graph repeatedly : host %.1f us/replay, incl. GPU drain %.1f us" % host_cost([graphs[0]] * 24))
