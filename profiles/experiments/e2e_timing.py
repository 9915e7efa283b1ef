"""Wall time per iteration of the e2e leg (host env) and the device-resident leg, a few repetitions."""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import PinnedHostEnv, SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner
dev = "cuda:0"
for name, env in (("host-env e2e", PinnedHostEnv(bench.NUM_ENVS, {"policy": bench.OBS_DIM}, bench.NUM_ACTIONS, pool=32, seed=0)),
                  ("device env", SynthVecEnv(bench.NUM_ENVS, {"policy": bench.OBS_DIM}, bench.NUM_ACTIONS, device=dev, seed=0))):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=dev)
        runner.learn(4)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); runner.learn(20); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) / 20 * 1e3)
    print(f"{name}: {min(ts):.2f} ms/iter (runs {[round(t, 2) for t in ts]})")
