"""Per-iteration wall / device times of learn(1) for a given gemm precision (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dev = "cuda:0"
torch.manual_seed(0)
env = SynthVecEnv(bench.NUM_ENVS, {"policy": bench.OBS_DIM}, bench.NUM_ACTIONS, device=dev, seed=0)
runner = OnPolicyRunner(env, bench.train_cfg(prec), log_dir=None, device=dev)
out = open(os.devnull, "w")
for it in range(12):
    torch.cuda.synchronize()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    so = sys.stdout; sys.stdout = out
    runner.learn(1)
    sys.stdout = so
    e1.record()
    torch.cuda.synchronize()
    print(f"{prec} iter {it}: wall {1e3 * (time.perf_counter() - t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms, "
          f"launches {nat.launch_count() - l0}, graph {runner.alg._graph is not None}, lr {runner.alg.learning_rate:.3e}")
runner.profile_phases = True
sys.stdout = out
runner.learn(3)
sys.stdout = so
print(dict(runner.phase_times))
