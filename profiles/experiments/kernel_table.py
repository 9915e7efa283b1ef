"""Device-time table (CUPTI through torch.profiler) of one C2 learn() iteration in a given GEMM precision.
usage: python profiles/experiments/kernel_table.py [precision] [config]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner
from torch.profiler import profile, ProfilerActivity

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
config = sys.argv[2] if len(sys.argv) > 2 else "c2"
c = bench.CONFIGS[config]
with contextlib.redirect_stdout(io.StringIO()):
    env = SynthVecEnv(c["envs"], dict(c["groups"]), bench.NUM_ACTIONS, device="cuda:0", seed=0)
    runner = OnPolicyRunner(env, bench.train_cfg(prec, config), log_dir=None, device="cuda:0")
    runner.learn(4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with contextlib.redirect_stdout(io.StringIO()):
        runner.learn(1)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"{config} {prec}: kernel table of one iteration: {tot / 1e3:.2f} ms of device time, {sum(e.count for e in rows)} launches")
for e in rows[:16]:
    print(f"{100 * e.device_time_total / tot:6.2f}%  {e.device_time_total / 1e3:9.3f} ms  n={e.count:5d}  "
          f"mean {e.device_time_total / e.count:8.2f} us  {e.key[:110]}")
