"""BASELINE config C3: 65536 envs/GPU, obs 235 (proprio 48 + height scan 187 as two observation groups),
act 12, 24 steps, MLP [512,256,128]: a few learn() iterations, timing and sanity."""
import os, sys, io, contextlib, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = "cuda:0"
cfg = bench.train_cfg(prec)
cfg["obs_groups"] = {"policy": ["policy", "height_scan"], "critic": ["policy", "height_scan"]}
with contextlib.redirect_stdout(io.StringIO()):
    env = SynthVecEnv(N, {"policy": 48, "height_scan": 187}, 12, device=dev, seed=0)
    runner = OnPolicyRunner(env, cfg, log_dir=None, device=dev)
    runner.learn(3)
torch.cuda.synchronize()
runner.profile_phases = True
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.cuda.profiler.start()   # ncu --profile-from-start off captures the K timed iterations only
t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    runner.learn(K)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
dt = (time.perf_counter() - t0) / K
p = torch.cat([q.detach().reshape(-1) for q in runner.alg.policy.parameters()])
print(f"C3 {prec} envs={N}: {dt*1e3:.1f} ms/iter, {N*24/dt/1e6:.2f} M env-steps/s, phases {runner.phase_times}, "
      f"lr {runner.alg.learning_rate:.3e}, params finite {bool(torch.isfinite(p).all())}, "
      f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
