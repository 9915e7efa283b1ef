"""Where does a learn() call spend host time in multi-GPU mode?  (torchrun, one rank per GPU)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = f"cuda:{local}"
torch.cuda.set_device(local)
out = sys.stdout
sys.stdout = open(os.devnull, "w")
env = SynthVecEnv(4096, {"policy": 48}, 12, device=dev, seed=rank)
runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=dev)
runner.learn(3)
alg = runner.alg
def t(fn, n=1):
    torch.cuda.synchronize(); torch.distributed.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n
res = dict(learn1=t(lambda: runner.learn(1), 5), learn10=t(lambda: runner.learn(10)) / 10,
           broadcast=t(alg.broadcast_parameters, 5), barrier=t(torch.distributed.barrier, 5),
           get_obs=t(lambda: env.get_observations().to(dev), 5), train_mode=t(runner.train_mode, 5))
if rank == 0:
    print({k: round(v, 3) for k, v in res.items()}, file=out)
alg.release_graphs()
torch.distributed.barrier()
os._exit(0)
