"""Multi-GPU consistency check of the data-parallel PPO update (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
Every rank owns its own synthetic envs; after `learn()` the parameters, Adam moments and learning rate must be
bit-identical on all ranks (reference post-condition), with the update replayed from a CUDA graph
(collective inside) and eagerly."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    out = sys.stdout
    sys.stdout = open(os.devnull, "w")
    results = {}
    for mode in ("graph", "eager"):
        torch.manual_seed(100 + rank)  # different initial weights per rank: the broadcast must align them
        env = SynthVecEnv(1024, {"policy": 48}, 12, device=dev, seed=rank)
        cfg = bench.train_cfg(os.environ.get("RSLRL_PRECISION", "fp16"), use_cuda_graph=(mode == "graph"))
        runner = OnPolicyRunner(env, cfg, log_dir=None, device=dev)
        runner.learn(5)
        torch.cuda.synchronize()
        alg = runner.alg
        flat = torch.cat([alg._group.values.detach().reshape(-1), alg._group.exp_avg.reshape(-1),
                          alg._group.exp_avg_sq.reshape(-1),
                          torch.tensor([alg.learning_rate], device=dev, dtype=torch.float32)])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        torch.distributed.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered[1:])
        graphed = alg._graph is not None
        alg.release_graphs()
        results[mode] = (same, graphed, alg.learning_rate, float(flat[:alg._group.numel].abs().sum()))
    sys.stdout = out
    if rank == 0:
        for mode, (same, graphed, lr, s) in results.items():
            print(f"mgpu_check world={world} mode={mode}: ranks identical={same} graph={graphed} lr={lr:.3e} |p|={s:.6f}")
        assert all(r[0] for r in results.values()), "parameters differ across ranks"
        assert results["graph"][1], "the multi-GPU update was not replayed from a graph"
        print("mgpu_check ok")
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
