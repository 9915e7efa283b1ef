"""In-kernel phase breakdown of `allreduce_clip_adam_kernel` at N ranks (run under torchrun with a trace build):

    python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_trace.so -DRSLRL_AR_TRACE -DRSLRL_GAE_TRACE
    RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_b200_trace.so python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_phase_trace.py

Each rank runs bench-shaped learn() iterations (C2, fp16), then reads the globaltimer stamps of the last 64 optimiser
launches: per launch the phase durations are taken as max over blocks of (stamp[k+1] - stamp[k]) and the launch duration
as last exit - first entry; rank 0 prints the medians over launches and the max over ranks."""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import bench
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.env import SynthVecEnv
from rsl_rl_b200.runners import OnPolicyRunner

PHASES = ["A wait for peers' backward", "1 reduce slice over ranks", "2 grid barrier + scalars", "3 Adam", "4 wait peers read + clear"]


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    out = sys.stdout
    sys.stdout = open(os.devnull, "w")
    torch.manual_seed(0)
    env = SynthVecEnv(4096, {"policy": 48}, 12, device=dev, seed=rank)
    runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=dev)
    runner.learn(8)
    torch.cuda.synchronize()
    lib = nat.lib()
    n_launch, n_blocks = 64, 148
    buf = (C.c_ulonglong * (n_launch * n_blocks * 8))()
    lib.rslrl_debug_read_trace_allreduce.restype = C.c_int
    assert lib.rslrl_debug_read_trace_allreduce(buf) == 0
    P4 = runner.alg._group.numel // 4
    grid = min(148, (P4 + 255) // 256)
    per_phase, total = [[] for _ in PHASES], []
    for l in range(n_launch):
        st = [[buf[(l * n_blocks + b) * 8 + s] for s in range(6)] for b in range(grid)]
        if any(x == 0 for row in st for x in row):
            continue
        for k in range(5):
            per_phase[k].append(max(row[k + 1] - row[k] for row in st) / 1e3)
        total.append((max(row[5] for row in st) - min(row[0] for row in st)) / 1e3)
    med = [statistics.median(p) for p in per_phase] + [statistics.median(total)]
    t = torch.tensor(med, dtype=torch.float64, device=dev)
    mx = t.clone()
    torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
    sys.stdout = out
    if rank == 0:
        print(f"allreduce_clip_adam_kernel phase breakdown, world={world}, grid={grid} blocks, P={runner.alg._group.numel} "
              f"(median over {len(total)} launches of max-over-blocks phase time; us; rank 0 | max over ranks)")
        for name, a, b in zip(PHASES + ["launch (first entry -> last exit)"], t.tolist(), mx.tolist()):
            print(f"  {name:36s} {a:8.2f} | {b:8.2f}")
    runner.alg.release_graphs()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
