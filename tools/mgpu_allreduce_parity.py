"""Multi-rank CUDA parity of the one-launch data-parallel optimiser step (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/mgpu_allreduce_parity.py

Every rank writes its OWN gradients and minibatch KL into the symmetric gradient buffer and calls
`FusedAdam.fused_step` (the W > 1 branch of `allreduce_clip_adam_kernel`: peer-memory all-reduce + clip + KL-adaptive LR +
Adam in one launch).  Rank 0 then checks, step by step, against the oracle applied to the MEAN of the rank gradients
(reference ppo.py:441-469 reduce_parameters, :272-290 KL rule with the fp32-rounded lr broadcast, :380 clip, :381 Adam),
and that parameters, Adam moments and lr are bit-identical on all ranks."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from oracle import ppo_oracle as O
from rsl_rl_b200.algorithms.flat_params import FlatParamGroup, FusedAdam


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(dev))
    torch.manual_seed(0)   # identical initial parameters on every rank (broadcast_parameters did this in learn())
    net = torch.nn.Sequential(torch.nn.Linear(48, 512), torch.nn.ELU(), torch.nn.Linear(512, 256), torch.nn.ELU(),
                              torch.nn.Linear(256, 13)).to(dev)
    names = [k for k, _ in net.named_parameters()]
    params = OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in net.named_parameters())
    group = FlatParamGroup(net.parameters(), dev)
    opt = FusedAdam(group, lr=1e-3)
    opt.enable_peer_allreduce(rank, world)
    assert opt.peer_allreduce
    adam, lr = O.adam_init(params), 1e-3
    kls = [0.05, 0.001, 0.03, 0.011, 0.0, 0.004, 0.1, 0.0049, 0.02, 0.0051]
    worst = dict(params=0.0, norm=0.0)
    for step, kl in enumerate(kls):
        # gradients and KL of every rank are generated on every rank (cheap) so that rank 0 can form the oracle's mean
        per_rank, per_kl = [], []
        for r in range(world):
            rng = np.random.default_rng(1000 * step + r)
            per_rank.append(OrderedDict((k, (rng.standard_normal(v.shape) * (3.0 if step % 2 else 0.01) * (1 + r)).astype(np.float32))
                                        for k, v in params.items()))
            per_kl.append(np.float32(kl * (0.5 + r)))
        for p, k in zip(net.parameters(), names):
            p.grad.copy_(torch.from_numpy(per_rank[rank][k]).to(dev))
        group.kl_slot.fill_(float(per_kl[rank]))
        opt.fused_step(world_size=world, max_grad_norm=1.0, adaptive=True, desired_kl=0.01, round_lr_fp32=True)
        torch.cuda.synchronize()
        # oracle: all_reduce(SUM) / world in fp32 (ppo.py:453-455), KL mean (ppo.py:273-274), lr through an fp32 tensor
        mean = OrderedDict()
        for k in names:
            acc = per_rank[0][k].copy()
            for r in range(1, world):
                acc = acc + per_rank[r][k]
            mean[k] = (acc / np.float32(world)).astype(np.float32)
        kl_mean = np.float32(sum(per_kl, np.float32(0)) / np.float32(world))
        lr = float(np.float32(O.adaptive_lr(lr, float(kl_mean), 0.01)))
        total = O.clip_grad_norm(mean, 1.0)
        O.adam_step(params, mean, adam, lr)
        info = opt.dev_state.read()
        flat = np.concatenate([params[k].reshape(-1) for k in names])
        mine = np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in net.parameters()])
        err = float(np.abs(mine.astype(np.float64) - flat).max() / np.abs(flat).max())
        worst["params"] = max(worst["params"], err)
        worst["norm"] = max(worst["norm"], abs(info["total_norm"] - total) / total)
        assert abs(info["lr"] - lr) < 1e-18 + 1e-12 * lr, (step, info["lr"], lr)
        assert abs(info["total_norm"] - total) < 1e-5 * total, (step, info["total_norm"], total)
        assert info["step"] == step + 1
        assert err < 1e-6, (step, err)
        assert float(group.grads.abs().max()) == 0.0, "gradient buffer not cleared after the peers read it"
        # bit-identical replicas
        blob = torch.cat([group.values, group.exp_avg, group.exp_avg_sq,
                          torch.tensor([info["lr"]], dtype=torch.float64, device=dev).view(torch.float32)])
        gathered = [torch.empty_like(blob) for _ in range(world)]
        dist.all_gather(gathered, blob)
        assert all(torch.equal(gathered[0].view(torch.int32), g.view(torch.int32)) for g in gathered[1:]), \
            f"step {step}: replicas differ"
    if rank == 0:
        print(f"mgpu_allreduce_parity ok: world={world}, {len(kls)} steps, worst param rel err {worst['params']:.2e}, "
              f"worst grad-norm rel err {worst['norm']:.2e}, ranks bit-identical")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
