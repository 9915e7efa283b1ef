"""CPU, world_size 2, gloo: host-side logic of the multi-GPU path.

The kernels cannot run here, so the test covers what the hosts do around them: the flat parameter
broadcast, the single in-place all-reduce of [gradients | KL] per minibatch, and that the scalars the
device-side rule consumes (sum of gradients, sum of KLs, world size) reproduce the reference's
average-then-clip-then-Adam semantics (rsl_rl/algorithms/ppo.py:272-290, 441-469) when fed to the oracle."""
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ppo_oracle as O
    from rsl_rl_b200.algorithms.flat_params import FlatParamGroup

    torch.manual_seed(100 + rank)  # different initial weights per rank, like independently constructed runners
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 3))
    group = FlatParamGroup(net.parameters(), "cpu")
    # ---- init: one flat broadcast from rank 0 (PPO.broadcast_parameters)
    dist.broadcast(group.values, src=0)
    ref = [torch.zeros_like(group.values) for _ in range(world)]
    dist.all_gather(ref, group.values)
    assert all(torch.equal(ref[0], r) for r in ref), "parameters differ across ranks after the broadcast"
    assert group.is_bound()

    # ---- per minibatch: local grads + local KL in the tail slot, ONE all-reduce
    rng = np.random.default_rng(rank)
    names = [k for k, _ in net.named_parameters()]
    local = OrderedDict((k, rng.standard_normal(tuple(p.shape)).astype(np.float32)) for k, p in net.named_parameters())
    for p, k in zip(net.parameters(), names):
        p.grad.copy_(torch.from_numpy(local[k]))
    kl_local = np.float32(0.004 + 0.03 * rank)
    group.kl_slot.fill_(float(kl_local))
    dist.all_reduce(group.grads_ext, op=dist.ReduceOp.SUM)

    # every rank now holds the same sums; emulate the device-side rule (optim.cu) on the host in fp64/fp32
    all_local = [None] * world
    dist.all_gather_object(all_local, (local, float(kl_local)))
    mean_grads = OrderedDict((k, sum(a[0][k] for a in all_local) / world) for k in names)
    for p, k in zip(net.parameters(), names):
        assert np.allclose(p.grad.numpy() / world, mean_grads[k], rtol=1e-6, atol=1e-7)
    kl_mean = float(group.kl_slot) / world
    assert abs(kl_mean - np.mean([a[1] for a in all_local])) < 1e-7
    lr = float(np.float32(O.adaptive_lr(1e-3, kl_mean, 0.01)))  # every rank applies the same rule -> no lr broadcast
    params = OrderedDict((k, p.detach().numpy().copy()) for k, p in net.named_parameters())
    O.clip_grad_norm(mean_grads, 1.0)
    O.adam_step(params, mean_grads, O.adam_init(params), lr)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([v.reshape(-1) for v in params.values()] + [[lr]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_and_identical_parameters(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b), "ranks disagree after the step"     # bit-identical, as the reference guarantees
    assert 1e-5 <= a[-1] <= 1e-2
