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


# ---------------------------------------------------------------------------------------------------------------------
# Distillation on two ranks (reference algorithms/distillation.py:156-185: broadcast_parameters / reduce_parameters)
# ---------------------------------------------------------------------------------------------------------------------

class _Setattr:
    """pytest's monkeypatch is not available inside a spawned worker; the process ends with the test."""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _distill_streams(N, T, A, dims):
    rng = np.random.default_rng(7)
    f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    obs = {k: f(T + 1, N, d) for k, d in dims.items()}
    return obs, f(T, N), torch.from_numpy((rng.random((T, N)) < 0.1).astype(np.int64)), f(T, N, A)


def _distill_run(rank, world, n_per_rank, seed):
    """One rollout + update of a feed-forward student on the env columns of `rank` (world 1: all columns)."""
    from rsl_rl_b200.algorithms import Distillation
    from rsl_rl_b200.modules import StudentTeacher
    from rsl_rl_b200.utils import TensorDictLite
    T, A, dims = 6, 3, {"policy": 7, "priv": 4}
    obs, rew, dones, z = _distill_streams(2 * n_per_rank, T, A, dims)
    cols = slice(rank * n_per_rank, (rank + 1) * n_per_rank) if world > 1 else slice(None)
    n = n_per_rank if world > 1 else 2 * n_per_rank
    td = lambda t: TensorDictLite({k: v[t, cols].contiguous() for k, v in obs.items()}, batch_size=[n])
    torch.manual_seed(seed)
    policy = StudentTeacher(td(0), {"policy": ["policy"], "teacher": ["policy", "priv"]}, A, student_hidden_dims=[16, 8],
                            teacher_hidden_dims=[16, 8])
    alg = Distillation(policy, num_learning_epochs=2, gradient_length=4, learning_rate=1e-3, max_grad_norm=0.02,
                       device="cpu", multi_gpu_cfg=None if world == 1 else dict(global_rank=rank, world_size=world))
    alg.init_storage("distillation", n, T, td(0), [A])
    if world > 1:
        alg.broadcast_parameters()     # rank 0's parameters (and buffers) everywhere, one flat broadcast
    zq = [z[t, cols].clone() for t in range(T)]
    policy._noise_fn = lambda like: zq.pop(0)
    with torch.inference_mode():
        for t in range(T):
            alg.act(td(t))
            alg.process_env_step(td(t + 1), rew[t, cols].contiguous(), dones[t, cols].contiguous(), {})
    loss = alg.update()
    return torch.cat([p.detach().reshape(-1) for p in policy.parameters()]).numpy().copy(), loss["behavior"]


def _distill_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import host_emulation
    host_emulation.install(_Setattr())
    params, loss = _distill_run(rank, world, 16, seed=100 + rank)   # different initial weights before the broadcast
    np.save(os.path.join(out_dir, f"distill_rank{rank}.npy"), np.concatenate([params, [loss]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_distillation_equals_one_rank_on_all_envs(tmp_path, monkeypatch):
    """Host logic of multi-GPU distillation above the emulated ABI (tests/host_emulation.py), gloo in place of NCCL: after
    the flat parameter broadcast, gradient averaging before the student-only clip and the fused step, (1) both ranks hold
    bit-identical parameters and (2) they equal ONE rank training on the union of the envs -- the behaviour-cloning loss is
    a mean over envs, so the mean of the two ranks' gradients is the gradient of the doubled batch (fp32 summation order
    aside)."""
    port = 29900 + os.getpid() % 90
    mp.spawn(_distill_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "distill_rank0.npy"), np.load(tmp_path / "distill_rank1.npy")
    assert np.array_equal(a[:-1], b[:-1]), "ranks disagree after the update"
    from tests import host_emulation
    host_emulation.install(monkeypatch)
    single, loss = _distill_run(0, 1, 16, seed=100)     # rank 0's initial weights, all 32 envs
    assert np.abs(a[:-1] - single).max() < 2e-6 * np.abs(single).max()
    assert abs(0.5 * (a[-1] + b[-1]) - loss) < 1e-6 * max(1.0, abs(loss))   # mean of the ranks' losses = loss on the union


# ---------------------------------------------------------------------------------------------------------------------
# PPO.update on two ranks, NCCL-path host logic (reference ppo.py:272-290, 428-469) against the oracle's world hook
# ---------------------------------------------------------------------------------------------------------------------

def _ppo_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RSLRL_ALLREDUCE="nccl")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import host_emulation
    host_emulation.install(_Setattr())
    from oracle import ppo_oracle as O
    from rsl_rl_b200.algorithms import PPO
    from rsl_rl_b200.modules import ActorCritic
    from rsl_rl_b200.utils import TensorDictLite

    N, T, D, A, hidden = 48, 6, 10, 3, [24, 12]
    cfg = dict(learning_rate=1e-3, num_learning_epochs=2, num_mini_batches=3, schedule="adaptive", value_loss_coef=1.0,
               clip_param=0.2, use_clipped_value_loss=True, desired_kl=0.01, entropy_coef=0.01, gamma=0.99, lam=0.95,
               max_grad_norm=0.5)
    rng = np.random.default_rng(50 + rank)                       # every rank owns its own envs, noise and permutation
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    obs, rewards, z = f(T + 1, N, D), f(T, N), f(T, N, A)
    dones = (rng.random((T, N)) < 0.1).astype(np.int64)
    time_outs = (rng.random((T, N)) < 0.5) & (dones > 0)
    perm = rng.permutation(N * T).astype(np.int64)
    td = lambda t: TensorDictLite({"policy": torch.from_numpy(obs[t])}, batch_size=[N])

    torch.manual_seed(200 + rank)                                # different initial weights before the broadcast
    policy = ActorCritic(td(0), {"policy": ["policy"], "critic": ["policy"]}, A, actor_hidden_dims=hidden,
                         critic_hidden_dims=hidden, activation="elu")
    alg = PPO(policy, device="cpu", multi_gpu_cfg=dict(global_rank=rank, world_size=world), gemm_precision="fp32",
              use_cuda_graph=False, rng_compat=False, **cfg)
    assert not alg.optimizer.peer_allreduce                      # the NCCL path: in-place all-reduce + two launches
    alg.init_storage("rl", N, T, td(0), [A])
    alg.broadcast_parameters()
    init = OrderedDict((k, v.detach().numpy().copy()) for k, v in policy.named_parameters())
    zq = [torch.from_numpy(z[t]) for t in range(T)]
    policy._noise_fn = lambda like: zq.pop(0)
    alg.storage._randperm = lambda n: torch.from_numpy(perm)
    with torch.inference_mode():
        for t in range(T):
            alg.act(td(t))
            alg.process_env_step(td(t + 1), torch.from_numpy(rewards[t]), torch.from_numpy(dones[t]),
                                 {"time_outs": torch.from_numpy(time_outs[t])})
        alg.compute_returns(td(T))
    alg.update()
    mine = np.concatenate([p.detach().numpy().reshape(-1) for p in policy.parameters()])

    # the oracle on this rank's rollout, gradients and KL averaged over the ranks where the reference does it
    def world_hook(grads, kl):
        flat = torch.from_numpy(np.concatenate([g.reshape(-1) for g in grads.values()] + [[np.float32(kl)]]))
        dist.all_reduce(flat)
        flat = (flat / world).numpy()
        out, off = OrderedDict(), 0
        for k, g in grads.items():
            out[k] = flat[off:off + g.size].reshape(g.shape).astype(np.float32)
            off += g.size
        return out, float(flat[-1])

    ocfg = dict(cfg, activation="elu", noise_std_type="scalar")
    params = OrderedDict((k, v.copy()) for k, v in init.items())
    st, last_values = O.rollout(params, obs, obs, rewards, dones, time_outs, z, ocfg)
    st["returns"], st["advantages"] = O.compute_returns(st["rewards"], st["values"], st["dones"], last_values, 0.99, 0.95)
    _, lr, trace = O.update(params, O.adam_init(params), 1e-3, st, perm, ocfg, world_hook=world_hook)
    want = np.concatenate([v.reshape(-1) for v in params.values()])
    err = float(np.abs(mine - want).max() / np.abs(want).max())
    np.save(os.path.join(out_dir, f"ppo_rank{rank}.npy"), np.concatenate([mine, [alg.learning_rate, lr, err]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ppo_update_matches_the_oracle_on_rank_averaged_gradients(tmp_path):
    """PPO.update with `multi_gpu_cfg` on two gloo ranks above the emulated ABI: per minibatch ONE in-place all-reduce of
    [gradients | KL], then norm / clip / KL-adaptive LR (fp32-rounded, as the reference's broadcast makes it) / Adam on
    the rank-averaged values.  Every rank keeps its own envs, advantage statistics and permutation.  Checked:
    replicas bit-identical, same LR, and each rank equal to the numpy oracle run on its own rollout with the world hook."""
    port = 29700 + os.getpid() % 150
    mp.spawn(_ppo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "ppo_rank0.npy"), np.load(tmp_path / "ppo_rank1.npy")
    assert np.array_equal(a[:-3], b[:-3]) and a[-3] == b[-3], "replicas disagree after update()"
    for r in (a, b):
        assert r[-1] < 1e-5, r[-1]                                  # product vs oracle, max-abs / max-abs
        assert abs(r[-3] / r[-2] - 1.0) < 1e-6 and r[-3] != 1e-3    # same LR trajectory (the rule fired at least once)
