#!/usr/bin/env python
"""Benchmark of the PPO learning iteration (BASELINE.json metric: env-steps/s per learn() iteration at
4096 envs/GPU; GAE HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

A "step" is one `OnPolicyRunner.learn` iteration (24 env steps of rollout on a seeded synthetic VecEnv,
GAE, 5 epochs x 4 minibatches of forward / loss / backward / clip / Adam) of config C2:
ActorCritic MLP [512,256,128] ELU, obs 48, act 12, 4096 envs per GPU.  For N > 1 launch under
torchrun (one rank per GPU); every rank owns its own envs (weak scaling) and the only collective is the
per-minibatch gradient all-reduce.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_ENVS, NUM_STEPS, OBS_DIM, NUM_ACTIONS = 4096, 24, 48, 12
HIDDEN = [512, 256, 128]
EPOCHS, MINIBATCHES = 5, 4
METRIC, UNIT = "ppo_env_steps_per_s_per_learn_iter_4096envs_per_gpu", "env-steps/s"
WORKLOAD = ("C2: PPO ActorCritic MLP[512,256,128] ELU, seeded synthetic VecEnv, 4096 envs/GPU, obs 48, act 12, "
            "24 steps, 5 epochs x 4 minibatches, adaptive LR")


def train_cfg(precision: str, **alg_extra):
    return dict(
        num_steps_per_env=NUM_STEPS, save_interval=10 ** 9, obs_groups={"policy": ["policy"], "critic": ["policy"]},
        policy=dict(class_name="ActorCritic", actor_hidden_dims=list(HIDDEN), critic_hidden_dims=list(HIDDEN),
                    activation="elu", init_noise_std=1.0, noise_std_type="scalar", actor_obs_normalization=False,
                    critic_obs_normalization=False),
        algorithm=dict(class_name="PPO", learning_rate=1e-3, num_learning_epochs=EPOCHS, num_mini_batches=MINIBATCHES,
                       schedule="adaptive", value_loss_coef=1.0, clip_param=0.2, use_clipped_value_loss=True,
                       desired_kl=0.01, entropy_coef=0.01, gamma=0.99, lam=0.95, max_grad_norm=1.0,
                       normalize_advantage_per_mini_batch=False, gemm_precision=precision, **alg_extra))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML every 100 ms while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # pragma: no cover
            self.err = str(exc)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


def gemm_flops_per_minibatch(mb: int):
    dims_a, dims_c = [OBS_DIM, *HIDDEN, NUM_ACTIONS], [OBS_DIM, *HIDDEN, 1]
    fwd = 2 * sum(a * b for d in (dims_a, dims_c) for a, b in zip(d[:-1], d[1:]))
    first = 2 * (dims_a[0] * dims_a[1] + dims_c[0] * dims_c[1])
    return (3 * fwd - first) * mb  # fwd + wgrad + dgrad (no dgrad into the observations)


def time_dominant_kernel(alg, reps: int = 2):
    """CUDA-event time of the GEMM launches (forward + backward of every minibatch of one update) on the
    launching stream; these are the dominant kernels of the iteration."""
    import torch

    eng, mb, M = alg._engine, alg._mb, alg.num_mini_batches
    xa, xc = alg._xa_raw[:, :alg._Oa], alg._xc_raw[:, :alg._Oc]
    from rsl_rl_b200 import _native as nat

    def dense(i):
        r = slice((i % M) * mb, (i % M + 1) * mb)
        eng.forward(xa[r], xc[r], eng.mu[:mb], eng.value[:mb])
        if eng.f16:  # the backward reads the fp16 copies of the minibatch observations
            eng.backward(alg._xa16[r, :alg._Oa], alg._xc16[r, :alg._Oc], eng.dmu[:mb], eng.dvalue[:mb])
        else:
            eng.backward(xa[r], xc[r], eng.dmu[:mb], eng.dvalue[:mb])

    # one CUDA graph per minibatch slice (host-side tensor-map encoding and ctypes overhead stay out of the timed
    # region, as in update() itself); replayed in update() order, each replay timed with events on its stream
    launches0 = nat.launch_count()
    dense(0)
    launches_per_mb = nat.launch_count() - launches0
    torch.cuda.synchronize()
    graphs = []
    for i in range(M):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            dense(i)
        graphs.append(g)
    total_ms, n_mb = 0.0, 0
    for _ in range(reps):
        for i in range(M * alg.num_learning_epochs):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graphs[i % M].replay()
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            n_mb += 1
    del graphs
    alg._group.zero_grad()
    flops = gemm_flops_per_minibatch(mb)
    return dict(ms_per_minibatch=total_ms / n_mb, launches_per_minibatch=launches_per_mb,
                avg_launch_us=1e3 * total_ms / n_mb / launches_per_mb, tflops=flops / (total_ms / n_mb * 1e-3) / 1e12,
                flops_per_launch=flops / launches_per_mb)


def time_gae(device, N: int, T: int = 24, rounds: int = 10):
    """Achieved HBM GB/s of the fused GAE + normalisation kernel (algorithmic bytes 17*T*N + 4*N).
    A rotating set of rollouts (> 3x the 126 MB L2 in total) is processed back to back from a CUDA graph, so
    every launch reads cold data and the figure is not bounded by Python's ~8 us per call."""
    import torch
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite

    nbytes = 17 * T * N + 4 * N
    nbuf = min(256, max(2, int(400e6 // nbytes) + 1))
    sts = []
    for _ in range(nbuf):
        st = RolloutStorage("rl", N, T, TensorDictLite({"policy": torch.zeros(N, 1, device=device)}, batch_size=[N]),
                            [1], device)
        st.rewards.normal_(); st.values.normal_()
        st.dones.copy_((torch.rand(T, N, 1, device=device) < 0.02).to(torch.uint8))
        sts.append(st)
    last = torch.randn(N, 1, device=device)
    for st in sts:
        st.compute_returns(last, 0.99, 0.95)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for st in sts:
                st.compute_returns(last, 0.99, 0.95)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (rounds * nbuf)
    return dict(envs=N, steps=T, bytes=nbytes, us=us, gbs=nbytes / us / 1e3, rollouts_in_rotation=nbuf)


def cpu_port_stepper(envs: int):
    """Returns step() -> seconds for one learn() iteration of the torch-CPU oracle port on `envs` envs
    (same nets, 24 steps, 5x4 minibatches); the synthetic rollout stream is generated once, outside."""
    from oracle import torch_cpu_port as P

    params, stream, z, perm = P.random_problem(envs, T=NUM_STEPS, obs_dim=OBS_DIM, A=NUM_ACTIONS, hidden=tuple(HIDDEN))
    state = dict(adam=P.adam_init(params), lr=1e-3)

    def step():
        t0 = time.perf_counter()
        _, state["lr"], _ = P.learn_iteration(params, state["adam"], state["lr"], stream, z, perm, P.DEFAULT_CFG)
        return time.perf_counter() - t0

    return step


def cpu_port_throughput(envs: int, iters: int = 1):
    import torch

    step = cpu_port_stepper(envs)
    dt = sum(step() for _ in range(iters)) / iters
    return envs * NUM_STEPS / dt, dt, torch.get_num_threads()


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference is pure Python on torch and cannot travel to the GPU box, so the arm times the
    oracle's torch-CPU port (same algorithm, same ATen CPU kernels) on the host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # bounded sample: size the per-step env count so that the whole run stays within a few minutes
    probe_envs = 512
    cpu_port_throughput(128)
    thr, dt, threads = cpu_port_throughput(probe_envs)
    budget = 150.0 / max(1, args.steps + args.warmup)
    envs = probe_envs
    while envs < NUM_ENVS and dt * (envs * 2 / probe_envs) < budget:
        envs *= 2
    step = cpu_port_stepper(envs)
    for _ in range(args.warmup):
        step()
    el = sum(step() for _ in range(args.steps))
    value = envs * NUM_STEPS * args.steps / el
    sample = f"{envs} of {NUM_ENVS} envs x {NUM_STEPS} steps per step (full 5x4 minibatch update), torch {torch.__version__} CPU"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * el / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference", config=dict(workload=WORKLOAD, parallelism="cpu"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_ours(args):
    import torch

    from rsl_rl_b200 import _native as nat
    from rsl_rl_b200.env import PinnedHostEnv, SynthVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    device = f"cuda:{local}"
    torch.cuda.set_device(local)
    K, W = args.steps, max(args.warmup, 3)
    # the modules print their architecture and NCCL prints its version banner on fd 1: keep stdout for the JSON line
    sys.stdout.flush()
    saved_fd = os.dup(1)
    devnull_fd = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull_fd, 1)
    stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")

    def restore_stdout():
        sys.stdout.flush()
        sys.stdout = stdout
        os.dup2(saved_fd, 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    # ---- device-resident run (value)
    torch.manual_seed(0)
    env = SynthVecEnv(NUM_ENVS, {"policy": OBS_DIM}, NUM_ACTIONS, device=device, seed=rank)
    runner = OnPolicyRunner(env, train_cfg(args.precision), log_dir=None, device=device)
    runner.learn(W)
    sampler = ClockSampler(local)
    sampler.start()              # (NVML initialisation takes milliseconds: keep it out of the rank-synchronised start)
    torch.cuda.profiler.start()  # lets `ncu --profile-from-start off` capture exactly the timed region
    barrier()
    launches0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runner.learn(K)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    clocks = sampler.finish()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = nat.launch_count() - launches0
    value = NUM_ENVS * NUM_STEPS * world * K / (ms * 1e-3)

    if args.only_timed:
        restore_stdout()
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W,
                                  ms_per_step=ms / K, gpu_launches=int(launches), clocks=clocks,
                                  note="--only-timed: profiling aid, not the bench line")))
        return

    # ---- per-phase device times (outside the timed region)
    runner.profile_phases = True
    runner.learn(3)
    phases = dict(runner.phase_times)
    runner.profile_phases = False

    # ---- end-to-end: env on the host (pinned), per-step H2D of obs/rewards/dones and D2H of actions + scalars
    torch.manual_seed(0)
    henv = PinnedHostEnv(NUM_ENVS, {"policy": OBS_DIM}, NUM_ACTIONS, pool=32, seed=rank)
    hrunner = OnPolicyRunner(henv, train_cfg(args.precision), log_dir=None, device=device)
    hrunner.learn(W)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    h0.record()
    hrunner.learn(K)
    h1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(h0.elapsed_time(h1), wall_ms))
    e2e_value = NUM_ENVS * NUM_STEPS * world * K / (e2e_ms * 1e-3)
    h2d = henv.h2d_bytes_per_step * NUM_STEPS + NUM_ENVS * OBS_DIM * 4  # + the observation fed to compute_returns
    d2h = henv.d2h_bytes_per_step * NUM_STEPS + 128                     # actions + the update() scalar read-back

    line = None
    if rank == 0:
        peaks = measured_peaks()
        dom = time_dominant_kernel(runner.alg)
        # DRAM bytes (read + write) per dense launch from the committed `ncu --set full` capture of one minibatch
        # (profiles/r1_ncu_full_update_fp16.txt: forward 38.1 MB, head backward 14.2, dgrad 38.6 + 88.9, grouped wgrad 172.3;
        # profiles/r1_ncu_full_update.txt for the tf32 build); None for modes without a capture
        traffic = {"fp16": 352.1e6 / 5}.get(args.precision)
        roofline = dict(bound="tensor", achieved=dom["tflops"], peak=peaks["tf_sustained"], unit="TFLOP/s",
                        frac=dom["tflops"] / peaks["tf_sustained"], traffic=traffic,
                        traffic_note="dram__bytes_read.sum + dram__bytes_write.sum averaged over the 5 dense launches of "
                                     "one minibatch (ncu --set full, profiles/r1_ncu_full_update_fp16.txt)",
                        kernel=("sgemm_kernel (fp32 CUDA-core FMA)" if args.precision == "fp32"
                                else f"mlp3_forward_kernel + tc_gemm_persistent_kernel<dgrad|wgrad> ({args.precision} tcgen05): "
                                     "all dense launches of one minibatch forward + backward"),
                        avg_launch_us=dom["avg_launch_us"], flops_per_launch=dom["flops_per_launch"],
                        peak_source=peaks["source"] + ", bf16 dense sustained")
        gae_c1, gae_c3 = time_gae(device, NUM_ENVS), time_gae(device, 65536)
        gae = dict(c1=gae_c1, c3=gae_c3, peak_gbs=peaks["hbm_gbs"], frac_c1=gae_c1["gbs"] / peaks["hbm_gbs"],
                   frac_c3=gae_c3["gbs"] / peaks["hbm_gbs"], note="rotating rollouts (> 3x L2 in total) processed back to back from a CUDA graph; every launch reads cold data")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_port_throughput(128)
            thr, dt, threads = cpu_port_throughput(NUM_ENVS)
            cpu = dict(value=thr, unit=UNIT, cores=threads, kind="port",
                       sample=f"1 full C2 iteration ({NUM_ENVS} envs x {NUM_STEPS} steps, 5x4 minibatches) of "
                              f"oracle/torch_cpu_port.py in {dt:.1f} s on {os.cpu_count()} host cpus")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype={"fp32": "f32", "tf32": "tf32", "fp16": "fp16 operands, fp32 accumulate"}[args.precision],
                    data="synthetic",
                    config=dict(workload=WORKLOAD, envs_per_gpu=NUM_ENVS, global_envs=NUM_ENVS * world,
                                parallelism=f"dp{world}", gemm_precision=args.precision,
                                l2="working set (rollout 35 MB + activations/gradients 0.4 GB) exceeds the 126 MB L2; "
                                   "no explicit flush between iterations",
                                cuda_graph=bool(runner.alg._graph is not None)),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=e2e_ms / K,
                             note="OnPolicyRunner.learn with the env on the host: pinned obs/rewards/dones copied in, "
                                  "actions and loss scalars copied out every step"),
                    gpu_launches=int(launches), roofline=roofline, gae=gae,
                    phases_ms=dict(phases, iteration=ms / K))
        if cpu is not None:
            line["cpu_baseline"] = cpu
    restore_stdout()
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1 and torch.distributed.is_initialized():
        # the captured update() graphs hold NCCL work: release them before the communicator goes away, and never let
        # the teardown of the process group keep a finished benchmark alive
        for r in (runner, hrunner):
            r.alg.release_graphs()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        torch.distributed.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RSLRL_PRECISION", "fp16"), choices=["fp32", "tf32", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-timed", action="store_true", help="profiling aid: warm-up + timed region only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
