#!/usr/bin/env python
"""Benchmark of the PPO learning iteration (BASELINE.json metric: env-steps/s per learn() iteration at
4096 envs/GPU; GAE HBM GB/s).

    python bench.py --gpus N --steps K --warmup W                # this repo's CUDA path (config C2)
    python bench.py --config c3 ...                              # large-batch config C3 (65 536 envs, obs 235)
    python bench.py --impl reference --gpus N --steps K ...      # the UNMODIFIED reference (baseline/_ref) on host cores

A "step" is one `OnPolicyRunner.learn` iteration (24 env steps of rollout on a seeded synthetic VecEnv, GAE,
5 epochs x 4 minibatches of forward / loss / backward / clip / Adam).  C2: ActorCritic MLP [512,256,128] ELU, obs 48,
act 12, 4096 envs per GPU.  For N > 1 launch under torchrun (one rank per GPU); every rank owns its own envs (weak
scaling) and the only collective is the per-minibatch gradient all-reduce.  Rank 0 prints ONE JSON line.
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

NUM_STEPS, NUM_ACTIONS = 24, 12
HIDDEN = [512, 256, 128]
EPOCHS, MINIBATCHES = 5, 4
UNIT = "env-steps/s"

CONFIGS = {
    "c2": dict(
        envs=4096, groups={"policy": 48}, obs_groups={"policy": ["policy"], "critic": ["policy"]},
        metric="ppo_env_steps_per_s_per_learn_iter_4096envs_per_gpu",
        workload="C2: PPO ActorCritic MLP[512,256,128] ELU, seeded synthetic VecEnv, 4096 envs/GPU, obs 48, act 12, "
                 "24 steps, 5 epochs x 4 minibatches, adaptive LR",
        l2="working set (rollout 35 MB + activations/gradients 0.2 GB per minibatch) exceeds the 126 MB L2; "
           "no explicit flush between iterations"),
    "c3": dict(
        envs=65536, groups={"policy": 48, "height_scan": 187},
        obs_groups={"policy": ["policy", "height_scan"], "critic": ["policy", "height_scan"]},
        metric="ppo_env_steps_per_s_per_learn_iter_65536envs_per_gpu",
        workload="C3: large-batch PPO ActorCritic MLP[512,256,128] ELU, seeded synthetic VecEnv, 65536 envs/GPU, "
                 "obs 235 (proprio 48 + height scan 187 as two observation groups), act 12, 24 steps, "
                 "5 epochs x 4 minibatches, adaptive LR",
        l2="working set (rollout 1.7 GB + permuted copy 1.7 GB) far exceeds the 126 MB L2; no explicit flush"),
}
# kept for callers of the round-1 module constants (profiles/experiments/*.py)
NUM_ENVS, OBS_DIM = CONFIGS["c2"]["envs"], 48
METRIC, WORKLOAD = CONFIGS["c2"]["metric"], CONFIGS["c2"]["workload"]


def train_cfg(precision: str, config: str = "c2", **alg_extra):
    c = CONFIGS[config]
    return dict(
        num_steps_per_env=NUM_STEPS, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in c["obs_groups"].items()},
        policy=dict(class_name="ActorCritic", actor_hidden_dims=list(HIDDEN), critic_hidden_dims=list(HIDDEN),
                    activation="elu", init_noise_std=1.0, noise_std_type="scalar", actor_obs_normalization=False,
                    critic_obs_normalization=False),
        algorithm=dict(class_name="PPO", learning_rate=1e-3, num_learning_epochs=EPOCHS, num_mini_batches=MINIBATCHES,
                       schedule="adaptive", value_loss_coef=1.0, clip_param=0.2, use_clipped_value_loss=True,
                       desired_kl=0.01, entropy_coef=0.01, gamma=0.99, lam=0.95, max_grad_norm=1.0,
                       normalize_advantage_per_mini_batch=False, gemm_precision=precision, **alg_extra))


def workload_config(config: str, world: int):
    """The `config` object of the JSON line: identical on both arms (ours / reference)."""
    c = CONFIGS[config]
    return dict(workload=c["workload"], envs_per_gpu=c["envs"], global_envs=c["envs"] * world, l2=c["l2"])


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


def gemm_flops_per_minibatch(mb: int, obs_dim: int):
    dims_a, dims_c = [obs_dim, *HIDDEN, NUM_ACTIONS], [obs_dim, *HIDDEN, 1]
    fwd = 2 * sum(a * b for d in (dims_a, dims_c) for a, b in zip(d[:-1], d[1:]))
    first = 2 * (dims_a[0] * dims_a[1] + dims_c[0] * dims_c[1])
    return (3 * fwd - first) * mb  # fwd + wgrad + dgrad (no dgrad into the observations)


def time_dominant_kernel(alg, obs_dim: int, reps: int = 2):
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
    flops = gemm_flops_per_minibatch(mb, obs_dim)
    return dict(ms_per_minibatch=total_ms / n_mb, launches_per_minibatch=launches_per_mb,
                avg_launch_us=1e3 * total_ms / n_mb / launches_per_mb, tflops=flops / (total_ms / n_mb * 1e-3) / 1e12,
                flops_per_launch=flops / launches_per_mb)


def dense_traffic(precision: str, config: str):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dense launches of ONE minibatch from the
    committed `ncu --set full` capture (profiles/dense_traffic.json is written by profiles/summarize_ncu.py from the
    .ncu-rep of this build); None when this build / mode has no capture."""
    path = os.path.join(ROOT, "profiles", "dense_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        e = t[f"{config}_{precision}"]
        return e["bytes_per_minibatch"] / e["launches_per_minibatch"], e.get("source")
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------------
# HBM-side kernels, timed alone on cold data: every call of a kernel works on its own buffer set, the sets of one
# measurement together exceed 2x the L2, and all calls are replayed back to back from ONE CUDA graph (no Python between
# launches); us = graph time / calls.
# ------------------------------------------------------------------------------------------------------
def _rotating_time(make_call, n_sets: int, rounds: int = 5):
    import torch

    calls = [make_call(i) for i in range(n_sets)]
    for c in calls:
        c()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for c in calls:
                c()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rounds * n_sets)


def _n_sets(nbytes: int, lo: int = 2, hi: int = 256, target: float = 300e6):
    return int(min(hi, max(lo, target // nbytes + 1)))


def time_gae(device, N: int, T: int = NUM_STEPS):
    """GAE reverse scan + advantage normalisation: algorithmic bytes 17*T*N + 4*N."""
    import torch
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite

    nbytes = 17 * T * N + 4 * N
    n = _n_sets(nbytes)
    sts = []
    for _ in range(n):
        st = RolloutStorage("rl", N, T, TensorDictLite({"policy": torch.zeros(N, 1, device=device)}, batch_size=[N]),
                            [1], device)
        st.rewards.normal_(); st.values.normal_()
        st.dones.copy_((torch.rand(T, N, 1, device=device) < 0.02).to(torch.uint8))
        sts.append(st)
    last = torch.randn(N, 1, device=device)
    us = _rotating_time(lambda i: (lambda st=sts[i]: st.compute_returns(last, 0.99, 0.95)), n)
    return dict(envs=N, steps=T, bytes=nbytes, us=us, gbs=nbytes / us / 1e3, sets=n)


def time_hbm_kernels(device, N: int, group_dims: dict, precision: str, peak_gbs: float):
    """gather / loss / optimiser step / store at the shape of `N` envs with the given observation groups."""
    import torch
    from rsl_rl_b200 import _native as nat
    from rsl_rl_b200.algorithms.flat_params import FlatParamGroup, FusedAdam
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite

    T, A, M = NUM_STEPS, NUM_ACTIONS, MINIBATCHES
    B, O = T * N, sum(group_dims.values())
    mb = B // M
    L = nat.lib()
    f32 = dict(dtype=torch.float32, device=device)
    out = {}

    def entry(nbytes, us, sets, what):
        return dict(bytes=nbytes, us=us, gbs=nbytes / us / 1e3, frac=nbytes / us / 1e3 / peak_gbs, sets=sets, what=what)

    # ---- store: one env step's row write (observations + rewards + dones; act() wrote the rest in place)
    obs0 = TensorDictLite({k: torch.randn(N, d, device=device) for k, d in group_dims.items()}, batch_size=[N])
    store_bytes = 2 * N * (4 * O + 4 + 1) + 4 * N
    n_st = max(1, min(8, int(300e6 // (store_bytes * T)) + 1))
    sts = [RolloutStorage("rl", N, T, obs0, [A], device) for _ in range(n_st)]
    rew, dones = torch.randn(N, device=device), (torch.rand(N, device=device) < 0.02).long()
    touts = dones.bool()

    def store_call(i):
        st = sts[i // T]

        def run():
            st.step = i % T
            st._store_in_place(obs0, rew, dones, touts, None, 0.99)
        return run

    us = _rotating_time(store_call, n_st * T)
    out["store"] = entry(store_bytes, us, n_st * T, "rollout-buffer row write of one env step (obs + rewards with the "
                                                   "time-out bootstrap + dones)")
    for st in sts:
        st.step = 0

    # ---- gather: permuted copy of the whole rollout (once per update)
    row = 4 * O + 12 * A + 16
    gather_bytes = B * (2 * row + 8)
    n_g = max(1, min(len(sts), int(300e6 // gather_bytes) + 1))
    perm = torch.randperm(B, device=device)
    pad4 = lambda d: (d + 3) // 4 * 4
    dsts = []
    for i in range(n_g):
        st = sts[i]
        for t in (st.actions, st.mu, st.sigma, st.values, st.advantages, st.returns, st.actions_log_prob):
            t.normal_()
        for k in group_dims:
            st.observations[k].normal_()
        dsts.append(dict(x=torch.zeros(B, pad4(O), **f32), a=torch.empty(B, A, **f32), m=torch.empty(B, A, **f32),
                         s=torch.empty(B, A, **f32), v=torch.empty(B, **f32), adv=torch.empty(B, **f32),
                         r=torch.empty(B, **f32), lp=torch.empty(B, **f32)))

    def gather_call(i):
        st, d = sts[i], dsts[i]
        fields, widths, outs, strides, cols = [], [], [], [], []
        col = 0
        for k, w in group_dims.items():
            fields.append(st.observations[k]); widths.append(w); outs.append(d["x"]); strides.append(d["x"].stride(0))
            cols.append(col)
            col += w
        for src, dst, w in ((st.actions, d["a"], A), (st.mu, d["m"], A), (st.sigma, d["s"], A), (st.values, d["v"], 1),
                            (st.advantages, d["adv"], 1), (st.returns, d["r"], 1), (st.actions_log_prob, d["lp"], 1)):
            fields.append(src); widths.append(w); outs.append(dst); strides.append(w); cols.append(0)
        return lambda: st._gather(fields, widths, perm, outs, strides, cols)

    us = _rotating_time(gather_call, n_g)
    out["gather"] = entry(gather_bytes, us, n_g, "permuted gather of the whole rollout with observation-group concat")

    # ---- loss: PPO objective forward + hand-derived backward of one minibatch
    loss_bytes = (20 * A + 24) * mb
    n_l = _n_sets(loss_bytes, hi=64)
    std = torch.ones(A, **f32)
    dstd = torch.zeros(A, **f32)
    ws = torch.zeros(int(L.rslrl_loss_workspace_bytes()), dtype=torch.uint8, device=device)
    sums = torch.zeros(3, dtype=torch.float64, device=device)
    kl = torch.zeros(1, **f32)
    f16 = precision == "fp16"
    lsets = []
    for _ in range(n_l):
        s = dict(mu=torch.randn(mb, A, **f32), v=torch.randn(mb, **f32), a=torch.randn(mb, A, **f32),
                 lp=torch.randn(mb, **f32) - 17.0, adv=torch.randn(mb, **f32), r=torch.randn(mb, **f32),
                 ov=torch.randn(mb, **f32), om=torch.randn(mb, A, **f32), os=torch.ones(mb, A, **f32),
                 dmu=torch.empty(mb, A, **f32), dv=torch.empty(mb, 4, **f32))
        if f16:
            s["dmu16"] = torch.empty(mb, 16, dtype=torch.float16, device=device)
            s["dv16"] = torch.empty(mb, 8, dtype=torch.float16, device=device)
        lsets.append(s)

    def loss_call(i):
        s = lsets[i]
        return lambda: nat.check(L.rslrl_ppo_loss_fwd_bwd(
            s["mu"].data_ptr(), std.data_ptr(), nat.STD_SCALAR, s["v"].data_ptr(), s["a"].data_ptr(), s["lp"].data_ptr(),
            s["adv"].data_ptr(), s["r"].data_ptr(), s["ov"].data_ptr(), s["om"].data_ptr(), s["os"].data_ptr(), None,
            s["dmu"].data_ptr(), A, s["dv"].data_ptr(), 4, dstd.data_ptr(), None, None, kl.data_ptr(), sums.data_ptr(),
            ws.data_ptr(), mb, mb, A, 0.2, 1.0, 0.01, 1,
            s["dmu16"].data_ptr() if f16 else None, 16 if f16 else 0, s["dv16"].data_ptr() if f16 else None,
            8 if f16 else 0, float(mb), None, None, nat.stream_ptr()))

    us = _rotating_time(loss_call, n_l)
    out["loss"] = entry(loss_bytes, us, n_l, "PPO objective forward + backward of one minibatch"
                        + (" (+ fp16 gradient copies for the tensor-core backward, not counted)" if f16 else ""))

    # ---- optimiser step: grad-norm + clip + KL-adaptive LR + Adam in one launch (32 P bytes)
    dims_a, dims_c = [O, *HIDDEN, A], [O, *HIDDEN, 1]
    shapes = [(A,)] + [s for d in (dims_a, dims_c) for a, b in zip(d[:-1], d[1:]) for s in ((b, a), (b,))]
    P = sum(int(torch.Size(s).numel()) for s in shapes)
    adam_bytes = 32 * P
    n_a = _n_sets(adam_bytes, hi=48)
    opts = []
    for _ in range(n_a):
        grp = FlatParamGroup([torch.nn.Parameter(torch.randn(*s, **f32) * 0.05) for s in shapes], device)
        grp.grads.normal_()
        if f16:
            grp.enable_shadow16()
        opts.append(FusedAdam(grp, lr=1e-3))

    def adam_call(i):
        o = opts[i]
        return lambda: o.fused_step(world_size=1, max_grad_norm=1.0, adaptive=True, desired_kl=0.01, zero_grad=False)

    us = _rotating_time(adam_call, n_a)
    out["adam"] = entry(adam_bytes, us, n_a, "grad-norm + clip + KL-adaptive LR + Adam, one launch"
                        + (" (+ fp16 weight shadow refresh, not counted)" if f16 else ""))
    return out


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the UNMODIFIED reference (pip-installed under baseline/_ref) with
    device='cpu' on the box's host cores, through its own OnPolicyRunner.learn, on this arm's config.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import reference_arm as R

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    why = R.available()
    if why:
        print(json.dumps(dict(impl="reference", unavailable=why)))
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)   # torchrun pins OMP_NUM_THREADS=1; the other ranks exit without work
    run = R.ReferenceRun(train_cfg("fp32", args.config), cfg["envs"], cfg["groups"], NUM_ACTIONS, device="cpu")
    # bounded sample: whole iterations at the full env count; warm-up and step counts shrink when one iteration is
    # so slow that K + W of them would not fit in a few minutes (stated in `sample`)
    first = run.iterate(1)
    budget = 200.0
    warm = max(0, min(args.warmup - 1, int(budget * 0.2 / first)))
    steps = max(1, min(args.steps, int(budget * 0.8 / first)))
    if warm:
        run.iterate(warm)
    el = run.iterate(steps)
    value = cfg["envs"] * NUM_STEPS * steps / el
    sample = (f"{steps} full learn() iterations ({cfg['envs']} envs x {NUM_STEPS} steps, {EPOCHS}x{MINIBATCHES} minibatches) "
              f"of the unmodified reference (rsl-rl-lib 3.1.0, baseline/_ref) on device='cpu' after {warm + 1} warm-up "
              f"iterations, torch {torch.__version__}, {torch.get_num_threads()} threads on {cores} host cpus; "
              f"log_dir=None, disable_logs=True (no logging on either arm)")
    line = dict(metric=cfg["metric"], value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm + 1,
                ms_per_step=1e3 * el / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference", config=workload_config(args.config, world),
                arm=dict(parallelism="cpu", device="cpu", requested_steps=args.steps, requested_warmup=args.warmup),
                cpu_baseline=dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind="reference", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def reference_on_cuda(config: str, device: str, iters: int = 10, warmup: int = 3):
    """The unmodified reference on cuda:0 (eager PyTorch: cuBLAS SGEMM, foreach Adam), same config and synthetic env --
    the in-box GPU comparison SURVEY section 8(d) asks for."""
    import torch

    from oracle import reference_arm as R

    why = R.available()
    if why:
        return dict(unavailable=why)
    cfg = CONFIGS[config]
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    run = R.ReferenceRun(train_cfg("fp32", config), cfg["envs"], cfg["groups"], NUM_ACTIONS, device=device)
    run.iterate(warmup)
    el = run.iterate(iters)
    del run
    torch.cuda.empty_cache()
    return dict(value=cfg["envs"] * NUM_STEPS * iters / el, unit=UNIT, ms_per_step=1e3 * el / iters, iters=iters,
                warmup=warmup, device=device, dtype="f32 (torch defaults: allow_tf32=%s)" % (tf32[0],),
                note="unmodified reference (baseline/_ref), eager PyTorch on the same GPU, device-resident synthetic env, "
                     "log_dir=None; wall clock around OnPolicyRunner.learn with a device synchronise on both sides")


def run_ours(args):
    import torch

    from rsl_rl_b200 import _native as nat
    from rsl_rl_b200.env import PinnedHostEnv, SynthVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner

    cfg = CONFIGS[args.config]
    envs, groups = cfg["envs"], cfg["groups"]
    obs_dim = sum(groups.values())
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
    env = SynthVecEnv(envs, groups, NUM_ACTIONS, device=device, seed=rank)
    runner = OnPolicyRunner(env, train_cfg(args.precision, args.config), log_dir=None, device=device)
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
    value = envs * NUM_STEPS * world * K / (ms * 1e-3)

    if args.only_timed:
        restore_stdout()
        if rank == 0:
            print(json.dumps(dict(metric=cfg["metric"], value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W,
                                  ms_per_step=ms / K, gpu_launches=int(launches), clocks=clocks,
                                  note="--only-timed: profiling aid, not the bench line")))
        return

    # ---- multi-GPU post-condition: parameters, Adam moments and learning rate bit-identical on all ranks
    ranks_identical = None
    if world > 1:
        g = runner.alg._group
        digest = torch.stack([g.values.double().sum(), g.values.double().abs().sum(), g.exp_avg.double().sum(),
                              g.exp_avg_sq.double().sum(),
                              torch.tensor(runner.alg.learning_rate, dtype=torch.float64, device=device)])
        lo, hi = digest.clone(), digest.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        bits = g.values.view(torch.int32).clone()
        bmin, bmax = bits.clone(), bits.clone()
        torch.distributed.all_reduce(bmin, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(bmax, op=torch.distributed.ReduceOp.MAX)
        ranks_identical = bool(torch.equal(lo, hi)) and bool(torch.equal(bmin, bmax))

    # ---- per-phase device times (outside the timed region)
    runner.profile_phases = True
    runner.learn(3)
    phases = dict(runner.phase_times)
    runner.profile_phases = False

    # ---- end-to-end: env on the host (pinned), per-step H2D of obs/rewards/dones and D2H of actions + scalars
    torch.manual_seed(0)
    henv = PinnedHostEnv(envs, groups, NUM_ACTIONS, pool=32, seed=rank)
    hrunner = OnPolicyRunner(henv, train_cfg(args.precision, args.config), log_dir=None, device=device)
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
    e2e_value = envs * NUM_STEPS * world * K / (e2e_ms * 1e-3)
    h2d = henv.h2d_bytes_per_step * NUM_STEPS + envs * obs_dim * 4  # + the observation fed to compute_returns
    d2h = henv.d2h_bytes_per_step * NUM_STEPS + 128                 # actions + the update() scalar read-back

    line = None
    if rank == 0:
        peaks = measured_peaks()
        dom = time_dominant_kernel(runner.alg, obs_dim)
        traffic, traffic_src = dense_traffic(args.precision, args.config)
        roofline = dict(bound="tensor", achieved=dom["tflops"], peak=peaks["tf_burst"], unit="TFLOP/s",
                        frac=dom["tflops"] / peaks["tf_burst"], frac_of_sustained_peak=dom["tflops"] / peaks["tf_sustained"],
                        traffic=traffic,
                        traffic_note=(f"dram__bytes_read.sum + dram__bytes_write.sum averaged over the dense launches of one "
                                      f"minibatch ({traffic_src})" if traffic is not None else
                                      "no ncu --set full capture of this build / mode under profiles/"),
                        kernel=("sgemm_kernel (fp32 CUDA-core FMA)" if args.precision == "fp32"
                                else ("mlp3_forward_kernel + dgrad_chain16_kernel + tc_gemm_persistent_kernel<wgrad> (fp16 tcgen05): "
                                      "all dense launches of one minibatch forward + backward" if args.precision == "fp16" else
                                      f"mlp3_forward_kernel / tc_gemm_persistent_kernel<fwd|dgrad|wgrad> ({args.precision} "
                                      "tcgen05): all dense launches of one minibatch forward + backward")),
                        avg_launch_us=dom["avg_launch_us"], flops_per_launch=dom["flops_per_launch"],
                        launches_per_minibatch=dom["launches_per_minibatch"], us_per_minibatch=1e3 * dom["ms_per_minibatch"],
                        peak_source=peaks["source"] + ", bf16 dense BURST peak (the dense launches are timed alone in "
                                                      "millisecond bursts of graph replays at boost clock)")
        # HBM-side kernels at both shapes (C1/C2: 4096 envs, obs 48; C3: 65536 envs, obs 235)
        hbm = {}
        for tag in ("c2", "c3"):
            c = CONFIGS[tag]
            gae = time_gae(device, c["envs"])
            h = time_hbm_kernels(device, c["envs"], c["groups"], args.precision, peaks["hbm_gbs"])
            h["gae"] = dict(bytes=gae["bytes"], us=gae["us"], gbs=gae["gbs"], frac=gae["gbs"] / peaks["hbm_gbs"],
                            sets=gae["sets"], what="GAE reverse scan + advantage normalisation (17*T*N + 4*N bytes)")
            hbm["c1" if tag == "c2" else tag] = h
            torch.cuda.empty_cache()
        gae = dict(c1=hbm["c1"]["gae"], c3=hbm["c3"]["gae"], peak_gbs=peaks["hbm_gbs"], frac_c1=hbm["c1"]["gae"]["frac"],
                   frac_c3=hbm["c3"]["gae"]["frac"])
        hbm["peak_gbs"] = peaks["hbm_gbs"]
        hbm["note"] = ("each kernel timed alone on cold data: every call works on its own buffer set (sets x bytes > 2x L2), "
                       "all calls replayed back to back from one CUDA graph; us = graph time / calls; bytes = ALGORITHMIC "
                       "bytes of SURVEY 8(d)")
        ref_cuda = None
        if world == 1 and not args.no_reference_cuda:
            try:
                ref_cuda = reference_on_cuda(args.config, device, iters=10 if args.config == "c2" else 2,
                                             warmup=3 if args.config == "c2" else 1)
                if "value" in ref_cuda:
                    ref_cuda["ours_over_reference_cuda"] = value / ref_cuda["value"]
            except Exception as exc:  # the bench line must not die with the comparison leg
                ref_cuda = dict(unavailable=f"{type(exc).__name__}: {exc}")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_sample(args.config)
        line = dict(metric=cfg["metric"], value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                    higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"fp32": "f32", "tf32": "tf32", "fp16": "fp16 operands, fp32 accumulate",
                           "tf32x3": "f32 (3xTF32 split-operand tensor-core products, fp32 accumulate)"}[args.precision],
                    data="synthetic", config=workload_config(args.config, world),
                    arm=dict(parallelism=f"dp{world}", gemm_precision=args.precision,
                             cuda_graph=bool(runner.alg._graph is not None)),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=e2e_ms / K,
                             note="OnPolicyRunner.learn with the env on the host: pinned obs/rewards/dones copied in, "
                                  "actions and loss scalars copied out every step"),
                    gpu_launches=int(launches), roofline=roofline, gae=gae, hbm_kernels=hbm,
                    phases_ms=dict(phases, iteration=ms / K))
        if ranks_identical is not None:
            line["ranks_identical"] = ranks_identical
        if ref_cuda is not None:
            line["reference_cuda"] = ref_cuda
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


def cpu_reference_sample(config: str):
    """cpu_baseline of the default line: a bounded sample (about 10-30 s) of the unmodified reference on the host cores
    (falls back to the oracle's torch-CPU port, kind "port", when baseline/_ref did not travel)."""
    import torch

    from oracle import reference_arm as R

    cfg = CONFIGS[config]
    if R.available() is None:
        run = R.ReferenceRun(train_cfg("fp32", config), cfg["envs"], cfg["groups"], NUM_ACTIONS, device="cpu")
        first = run.iterate(1)
        n = max(1, min(4, int(20.0 / first)))
        el = run.iterate(n)
        return dict(value=cfg["envs"] * NUM_STEPS * n / el, unit=UNIT, cores=torch.get_num_threads(), kind="reference",
                    sample=f"{n} full learn() iterations ({cfg['envs']} envs) of the unmodified reference (baseline/_ref) on "
                           f"device='cpu' after 1 warm-up iteration, {el / n:.2f} s each on {os.cpu_count()} host cpus")
    from oracle import torch_cpu_port as P

    obs_dim = sum(cfg["groups"].values())
    params, stream, z, perm = P.random_problem(cfg["envs"], T=NUM_STEPS, obs_dim=obs_dim, A=NUM_ACTIONS, hidden=tuple(HIDDEN))
    adam = P.adam_init(params)
    t0 = time.perf_counter()
    P.learn_iteration(params, adam, 1e-3, stream, z, perm, P.DEFAULT_CFG)
    dt = time.perf_counter() - t0
    return dict(value=cfg["envs"] * NUM_STEPS / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"1 full iteration of oracle/torch_cpu_port.py in {dt:.1f} s (baseline/_ref missing)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("RSLRL_BENCH_CONFIG", "c2"), choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("RSLRL_PRECISION", "fp16"), choices=["fp32", "tf32", "fp16", "tf32x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--only-timed", action="store_true", help="profiling aid: warm-up + timed region only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
