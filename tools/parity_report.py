#!/usr/bin/env python
"""Measured parity of the CUDA path against the numpy oracle at benchmark shapes, per GEMM precision.

    python tools/parity_report.py [--out profiles/parity_r2.json] [--shapes c1,c3] [--precisions tf32x3,fp32,tf32,fp16]

For each (shape, precision): one full learn() iteration (rollout on a recorded synthetic stream, GAE, 5 x 4 minibatches)
with the noise and the permutation shared with the oracle; errors are max-abs error / max-abs reference value (SURVEY
section 8 A8).  The per-minibatch parameter trace is compared up to the first KL-adaptive LR decision that differs.
Also imported by tests/test_gpu_ppo_parity.py (the thresholds live in the test)."""
from __future__ import annotations

import argparse
import json
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPES = {
    "c1": dict(N=4096, T=24, groups={"policy": 48}, A=12, hidden=[512, 256, 128], seed=9),
    # C3's observation layout (proprio 48 + height scan 187 = 235 columns, two groups) at 8192 envs
    "c3": dict(N=8192, T=24, groups={"policy": 48, "height_scan": 187}, A=12, hidden=[512, 256, 128], seed=19),
}
_ORACLE_CACHE = {}


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def problem(shape: str, reward_scale: float = 1.0):
    s = SHAPES[shape]
    rng = np.random.default_rng(s["seed"])
    N, T, A = s["N"], s["T"], s["A"]
    f = lambda *d: rng.standard_normal(d).astype(np.float32)
    obs = {k: f(T + 1, N, d) for k, d in s["groups"].items()}
    rewards = (f(T, N) * np.float32(reward_scale)).astype(np.float32)
    dones = (rng.random((T, N)) < 0.02).astype(np.int64)
    time_outs = (rng.random((T, N)) < 0.5) & (dones > 0)
    z = f(T, N, A)
    perm = rng.permutation(N * T).astype(np.int64)
    keys = list(s["groups"])
    meta = dict(num_envs=N, num_steps=T, num_actions=A, hidden=s["hidden"], activation="elu", noise_std_type="scalar",
                obs_groups={"policy": keys, "critic": keys}, group_dims=dict(s["groups"]),
                alg=dict(learning_rate=1e-3, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive",
                         value_loss_coef=1.0, clip_param=0.2, use_clipped_value_loss=True, desired_kl=0.01,
                         entropy_coef=0.01, gamma=0.99, lam=0.95, max_grad_norm=1.0,
                         normalize_advantage_per_mini_batch=False))
    return meta, obs, rewards, dones, time_outs, z, perm


def init_params(meta, seed=0):
    import torch
    from rsl_rl_b200.modules import ActorCritic
    from rsl_rl_b200.utils import TensorDictLite

    torch.manual_seed(seed)
    obs = TensorDictLite({k: torch.zeros(2, d) for k, d in meta["group_dims"].items()}, batch_size=[2])
    pol = ActorCritic(obs, meta["obs_groups"], meta["num_actions"], actor_hidden_dims=meta["hidden"],
                      critic_hidden_dims=meta["hidden"], activation=meta["activation"])
    return OrderedDict((k, v.detach().numpy().copy()) for k, v in pol.named_parameters())


def oracle_run(shape: str, reward_scale: float = 1.0, gemm_round=None):
    """gemm_round="tf32": the oracle with every matmul operand rounded to tf32 (noise floor of the tensor-core class)."""
    from oracle import ppo_oracle as O

    key = (shape, reward_scale, gemm_round)
    if key not in _ORACLE_CACHE:
        O.GEMM_ROUND = gemm_round
        meta, obs, rewards, dones, time_outs, z, perm = problem(shape, reward_scale)
        params = init_params(meta)
        cfg = dict(meta["alg"], activation="elu", noise_std_type="scalar")
        cat = np.concatenate([obs[k] for k in meta["obs_groups"]["policy"]], axis=-1)
        stream = dict(obs_actor=cat, obs_critic=cat, rewards=rewards, dones=dones, time_outs=time_outs)
        st, last_values, loss_dict, lr, trace = O.learn_iteration(params, O.adam_init(params), 1e-3, stream, z, perm, cfg)
        O.GEMM_ROUND = None
        _ORACLE_CACHE[key] = dict(st=st, lr=lr, trace=trace, loss=loss_dict,
                                  final=np.concatenate([v.reshape(-1) for v in params.values()]))
    return _ORACLE_CACHE[key]


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def tf32_noise_floor(shape: str):
    """fp32 oracle vs the same oracle with tf32-rounded matmul operands: what the arithmetic class itself costs."""
    ref, emu = oracle_run(shape), oracle_run(shape, gemm_round="tf32")
    n = 0
    while n < len(ref["trace"]) and abs(emu["trace"][n]["lr"] - ref["trace"][n]["lr"]) <= 1e-12 * ref["trace"][n]["lr"]:
        n += 1
    per_step = [rel_err(emu["trace"][i]["params_after"], ref["trace"][i]["params_after"]) for i in range(n)]
    l2 = [rel_l2(emu["trace"][i]["params_after"], ref["trace"][i]["params_after"]) for i in range(n)]
    return dict(shape=shape, what="oracle(tf32-rounded matmul operands) vs oracle(fp32)",
                grad0_max_norm=rel_err(emu["trace"][0]["grads"], ref["trace"][0]["grads"]), lr_steps_in_agreement=n,
                params_after_max_norm=max(per_step) if per_step else None, params_after_step1=per_step[0] if per_step else None,
                params_after_rel_l2=max(l2) if l2 else None)


def cuda_run(shape: str, precision: str, reward_scale: float = 1.0, dev="cuda:0"):
    import torch
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner

    meta, obs, rewards, dones, time_outs, z, perm = problem(shape, reward_scale)
    init = init_params(meta)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    env = ReplayVecEnv({k: t(v) for k, v in obs.items()}, t(rewards), t(dones), t(time_outs), meta["num_actions"], device=dev)
    cfg = dict(num_steps_per_env=meta["num_steps"], save_interval=10 ** 9, obs_groups=meta["obs_groups"],
               policy=dict(class_name="ActorCritic", actor_hidden_dims=list(meta["hidden"]),
                           critic_hidden_dims=list(meta["hidden"]), activation="elu", init_noise_std=1.0),
               algorithm=dict(meta["alg"], class_name="PPO", gemm_precision=precision, use_cuda_graph=False))
    runner = OnPolicyRunner(env, cfg, log_dir=None, device=dev)
    alg = runner.alg
    alg.policy.load_state_dict({k: t(v).to(dev) for k, v in init.items()})
    assert alg.fused, alg._engine_reason
    zq = [t(x).to(dev) for x in z]
    alg.policy._noise_fn = lambda like: zq.pop(0)
    pt = t(perm).to(dev)
    alg.storage._randperm = lambda n: pt
    grp = alg._group
    pick = torch.cat([torch.arange(o, o + q.numel()) for o, q in zip(grp.offsets, grp.params)]).to(dev)
    grads, before, lrs = [], [], []

    def hook(a):   # after the backward of minibatch i, before its optimiser step
        grads.append(a._group.grads[pick].cpu().numpy().copy())
        before.append(a._group.values[pick].cpu().numpy().copy())
        lrs.append(float(a.optimizer.dev_state.buf[0]))      # lr in force BEFORE this minibatch's decision

    alg._mb_hook = hook
    runner.learn(1)
    final = grp.values[pick].cpu().numpy().copy()
    params_after = before[1:] + [final]
    lr_after = lrs[1:] + [alg.learning_rate]
    s = alg.storage
    storage = {k: getattr(s, k).cpu().numpy().copy() for k in ("mu", "values", "actions_log_prob", "returns", "advantages", "dones")}
    info = dict(alg.last_update_info)
    info["fp16_saturation_events"] = alg._fp16_saturations
    del runner, alg, env
    torch.cuda.empty_cache()
    return dict(grads=grads, params_after=params_after, lr_after=lr_after, storage=storage, final=final, init=init,
                info=info)


def compare(shape: str, precision: str, reward_scale: float = 1.0):
    ref = oracle_run(shape, reward_scale)
    mine = cuda_run(shape, precision, reward_scale)
    st, trace = ref["st"], ref["trace"]
    out = dict(shape=shape, precision=precision, reward_scale=reward_scale)
    for k in ("mu", "values", "actions_log_prob", "returns", "advantages"):
        out[f"rollout.{k}"] = rel_err(mine["storage"][k], st[k])
    out["rollout.dones_equal"] = bool(np.array_equal(mine["storage"]["dones"], st["dones"]))
    g0, g0r = mine["grads"][0].astype(np.float64), trace[0]["grads"].astype(np.float64)
    out["grad0.max_norm"] = rel_err(g0, g0r)
    out["grad0.rel_l2"] = float(np.linalg.norm(g0 - g0r) / np.linalg.norm(g0r))
    lr_ref = [tr["lr"] for tr in trace]
    agree = 0
    while agree < len(lr_ref) and abs(mine["lr_after"][agree] - lr_ref[agree]) <= 1e-12 * lr_ref[agree]:
        agree += 1
    out["lr.steps_in_agreement"] = agree
    out["lr.mine"], out["lr.ref"] = mine["lr_after"][-1], lr_ref[-1]
    per_step = [rel_err(mine["params_after"][i], trace[i]["params_after"]) for i in range(agree)]
    out["params_after.max_norm_per_step"] = per_step
    l2 = [rel_l2(mine["params_after"][i], trace[i]["params_after"]) for i in range(agree)]
    out["params_after.rel_l2_until_lr_divergence"] = max(l2) if l2 else None
    out["params_after.max_norm_until_lr_divergence"] = max(per_step) if per_step else None
    out["params_after.step1"] = per_step[0] if per_step else None
    init = np.concatenate([v.reshape(-1) for v in mine["init"].values()]).astype(np.float64)
    if agree == len(lr_ref):
        upd, upd_ref = mine["final"].astype(np.float64) - init, ref["final"].astype(np.float64) - init
        out["update.rel_l2"] = float(np.linalg.norm(upd - upd_ref) / np.linalg.norm(upd_ref))
        out["params_final.max_norm"] = rel_err(mine["final"], ref["final"])
    out["finite"] = bool(np.isfinite(mine["final"]).all())
    out["fp16_saturation_events"] = mine["info"]["fp16_saturation_events"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_r2.json"))
    ap.add_argument("--shapes", default="c1,c3")
    ap.add_argument("--precisions", default="tf32x3,fp32,tf32,fp16")
    args = ap.parse_args()
    rows = []
    for shape in args.shapes.split(","):
        for prec in args.precisions.split(","):
            r = compare(shape, prec)
            rows.append(r)
            print(json.dumps({k: v for k, v in r.items() if k != "params_after.max_norm_per_step"}), flush=True)
    floors = []
    if "tf32" in args.precisions.split(",") or "fp16" in args.precisions.split(","):
        for shape in args.shapes.split(","):
            floors.append(tf32_noise_floor(shape))
            print(json.dumps(floors[-1]), flush=True)
    with open(args.out, "w") as fh:
        json.dump(dict(note="max-abs error / max-abs reference value against oracle/ppo_oracle.py (numpy fp32); one learn() "
                            "iteration with shared noise and permutation; shapes: " + json.dumps(SHAPES),
                       rows=rows, tf32_class_noise_floor=floors), fh, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
