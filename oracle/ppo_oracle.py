"""TEST INFRASTRUCTURE -- CPU oracle for the rsl_rl PPO learning iteration (numpy, fp32).

This file is a from-scratch *restatement* of the reference algorithm, used only as the checker in
`tests/`, in `__graft_entry__.smoke()` and as `bench.py`'s CPU baseline leg.  It is never imported by
the product package `rsl_rl_b200`.  All `file:line` citations are relative to the reference checkout
(kaixi287/rsl_rl v3.1.0).

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference itself, produced in the build container by `oracle/make_golden.py`
(unmodified reference + `oracle/_shim/tensordict`) and committed under `tests/golden/`;
`tests/test_oracle_vs_golden.py` is the check.

Arithmetic that the reference delegates to PyTorch (unpinned `torch>=2.6`, here 2.11.0) is restated
from PyTorch's published semantics:
  * nn.Linear / nn.ELU etc.           -> `mlp_forward` / `mlp_backward`
  * torch.distributions.Normal         -> `gaussian_*`   (torch/distributions/normal.py: sample =
                                          normal_(0,1)*std+mean, log_prob, entropy)
  * autograd of maximum / clamp / pow  -> `ppo_loss_fwd_bwd` (tie rule: grad/2 to both sides)
  * nn.utils.clip_grad_norm_           -> `clip_grad_norm`  (coef = max_norm/(norm+1e-6), clamped to 1)
  * torch.optim.Adam (defaults)        -> `adam_step`
  * split / pad_sequence               -> `split_and_pad_trajectories`

Randomness (action noise z ~ N(0,1) and the minibatch permutation) is an INPUT of every function
here; the reference draws it from torch generators whose streams are device specific.
"""

from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------------
# activations  (rsl_rl/utils/utils.py:18-49 resolve_nn_activation -> torch.nn modules)
# --------------------------------------------------------------------------------------------------

_SELU_ALPHA = 1.6732632423543772848170429916717
_SELU_SCALE = 1.0507009873554804934193349852946


def act_forward(name: str, z: np.ndarray) -> np.ndarray:
    name = name.lower()
    if name in ("elu", "crelu"):  # nn.ELU(alpha=1) and nn.CELU(alpha=1) coincide
        return np.where(z > 0, z, np.expm1(np.minimum(z, 0))).astype(F32)
    if name == "selu":
        return (F32(_SELU_SCALE) * np.where(z > 0, z, F32(_SELU_ALPHA) * np.expm1(np.minimum(z, 0)))).astype(F32)
    if name == "relu":
        return np.maximum(z, 0).astype(F32)
    if name == "lrelu":
        return np.where(z > 0, z, F32(0.01) * z).astype(F32)
    if name == "tanh":
        return np.tanh(z).astype(F32)
    if name == "sigmoid":
        return (1.0 / (1.0 + np.exp(-z))).astype(F32)
    if name == "softplus":  # beta=1, threshold=20
        return np.where(z > 20, z, np.log1p(np.exp(np.minimum(z, 20)))).astype(F32)
    if name == "identity":
        return z
    raise ValueError(f"oracle: unsupported activation {name!r}")


def act_backward(name: str, z: np.ndarray, h: np.ndarray, dh: np.ndarray) -> np.ndarray:
    """dL/dz from dL/dh; z = pre-activation, h = act(z)."""
    name = name.lower()
    if name in ("elu", "crelu"):
        return np.where(z > 0, dh, dh * np.exp(np.minimum(z, 0))).astype(F32)
    if name == "selu":
        return np.where(z > 0, F32(_SELU_SCALE) * dh,
                        dh * F32(_SELU_SCALE * _SELU_ALPHA) * np.exp(np.minimum(z, 0))).astype(F32)
    if name == "relu":
        return np.where(z > 0, dh, 0).astype(F32)
    if name == "lrelu":
        return np.where(z > 0, dh, F32(0.01) * dh).astype(F32)
    if name == "tanh":
        return (dh * (1 - h * h)).astype(F32)
    if name == "sigmoid":
        return (dh * h * (1 - h)).astype(F32)
    if name == "softplus":
        return np.where(z > 20, dh, dh / (1.0 + np.exp(-np.minimum(z, 20)))).astype(F32)
    if name == "identity":
        return dh
    raise ValueError(f"oracle: unsupported activation {name!r}")


# --------------------------------------------------------------------------------------------------
# MLP  (rsl_rl/networks/mlp.py:58-83 construction, :106-114 forward)
# --------------------------------------------------------------------------------------------------

def mlp_layers(params: dict, prefix: str):
    """[(W[out,in], b[out]), ...] in nn.Sequential order; Linear layers sit at even indices
    (rsl_rl/networks/mlp.py:60-69,81-83)."""
    out = []
    i = 0
    while f"{prefix}.{i}.weight" in params:
        out.append((params[f"{prefix}.{i}.weight"], params[f"{prefix}.{i}.bias"]))
        i += 2
    return out


# Optional emulation of tensor-core operand rounding (tests only): with GEMM_ROUND = "tf32" every matmul operand is
# rounded to tf32 (10-bit mantissa, nearest, ties away) before an fp32-accumulated product -- the arithmetic class of
# cuBLAS TF32 / tcgen05 kind::tf32.  It measures the NOISE FLOOR of that class against the fp32 oracle (e.g. how far one
# Adam step moves parameters whose gradient sign is decided by rounding noise); it is not a model of any kernel.
GEMM_ROUND = None


def _r(a: np.ndarray) -> np.ndarray:
    if GEMM_ROUND is None:
        return a
    if GEMM_ROUND == "tf32":
        u = np.ascontiguousarray(a, dtype=F32).view(np.uint32)
        return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(F32)
    raise ValueError(GEMM_ROUND)


def mlp_forward(layers, x: np.ndarray, activation: str = "elu"):
    """Returns (y, cache) where cache holds every layer input / pre-activation / output."""
    cache = []
    h = x.astype(F32, copy=False)
    n = len(layers)
    for li, (W, b) in enumerate(layers):
        z = _r(h) @ _r(W).T + b
        if li < n - 1:
            hn = act_forward(activation, z)
        else:
            hn = z
        cache.append((h, z, hn))
        h = hn
    return h, cache


def mlp_backward(layers, cache, dy: np.ndarray, activation: str = "elu", need_dx: bool = False):
    """Returns ([(dW, db), ...], dx or None)."""
    grads = [None] * len(layers)
    n = len(layers)
    d = dy.astype(F32, copy=False)
    for li in range(n - 1, -1, -1):
        x_in, z, h = cache[li]
        if li < n - 1:
            d = act_backward(activation, z, h, d)
        W, _ = layers[li]
        grads[li] = ((_r(d).T @ _r(x_in)).astype(F32), d.sum(axis=0, dtype=np.float64).astype(F32))
        if li > 0 or need_dx:
            d = (_r(d) @ _r(W)).astype(F32)
    return grads, (d if need_dx else None)


# --------------------------------------------------------------------------------------------------
# Gaussian policy head  (rsl_rl/modules/actor_critic.py:106-146,170-171; torch/distributions/normal.py)
# --------------------------------------------------------------------------------------------------

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
_ENTROPY_CONST = 0.5 + 0.5 * math.log(2 * math.pi)


def gaussian_sample(mu, sigma, z):
    """Normal.sample(): normal_(0,1) then mul_(std) then add_(mean) -- two roundings, no FMA."""
    return (z.astype(F32) * sigma.astype(F32)).astype(F32) + mu.astype(F32)


def gaussian_log_prob(mu, sigma, a):
    var = sigma * sigma
    lp = -((a - mu) ** 2) / (2 * var) - np.log(sigma) - F32(_LOG_SQRT_2PI)
    return lp.sum(axis=-1, dtype=F32)


def gaussian_entropy(sigma):
    return (F32(_ENTROPY_CONST) + np.log(sigma)).sum(axis=-1, dtype=F32)


def policy_sigma(params, mu, noise_std_type="scalar"):
    """actor_critic.py:133-138 -- sigma broadcast to mu's shape."""
    if noise_std_type == "scalar":
        return np.broadcast_to(params["std"], mu.shape)
    if noise_std_type == "log":
        return np.broadcast_to(np.exp(params["log_std"]), mu.shape)
    raise ValueError(noise_std_type)


# --------------------------------------------------------------------------------------------------
# Rollout side  (rsl_rl/algorithms/ppo.py:129-169)
# --------------------------------------------------------------------------------------------------

def act(params, obs_actor, obs_critic, z, activation="elu", noise_std_type="scalar"):
    """PPO.act (ppo.py:129-140): returns actions, values[N,1], log_prob[N], mu, sigma."""
    mu, _ = mlp_forward(mlp_layers(params, "actor"), obs_actor, activation)
    sigma = policy_sigma(params, mu, noise_std_type)
    a = gaussian_sample(mu, sigma, z)
    v, _ = mlp_forward(mlp_layers(params, "critic"), obs_critic, activation)
    lp = gaussian_log_prob(mu, sigma, a)
    return a, v, lp, mu, np.ascontiguousarray(sigma)


def bootstrap_rewards(rewards, values, time_outs, gamma, intrinsic=None):
    """PPO.process_env_step (ppo.py:150-164): r (+ r_int) + gamma * squeeze(V * time_outs)."""
    r = rewards.astype(F32).copy()
    if intrinsic is not None:
        r = r + intrinsic.astype(F32)
    if time_outs is not None:
        r = r + F32(gamma) * (values.reshape(-1).astype(F32) * time_outs.astype(F32))
    return r.astype(F32)


# --------------------------------------------------------------------------------------------------
# GAE  (rsl_rl/storage/rollout_storage.py:127-149)
# --------------------------------------------------------------------------------------------------

def compute_returns(rewards, values, dones, last_values, gamma, lam, normalize_advantage=True):
    """rewards/values/dones: [T,N,1]; last_values [N,1].  Returns (returns, advantages) [T,N,1] fp32.

    The scan reproduces the reference's fp32 operation order exactly
    ((nnt*gamma)*next_v, ((nnt*gamma)*lam)*adv); mean / unbiased std are taken in fp64.
    """
    T = rewards.shape[0]
    rewards = rewards.astype(F32, copy=False)
    values = values.astype(F32, copy=False)
    returns = np.zeros_like(values)
    adv = np.zeros_like(values[0])
    g, l = F32(gamma), F32(lam)
    for t in range(T - 1, -1, -1):
        next_values = last_values.astype(F32) if t == T - 1 else values[t + 1]
        nnt = F32(1.0) - dones[t].astype(F32)
        delta = rewards[t] + nnt * g * next_values - values[t]
        adv = delta + nnt * g * l * adv
        returns[t] = adv + values[t]
    advantages = returns - values
    if normalize_advantage:
        a64 = advantages.astype(np.float64)
        mean = a64.mean()
        std = a64.std(ddof=1)
        advantages = ((advantages - F32(mean)) / (F32(std) + F32(1e-8))).astype(F32)
    return returns, advantages


# --------------------------------------------------------------------------------------------------
# Minibatch gather  (rsl_rl/storage/rollout_storage.py:160-203)
# --------------------------------------------------------------------------------------------------

def mini_batch_slices(perm: np.ndarray, num_mini_batches: int):
    """indices[start:end] for each minibatch; perm has num_mini_batches*mini_batch_size entries."""
    mb = perm.shape[0] // num_mini_batches
    return [perm[i * mb:(i + 1) * mb] for i in range(num_mini_batches)]


def gather_minibatch(fields: dict, idx: np.ndarray) -> dict:
    """fields: name -> [T,N,d]; flatten(0,1) then row gather (index = t*N + n)."""
    return {k: v.reshape(-1, v.shape[-1])[idx] for k, v in fields.items()}


# --------------------------------------------------------------------------------------------------
# PPO loss forward + hand-derived backward  (rsl_rl/algorithms/ppo.py:260-315)
# --------------------------------------------------------------------------------------------------

def kl_mean(mu, sigma, old_mu, old_sigma):
    """ppo.py:262-269."""
    kl = (np.log(sigma / old_sigma + F32(1.0e-5))
          + (old_sigma ** 2 + (old_mu - mu) ** 2) / (F32(2.0) * sigma ** 2) - F32(0.5)).sum(axis=-1)
    return F32(kl.astype(np.float64).mean())


def adaptive_lr(lr: float, kl: float, desired_kl: float) -> float:
    """ppo.py:280-284; the comparison is done in fp32 (0-dim fp32 tensor vs python scalar)."""
    kl = F32(kl)
    if kl > F32(desired_kl * 2.0):
        return max(1e-5, lr / 1.5)
    if kl < F32(desired_kl / 2.0) and kl > F32(0.0):
        return min(1e-2, lr * 1.5)
    return lr


def ppo_loss_fwd_bwd(mu, sigma, value, actions, old_logp, advantages, returns, old_values,
                     clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01,
                     use_clipped_value_loss=True, n_entropy=None):
    """Forward value of the PPO objective and its gradient w.r.t. (mu, sigma, value).

    mu/sigma/actions [mb,A]; value/old_values/returns/advantages [mb,1] (or [mb]); old_logp [mb,1]|[mb].
    `n_entropy`: number of leading rows that contribute entropy (symmetry augmentation keeps only
    the original block, ppo.py:255-257); defaults to all rows.
    Returns dict(surrogate, value_loss, entropy, loss, dmu, dsigma, dvalue).
    """
    mb = mu.shape[0]
    n_ent = mb if n_entropy is None else n_entropy
    adv = advantages.reshape(-1).astype(F32)
    ret = returns.reshape(-1).astype(F32)
    v_old = old_values.reshape(-1).astype(F32)
    v = value.reshape(-1).astype(F32)
    lp_old = old_logp.reshape(-1).astype(F32)
    eps = F32(clip_param)

    logp = gaussian_log_prob(mu, sigma, actions)
    ratio = np.exp(logp - lp_old)
    s1 = -adv * ratio
    clamped = np.clip(ratio, F32(1.0) - eps, F32(1.0) + eps)
    s2 = -adv * clamped
    surrogate = np.maximum(s1, s2)
    surrogate_loss = surrogate.astype(np.float64).mean()

    if use_clipped_value_loss:
        dv = v - v_old
        v_clip = v_old + np.clip(dv, -eps, eps)
        l1 = (v - ret) ** 2
        l2 = (v_clip - ret) ** 2
        value_loss = np.maximum(l1, l2).astype(np.float64).mean()
    else:
        value_loss = ((ret - v) ** 2).astype(np.float64).mean()

    ent = gaussian_entropy(sigma[:n_ent])
    entropy = ent.astype(np.float64).mean()
    loss = surrogate_loss + value_loss_coef * value_loss - entropy_coef * entropy

    # ---- backward -------------------------------------------------------------------------------
    g = F32(1.0 / mb)
    # maximum: grad -> the larger side, split in halves on ties (aten derivatives.yaml: maximum)
    g1 = np.where(s1 > s2, g, np.where(s1 == s2, g * F32(0.5), F32(0))).astype(F32)
    g2 = np.where(s2 > s1, g, np.where(s1 == s2, g * F32(0.5), F32(0))).astype(F32)
    in_range = ((ratio >= F32(1.0) - eps) & (ratio <= F32(1.0) + eps)).astype(F32)  # clamp backward mask
    g_ratio = g1 * (-adv) + g2 * (-adv) * in_range
    g_logp = (g_ratio * ratio).astype(F32)[:, None]
    diff = actions - mu
    var = sigma * sigma
    dmu = (g_logp * diff / var).astype(F32)
    dsigma = (g_logp * (diff * diff / (var * sigma) - F32(1.0) / sigma)).astype(F32)
    # entropy bonus: -c_e * mean_i sum_a log sigma
    dsigma[:n_ent] += (F32(-entropy_coef / n_ent) / sigma[:n_ent]).astype(F32)

    if use_clipped_value_loss:
        ga = np.where(l1 > l2, g, np.where(l1 == l2, g * F32(0.5), F32(0))).astype(F32)
        gb = np.where(l2 > l1, g, np.where(l1 == l2, g * F32(0.5), F32(0))).astype(F32)
        in_v = ((dv >= -eps) & (dv <= eps)).astype(F32)
        dvalue = F32(value_loss_coef) * (ga * F32(2.0) * (v - ret) + gb * F32(2.0) * (v_clip - ret) * in_v)
    else:
        dvalue = F32(value_loss_coef) * g * F32(-2.0) * (ret - v)

    return dict(surrogate=float(surrogate_loss), value_loss=float(value_loss), entropy=float(entropy),
                loss=float(loss), dmu=dmu, dsigma=dsigma, dvalue=dvalue.astype(F32).reshape(-1, 1),
                logp=logp, ratio=ratio)


# --------------------------------------------------------------------------------------------------
# Gradient clipping + Adam  (ppo.py:380-381; torch/nn/utils/clip_grad.py; torch/optim/adam.py)
# --------------------------------------------------------------------------------------------------

def clip_grad_norm(grads: "OrderedDict[str, np.ndarray]", max_norm: float):
    """In place.  Returns the total norm (before clipping)."""
    norms = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()) for g in grads.values()])
    total = float(np.sqrt((norms ** 2).sum()))
    coef = min(max_norm / (total + 1e-6), 1.0)
    for k in grads:
        grads[k] = (grads[k] * F32(coef)).astype(F32)
    return total


def adam_init(params):
    return dict(step=0,
                exp_avg=OrderedDict((k, np.zeros_like(v)) for k, v in params.items()),
                exp_avg_sq=OrderedDict((k, np.zeros_like(v)) for k, v in params.items()))


def adam_step(params, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (weight_decay=0, amsgrad=False, maximize=False); in place."""
    state["step"] += 1
    t = state["step"]
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, p in params.items():
        g = grads[k]
        m = state["exp_avg"][k]
        v = state["exp_avg_sq"][k]
        m += (g - m) * F32(1 - beta1)  # lerp_
        v *= F32(beta2)
        v += F32(1 - beta2) * g * g
        denom = np.sqrt(v) / F32(bc2_sqrt) + F32(eps)
        p += F32(-step_size) * (m / denom)


# --------------------------------------------------------------------------------------------------
# One minibatch step and the whole update()  (rsl_rl/algorithms/ppo.py:178-422)
# --------------------------------------------------------------------------------------------------

def minibatch_grads(params, batch, cfg):
    """Forward + loss + backward on one minibatch.  batch keys: obs_actor, obs_critic, actions,
    values, advantages, returns, old_logp, old_mu, old_sigma.  Returns (grads OrderedDict in
    parameter order, info dict)."""
    activation = cfg.get("activation", "elu")
    nst = cfg.get("noise_std_type", "scalar")
    actor, critic = mlp_layers(params, "actor"), mlp_layers(params, "critic")
    mu, cache_a = mlp_forward(actor, batch["obs_actor"], activation)
    sigma = np.ascontiguousarray(policy_sigma(params, mu, nst))
    value, cache_c = mlp_forward(critic, batch["obs_critic"], activation)
    kl = kl_mean(mu, sigma, batch["old_mu"], batch["old_sigma"])
    out = ppo_loss_fwd_bwd(mu, sigma, value, batch["actions"], batch["old_logp"], batch["advantages"],
                           batch["returns"], batch["values"], cfg["clip_param"], cfg["value_loss_coef"],
                           cfg["entropy_coef"], cfg["use_clipped_value_loss"])
    ga, _ = mlp_backward(actor, cache_a, out["dmu"], activation)
    gc, _ = mlp_backward(critic, cache_c, out["dvalue"], activation)
    grads = OrderedDict()
    for k in params:
        if k == "std":
            grads[k] = out["dsigma"].sum(axis=0, dtype=np.float64).astype(F32)
        elif k == "log_std":
            grads[k] = (out["dsigma"] * sigma).sum(axis=0, dtype=np.float64).astype(F32)
        else:
            net, idx, kind = k.split(".")
            gw, gb = (ga if net == "actor" else gc)[int(idx) // 2]
            grads[k] = gw if kind == "weight" else gb
    info = dict(kl=float(kl), surrogate=out["surrogate"], value_loss=out["value_loss"],
                entropy=out["entropy"], mu=mu, value=value)
    return grads, info


def update(params, adam_state, lr, storage, perm, cfg, world_hook=None):
    """PPO.update for the feed-forward policy.  `storage`: dict with obs_actor/obs_critic/actions/
    values/returns/advantages/actions_log_prob/mu/sigma, each [T,N,d].  `perm`: int64 [M*mb].
    Mutates params/adam_state in place.  Returns (loss_dict, lr, trace) where trace is the per
    minibatch record list.  `world_hook(grads, kl) -> (grads, kl)` emulates the multi-GPU averaging
    (ppo.py:272-274, 441-469)."""
    M, E = cfg["num_mini_batches"], cfg["num_learning_epochs"]
    fields = dict(obs_actor=storage["obs_actor"], obs_critic=storage["obs_critic"],
                  actions=storage["actions"], values=storage["values"], returns=storage["returns"],
                  advantages=storage["advantages"], old_logp=storage["actions_log_prob"],
                  old_mu=storage["mu"], old_sigma=storage["sigma"])
    slices = mini_batch_slices(perm, M)
    sums = dict(value_function=0.0, surrogate=0.0, entropy=0.0)
    trace = []
    for _ in range(E):
        for idx in slices:
            batch = gather_minibatch(fields, idx)
            grads, info = minibatch_grads(params, batch, cfg)
            kl = info["kl"]
            if world_hook is not None:
                grads, kl = world_hook(grads, kl)
            if cfg.get("desired_kl") is not None and cfg.get("schedule") == "adaptive":
                lr = adaptive_lr(lr, kl, cfg["desired_kl"])
            raw = np.concatenate([g.reshape(-1) for g in grads.values()])
            total_norm = clip_grad_norm(grads, cfg["max_grad_norm"])
            adam_step(params, grads, adam_state, lr)
            sums["value_function"] += info["value_loss"]
            sums["surrogate"] += info["surrogate"]
            sums["entropy"] += info["entropy"]
            trace.append(dict(kl=kl, lr=lr, total_norm=total_norm, grads=raw,
                              surrogate=info["surrogate"], value_loss=info["value_loss"],
                              entropy=info["entropy"],
                              params_after=np.concatenate([v.reshape(-1) for v in params.values()])))
    n = E * M
    return {k: v / n for k, v in sums.items()}, lr, trace


def rollout(params, obs_actor_seq, obs_critic_seq, rewards, dones, time_outs, z, cfg):
    """T env steps of PPO.act + process_env_step against a pre-recorded env stream.
    obs_*_seq: [T+1,N,O] (row t is what the policy sees at step t, row T feeds compute_returns).
    Returns the filled storage dict (all [T,N,d]) and last_values [N,1]."""
    T, N = rewards.shape[:2]
    A = z.shape[-1]
    st = dict(obs_actor=obs_actor_seq[:T].astype(F32), obs_critic=obs_critic_seq[:T].astype(F32),
              actions=np.zeros((T, N, A), F32), values=np.zeros((T, N, 1), F32),
              actions_log_prob=np.zeros((T, N, 1), F32), mu=np.zeros((T, N, A), F32),
              sigma=np.zeros((T, N, A), F32), rewards=np.zeros((T, N, 1), F32),
              dones=np.zeros((T, N, 1), np.uint8))
    act_name, nst = cfg.get("activation", "elu"), cfg.get("noise_std_type", "scalar")
    for t in range(T):
        a, v, lp, mu, sigma = act(params, obs_actor_seq[t], obs_critic_seq[t], z[t], act_name, nst)
        st["actions"][t], st["values"][t], st["actions_log_prob"][t, :, 0] = a, v, lp
        st["mu"][t], st["sigma"][t] = mu, sigma
        to = None if time_outs is None else time_outs[t]
        st["rewards"][t, :, 0] = bootstrap_rewards(rewards[t], v, to, cfg["gamma"])
        st["dones"][t, :, 0] = dones[t].astype(np.uint8)
    last_values, _ = mlp_forward(mlp_layers(params, "critic"), obs_critic_seq[T], act_name)
    return st, last_values


def learn_iteration(params, adam_state, lr, stream, z, perm, cfg):
    """One OnPolicyRunner.learn iteration (on_policy_runner.py:97-149) on a recorded env stream."""
    st, last_values = rollout(params, stream["obs_actor"], stream["obs_critic"], stream["rewards"],
                              stream["dones"], stream.get("time_outs"), z, cfg)
    st["returns"], st["advantages"] = compute_returns(
        st["rewards"], st["values"], st["dones"], last_values, cfg["gamma"], cfg["lam"],
        normalize_advantage=not cfg.get("normalize_advantage_per_mini_batch", False))
    loss_dict, lr, trace = update(params, adam_state, lr, st, perm, cfg)
    return st, last_values, loss_dict, lr, trace


# --------------------------------------------------------------------------------------------------
# Trajectory split / pad / unpad  (rsl_rl/utils/utils.py:78-141; rollout_storage.py:218-250)
# --------------------------------------------------------------------------------------------------

def split_and_pad_trajectories(tensor: np.ndarray, dones: np.ndarray):
    """tensor [T,N,...], dones [T,N,1] -> (padded [T,n_traj,...], masks bool [T,n_traj]).

    Trajectories are enumerated env-major (env 0's pieces first); each piece ends at a done (the
    last step of every env is forced done, utils.py:97-98)."""
    T, N = dones.shape[:2]
    d = dones.reshape(T, N).astype(bool).copy()
    d[-1] = True
    lengths, starts = [], []
    for n in range(N):
        s = 0
        for t in range(T):
            if d[t, n]:
                lengths.append(t - s + 1)
                starts.append((n, s))
                s = t + 1
    n_traj = len(lengths)
    padded = np.zeros((T, n_traj) + tensor.shape[2:], tensor.dtype)
    masks = np.zeros((T, n_traj), bool)
    for j, ((n, s), L) in enumerate(zip(starts, lengths)):
        padded[:L, j] = tensor[s:s + L, n]
        masks[:L, j] = True
    return padded, masks


def unpad_trajectories(trajectories: np.ndarray, masks: np.ndarray):
    """utils.py:134-141: [T,n_traj,H] -> [T,N,H]."""
    T = trajectories.shape[0]
    flat = trajectories.transpose(1, 0, 2)[masks.T]
    return flat.reshape(-1, T, trajectories.shape[-1]).transpose(1, 0, 2)


def last_was_done(dones: np.ndarray):
    """rollout_storage.py:218-221: [T,N] bool, row 0 True, rows 1.. = dones[:-1]."""
    d = dones.reshape(dones.shape[0], dones.shape[1]).astype(bool)
    out = np.zeros_like(d)
    out[1:] = d[:-1]
    out[0] = True
    return out


def hidden_at_trajectory_starts(saved: np.ndarray, dones: np.ndarray):
    """rollout_storage.py:238-244: saved [T,L,N,H] -> [L, n_traj, H], env-major trajectory order."""
    lwd = last_was_done(dones)  # [T,N]
    x = saved.transpose(2, 0, 1, 3)  # [N,T,L,H]
    return x[lwd.T].transpose(1, 0, 2)
