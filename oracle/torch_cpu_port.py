"""TEST / BENCH INFRASTRUCTURE -- the oracle's learn() iteration restated on torch CPU tensors.

`oracle/ppo_oracle.py` (numpy) is the checker.  Its elementwise parts are single threaded, which makes
it ~6x slower than the reference's own CPU path (torch ATen kernels: MKL sgemm + vectorised,
multi-threaded elementwise ops) and would overstate any GPU/CPU ratio.  This module restates the SAME
hand-derived algorithm (no autograd) with torch CPU ops so that `bench.py`'s CPU baseline and
`--impl reference` arm time the arithmetic library the reference really runs on.  It is pinned to the
numpy oracle by `tests/test_oracle_vs_golden.py::test_torch_cpu_port_matches_numpy_oracle`.
Never imported by the product package.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
_ENTROPY_CONST = 0.5 + 0.5 * math.log(2 * math.pi)


def _layers(params, prefix):
    out, i = [], 0
    while f"{prefix}.{i}.weight" in params:
        out.append((params[f"{prefix}.{i}.weight"], params[f"{prefix}.{i}.bias"]))
        i += 2
    return out


def mlp_forward(layers, x):
    """ELU MLP (reference rsl_rl/networks/mlp.py:106-114); cache = inputs and outputs of every layer."""
    cache, h, n = [], x, len(layers)
    for li, (W, b) in enumerate(layers):
        z = torch.addmm(b, h, W.t())
        hn = torch.nn.functional.elu(z) if li < n - 1 else z
        cache.append((h, hn))
        h = hn
    return h, cache


def mlp_backward(layers, cache, dy):
    grads, d, n = [None] * len(layers), dy, len(layers)
    for li in range(n - 1, -1, -1):
        x_in, h = cache[li]
        if li < n - 1:
            d = torch.ops.aten.elu_backward(d, 1.0, 1.0, 1.0, True, h)  # ELU' from the output (ATen's own kernel)
        grads[li] = (d.t() @ x_in, d.sum(0))
        if li > 0:
            d = d @ layers[li][0]
    return grads


def log_prob(mu, sigma, a):
    return (-((a - mu) ** 2) / (2 * sigma * sigma) - torch.log(sigma) - _LOG_SQRT_2PI).sum(-1)


def compute_returns(rewards, values, dones, last_values, gamma, lam):
    """rollout_storage.py:127-149 on [T,N,1] tensors."""
    T = rewards.shape[0]
    returns = torch.zeros_like(values)
    adv = torch.zeros_like(values[0])
    for t in range(T - 1, -1, -1):
        nv = last_values if t == T - 1 else values[t + 1]
        nnt = 1.0 - dones[t].float()
        delta = rewards[t] + nnt * gamma * nv - values[t]
        adv = delta + nnt * gamma * lam * adv
        returns[t] = adv + values[t]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


def loss_fwd_bwd(mu, sigma, value, actions, old_logp, adv, ret, v_old, clip, c_v, c_e):
    """ppo.py:297-315 forward and hand-derived backward (see ppo_oracle.ppo_loss_fwd_bwd)."""
    mb = mu.shape[0]
    lp = log_prob(mu, sigma, actions)
    ratio = torch.exp(lp - old_logp)
    s1 = -adv * ratio
    s2 = -adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)
    surrogate = torch.maximum(s1, s2).mean()
    dv = value - v_old
    v_clip = v_old + dv.clamp(-clip, clip)
    l1, l2 = (value - ret) ** 2, (v_clip - ret) ** 2
    value_loss = torch.maximum(l1, l2).mean()
    entropy = (_ENTROPY_CONST + torch.log(sigma)).sum(-1).mean()
    g = 1.0 / mb
    tie = s1 == s2
    g1 = torch.where(s1 > s2, g, 0.0) + tie * (0.5 * g)
    g2 = torch.where(s2 > s1, g, 0.0) + tie * (0.5 * g)
    in_range = ((ratio >= 1.0 - clip) & (ratio <= 1.0 + clip)).float()
    g_lp = ((g1 * (-adv) + g2 * (-adv) * in_range) * ratio).unsqueeze(1)
    diff, var = actions - mu, sigma * sigma
    dmu = g_lp * diff / var
    dsigma = g_lp * (diff * diff / (var * sigma) - 1.0 / sigma) - (c_e / mb) / sigma
    tie_v = l1 == l2
    ga = torch.where(l1 > l2, g, 0.0) + tie_v * (0.5 * g)
    gb = torch.where(l2 > l1, g, 0.0) + tie_v * (0.5 * g)
    in_v = ((dv >= -clip) & (dv <= clip)).float()
    dvalue = c_v * (ga * 2.0 * (value - ret) + gb * 2.0 * (v_clip - ret) * in_v)
    return surrogate, value_loss, entropy, dmu, dsigma, dvalue


def learn_iteration(params, adam, lr, stream, z, perm, cfg):
    """One OnPolicyRunner.learn iteration on a recorded env stream; tensors are torch CPU fp32.
    params: OrderedDict name -> tensor (mutated); adam: dict(step, m, v); returns (loss_dict, lr)."""
    gamma, lam, clip = cfg["gamma"], cfg["lam"], cfg["clip_param"]
    obs, rewards, dones, time_outs = stream["obs"], stream["rewards"], stream["dones"], stream["time_outs"]
    T, N = rewards.shape
    A = z.shape[-1]
    actor, critic = _layers(params, "actor"), _layers(params, "critic")
    f = lambda *s: torch.zeros(*s)
    st = dict(actions=f(T, N, A), values=f(T, N, 1), logp=f(T, N, 1), mu=f(T, N, A), sigma=f(T, N, A),
              rewards=f(T, N, 1), dones=torch.zeros(T, N, 1, dtype=torch.uint8))
    for t in range(T):  # PPO.act + process_env_step (ppo.py:129-169)
        mu, _ = mlp_forward(actor, obs[t])
        sigma = params["std"].expand_as(mu)
        a = z[t] * sigma + mu
        v, _ = mlp_forward(critic, obs[t])
        st["actions"][t], st["values"][t], st["mu"][t], st["sigma"][t] = a, v, mu, sigma
        st["logp"][t, :, 0] = log_prob(mu, sigma, a)
        st["rewards"][t, :, 0] = rewards[t] + gamma * (v.squeeze(1) * time_outs[t])
        st["dones"][t, :, 0] = dones[t]
    last_values, _ = mlp_forward(critic, obs[T])
    returns, advantages = compute_returns(st["rewards"], st["values"], st["dones"], last_values, gamma, lam)

    M, E = cfg["num_mini_batches"], cfg["num_learning_epochs"]
    flat = lambda x: x.flatten(0, 1)
    fields = [flat(obs[:T]), flat(st["actions"]), flat(st["values"]), flat(advantages), flat(returns), flat(st["logp"]),
              flat(st["mu"]), flat(st["sigma"])]
    mb = perm.shape[0] // M
    sums = [0.0, 0.0, 0.0]
    b1, b2, eps = 0.9, 0.999, 1e-8
    names = list(params.keys())
    for _ in range(E):
        for i in range(M):
            idx = perm[i * mb:(i + 1) * mb]
            ob, actions, v_old, adv, ret, old_logp, old_mu, old_sigma = [x[idx] for x in fields]
            mu, cache_a = mlp_forward(actor, ob)
            sigma = params["std"].expand_as(mu)
            value, cache_c = mlp_forward(critic, ob)
            kl = (torch.log(sigma / old_sigma + 1e-5) + (old_sigma ** 2 + (old_mu - mu) ** 2) / (2.0 * sigma ** 2)
                  - 0.5).sum(-1).mean()
            if cfg.get("desired_kl") is not None and cfg.get("schedule") == "adaptive":
                if kl > cfg["desired_kl"] * 2.0:
                    lr = max(1e-5, lr / 1.5)
                elif kl < cfg["desired_kl"] / 2.0 and kl > 0.0:
                    lr = min(1e-2, lr * 1.5)
            s, vl, ent, dmu, dsigma, dvalue = loss_fwd_bwd(
                mu, sigma, value.squeeze(1), actions, old_logp.squeeze(1), adv.squeeze(1), ret.squeeze(1),
                v_old.squeeze(1), clip, cfg["value_loss_coef"], cfg["entropy_coef"])
            ga, gc = mlp_backward(actor, cache_a, dmu), mlp_backward(critic, cache_c, dvalue.unsqueeze(1))
            grads = OrderedDict()
            for k in names:
                if k == "std":
                    grads[k] = dsigma.sum(0)
                else:
                    net, li, kind = k.split(".")
                    gw, gb = (ga if net == "actor" else gc)[int(li) // 2]
                    grads[k] = gw if kind == "weight" else gb
            total = math.sqrt(sum(float((g_ * g_).sum()) for g_ in grads.values()))
            coef = min(cfg["max_grad_norm"] / (total + 1e-6), 1.0)
            adam["step"] += 1
            t_ = adam["step"]
            step_size, bc2_sqrt = lr / (1 - b1 ** t_), math.sqrt(1 - b2 ** t_)
            for k in names:
                g_ = grads[k] * coef
                m, v = adam["m"][k], adam["v"][k]
                m.lerp_(g_, 1 - b1)
                v.mul_(b2).addcmul_(g_, g_, value=1 - b2)
                params[k].addcdiv_(m, (v.sqrt() / bc2_sqrt).add_(eps), value=-step_size)
            sums[0] += float(vl); sums[1] += float(s); sums[2] += float(ent)
    n = E * M
    return dict(value_function=sums[0] / n, surrogate=sums[1] / n, entropy=sums[2] / n), lr, dict(returns=returns,
                                                                                                 advantages=advantages)


def adam_init(params):
    return dict(step=0, m=OrderedDict((k, torch.zeros_like(v)) for k, v in params.items()),
                v=OrderedDict((k, torch.zeros_like(v)) for k, v in params.items()))


def random_problem(N, T=24, obs_dim=48, A=12, hidden=(512, 256, 128), seed=0):
    """Synthetic rollout stream + torch-default-initialised parameters of the named architecture."""
    g = torch.Generator().manual_seed(seed)
    params = OrderedDict(std=torch.ones(A))
    for net, out in (("actor", A), ("critic", 1)):
        dims = [obs_dim, *hidden, out]
        for i in range(len(dims) - 1):
            k = 1.0 / math.sqrt(dims[i])
            params[f"{net}.{2 * i}.weight"] = (torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * k
            params[f"{net}.{2 * i}.bias"] = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * k
    dones = (torch.rand(T, N, generator=g) < 0.02).to(torch.uint8)
    stream = dict(obs=torch.randn(T + 1, N, obs_dim, generator=g), rewards=torch.randn(T, N, generator=g), dones=dones,
                  time_outs=((torch.rand(T, N, generator=g) < 0.5) & (dones > 0)).float())
    z = torch.randn(T, N, A, generator=g)
    perm = torch.randperm(N * T, generator=g)
    return params, stream, z, perm


DEFAULT_CFG = dict(learning_rate=1e-3, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive",
                   value_loss_coef=1.0, clip_param=0.2, use_clipped_value_loss=True, desired_kl=0.01, entropy_coef=0.01,
                   gamma=0.99, lam=0.95, max_grad_norm=1.0)
