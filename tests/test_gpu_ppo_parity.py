"""GPU: whole learn() iterations of the CUDA path against (a) the golden traces recorded from the
unmodified reference and (b) the numpy oracle on seeded synthetic rollouts at benchmark sizes."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from tests.golden_util import cfg_of, flat, load, meta_of, params_from, rel_err, stream_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _train_cfg(meta, **alg_extra):
    alg = dict(meta["alg"], class_name="PPO", **alg_extra)
    return dict(num_steps_per_env=meta["num_steps"], save_interval=10 ** 9,
                obs_groups={k: list(v) for k, v in meta["obs_groups"].items()},
                policy=dict(class_name="ActorCritic", actor_hidden_dims=list(meta["hidden"]),
                            critic_hidden_dims=list(meta["hidden"]), activation=meta["activation"],
                            init_noise_std=1.0, noise_std_type=meta["noise_std_type"], **meta.get("policy_extra", {})),
                algorithm=alg)


def _make_runner(meta, obs_streams, rewards, dones, time_outs, init_params, log_dir=None, device=DEV, **alg_extra):
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    env = ReplayVecEnv({k: t(v) for k, v in obs_streams.items()}, t(rewards), t(dones), t(time_outs),
                       meta["num_actions"], device=device)
    runner = OnPolicyRunner(env, _train_cfg(meta, **alg_extra), log_dir=log_dir, device=device)
    runner.alg.policy.load_state_dict({k: t(v).to(device) for k, v in init_params.items()})
    assert runner.alg.fused, runner.alg._engine_reason
    return runner, env


@pytest.mark.parametrize("name", ["ppo_mlp_small", "ppo_mlp_groups"])
def test_learn_iterations_match_reference_golden(name):
    check_learn_iterations_against_golden(name, DEV)


def check_learn_iterations_against_golden(name, DEV, **alg_extra):
    """The feed-forward PPO path (fused engine) replaying a golden trace of the unmodified reference on device `DEV`
    (also driven by tests/test_host_logic_engine.py above the emulated ABI)."""
    g = load(name)
    meta = meta_of(g)
    names = meta["param_names"]
    T, iters = meta["num_steps"], meta["iterations"]
    # the env replays the concatenated recorded stream of all iterations
    obs = {k: np.concatenate([g[f"it{i}.stream.obs.{k}"][:T + (i == iters - 1)] for i in range(iters)])
           for k in meta["group_dims"]}
    cat = lambda key: np.concatenate([g[f"it{i}.stream.{key}"] for i in range(iters)])
    init = OrderedDict((k, g[f"init.{k}"]) for k in names)
    runner, env = _make_runner(meta, obs, cat("rewards"), cat("dones"), cat("time_outs"), init, device=DEV,
                               use_cuda_graph=False, **alg_extra)
    alg = runner.alg
    for it in range(iters):
        p = f"it{it}."
        zq = [torch.from_numpy(z).to(DEV) for z in g[p + "z_rollout"]]
        alg.policy._noise_fn = lambda like, q=zq: q.pop(0)
        perm = torch.from_numpy(g[p + "perm"]).to(DEV)
        alg.storage._randperm = lambda n, perm=perm: perm
        raw = []
        alg._mb_hook = lambda a: raw.append(a._group.grads.detach().clone())
        st = alg.storage
        snap = {}
        orig_update = alg.update

        def update_and_snapshot():
            for k in ("actions", "values", "actions_log_prob", "mu", "sigma", "rewards", "dones", "returns", "advantages"):
                snap[k] = getattr(st, k).cpu().numpy().copy()
            for k in meta["group_dims"]:
                snap["obs." + k] = st.observations[k].cpu().numpy().copy()
            return orig_update()

        alg.update = update_and_snapshot
        runner.learn(1)
        alg.update = orig_update
        # ---- rollout products
        for k in meta["group_dims"]:
            assert np.array_equal(snap["obs." + k], g[p + "storage.obs." + k])
        assert np.array_equal(snap["dones"], g[p + "storage.dones"])
        assert rel_err(snap["mu"], g[p + "storage.mu"]) < 1e-5
        assert rel_err(snap["actions"], g[p + "storage.actions"]) < 1e-5
        assert rel_err(snap["values"], g[p + "storage.values"]) < 1e-5
        assert rel_err(snap["actions_log_prob"], g[p + "storage.actions_log_prob"]) < 1e-5
        assert rel_err(snap["sigma"], g[p + "storage.sigma"]) < 1e-6
        assert rel_err(snap["rewards"], g[p + "storage.rewards"]) < 1e-5
        assert rel_err(snap["returns"], g[p + "storage.returns"]) < 1e-5
        assert rel_err(snap["advantages"], g[p + "storage.advantages"]) < 1e-5
        # ---- per-minibatch gradients (padded flat layout -> reference order)
        grp = alg._group
        pick = torch.cat([torch.arange(o, o + q.numel()) for o, q in zip(grp.offsets, grp.params)]).to(DEV)
        ref_raw = g[p + "mb.raw_grads"]
        assert len(raw) == ref_raw.shape[0]
        assert rel_err(raw[0][pick].cpu().numpy(), ref_raw[0]) < 1e-5
        for i in range(len(raw)):
            assert rel_err(raw[i][pick].cpu().numpy(), ref_raw[i]) < 5e-4, i
        # ---- post-update parameters, Adam state, learning rate
        mine = torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()]).cpu().numpy()
        assert rel_err(mine, g[p + "mb.params_after"][-1]) < 1e-5
        assert rel_err(grp.exp_avg[pick].cpu().numpy(), g[p + "adam.exp_avg"]) < 1e-4
        assert rel_err(grp.exp_avg_sq[pick].cpu().numpy(), g[p + "adam.exp_avg_sq"]) < 1e-4
        assert alg.last_update_info["step"] == int(g[p + "adam.step"])
        assert abs(alg.learning_rate - float(g[p + "learning_rate"])) < 1e-15
        sd = alg.optimizer.state_dict()
        assert float(sd["state"][0]["step"]) == float(g[p + "adam.step"]) and sd["param_groups"][0]["lr"] == alg.learning_rate
    return runner


def _synthetic_problem(N, T, O_dim, A, hidden, iters, seed):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    obs = f(iters * T + 1, N, O_dim)
    rewards = f(iters * T, N)
    dones = (rng.random((iters * T, N)) < 0.02).astype(np.int64)
    time_outs = (rng.random((iters * T, N)) < 0.5) & (dones > 0)
    z = f(iters, T, N, A)
    B = N * T
    perms = np.stack([rng.permutation(B) for _ in range(iters)]).astype(np.int64)
    meta = dict(num_envs=N, num_steps=T, num_actions=A, hidden=hidden, activation="elu", noise_std_type="scalar",
                obs_groups={"policy": ["policy"], "critic": ["policy"]}, group_dims={"policy": O_dim},
                alg=dict(learning_rate=1e-3, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive",
                         value_loss_coef=1.0, clip_param=0.2, use_clipped_value_loss=True, desired_kl=0.01,
                         entropy_coef=0.01, gamma=0.99, lam=0.95, max_grad_norm=1.0,
                         normalize_advantage_per_mini_batch=False))
    return meta, obs, rewards, dones, time_outs, z, perms


def _init_params_like_torch(meta, seed):
    """Random-init weights of the named architecture (torch default init), as numpy."""
    from rsl_rl_b200.modules import ActorCritic
    from rsl_rl_b200.utils import TensorDictLite
    torch.manual_seed(seed)
    N = 2
    obs = TensorDictLite({k: torch.zeros(N, d) for k, d in meta["group_dims"].items()}, batch_size=[N])
    pol = ActorCritic(obs, meta["obs_groups"], meta["num_actions"], actor_hidden_dims=meta["hidden"],
                      critic_hidden_dims=meta["hidden"], activation=meta["activation"])
    return OrderedDict((k, v.detach().numpy().copy()) for k, v in pol.named_parameters())


def _run_both(meta, obs, rewards, dones, time_outs, z, perms, iters, capture_grads=False, **alg_extra):
    init = _init_params_like_torch(meta, 0)
    T = meta["num_steps"]
    runner, env = _make_runner(meta, {"policy": obs}, rewards, dones, time_outs, init, **alg_extra)
    alg = runner.alg
    raw = []
    if capture_grads:
        alg._mb_hook = lambda a: raw.append(a._group.grads.detach().clone())
    params = OrderedDict((k, v.copy()) for k, v in init.items())
    adam, lr, cfg = O.adam_init(params), meta["alg"]["learning_rate"], cfg_of(meta)
    out = []
    for it in range(iters):
        zq = [torch.from_numpy(zz).to(DEV) for zz in z[it]]
        alg.policy._noise_fn = lambda like, q=zq: q.pop(0)
        perm = torch.from_numpy(perms[it]).to(DEV)
        alg.storage._randperm = lambda n, perm=perm: perm
        runner.learn(1)
        sl = slice(it * T, (it + 1) * T)
        stream = dict(obs_actor=obs[it * T:(it + 1) * T + 1], obs_critic=obs[it * T:(it + 1) * T + 1],
                      rewards=rewards[sl], dones=dones[sl], time_outs=time_outs[sl])
        st, last_values, loss_dict, lr, trace = O.learn_iteration(params, adam, lr, stream, z[it], perms[it], cfg)
        mine = torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()]).cpu().numpy()
        out.append(dict(mine=mine, ref=flat(params), lr_mine=alg.learning_rate, lr_ref=lr, st=st, alg=alg, raw=raw,
                        trace=trace))
    return out


def test_three_iterations_with_cuda_graph_replay_vs_oracle():
    """Iteration 0 runs eagerly, iteration 1 captures the update() graph, iteration 2 replays it."""
    meta, obs, rewards, dones, time_outs, z, perms = _synthetic_problem(256, 8, 20, 6, [64, 32], 3, seed=5)
    res = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 3, use_cuda_graph=True)
    for it, r in enumerate(res):
        assert abs(r["lr_mine"] - r["lr_ref"]) < 1e-15, it
        assert rel_err(r["mine"], r["ref"]) < 2e-5, it
    assert res[-1]["alg"]._graph is not None


def test_c1_shape_full_iteration_vs_oracle():
    """BASELINE config C1/C2 per GPU: 4096 envs x 24 steps, obs 48, act 12, MLP [512,256,128], 5x4 minibatches.

    Tolerances: everything up to and including the first minibatch's gradient is held to the fp32 class
    (1e-5).  The parameters after all 20 Adam steps are held to 1e-4: Adam's m/sqrt(v) normalisation
    amplifies summation-order noise of the 24576-row GEMMs -- the unmodified reference (torch CPU, MKL)
    and the numpy oracle (OpenBLAS) themselves differ by 2.2e-5 on this config (parity notes)."""
    meta, obs, rewards, dones, time_outs, z, perms = _synthetic_problem(4096, 24, 48, 12, [512, 256, 128], 1, seed=9)
    r = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 1, capture_grads=True, use_cuda_graph=False,
                  gemm_precision="fp32")[0]   # the CUDA-core FMA mode (the default is the 3xTF32 tensor-core mode below)
    st, alg = r["st"], r["alg"]
    s = alg.storage
    assert rel_err(s.mu.cpu().numpy(), st["mu"]) < 1e-5
    assert rel_err(s.values.cpu().numpy(), st["values"]) < 1e-5
    assert rel_err(s.actions_log_prob.cpu().numpy(), st["actions_log_prob"]) < 1e-5
    assert np.array_equal(s.dones.cpu().numpy(), st["dones"])
    assert rel_err(s.returns.cpu().numpy(), st["returns"]) < 1e-5
    assert rel_err(s.advantages.cpu().numpy(), st["advantages"]) < 1e-5
    grp = alg._group
    pick = torch.cat([torch.arange(o, o + q.numel()) for o, q in zip(grp.offsets, grp.params)]).to(DEV)
    assert len(r["raw"]) == 20
    assert rel_err(r["raw"][0][pick].cpu().numpy(), r["trace"][0]["grads"]) < 1e-5
    assert abs(alg.last_update_info["total_norm"] - r["trace"][-1]["total_norm"]) < 1e-3 * r["trace"][-1]["total_norm"]
    assert abs(r["lr_mine"] - r["lr_ref"]) < 1e-15
    assert rel_err(r["mine"], r["ref"]) < 1e-4


def test_c1_shape_full_iteration_vs_oracle_3xtf32():
    """gemm_precision='tf32x3' (three split-operand tcgen05 products per GEMM, fp32 accumulation in TMEM) at the C1/C2
    shape: the fp32 class of the CUDA-core mode above -- 1e-5 up to the first minibatch's gradient; the parameters after the
    20 Adam steps are held to 2e-4 (measured 1.2e-4, see below)."""
    meta, obs, rewards, dones, time_outs, z, perms = _synthetic_problem(4096, 24, 48, 12, [512, 256, 128], 1, seed=9)
    r = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 1, capture_grads=True, use_cuda_graph=False,
                  gemm_precision="tf32x3")[0]
    st, alg = r["st"], r["alg"]
    s = alg.storage
    assert rel_err(s.mu.cpu().numpy(), st["mu"]) < 1e-5
    assert rel_err(s.values.cpu().numpy(), st["values"]) < 1e-5
    assert rel_err(s.actions_log_prob.cpu().numpy(), st["actions_log_prob"]) < 1e-5
    assert rel_err(s.returns.cpu().numpy(), st["returns"]) < 1e-5
    assert rel_err(s.advantages.cpu().numpy(), st["advantages"]) < 1e-5
    grp = alg._group
    pick = torch.cat([torch.arange(o, o + q.numel()) for o, q in zip(grp.offsets, grp.params)]).to(DEV)
    e_grad = rel_err(r["raw"][0][pick].cpu().numpy(), r["trace"][0]["grads"])
    e_par = rel_err(r["mine"], r["ref"])
    print(f"tf32x3 C1: first-minibatch gradient {e_grad:.3e}, parameters after 20 steps {e_par:.3e}")
    assert e_grad < 1e-5
    assert abs(r["lr_mine"] - r["lr_ref"]) < 1e-15
    # measured 1.1e-4 (CUDA-core fp32 mode: 1.0e-4 in the same max-norm metric, tools/parity_report.py): the first gradient
    # is 1e-6 from the oracle instead of 5e-7 (the tensor core truncates when it adds into the fp32 accumulator), and 20
    # Adam steps amplify that as they do any summation-order noise
    assert e_par < 2e-4


def _check_tensor_core_class(r, floor):
    """North-star class for TF32 / 16-bit GEMMs (1e-3): rollout products and the first-minibatch gradient are held to
    1e-3 in MAX norm (max-abs error / max-abs value).
    The per-minibatch parameter trace (compared up to the first KL-adaptive LR decision that differs) is held to the
    NOISE FLOOR OF THE ARITHMETIC CLASS: Adam divides every gradient element by its own magnitude, so an element whose
    sign is decided by operand rounding moves its parameter by +-lr either way.  The fp32 oracle re-run with
    tf32-rounded matmul operands (tools/parity_report.tf32_noise_floor -- what a correctly rounded cuBLAS-TF32 reference
    would do) is itself 2.9e-3 (max norm, = 2 * lr after ONE step) / 1.1e-3 (relative L2) away from the fp32 oracle at the
    C1 shape; the CUDA path must not be worse than 1.25x that floor (measured: 2.9e-3 / 1.12e-3, profiles/parity_r2.json)."""
    for k in ("rollout.mu", "rollout.values", "rollout.actions_log_prob", "rollout.returns", "rollout.advantages"):
        assert r[k] < 1e-3, (k, r[k])
    assert r["rollout.dones_equal"] and r["finite"]
    assert r["grad0.max_norm"] < 1e-3, r["grad0.max_norm"]
    assert r["lr.steps_in_agreement"] >= 1
    assert r["params_after.rel_l2_until_lr_divergence"] < 1.25 * floor["params_after_rel_l2"] + 1e-4, \
        (r["params_after.rel_l2_until_lr_divergence"], floor["params_after_rel_l2"])
    assert r["params_after.max_norm_until_lr_divergence"] < 1.25 * floor["params_after_max_norm"] + 1e-4, \
        (r["params_after.max_norm_until_lr_divergence"], floor["params_after_max_norm"])
    assert floor["params_after_max_norm"] < 5e-3 and floor["params_after_rel_l2"] < 2e-3   # the floor itself is ~1e-3
    if r["lr.steps_in_agreement"] < 20:   # a KL within rounding noise of a threshold flipped one decision: 1.5x steps
        ratio = r["lr.mine"] / r["lr.ref"]
        assert min(abs(ratio - 1.5 ** k) for k in range(-4, 5)) < 1e-6, "lr trajectories differ by more than threshold flips"


@pytest.mark.parametrize("prec", ["tf32", "fp16"])
def test_tensor_core_iteration_vs_oracle_c1(prec):
    """gemm_precision='tf32' / 'fp16' (tcgen05) at the C1/C2 shape (4096 envs, obs 48, MLP [512,256,128])."""
    from tools import parity_report as R
    r = R.compare("c1", prec)
    print({k: v for k, v in r.items() if k != "params_after.max_norm_per_step"})
    _check_tensor_core_class(r, R.tf32_noise_floor("c1"))


def test_tensor_core_iteration_vs_oracle_c3_layout():
    """The benchmark default (fp16) on C3's observation layout: two groups 48 + 187 = 235 columns (rows not 16-byte
    aligned: padded first-layer shadow, staged rollout inputs, vector-atomic first-layer weight gradient), 8192 envs."""
    from tools import parity_report as R
    r = R.compare("c3", "fp16")
    print({k: v for k, v in r.items() if k != "params_after.max_norm_per_step"})
    _check_tensor_core_class(r, R.tf32_noise_floor("c3"))


def test_3xtf32_iteration_vs_oracle_c3_layout():
    """The DEFAULT precision (tf32x3) on C3's observation layout (two groups 48 + 187 = 235 columns, rows not 16-byte aligned:
    zero-padded first-layer shadow on the tensor-core path), 8192 envs: rollout products and the first-minibatch gradient
    in the fp32 class (1e-5), the parameter trace over the 20 Adam steps within 6e-4 with every LR decision agreeing."""
    from tools import parity_report as R
    r = R.compare("c3", "tf32x3")
    print({k: v for k, v in r.items() if k != "params_after.max_norm_per_step"})
    for k in ("rollout.mu", "rollout.values", "rollout.actions_log_prob", "rollout.returns", "rollout.advantages"):
        assert r[k] < 1e-5, (k, r[k])
    assert r["rollout.dones_equal"] and r["finite"]
    assert r["grad0.max_norm"] < 1e-5, r["grad0.max_norm"]
    assert r["lr.steps_in_agreement"] == 20
    # (measured 3.5e-4; the CUDA-core fp32 mode: 9e-5 -- Adam turns the 4e-6 relative-L2 gradient difference of the first
    #  minibatch into sign flips of near-zero gradient elements, profiles/parity_r2.json)
    assert r["params_after.max_norm_until_lr_divergence"] < 6e-4, r["params_after.max_norm_until_lr_divergence"]


@pytest.mark.parametrize("reward_scale", [1e4, 1e-4])
def test_fp16_gradient_range_is_tracked(reward_scale):
    """fp16 storage of the layer gradients under extreme reward magnitudes (value-loss gradients ~ reward scale): the
    device-resident dynamic power-of-two scale is calibrated on the first minibatch and then follows the gradient
    magnitude; nothing saturates, results stay finite and in the tensor-core class."""
    from tools import parity_report as R
    R.SHAPES["range"] = dict(N=1024, T=8, groups={"policy": 48}, A=12, hidden=[128, 128, 64], seed=31)
    r = R.compare("range", "fp16", reward_scale=reward_scale)
    print({k: v for k, v in r.items() if k != "params_after.max_norm_per_step"})
    assert r["finite"]
    for k in ("rollout.mu", "rollout.values", "rollout.returns"):
        assert r[k] < 1e-3, (k, r[k])
    assert r["fp16_saturation_events"] == 0, r["fp16_saturation_events"]
    assert r["grad0.max_norm"] < 2e-3, r["grad0.max_norm"]
    assert r["params_after.rel_l2_until_lr_divergence"] is None or r["params_after.rel_l2_until_lr_divergence"] < 3e-3


def test_tf32_unaligned_observation_width_vs_oracle():
    """Observation width 10 (row pitch not a multiple of 16 bytes): the tensor-core path reads a zero-padded shadow
    of the first-layer weights and padded inputs."""
    meta, obs, rewards, dones, time_outs, z, perms = _synthetic_problem(512, 8, 10, 6, [64, 32], 3, seed=11)
    # one iteration, eager: forward parity of the padded path and the first update
    r = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 1, use_cuda_graph=False, gemm_precision="tf32")[0]
    s, st = r["alg"].storage, r["st"]
    assert r["alg"]._engine.w0_pad, "the padded shadow weights were expected to be in use"
    assert rel_err(s.mu.cpu().numpy(), st["mu"]) < 1e-3
    assert rel_err(s.values.cpu().numpy(), st["values"]) < 1e-3
    init = np.concatenate([v.reshape(-1) for v in _init_params_like_torch(meta, 0).values()]).astype(np.float64)
    upd, upd_ref = r["mine"].astype(np.float64) - init, r["ref"].astype(np.float64) - init
    if abs(r["lr_mine"] - r["lr_ref"]) < 1e-12:
        assert np.linalg.norm(upd - upd_ref) / np.linalg.norm(upd_ref) < 5e-2
    # three iterations with the rollout / update CUDA graphs active: the shadow must follow the parameters
    res = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 3, use_cuda_graph=True, gemm_precision="tf32")
    alg = res[-1]["alg"]
    w = alg.policy.actor[0].weight
    assert torch.equal(alg._engine.w0_pad["a"][:, :w.shape[1]], w.detach())
    assert float(alg._engine.w0_pad["a"][:, w.shape[1]:].abs().max()) == 0.0
    ratio = res[-1]["lr_mine"] / res[-1]["lr_ref"]
    assert min(abs(ratio - 1.5 ** k) for k in range(-6, 7)) < 1e-6
    if abs(ratio - 1.0) < 1e-9:
        u, ur = res[-1]["mine"].astype(np.float64) - init, res[-1]["ref"].astype(np.float64) - init
        assert np.linalg.norm(u - ur) / np.linalg.norm(ur) < 0.15


def test_fp16_three_iterations_with_cuda_graphs_shadow_follows_parameters():
    """gemm_precision='fp16' with the rollout / update CUDA graphs active over three iterations: the fp16 weight
    shadow must equal fp16(parameters) after every update and the LR trajectory may differ from the oracle's only by
    threshold flips of the KL rule."""
    meta, obs, rewards, dones, time_outs, z, perms = _synthetic_problem(1024, 8, 48, 12, [128, 128, 64], 3, seed=21)
    res = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 3, use_cuda_graph=True, gemm_precision="fp16")
    alg = res[-1]["alg"]
    assert alg._engine.f16 and alg._graph is not None
    alg._engine.refresh_padded()
    grp = alg._group
    assert torch.equal(grp.values16, grp.values.half())
    s, st = alg.storage, res[-1]["st"]
    ratio = res[-1]["lr_mine"] / res[-1]["lr_ref"]
    assert min(abs(ratio - 1.5 ** k) for k in range(-6, 7)) < 1e-6
    init = np.concatenate([v.reshape(-1) for v in _init_params_like_torch(meta, 0).values()]).astype(np.float64)
    if abs(ratio - 1.0) < 1e-9:
        u, ur = res[-1]["mine"].astype(np.float64) - init, res[-1]["ref"].astype(np.float64) - init
        assert np.linalg.norm(u - ur) / np.linalg.norm(ur) < 0.15


def test_fp16_unaligned_observation_width_vs_oracle():
    """gemm_precision='fp16' with observation width 10 (rows not 16-byte aligned): zero-padded first-layer shadow,
    staged inputs for the last-value forward, padded fp16 observation copy and the vector-atomic epilogue for the first
    layer's weight gradient."""
    meta, obs, rewards, dones, time_outs, z, perms = _synthetic_problem(512, 8, 10, 6, [128, 128, 64], 3, seed=13)
    r = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 1, use_cuda_graph=False, gemm_precision="fp16")[0]
    s, st = r["alg"].storage, r["st"]
    assert r["alg"]._engine.f16 and r["alg"]._engine.w0_pad
    assert rel_err(s.mu.cpu().numpy(), st["mu"]) < 1e-3
    assert rel_err(s.values.cpu().numpy(), st["values"]) < 1e-3
    assert rel_err(s.returns.cpu().numpy(), st["returns"]) < 1e-3
    init = np.concatenate([v.reshape(-1) for v in _init_params_like_torch(meta, 0).values()]).astype(np.float64)
    if abs(r["lr_mine"] - r["lr_ref"]) < 1e-12:
        upd, upd_ref = r["mine"].astype(np.float64) - init, r["ref"].astype(np.float64) - init
        assert np.linalg.norm(upd - upd_ref) / np.linalg.norm(upd_ref) < 5e-2
    # three iterations with the rollout / update graphs active (the LR may hit its clamp at 1e-2 here, so the two
    # trajectories are only compared when they coincide)
    res = _run_both(meta, obs, rewards, dones, time_outs, z, perms, 3, use_cuda_graph=True, gemm_precision="fp16")
    assert 1e-5 <= res[-1]["lr_mine"] <= 1e-2 and np.isfinite(res[-1]["mine"]).all()
    if abs(res[-1]["lr_mine"] / res[-1]["lr_ref"] - 1.0) < 1e-9:
        u, ur = res[-1]["mine"].astype(np.float64) - init, res[-1]["ref"].astype(np.float64) - init
        assert np.linalg.norm(u - ur) / np.linalg.norm(ur) < 0.15


# ----------------------------------------------------------------------------------------------------------------------
# observation normalisers, logging path, checkpoints -- against the unmodified reference's golden run
# ----------------------------------------------------------------------------------------------------------------------
def _golden_stream(g, meta):
    T, iters = meta["num_steps"], meta["iterations"]
    obs = {k: np.concatenate([g[f"it{i}.stream.obs.{k}"][:T + (i == iters - 1)] for i in range(iters)])
           for k in meta["group_dims"]}
    cat = lambda key: np.concatenate([g[f"it{i}.stream.{key}"] for i in range(iters)])
    return obs, cat("rewards"), cat("dones"), cat("time_outs")


def _drive_golden(g, meta, runner, iters, on_iter=None, DEV=DEV):
    alg = runner.alg
    for it in range(iters):
        p = f"it{it}."
        zq = [torch.from_numpy(z).to(DEV) for z in g[p + "z_rollout"]]
        alg.policy._noise_fn = lambda like, q=zq: q.pop(0)
        perm = torch.from_numpy(g[p + "perm"]).to(DEV)
        alg.storage._randperm = lambda n, perm=perm: perm
        runner.learn(1)
        if on_iter is not None:
            on_iter(it, p)


@pytest.mark.parametrize("use_graph", [False, True])
def test_observation_normalisers_match_reference_golden(use_graph):
    """actor/critic_obs_normalization=True (EmpiricalNormalization, reference normalization.py:44-66) over four
    iterations: the running statistics change every env step while act() launch plans and -- with use_graph -- the
    captured update() graph hold the buffers' addresses (eager -> capture -> replay)."""
    alg = check_normaliser_golden(DEV, use_graph).alg
    assert (alg._graph is not None) == use_graph


def check_normaliser_golden(DEV, use_graph, **alg_extra):
    """(also driven by tests/test_host_logic_engine.py above the emulated ABI, without graphs)"""
    g = load("ppo_mlp_norm")
    meta = meta_of(g)
    obs, rewards, dones, time_outs = _golden_stream(g, meta)
    init = OrderedDict((k[len("init."):], g[k]) for k in g.files if k.startswith("init."))
    runner, env = _make_runner(meta, obs, rewards, dones, time_outs, init, device=DEV, use_cuda_graph=use_graph,
                               **alg_extra)
    alg = runner.alg
    names = meta["param_names"]

    def check(it, p):
        sd = alg.policy.state_dict()
        for key in ("actor_obs_normalizer._mean", "actor_obs_normalizer._var", "actor_obs_normalizer._std",
                    "critic_obs_normalizer._mean", "critic_obs_normalizer._var", "critic_obs_normalizer._std"):
            assert rel_err(sd[key].cpu().numpy(), g[p + "final." + key]) < 1e-5, (it, key)
        assert int(sd["actor_obs_normalizer.count"]) == int(g[p + "final.actor_obs_normalizer.count"])
        assert rel_err(alg.storage.mu.cpu().numpy(), g[p + "storage.mu"]) < 2e-5, it
        assert rel_err(alg.storage.values.cpu().numpy(), g[p + "storage.values"]) < 2e-5, it
        assert rel_err(alg.storage.returns.cpu().numpy(), g[p + "storage.returns"]) < 2e-5, it
        mine = np.concatenate([sd[k].cpu().numpy().reshape(-1) for k in names])
        assert abs(alg.learning_rate - float(g[p + "learning_rate"])) < 1e-15, it
        assert rel_err(mine, g[p + "mb.params_after"][-1]) < 5e-5, it

    _drive_golden(g, meta, runner, meta["iterations"], check, DEV=DEV)
    return runner


def test_logging_path_matches_reference_buffers(tmp_path):
    """log_dir set: the device-side episode statistics, flushed once per iteration, reproduce the
    reference's rewbuffer / lenbuffer contents (on_policy_runner.py:113-139) and the logged mean losses."""
    check_logging_path(DEV, tmp_path)


def check_logging_path(DEV, tmp_path, **alg_extra):
    """(also driven by tests/test_host_logic_engine.py above the emulated ABI)"""
    g = load("ppo_mlp_small")
    meta = meta_of(g)
    obs, rewards, dones, time_outs = _golden_stream(g, meta)
    init = OrderedDict((k, g[f"init.{k}"]) for k in meta["param_names"])
    runner, env = _make_runner(meta, obs, rewards, dones, time_outs, init, log_dir=str(tmp_path), device=DEV,
                               use_cuda_graph=False, **alg_extra)
    seen = []
    orig_log = runner.log

    def log_and_record(locs, *a, **k):
        seen.append(dict(rew=list(locs["rewbuffer"]), length=list(locs["lenbuffer"]), loss=dict(locs["loss_dict"])))
        return orig_log(locs, *a, **k)

    runner.log = log_and_record

    def check(it, p):
        ref_rew, ref_len = g[p + "log.rewbuffer"], g[p + "log.lenbuffer"]
        assert len(seen) == it + 1
        assert np.array_equal(np.asarray(seen[-1]["length"], dtype=np.float64), ref_len), it
        assert len(seen[-1]["rew"]) == len(ref_rew)
        if len(ref_rew):
            assert rel_err(np.asarray(seen[-1]["rew"]), ref_rew) < 1e-5, it
        for k in ("value_function", "surrogate", "entropy"):
            ref = float(g[p + "log.loss." + k])
            assert abs(seen[-1]["loss"][k] - ref) < 1e-4 * max(1.0, abs(ref)), (it, k)

    _drive_golden(g, meta, runner, meta["iterations"], check, DEV=DEV)
    import os
    assert any(f.startswith("events.out.tfevents") for f in os.listdir(tmp_path)), "no TensorBoard event file written"
    assert any(f.startswith("model_") for f in os.listdir(tmp_path)), "no checkpoint written at the end of learn()"


def test_reference_checkpoint_round_trip(tmp_path):
    """a checkpoint WRITTEN BY THE REFERENCE (tests/golden/ppo_mlp_norm_ref_checkpoint.pt, saved by
    oracle/make_golden.py after four iterations, on_policy_runner.py:289-305) loads into this runner (parameters,
    normaliser buffers, Adam moments / step / lr, iteration counter), training resumes from it, and a checkpoint written
    here has the reference's structure (and loads into the reference when its install travelled with the repo)."""
    check_reference_checkpoint_round_trip(DEV, tmp_path)


def check_reference_checkpoint_round_trip(DEV, tmp_path, **alg_extra):
    """(also driven by tests/test_host_logic_engine.py above the emulated ABI)"""
    import os
    from tests.golden_util import GOLDEN
    g = load("ppo_mlp_norm")
    meta = meta_of(g)
    obs, rewards, dones, time_outs = _golden_stream(g, meta)
    init = OrderedDict((k[len("init."):], g[k]) for k in g.files if k.startswith("init."))
    runner, env = _make_runner(meta, obs, rewards, dones, time_outs, init, log_dir=str(tmp_path), device=DEV,
                               use_cuda_graph=False, **alg_extra)
    path = os.path.join(GOLDEN, "ppo_mlp_norm_ref_checkpoint.pt")
    ref_ckpt = torch.load(path, map_location="cpu", weights_only=False)
    runner.load(path)
    alg = runner.alg
    last = f"it{meta['iterations'] - 1}."
    sd = alg.policy.state_dict()
    for k in sd:
        assert np.array_equal(sd[k].cpu().numpy(), g[last + "final." + k]), k
    grp = alg._group
    pick = torch.cat([torch.arange(o, o + q.numel()) for o, q in zip(grp.offsets, grp.params)]).to(DEV)
    assert np.array_equal(grp.exp_avg[pick].cpu().numpy(), g[last + "adam.exp_avg"])
    assert np.array_equal(grp.exp_avg_sq[pick].cpu().numpy(), g[last + "adam.exp_avg_sq"])
    assert alg.optimizer._host_step == int(g[last + "adam.step"])
    assert runner.current_learning_iteration == ref_ckpt["iter"]
    # resume: one more iteration on fresh data runs on the loaded state (Adam step counter keeps counting)
    alg.learning_rate = ref_ckpt["optimizer_state_dict"]["param_groups"][0]["lr"]
    env.t = 0
    runner.learn(1)
    assert alg.last_update_info["step"] == int(g[last + "adam.step"]) + 8
    assert all(bool(torch.isfinite(v).all()) for v in alg.policy.state_dict().values())
    # save: same top-level structure, same keys, shapes and dtypes as the reference's file
    mine_path = os.path.join(str(tmp_path), "mine.pt")
    runner.save(mine_path)
    mine = torch.load(mine_path, map_location="cpu", weights_only=False)
    assert set(ref_ckpt.keys()) <= set(mine.keys()), (sorted(ref_ckpt.keys()), sorted(mine.keys()))
    for k, v in ref_ckpt["model_state_dict"].items():
        assert mine["model_state_dict"][k].shape == v.shape and mine["model_state_dict"][k].dtype == v.dtype, k
    ro, mo = ref_ckpt["optimizer_state_dict"], mine["optimizer_state_dict"]
    assert set(ro["state"].keys()) == set(mo["state"].keys())
    for i in ro["state"]:
        assert set(ro["state"][i].keys()) == set(mo["state"][i].keys())
        assert ro["state"][i]["exp_avg"].shape == mo["state"][i]["exp_avg"].shape
    assert {k for k in ro["param_groups"][0]} <= {k for k in mo["param_groups"][0]}
    # ... and the reference itself loads it (CPU), when baseline/_ref is present
    from oracle import reference_arm as R
    if R.available() is None:
        R.import_reference()
        from rsl_rl.runners import OnPolicyRunner as RefRunner
        import contextlib, io
        cfg = _train_cfg(meta)
        for k in ("gemm_precision", "use_cuda_graph", "rng_compat"):
            cfg["algorithm"].pop(k, None)
        renv = R.make_env(meta["num_envs"], meta["group_dims"], meta["num_actions"], "cpu")
        with contextlib.redirect_stdout(io.StringIO()):
            ref_runner = RefRunner(renv, cfg, log_dir=None, device="cpu")
            ref_runner.load(mine_path)
        rsd = ref_runner.alg.policy.state_dict()
        for k, v in alg.policy.state_dict().items():
            assert torch.equal(rsd[k], v.cpu()), k
        assert ref_runner.current_learning_iteration == runner.current_learning_iteration


@pytest.mark.parametrize("direct", ["0", "1"])
def test_host_resident_env_receives_the_actions_written_by_the_rollout_kernel(direct, monkeypatch):
    """Env on the host (pinned buffers).  direct=1 (opt-in, RSLRL_HOST_ACTIONS=1): from the second env step on the one-launch
    rollout kernel writes the actions into a pinned ring as well and the runner hands that ring slot to the env after one
    stream synchronisation (no device-to-host copy); direct=0: pinned copy ring.  Either way what the env receives must be
    exactly what the rollout buffer holds."""
    monkeypatch.setenv("RSLRL_HOST_ACTIONS", direct)
    from rsl_rl_b200.env import PinnedHostEnv
    from rsl_rl_b200.runners import OnPolicyRunner
    import bench

    class Recording(PinnedHostEnv):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.seen = []

        def step(self, actions):
            assert not actions.is_cuda
            self.seen.append(actions.clone())
            return super().step(actions)

    env = Recording(512, {"policy": 48}, 12, pool=8, seed=1)
    runner = OnPolicyRunner(env, bench.train_cfg("fp16"), log_dir=None, device=DEV)
    runner.learn(2)
    alg = runner.alg
    assert (alg._host_actions is not None and alg.last_host_actions is not None) == (direct == "1")
    T = alg.storage.num_transitions_per_env
    got = torch.stack(env.seen[-T:])                      # the second iteration
    assert torch.equal(got, alg.storage.actions.cpu())
