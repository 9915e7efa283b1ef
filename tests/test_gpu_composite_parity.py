"""GPU: the configurations that go through PPO._update_composite -- recurrent policy (LSTM memory,
recurrent_mini_batch_generator, split/pad/unpad) and RND + symmetry augmentation + mirror loss (C5) -- against
golden traces recorded from the unmodified reference (oracle/make_golden.py)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from tests.golden_util import load, meta_of, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _runner_from_golden(g, meta, DEV=DEV, **alg_extra):
    from oracle.symmetry_example import mirror_augmentation
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner
    T, iters = meta["num_steps"], meta["iterations"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    obs = {k: t(np.concatenate([g[f"it{i}.stream.obs.{k}"][:T + (i == iters - 1)] for i in range(iters)]))
           for k in meta["group_dims"]}
    cat = lambda key: t(np.concatenate([g[f"it{i}.stream.{key}"] for i in range(iters)]))
    env = ReplayVecEnv(obs, cat("rewards"), cat("dones"), cat("time_outs"), meta["num_actions"], device=DEV)
    alg = dict(meta["alg"], class_name="PPO", use_cuda_graph=False, **alg_extra)
    if meta["rnd_cfg"] is not None:
        alg["rnd_cfg"] = dict(meta["rnd_cfg"])
    if meta["symmetry"]:
        flags = meta.get("symmetry_cfg", dict(use_data_augmentation=True, use_mirror_loss=True, mirror_loss_coeff=0.5))
        alg["symmetry_cfg"] = dict(flags, data_augmentation_func=mirror_augmentation)
    cfg = dict(num_steps_per_env=T, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in meta["obs_groups"].items()},
               policy=dict(class_name=meta["policy_class"], actor_hidden_dims=list(meta["hidden"]),
                           critic_hidden_dims=list(meta["hidden"]), activation=meta["activation"], init_noise_std=1.0,
                           noise_std_type=meta["noise_std_type"], **meta["policy_extra"]),
               algorithm=alg)
    runner = OnPolicyRunner(env, cfg, log_dir=None, device=DEV)
    sd = {k[len("init."):]: t(g[k]).to(DEV) for k in g.files if k.startswith("init.")}
    runner.alg.policy.load_state_dict(sd)
    if runner.alg.rnd:
        runner.alg.rnd.load_state_dict({k[len("init_rnd."):]: t(g[k]).to(DEV) for k in g.files if k.startswith("init_rnd.")})
    assert not runner.alg.fused
    return runner


@pytest.mark.parametrize("name", ["ppo_lstm_small", "ppo_rnd_sym_small", "ppo_rnd_norm_small"])
def test_composite_learn_iterations_match_reference_golden(name):
    torch.backends.cudnn.allow_tf32 = False   # the fp32 parity class
    check_composite_golden(name, DEV)


def check_composite_golden(name, DEV, **alg_extra):
    """A recurrent / RND / symmetry configuration replaying a golden trace of the unmodified reference on device `DEV`
    (also driven by tests/test_host_logic_engine.py above the emulated ABI)."""
    g = load(name)
    meta = meta_of(g)
    runner = _runner_from_golden(g, meta, DEV, **alg_extra)
    alg, st = runner.alg, runner.alg.storage
    if name == "ppo_lstm_small":   # the recurrent configuration runs on the engine (native LSTM + fused MLP path), no autograd
        assert alg._recurrent_fused and alg.rollout_fused and not alg.fused
    for it in range(meta["iterations"]):
        p = f"it{it}."
        zq = [torch.from_numpy(z).to(DEV) for z in g[p + "z_rollout"]]
        alg.policy._noise_fn = lambda like, q=zq: q.pop(0) if q else torch.zeros_like(like)
        if p + "perm" in g.files:
            perm = torch.from_numpy(g[p + "perm"]).to(DEV)
            st._randperm = lambda n, perm=perm: perm
        snap = {}
        orig_update = alg.update

        def update_and_snapshot():
            for k in ("actions", "values", "actions_log_prob", "mu", "sigma", "rewards", "dones", "returns", "advantages"):
                snap[k] = getattr(st, k).cpu().numpy().copy()
            if st.saved_hidden_states_a is not None:
                snap["hidden_a0"] = st.saved_hidden_states_a[0].cpu().numpy().copy()
                snap["hidden_c1"] = st.saved_hidden_states_c[1].cpu().numpy().copy()
            return orig_update()

        alg.update = update_and_snapshot
        runner.learn(1)
        alg.update = orig_update
        assert np.array_equal(snap["dones"], g[p + "storage.dones"])
        for k in ("mu", "actions", "values", "actions_log_prob", "rewards", "returns", "advantages"):
            assert rel_err(snap[k], g[p + "storage." + k]) < 2e-5, (it, k)
        if "hidden_a0" in snap:
            assert rel_err(snap["hidden_a0"], g[p + "storage.hidden_a0"]) < 2e-5
            assert rel_err(snap["hidden_c1"], g[p + "storage.hidden_c1"]) < 2e-5
        mine = torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()]).cpu().numpy()
        assert abs(alg.learning_rate - float(g[p + "learning_rate"])) < 1e-15, it
        assert rel_err(mine, g[p + "mb.params_after"][-1]) < 5e-5, it
        if alg.rnd:
            for k, v in alg.rnd.state_dict().items():
                assert rel_err(v.cpu().numpy(), g[p + "final_rnd." + k]) < 5e-5, k
        assert alg.last_update_info["step"] == int(g[p + "adam.step"])
    return runner


def _small_rnd_sym_runner(aug, use_graph):
    from rsl_rl_b200.env import SynthVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner
    torch.manual_seed(0)
    env = SynthVecEnv(256, {"policy": 16}, 4, device=DEV, seed=3)
    cfg = dict(num_steps_per_env=8, save_interval=10 ** 9,
               obs_groups={"policy": ["policy"], "critic": ["policy"], "rnd_state": ["policy"]},
               policy=dict(class_name="ActorCritic", actor_hidden_dims=[64, 32], critic_hidden_dims=[64, 32],
                           activation="elu", init_noise_std=1.0),
               algorithm=dict(class_name="PPO", learning_rate=1e-3, num_learning_epochs=2, num_mini_batches=2,
                              schedule="adaptive", desired_kl=0.01, use_cuda_graph=use_graph,
                              rnd_cfg=dict(weight=1.0, weight_schedule=None, reward_normalization=False,
                                           state_normalization=False, learning_rate=1e-3, num_outputs=1,
                                           predictor_hidden_dims=[-1], target_hidden_dims=[-1]),
                              symmetry_cfg=dict(use_data_augmentation=True, use_mirror_loss=True, mirror_loss_coeff=0.5,
                                                data_augmentation_func=aug)))
    return OnPolicyRunner(env, cfg, log_dir=None, device=DEV)


def test_composite_update_cuda_graph_matches_eager_and_falls_back():
    """Feed-forward composite update (RND + symmetry): the minibatch loop replayed from one CUDA graph gives the same
    parameters as the eager loop; an augmentation function that cannot be captured (uploads host data on every call)
    falls back to eager instead of failing."""
    from oracle.symmetry_example import mirror_augmentation as uploading_aug   # builds its index tensors from numpy per call
    cache = {}

    def cached_aug(obs=None, actions=None, env=None):
        def mirror(x):
            d = x.shape[-1]
            if d not in cache:
                g = torch.Generator().manual_seed(100 + d)
                cache[d] = (torch.randperm(d, generator=g).to(DEV), (torch.randint(0, 2, (d,), generator=g) * 2 - 1).float().to(DEV))
            perm, sign = cache[d]
            return x[..., perm] * sign
        obs_out = None
        if obs is not None:
            n = obs.batch_size[0]
            obs_out = type(obs)({k: torch.cat([v, mirror(v)], dim=0) for k, v in obs.items()}, batch_size=[2 * n])
        return obs_out, (None if actions is None else torch.cat([actions, mirror(actions)], dim=0))

    finals = {}
    for mode in ("eager", "graph"):
        runner = _small_rnd_sym_runner(cached_aug, mode == "graph")
        torch.manual_seed(5)
        runner.learn(5)
        alg = runner.alg
        assert not alg.fused
        assert (alg._cgraph is not None) == (mode == "graph")
        finals[mode] = (torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()]).cpu().numpy(),
                        alg.learning_rate, alg.last_update_info["step"])
    assert finals["eager"][2] == finals["graph"][2] == 5 * 2 * 2
    assert abs(finals["eager"][1] - finals["graph"][1]) < 1e-15
    assert rel_err(finals["graph"][0], finals["eager"][0]) < 1e-4
    # fallback
    runner = _small_rnd_sym_runner(uploading_aug, True)
    runner.learn(4)
    alg = runner.alg
    assert alg._cgraph is None and alg._cgraph_failed
    p = torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()])
    assert bool(torch.isfinite(p).all()) and alg.last_update_info["step"] == 4 * 2 * 2


def _lstm_runner(use_graph, hidden=(64, 32), envs=256, rnn=32):
    from rsl_rl_b200.env import SynthVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner
    torch.manual_seed(0)
    env = SynthVecEnv(envs, {"policy": 16}, 4, device=DEV, seed=3, p_done=0.05)
    cfg = dict(num_steps_per_env=8, save_interval=10 ** 9, obs_groups={"policy": ["policy"], "critic": ["policy"]},
               policy=dict(class_name="ActorCriticRecurrent", actor_hidden_dims=list(hidden), critic_hidden_dims=list(hidden),
                           activation="elu", init_noise_std=1.0, rnn_type="lstm", rnn_hidden_dim=rnn, rnn_num_layers=1),
               algorithm=dict(class_name="PPO", learning_rate=1e-3, num_learning_epochs=3, num_mini_batches=2,
                              schedule="adaptive", desired_kl=0.01, use_cuda_graph=use_graph, gemm_precision="fp32"))
    return OnPolicyRunner(env, cfg, log_dir=None, device=DEV)


def test_recurrent_update_on_the_engine_matches_the_autograd_path_and_graph_replay(monkeypatch):
    """ActorCriticRecurrent: (a) the hand-written recurrent update on the engine (native LSTM forward / backward through
    time, fused MLP path, no autograd) gives the parameters of the autograd composite path; (b) replaying every
    minibatch of the later epochs from its CUDA graph gives the parameters of the eager loop."""
    finals = {}
    for mode in ("autograd", "eager", "graph"):
        monkeypatch.setenv("RSLRL_RECURRENT_FUSED", "0" if mode == "autograd" else "1")
        runner = _lstm_runner(mode == "graph")
        alg = runner.alg
        assert alg._recurrent_fused == (mode != "autograd")
        torch.manual_seed(5)
        runner.learn(4)
        finals[mode] = (torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()]).cpu().numpy(), alg.learning_rate)
    assert finals["eager"][1] == finals["graph"][1]
    assert rel_err(finals["graph"][0], finals["eager"][0]) < 1e-6
    assert abs(finals["eager"][1] - finals["autograd"][1]) < 1e-12
    assert rel_err(finals["eager"][0], finals["autograd"][0]) < 5e-5
