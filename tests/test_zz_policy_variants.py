"""Policy variants of SURVEY 8 (f4) on the PPO path -- the state-dependent standard deviation head
(`state_dependent_std=True`, reference modules/actor_critic.py:119-128, both noise parametrisations) and GRU memories
(`rnn_type="gru"`, reference networks/memory.py:23) -- against golden traces recorded from the unmodified reference
(oracle/make_golden.py).  They run through `PPO._update_composite` (dense layers on the library's kernels through their
autograd functions, rollout buffer / GAE / gather / trajectory kernels / fused clip + Adam as everywhere else).

As in test_zz_distillation.py there is a `-m gpu` leg (the product path) and, for the feed-forward cases, a CPU leg above the
host-memory emulation of the C ABI (tests/host_emulation.py; pins the host logic only).
"""
import numpy as np
import pytest
import torch

from tests.golden_util import load, meta_of, rel_err


def _runner_from_golden(g, meta, device, **alg_extra):
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import OnPolicyRunner
    T, iters = meta["num_steps"], meta["iterations"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    obs = {k: t(np.concatenate([g[f"it{i}.stream.obs.{k}"][:T + (i == iters - 1)] for i in range(iters)]))
           for k in meta["group_dims"]}
    cat = lambda key: t(np.concatenate([g[f"it{i}.stream.{key}"] for i in range(iters)]))
    env = ReplayVecEnv(obs, cat("rewards"), cat("dones"), cat("time_outs"), meta["num_actions"], device=device)
    cfg = dict(num_steps_per_env=T, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in meta["obs_groups"].items()},
               policy=dict(class_name=meta["policy_class"], actor_hidden_dims=list(meta["hidden"]),
                           critic_hidden_dims=list(meta["hidden"]), activation=meta["activation"], init_noise_std=1.0,
                           noise_std_type=meta["noise_std_type"], **meta["policy_extra"]),
               algorithm=dict(meta["alg"], class_name="PPO", use_cuda_graph=False, **alg_extra))
    runner = OnPolicyRunner(env, cfg, log_dir=None, device=device)
    sd = {k[len("init."):]: t(g[k]).to(device) for k in g.files if k.startswith("init.")}
    runner.alg.policy.load_state_dict(sd)
    return runner


def _check_ppo_variant(name, device, general_path=True, **alg_extra):
    g = load(name)
    meta = meta_of(g)
    runner = _runner_from_golden(g, meta, device, **alg_extra)
    alg, st = runner.alg, runner.alg.storage
    if general_path:   # modules + autograd functions on the kernels (no fused engine)
        assert not alg.fused and not alg.rollout_fused
    for it in range(meta["iterations"]):
        p = f"it{it}."
        zq = [torch.from_numpy(z).to(device) for z in g[p + "z_rollout"]]
        alg.policy._noise_fn = lambda like, q=zq: q.pop(0) if q else torch.zeros_like(like)
        if p + "perm" in g.files:
            perm = torch.from_numpy(g[p + "perm"]).to(device)
            st._randperm = lambda n, perm=perm: perm
        snap = {}
        orig_update = alg.update

        def update_and_snapshot():
            for k in ("actions", "values", "actions_log_prob", "mu", "sigma", "rewards", "dones", "returns", "advantages"):
                snap[k] = getattr(st, k).cpu().numpy().copy()
            if st.saved_hidden_states_a is not None:
                snap["hidden_a0"] = st.saved_hidden_states_a[0].cpu().numpy().copy()
                snap["hidden_c0"] = st.saved_hidden_states_c[0].cpu().numpy().copy()
            return orig_update()

        alg.update = update_and_snapshot
        runner.learn(1)
        alg.update = orig_update
        assert np.array_equal(snap["dones"], g[p + "storage.dones"])
        for k in ("mu", "sigma", "actions", "values", "actions_log_prob", "rewards", "returns", "advantages"):
            assert rel_err(snap[k], g[p + "storage." + k]) < 2e-5, (it, k)
        if "hidden_a0" in snap:   # GRU: one hidden tensor per memory, LSTM: two (reference rollout_storage.py:108-110)
            assert len(st.saved_hidden_states_a) == (1 if meta["policy_extra"]["rnn_type"] == "gru" else 2)
            assert rel_err(snap["hidden_a0"], g[p + "storage.hidden_a0"]) < 2e-5
            assert rel_err(snap["hidden_c0"], g[p + "storage.hidden_c0"]) < 2e-5
        mine = torch.cat([q.detach().reshape(-1) for q in alg.policy.parameters()]).cpu().numpy()
        assert abs(alg.learning_rate - float(g[p + "learning_rate"])) < 1e-15, it
        assert rel_err(mine, g[p + "mb.params_after"][-1]) < 5e-5, it
        assert alg.last_update_info["step"] == int(g[p + "adam.step"])
    return runner


@pytest.mark.parametrize("name", ["ppo_sdstd_log_small", "ppo_sdstd_scalar_small"])
def test_state_dependent_std_host_logic_matches_reference_golden_on_the_emulated_abi(name, monkeypatch):
    from tests import host_emulation
    lib = host_emulation.install(monkeypatch)
    runner = _check_ppo_variant(name, "cpu")
    assert runner.alg.policy.state_dependent_std and runner.alg.storage.sigma.std() > 0   # per-sample sigma was stored
    assert {"gae_norm", "permute_gather", "store_transition", "allreduce_clip_adam"} <= set(lib.calls)


def test_gru_policy_host_logic_matches_reference_golden_on_the_emulated_abi(monkeypatch):
    """Recurrent generator with single-tensor (GRU) hidden states: trajectory index, padded observations, masks, hidden
    states at trajectory starts, unpad -- host logic above the emulated trajectory kernels; the GRU itself is torch's."""
    from tests import host_emulation
    lib = host_emulation.install(monkeypatch)
    runner = _check_ppo_variant("ppo_gru_small", "cpu")
    assert isinstance(runner.alg.policy.memory_a.rnn, torch.nn.GRU)
    assert {"trajectory_index", "split_and_pad", "trajectory_masks", "hidden_at_starts"} <= set(lib.calls)


def test_lstm_policy_autograd_path_on_the_emulated_abi_reproduces_the_reference_golden(monkeypatch):
    """Cross-check of the emulation itself on a path the GPU tests already pin (tests/test_gpu_composite_parity.py): the
    LSTM policy through the autograd functions of networks/lstm.py (RSLRL_RECURRENT_FUSED=0) -- hoisted input projection,
    time loop, backward through time over 8 steps, weight gradients over all steps -- lands on the reference's golden."""
    from tests import host_emulation
    monkeypatch.setenv("RSLRL_RECURRENT_FUSED", "0")
    lib = host_emulation.install(monkeypatch)
    runner = _check_ppo_variant("ppo_lstm_small", "cpu")
    assert runner.alg.policy.memory_a.native and {"lstm_step_forward", "lstm_step_backward"} <= set(lib.calls)


def test_lstm_policy_on_the_fused_engine_host_logic_on_the_emulated_abi(monkeypatch):
    """The recurrent policy's DEFAULT path (PPO._recurrent_minibatches_fused: LSTM step + engine rollout, per minibatch
    native LSTM over the padded trajectories, un-pad, engine forward / loss / hand-written backward including the gradient
    into the LSTM outputs, pad, native BPTT into the flat gradient buffer, fused step; no autograd) above the emulated ABI
    lands on the reference's golden like the autograd path does."""
    from tests import host_emulation
    lib = host_emulation.install(monkeypatch)
    runner = _check_ppo_variant("ppo_lstm_small", "cpu", general_path=False, rng_compat=False)
    alg = runner.alg
    assert alg._recurrent_fused and alg.rollout_fused and not alg.fused
    assert {"lstm_step_forward", "lstm_step_backward", "unpad", "split_and_pad", "ppo_loss_fwd_bwd", "store_step"} <= set(lib.calls)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ppo_sdstd_log_small", "ppo_sdstd_scalar_small", "ppo_gru_small"])
def test_policy_variants_match_reference_golden(name):
    from rsl_rl_b200 import _native as nat
    torch.backends.cudnn.allow_tf32 = False   # GRU memories run in cuDNN (fp32 parity class)
    before = nat.launch_count()
    runner = _check_ppo_variant(name, "cuda:0")
    assert nat.launch_count() > before
    if name == "ppo_gru_small":
        assert isinstance(runner.alg.policy.memory_a.rnn, torch.nn.GRU) and not runner.alg.policy.memory_a.native
