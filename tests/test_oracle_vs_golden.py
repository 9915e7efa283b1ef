"""CPU: pins oracle/ppo_oracle.py against vectors produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import ppo_oracle as O
from tests.golden_util import cfg_of, flat, load, meta_of, params_from, rel_err, stream_of


def test_gae_scan_is_bit_exact_and_normalisation_1e6():
    g = load("gae_cases")
    names = sorted({k.rsplit(".", 1)[0] for k in g.files if k.endswith(".returns")})
    assert len(names) >= 13
    for n in names:
        gamma, lam = (0.9, 0.8) if n.startswith("odd") else (0.99, 0.95)
        ret, adv = O.compute_returns(g[n + ".rewards"], g[n + ".values"], g[n + ".dones"], g[n + ".last_values"],
                                     gamma, lam, normalize_advantage=n.endswith("norm"))
        assert np.array_equal(ret, g[n + ".returns"]), n  # same fp32 op order -> identical bits
        if n.endswith("raw"):
            assert np.array_equal(adv, g[n + ".advantages"]), n
        else:
            assert rel_err(adv, g[n + ".advantages"]) < 1e-6, n


@pytest.mark.parametrize("tag", ["clipped", "unclipped"])
def test_loss_values_and_hand_derived_gradients(tag):
    g = load("loss_terms")
    f = lambda k: g[f"{tag}.{k}"]
    out = O.ppo_loss_fwd_bwd(f("mu"), f("sigma"), f("value"), f("actions"), f("old_logp"), f("advantages"),
                             f("returns"), f("old_values"), 0.2, 1.0, 0.01, tag == "clipped")
    assert abs(out["surrogate"] - float(f("surrogate"))) < 1e-6 * max(1, abs(float(f("surrogate"))))
    assert abs(out["value_loss"] - float(f("value_loss"))) < 1e-5 * abs(float(f("value_loss")))
    assert abs(out["entropy"] - float(f("entropy"))) < 1e-6 * abs(float(f("entropy")))
    assert rel_err(out["logp"], f("logp")) < 1e-6
    assert rel_err(out["dmu"], f("dmu")) < 1e-5
    assert rel_err(out["dsigma"], f("dsigma")) < 1e-5
    assert rel_err(out["dvalue"], f("dvalue")) < 1e-5
    kl = O.kl_mean(f("mu"), f("sigma"), f("old_mu"), f("old_sigma"))
    assert abs(kl - float(f("kl"))) < 1e-5 * abs(float(f("kl")))
    # the ratio must leave the clip range on both sides for the test to mean anything
    assert (out["ratio"] > 1.2).any() and (out["ratio"] < 0.8).any()


@pytest.mark.parametrize("name", ["ppo_mlp_small", "ppo_mlp_groups"])
def test_full_learn_iterations_match_reference(name):
    g = load(name)
    meta = meta_of(g)
    cfg = cfg_of(meta)
    names = meta["param_names"]
    params = params_from(g, "init", names)
    adam = O.adam_init(params)
    lr = cfg["learning_rate"]
    for it in range(meta["iterations"]):
        p = f"it{it}."
        st, last_values, loss_dict, lr, trace = O.learn_iteration(
            params, adam, lr, stream_of(g, it, meta), g[p + "z_rollout"], g[p + "perm"], cfg)
        # rollout products
        assert rel_err(st["actions"], g[p + "storage.actions"]) < 1e-5
        assert rel_err(st["values"], g[p + "storage.values"]) < 1e-5
        assert rel_err(st["actions_log_prob"], g[p + "storage.actions_log_prob"]) < 1e-5
        assert rel_err(st["mu"], g[p + "storage.mu"]) < 1e-5
        assert rel_err(st["sigma"], g[p + "storage.sigma"]) < 1e-6
        assert rel_err(st["rewards"], g[p + "storage.rewards"]) < 1e-6
        assert np.array_equal(st["dones"], g[p + "storage.dones"])
        assert rel_err(st["returns"], g[p + "storage.returns"]) < 1e-5
        assert rel_err(st["advantages"], g[p + "storage.advantages"]) < 1e-5
        # per-minibatch trace
        n_mb = len(trace)
        assert n_mb == g[p + "mb.lr"].shape[0]
        assert np.allclose([t["lr"] for t in trace], g[p + "mb.lr"], rtol=1e-12), "KL-adaptive lr sequence"
        assert rel_err(trace[0]["grads"], g[p + "mb.raw_grads"][0]) < 1e-5
        for i in range(n_mb):
            assert rel_err(trace[i]["grads"], g[p + "mb.raw_grads"][i]) < 2e-4, i
            assert abs(trace[i]["total_norm"] - g[p + "mb.total_norm"][i]) < 2e-4 * g[p + "mb.total_norm"][i]
        # post-update state
        assert rel_err(flat(params), g[p + "mb.params_after"][-1]) < 1e-5
        assert rel_err(np.concatenate([v.reshape(-1) for v in adam["exp_avg"].values()]), g[p + "adam.exp_avg"]) < 1e-4
        assert rel_err(np.concatenate([v.reshape(-1) for v in adam["exp_avg_sq"].values()]), g[p + "adam.exp_avg_sq"]) < 1e-4
        assert adam["step"] == int(g[p + "adam.step"])
        assert abs(lr - float(g[p + "learning_rate"])) < 1e-15


def test_trajectory_split_pad_unpad_bit_exact():
    g = load("trajectory_cases")
    for n in ("small", "none", "all", "c4like"):
        padded, masks = O.split_and_pad_trajectories(g[n + ".x"], g[n + ".dones"])
        assert np.array_equal(masks, g[n + ".masks"]), n
        assert np.array_equal(padded, g[n + ".padded"]), n
        assert np.array_equal(O.unpad_trajectories(padded, masks), g[n + ".unpadded"]), n
        assert np.array_equal(O.last_was_done(g[n + ".dones"]), g[n + ".last_was_done"]), n
        assert np.array_equal(O.hidden_at_trajectory_starts(g[n + ".saved_hidden"], g[n + ".dones"]),
                              g[n + ".hidden_at_starts"]), n
        assert masks.sum() == g[n + ".x"].shape[0] * g[n + ".x"].shape[1]


def test_torch_cpu_port_matches_numpy_oracle():
    """oracle/torch_cpu_port.py (the CPU-baseline leg of bench.py) computes the same iteration as the
    pinned numpy oracle."""
    import torch
    from collections import OrderedDict
    from oracle import torch_cpu_port as P

    params, stream, z, perm = P.random_problem(96, T=8, obs_dim=12, A=4, hidden=(32, 16), seed=3)
    cfg = dict(P.DEFAULT_CFG, num_learning_epochs=2)
    np_params = OrderedDict((k, v.numpy().copy()) for k, v in params.items())
    np_stream = dict(obs_actor=stream["obs"].numpy(), obs_critic=stream["obs"].numpy(), rewards=stream["rewards"].numpy(),
                     dones=stream["dones"].numpy(), time_outs=stream["time_outs"].numpy() > 0)
    ocfg = dict(cfg, activation="elu", noise_std_type="scalar")
    st, _, loss_o, lr_o, _ = O.learn_iteration(np_params, O.adam_init(np_params), 1e-3, np_stream, z.numpy(),
                                               perm.numpy(), ocfg)
    loss_t, lr_t, extra = P.learn_iteration(params, P.adam_init(params), 1e-3, stream, z, perm, cfg)
    assert lr_t == lr_o
    assert rel_err(extra["returns"].numpy(), st["returns"]) < 1e-6
    assert rel_err(extra["advantages"].numpy(), st["advantages"]) < 1e-5
    for k in ("value_function", "surrogate", "entropy"):
        assert abs(loss_t[k] - loss_o[k]) < 1e-5 * max(1.0, abs(loss_o[k]))
    a = np.concatenate([v.numpy().reshape(-1) for v in params.values()])
    b = np.concatenate([v.reshape(-1) for v in np_params.values()])
    assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("name", ["distill_mlp_small", "distill_mlp_adamw"])
def test_distillation_oracle_matches_reference(name):
    """oracle/distill_oracle.py (feed-forward student: rollout with both normalisers, windows of `gradient_length` steps
    that carry across the epoch boundary, student-only clipping, Adam / AdamW) against the reference's recorded trace:
    rollout buffer, the student's parameters after EVERY optimiser step, the reported loss, the normaliser statistics."""
    from oracle import distill_oracle as D
    g = load(name)
    meta = meta_of(g)
    cfg = dict(meta["alg"], activation=meta["activation"])
    state = {k[len("init."):]: g[k].copy() for k in g.files if k.startswith("init.")}
    student = [k for k in meta["param_names"] if k.startswith("student.")]
    adam = O.adam_init({k: state[k] for k in student})
    step = None
    if cfg["optimizer"] == "adamw":
        step = lambda params, grads: D.adamw_step(params, grads, adam, cfg["learning_rate"])
    # where the student sits inside the reference's flat parameter trace (policy.parameters() order)
    sizes = [int(np.prod(g["init." + k].shape)) for k in meta["param_names"]]
    first = sum(s for k, s in zip(meta["param_names"], sizes) if meta["param_names"].index(k) < meta["param_names"].index(student[0]))
    span = slice(first, first + sum(s for k, s in zip(meta["param_names"], sizes) if k in student))
    cat = lambda p, groups: np.concatenate([g[f"{p}stream.obs.{k}"] for k in groups], axis=-1)
    parts = [f"it{i}." for i in range(meta["iterations"])] + (["resume."] if "resume.z_rollout" in g.files else [])
    for p in parts:
        st = D.rollout(state, cat(p, meta["obs_groups"]["policy"]), cat(p, meta["obs_groups"]["teacher"]), g[p + "z_rollout"],
                       g[p + "stream.rewards"], g[p + "stream.dones"], cfg)
        assert rel_err(st["actions"], g[p + "storage.actions"]) < 1e-5
        assert rel_err(st["privileged_actions"], g[p + "storage.privileged_actions"]) < 1e-5
        assert np.array_equal(st["dones"], g[p + "storage.dones"]) and np.array_equal(st["rewards"], g[p + "storage.rewards"])
        loss, trace = D.update(state, adam, st, cfg, optimizer_step=step)
        ref = g[p + "mb.params_after"]
        assert len(trace) == len(ref)
        for i, (mine, want) in enumerate(zip(trace, ref)):
            assert rel_err(mine, want[span]) < 1e-5, (p, i)
        assert abs(loss - float(g[p + "loss.behavior"])) < 1e-6 * max(1.0, abs(loss))
        for k in state:
            if "normalizer" in k and not k.endswith("count"):
                assert rel_err(np.asarray(state[k]).reshape(-1), g[p + "final." + k].reshape(-1)) < 1e-5, k
    if "mb.total_norm" in " ".join(g.files):   # the clip saw the student's gradient only
        assert g["it0.mb.raw_grads"].shape[1] == span.stop - span.start
