#!/usr/bin/env python
"""TEST INFRASTRUCTURE (build container only: needs the reference checkout) -- differential check of the HOST LOGIC over a
matrix of configurations that the committed goldens do not cover one by one.

For every case the UNMODIFIED reference is run on CPU with `oracle/make_golden.py`'s recorders (a throw-away golden in a
temporary directory), then this library's host code replays the same env stream / noise / permutation above the
host-memory emulation of the C ABI (`tests/host_emulation.py`) with the assertions of the parity tests (rollout buffer,
LR sequence, parameters after the update(s), normaliser statistics, hidden states).  It exercises launch order, buffer
wiring, autograd wiring and configuration handling -- not the kernels (those are the `-m gpu` tests).

    python tools/differential_fuzz.py [--ref /root/reference] [--only NAME ...]

Matrix: PPO feed-forward engine: 6 activations, per-minibatch advantage normalisation, fixed schedule, no KL
target, env / minibatch counts that do not divide, observation groups, log-std + unclipped value loss, 1 and 4 hidden
layers; recurrent: 2-layer LSTM (engine), 2-layer GRU (general path), LSTM + normalisers; RND / symmetry: every
combination of the two symmetry switches, RND alone, step / linear weight schedules, reward normaliser, wide embedding
networks, a separate `rnd_state` group, RND + symmetry + per-minibatch normalisation; distillation (both update paths
each): no optimiser step at all (window longer than the update), windows of 1, 5 over 3 epochs with Huber + clip,
ReLU + log-std, teacher normaliser only, LSTM student with a feed-forward teacher and both normalisers, LSTM windows of
1 and 10, no dones, all dones.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

PPO_BASE = dict(num_envs=48, num_steps=6, group_dims={"policy": 11}, obs_groups={"policy": ["policy"], "critic": ["policy"]},
                num_actions=4, hidden=[24, 12], activation="elu", noise_std_type="scalar", alg_overrides={}, iterations=2)
REC = dict(policy_class="ActorCriticRecurrent", num_envs=32, p_done=0.15)
RND = dict(weight=1.0, weight_schedule=None, reward_normalization=False, state_normalization=False, learning_rate=1e-3,
           num_outputs=1, predictor_hidden_dims=[-1], target_hidden_dims=[-1])
E2 = dict(num_learning_epochs=2)
PPO_CASES = {
    **{f"act_{a}": dict(activation=a) for a in ("relu", "tanh", "selu", "sigmoid", "lrelu", "softplus")},
    "adv_norm_per_minibatch": dict(alg_overrides=dict(normalize_advantage_per_mini_batch=True)),
    "fixed_schedule": dict(alg_overrides=dict(schedule="fixed")),
    "no_kl_target": dict(alg_overrides=dict(desired_kl=None)),
    "odd_counts_and_groups": dict(num_envs=50, num_steps=5, alg_overrides=dict(num_mini_batches=3),
                                  group_dims={"policy": 7, "extra": 3},
                                  obs_groups={"policy": ["policy"], "critic": ["policy", "extra"]}),
    "log_std_unclipped_value": dict(noise_std_type="log", alg_overrides=dict(use_clipped_value_loss=False, max_grad_norm=0.3)),
    "wide_action_head": dict(num_actions=17),            # beyond the narrow-head kernel (N <= 12): general wgrad / dgrad
    "single_minibatch_single_epoch": dict(alg_overrides=dict(num_learning_epochs=1, num_mini_batches=1)),
    "hidden_dim_minus_one": dict(hidden=[-1, 16]),       # -1 = the input width (mlp.py:58-60)
    "one_hidden_layer": dict(hidden=[16]),
    "four_hidden_layers": dict(hidden=[32, 24, 16, 8]),
    # tf32 mode with shapes the one-launch fused forward covers: rollout launch plans with the Gaussian head in the
    # epilogue (observation rows 16-byte aligned: read in place; 11 columns: zero-padded staging + padded first-layer
    # weights), fused forward of the update storing the activations
    "tf32_fused_forward_aligned": dict(hidden=[128, 128, 32], group_dims={"policy": 12}, alg_extra=dict(gemm_precision="tf32"),
                                       alg_overrides=dict(num_learning_epochs=1, num_mini_batches=2)),
    "tf32_fused_forward_unaligned": dict(hidden=[128, 128, 32], alg_extra=dict(gemm_precision="tf32"),
                                         alg_overrides=dict(num_learning_epochs=1, num_mini_batches=2)),
    # the 16-bit storage chain (the bench default): fp16 weight shadow, activations, observations and layer gradients,
    # fused forward, loss with the dynamic power-of-two gradient scale, fused data-gradient chain, grouped weight
    # gradients, shadow refresh inside the optimiser step -- the emulation keeps an exact side copy of every fp16 buffer,
    # so a correct wiring reproduces the fp32 reference to fp32 accuracy
    "fp16_storage_chain": dict(hidden=[128, 128, 64], group_dims={"policy": 12}, alg_extra=dict(gemm_precision="fp16"),
                               alg_overrides=dict(num_learning_epochs=2, num_mini_batches=2)),
    "fp16_storage_chain_unaligned_obs": dict(hidden=[256, 128, 64], alg_extra=dict(gemm_precision="fp16"),
                                             alg_overrides=dict(num_learning_epochs=1, num_mini_batches=2)),
    "lstm_two_layers": dict(REC, policy_extra=dict(rnn_type="lstm", rnn_hidden_dim=12, rnn_num_layers=2)),
    "gru_two_layers": dict(REC, policy_extra=dict(rnn_type="gru", rnn_hidden_dim=12, rnn_num_layers=2)),
    "state_dependent_std_lstm": dict(REC, noise_std_type="log", general=True,
                                     policy_extra=dict(rnn_type="lstm", rnn_hidden_dim=12, state_dependent_std=True)),
    "state_dependent_std_groups_relu": dict(activation="relu", general=True, policy_extra=dict(state_dependent_std=True),
                                            group_dims={"policy": 7, "extra": 3},
                                            obs_groups={"policy": ["policy", "extra"], "critic": ["policy"]}),
    "lstm_normalisers": dict(REC, policy_extra=dict(rnn_type="lstm", rnn_hidden_dim=12, rnn_num_layers=1,
                                                    actor_obs_normalization=True, critic_obs_normalization=True)),
}
COMPOSITE_CASES = {
    "symmetry_only": dict(symmetry=True),
    "mirror_loss_without_augmentation": dict(symmetry=dict(use_data_augmentation=False, use_mirror_loss=True)),
    "augmentation_without_mirror_loss": dict(symmetry=dict(use_data_augmentation=True, use_mirror_loss=False)),
    "symmetry_for_logging_only": dict(symmetry=dict(use_data_augmentation=False, use_mirror_loss=False)),
    "rnd_only": dict(rnd_cfg=dict(RND)),
    "rnd_step_schedule": dict(rnd_cfg=dict(RND, weight_schedule=dict(mode="step", final_step=4, final_value=0.3))),
    "rnd_linear_schedule": dict(rnd_cfg=dict(RND, weight_schedule=dict(mode="linear", initial_step=2, final_step=9,
                                                                       final_value=0.1))),
    "rnd_reward_normaliser": dict(rnd_cfg=dict(RND, reward_normalization=True)),
    "rnd_wide_networks": dict(rnd_cfg=dict(RND, num_outputs=5, predictor_hidden_dims=[16, 8], target_hidden_dims=[12])),
    "rnd_state_group": dict(rnd_cfg=dict(RND), group_dims={"policy": 11, "rs": 5},
                            obs_groups={"policy": ["policy"], "critic": ["policy"], "rnd_state": ["rs"]}),
    "rnd_symmetry_adv_norm_per_minibatch": dict(rnd_cfg=dict(RND), symmetry=True,
                                                alg_overrides=dict(E2, normalize_advantage_per_mini_batch=True)),
}
DISTILL_BASE = dict(num_envs=24, num_steps=7, group_dims={"policy": 6, "priv": 3},
                    obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]}, teacher_groups=["policy", "priv"],
                    num_actions=2, hidden=[16, 8], activation="elu", policy_class="StudentTeacher", policy_extra={},
                    teacher_class="ActorCritic", teacher_extra={}, alg_overrides={}, iterations=2)
LSTM_PAIR = dict(policy_class="StudentTeacherRecurrent",
                 policy_extra=dict(rnn_type="lstm", rnn_hidden_dim=8, teacher_recurrent=True),
                 teacher_class="ActorCriticRecurrent", teacher_extra=dict(rnn_type="lstm", rnn_hidden_dim=8), p_done=0.25)
DISTILL_CASES = {
    "window_longer_than_update": dict(),
    "window_of_one": dict(alg_overrides=dict(gradient_length=1, num_learning_epochs=2)),
    "window_5_three_epochs_huber_clip": dict(alg_overrides=dict(gradient_length=5, num_learning_epochs=3, max_grad_norm=0.1,
                                                                loss_type="huber")),
    "relu_log_std": dict(activation="relu", policy_extra=dict(noise_std_type="log"), alg_overrides=dict(gradient_length=3)),
    "teacher_normaliser_only": dict(policy_extra=dict(teacher_obs_normalization=True),
                                    teacher_extra=dict(actor_obs_normalization=True), alg_overrides=dict(gradient_length=3)),
    "lstm_student_ff_teacher_normalisers": dict(
        policy_class="StudentTeacherRecurrent",
        policy_extra=dict(rnn_type="lstm", rnn_hidden_dim=8, rnn_num_layers=1, teacher_recurrent=False,
                          student_obs_normalization=True, teacher_obs_normalization=True),
        teacher_extra=dict(actor_obs_normalization=True), alg_overrides=dict(gradient_length=4, num_learning_epochs=2),
        p_done=0.25),
    "lstm_window_of_one": dict(LSTM_PAIR, alg_overrides=dict(gradient_length=1)),
    "lstm_window_10_three_epochs": dict(LSTM_PAIR, alg_overrides=dict(gradient_length=10, num_learning_epochs=3,
                                                                      max_grad_norm=0.2)),
    "lstm_no_dones": dict(LSTM_PAIR, alg_overrides=dict(gradient_length=3, num_learning_epochs=2), p_done=0.0),
    "lstm_all_dones": dict(LSTM_PAIR, alg_overrides=dict(gradient_length=3, num_learning_epochs=2), p_done=1.0),
}


class _Setattr:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    import make_golden as mg
    mg._import_reference(args.ref)
    out = tempfile.mkdtemp(prefix="rslrl_fuzz_")
    import tests.golden_util as golden_util
    golden_util.GOLDEN = out                                   # the parity checks read their traces from here
    from tests import host_emulation
    from tests import test_zz_distillation, test_zz_policy_variants
    from tests.test_gpu_composite_parity import check_composite_golden
    from tests.test_gpu_ppo_parity import check_learn_iterations_against_golden

    def ppo(name, over, seed):
        over = dict(over)
        general = over.pop("general", False)     # configurations the engines do not take: modules + autograd functions
        alg_extra = over.pop("alg_extra", {})
        mg.scenario_ppo_mlp(out, name, seed=seed, **dict(PPO_BASE, **over))
        lib = host_emulation.install(_Setattr())
        if "policy_class" in over or general:
            r = test_zz_policy_variants._check_ppo_variant(name, "cpu", general_path=False, rng_compat=False)
        else:
            r = check_learn_iterations_against_golden(name, "cpu", rng_compat=False, **alg_extra)
        path = "engine" if r.alg.fused else ("recurrent engine" if r.alg._recurrent_fused else "general path")
        if "mlp3_plan_run" in lib.calls:
            path += f" (one-launch rollout plan x{lib.calls.count('mlp3_plan_run')}, fused forward x{lib.calls.count('mlp3_forward')})"
        return path

    def composite(name, over, seed):
        mg.scenario_ppo_mlp(out, name, seed=seed, **dict(dict(PPO_BASE, alg_overrides=E2), **over))
        host_emulation.install(_Setattr())
        r = check_composite_golden(name, "cpu", rng_compat=False)
        return "composite engine" if r.alg._composite_fusable() else "general path"

    def distill(name, over, seed):
        mg.scenario_distillation(out, name, seed=seed, **dict(DISTILL_BASE, **over))
        paths = []
        for fast in ("1", "0"):
            os.environ["RSLRL_DISTILL_BATCHED"] = os.environ["RSLRL_DISTILL_WINDOWED"] = fast
            host_emulation.install(_Setattr())
            r = test_zz_distillation._check_distillation(name, "cpu", 2e-5)
            paths.append("batched" if r.alg._batched_update else ("windowed" if r.alg._windowed_recurrent else "step loop"))
        return " + ".join(paths)

    failed = 0
    for family, fn, cases in (("ppo", ppo, PPO_CASES), ("composite", composite, COMPOSITE_CASES),
                              ("distill", distill, DISTILL_CASES)):
        for name, over in cases.items():
            full = f"{family}_{name}"
            if args.only and full not in args.only and name not in args.only:
                continue
            # The clipped objective is discontinuous: a recorded trace in which one sample's probability ratio sits on the
            # clip boundary (to ~1e-7) replays differently under fp32 summation-order noise -- observed once in this
            # matrix (one minibatch gradient off by a single sample's contribution, every other step exact to 1e-7, and
            # 1e-7 perturbations of the initial weights toggle it).  A case therefore gets up to three differently
            # seeded reference traces; a host-logic difference fails all of them.
            errors = []
            for seed in (len(name) * 7, len(name) * 7 + 1, len(name) * 7 + 2):
                try:
                    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                        path = fn(full, over, seed)
                    note = f"  (seed {seed}; knife-edge trace(s) before: {len(errors)})" if errors else ""
                    print(f"PASS  {full:55s} {path}{note}", flush=True)
                    break
                except AssertionError:
                    errors.append(traceback.format_exc())
                except Exception:
                    errors = [traceback.format_exc()] * 3
                    break
            else:
                pass
            if len(errors) >= 3:
                failed += 1
                last = errors[-1].strip().splitlines()[-8:]
                print(f"FAIL  {full:55s} " + " | ".join(l.strip() for l in last), flush=True)
    print(f"{failed} failed")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
