"""CPU: host logic of the feed-forward PPO path on the fused engine (rollout launch plans, in-place rollout rows, the
one-gather-per-update permuted copy, per-minibatch forward / loss / backward / fused optimiser step sequencing, observation
groups and normalisers) above the host-memory emulation of the C ABI (tests/host_emulation.py), replaying the golden traces
of the unmodified reference with the same checks the `-m gpu` tests apply to the real kernels
(tests/test_gpu_ppo_parity.py::check_learn_iterations_against_golden).  CUDA graphs and the CUDA generator are not
available here: `use_cuda_graph=False`, `rng_compat=False` (the discarded draws of update() only advance the generator)."""
import pytest


@pytest.mark.parametrize("name", ["ppo_mlp_small", "ppo_mlp_groups"])
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_fused_engine_host_logic_replays_the_reference_golden(name, precision, monkeypatch):
    from tests import host_emulation
    from tests.test_gpu_ppo_parity import check_learn_iterations_against_golden
    lib = host_emulation.install(monkeypatch)
    runner = check_learn_iterations_against_golden(name, "cpu", gemm_precision=precision, rng_compat=False)
    alg = runner.alg
    assert alg.fused and alg.rollout_fused
    calls = set(lib.calls)
    assert {"gaussian_act", "store_step", "gae_norm", "permute_gather", "ppo_loss_fwd_bwd", "linear_backward_head",
            "allreduce_clip_adam"} <= calls
    assert "store_transition" not in calls      # every rollout row was written in place by act() + the per-step plan
    M, E = alg.num_mini_batches, alg.num_learning_epochs
    assert lib.calls.count("ppo_loss_fwd_bwd") == lib.calls.count("allreduce_clip_adam")   # one loss, one step per minibatch


@pytest.mark.parametrize("name", ["ppo_rnd_sym_small", "ppo_rnd_norm_small", "ppo_lstm_small"])
def test_composite_paths_host_logic_replays_the_reference_golden(name, monkeypatch):
    """RND + symmetry augmentation + mirror loss, RND with state / reward normalisers (both on the fused engine: one
    permuted gather per update, loss over k * minibatch rows without a materialised repeat, mirror-loss gradient added to
    dL/dmu, the RND predictor's MSE under autograd, two fused optimiser steps per minibatch) and the recurrent policy, with
    the checks of tests/test_gpu_composite_parity.py::check_composite_golden."""
    from tests import host_emulation
    from tests.test_gpu_composite_parity import check_composite_golden
    lib = host_emulation.install(monkeypatch)
    runner = check_composite_golden(name, "cpu", rng_compat=False)
    calls = set(lib.calls)
    if "rnd" in name:
        assert runner.alg._composite_fusable() and {"rnd_reward", "ppo_loss_fwd_bwd", "permute_gather"} <= calls
        assert lib.calls.count("allreduce_clip_adam") == 2 * lib.calls.count("ppo_loss_fwd_bwd")   # policy + RND predictor


def test_observation_normalisers_on_the_engine_path_replay_the_reference_golden(monkeypatch):
    """actor / critic observation normalisers over four iterations: the running statistics change every env step while the
    rollout launch plans hold the buffers' addresses; the update normalises the gathered minibatch observations straight
    into the padded input buffers (`normalizer(x, out=...)`)."""
    from tests import host_emulation
    from tests.test_gpu_ppo_parity import check_normaliser_golden
    lib = host_emulation.install(monkeypatch)
    runner = check_normaliser_golden("cpu", False, rng_compat=False)
    assert runner.alg.fused and {"empirical_norm_update", "empirical_normalize"} <= set(lib.calls)


def test_logging_path_reproduces_the_reference_buffers(monkeypatch, tmp_path):
    """SURVEY 8 f1 above the emulated ABI: device-side episode statistics flushed once per iteration against the reference's
    `rewbuffer` / `lenbuffer` and logged loss means, TensorBoard events, the final checkpoint."""
    pytest.importorskip("tensorboard")
    from tests import host_emulation
    from tests.test_gpu_ppo_parity import check_logging_path
    host_emulation.install(monkeypatch)
    check_logging_path("cpu", tmp_path, rng_compat=False)


def test_reference_written_checkpoint_round_trip(monkeypatch, tmp_path):
    """SURVEY 8 f3 above the emulated ABI: the checkpoint the reference wrote loads (parameters, normaliser buffers, Adam
    moments / step / lr, iteration), training resumes, the checkpoint written here has the reference's structure -- and the
    reference loads it when its install is present."""
    pytest.importorskip("tensorboard")
    from tests import host_emulation
    from tests.test_gpu_ppo_parity import check_reference_checkpoint_round_trip
    host_emulation.install(monkeypatch)
    check_reference_checkpoint_round_trip("cpu", tmp_path, rng_compat=False)


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_fused_forward_and_16_bit_storage_chain_host_logic(precision, monkeypatch):
    """The tensor-core modes' own host paths at a shape they cover (golden `ppo_mlp_fused_shape`: MLP [128, 128, 64]):
    one-launch rollout plans with the Gaussian head in the epilogue and the fused forward (`tf32`), and on top of that the
    16-bit storage chain of the bench default (`fp16`): fp16 weight shadow refreshed inside the optimiser step, fp16
    observations / activations / layer gradients, the loss launch's dynamic power-of-two gradient scale and its inverse in
    the weight-gradient launch, the fused data-gradient chain, all weight gradients in one grouped launch.  The emulation
    computes in fp32 and keeps an exact side copy of every fp16 buffer (tests/host_emulation._HalfStore), so a correct wiring
    lands on the fp32 reference at fp32 tolerances; what the storage rounding itself costs is the GPU tests' subject."""
    from collections import Counter
    from tests import host_emulation
    from tests.test_gpu_ppo_parity import check_learn_iterations_against_golden
    lib = host_emulation.install(monkeypatch)
    runner = check_learn_iterations_against_golden("ppo_mlp_fused_shape", "cpu", gemm_precision=precision, rng_compat=False)
    eng, n = runner.alg._engine, Counter(lib.calls)
    assert eng.fused_fwd and n["mlp3_plan_run"] == 6 and n["gaussian_act"] == 0      # one launch per env step
    assert n["mlp3_forward"] == 4 + 1                                               # 4 minibatches + the last values
    if precision == "fp16":
        assert eng.f16 and eng.head16 and eng.dgrad_chain and eng.scale_state is not None
        assert n["dgrad_chain16"] == 4 and n["linear_backward_weight"] == 4 and n["linear_backward_head"] == 0
        assert n["ppo_loss_fwd_bwd"] == 4 + 1                                       # + the one-off scale calibration
        assert n["convert_f32_f16"] >= 2                                            # weight shadow, minibatch observations
        assert float(eng.scale_state[0]) * float(eng.scale_state[1]) == 1.0 and float(eng.scale_state[0]) >= 1.0
