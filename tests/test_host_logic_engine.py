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
