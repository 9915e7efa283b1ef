"""Teacher-student distillation (Distillation, StudentTeacher, StudentTeacherRecurrent, DistillationRunner)
against golden traces recorded from the unmodified reference (oracle/make_golden.py: scenario_distillation).

Two legs over the same checks:
  * `-m gpu`: the product path on cuda:0 (GEMM / LSTM / store / optimiser kernels of the library);
  * CPU: the same host code above a host-memory emulation of the C ABI (tests/host_emulation.py) -- this pins the host
    logic (what is launched on which buffers in which order, autograd wiring, hidden-state carrying, checkpoint rules)
    to the reference's results in the build container, where no GPU exists.  It is not a product path.
(The file name sorts after every other test module on purpose: these are the newest GPU tests.)
"""
import numpy as np
import pytest
import torch

from tests.golden_util import load, meta_of, rel_err

CASES = ["distill_mlp_small", "distill_lstm_small"]


def _runner(g, meta, device, parts=None):
    """DistillationRunner on an env that replays the recorded stream of `parts` (default: the regular iterations)."""
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import DistillationRunner
    T = meta["num_steps"]
    parts = parts or [f"it{i}." for i in range(meta["iterations"])]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    obs = {k: t(np.concatenate([g[f"{p}stream.obs.{k}"][:T + (p == parts[-1])] for p in parts]))
           for k in meta["group_dims"]}
    cat = lambda key: t(np.concatenate([g[f"{p}stream.{key}"] for p in parts]))
    env = ReplayVecEnv(obs, cat("rewards"), cat("dones"), None, meta["num_actions"], device=device)
    cfg = dict(num_steps_per_env=T, save_interval=10 ** 9, obs_groups={k: list(v) for k, v in meta["obs_groups"].items()},
               policy=dict(class_name=meta["policy_class"], student_hidden_dims=list(meta["hidden"]),
                           teacher_hidden_dims=list(meta["hidden"]), activation=meta["activation"], init_noise_std=0.1,
                           **meta["policy_extra"]),
               algorithm=dict(meta["alg"], class_name="Distillation"))
    return DistillationRunner(env, cfg, log_dir=None, device=device)


def _state(g, prefix, device):
    return {k[len(prefix):]: torch.from_numpy(g[k]).to(device) for k in g.files if k.startswith(prefix)}


def _check_distillation(name, device, tol):
    g = load(name)
    meta = meta_of(g)
    runner = _runner(g, meta, device)
    alg, pol, st = runner.alg, runner.alg.policy, runner.alg.storage
    assert [k for k, _ in pol.named_parameters()] == meta["param_names"]

    # the runner refuses to train before a teacher is loaded (reference distillation_runner.py:59-61)
    with pytest.raises(ValueError, match="Teacher model parameters not loaded"):
        runner.learn(1)
    # a PPO checkpoint fills the teacher only and does not resume (reference student_teacher.py:184-198)
    assert pol.load_state_dict(_state(g, "teacher_ckpt.", device)) is False and pol.loaded_teacher
    init = _state(g, "init.", device)
    for k, v in pol.state_dict().items():
        if k.startswith(("teacher", "memory_t")):
            assert torch.equal(v, init[k]), k
    assert not pol.teacher.training
    # a distillation checkpoint fills everything and resumes (:199-204); this also sets the recorded student
    assert pol.load_state_dict(init) is True
    with pytest.raises(ValueError, match="does not contain student or teacher"):
        pol.load_state_dict({"foo.weight": torch.zeros(1)})

    for it in range(meta["iterations"]):
        p = f"it{it}."
        zq = [torch.from_numpy(z).to(device) for z in g[p + "z_rollout"]]
        pol._noise_fn = lambda like, q=zq: q.pop(0)
        snap, trace = {}, []
        orig_update = alg.update

        def update_and_snapshot():
            for k in ("actions", "privileged_actions", "rewards", "dones"):
                snap[k] = getattr(st, k).cpu().numpy().copy()
            for k in meta["group_dims"]:
                snap["obs." + k] = st.observations[k].cpu().numpy().copy()
            return orig_update()

        alg.update = update_and_snapshot
        alg._mb_hook = lambda: trace.append(torch.cat([q.detach().reshape(-1) for q in pol.parameters()]).cpu().numpy())
        losses = {}
        orig_update_fn = alg.update

        def update_and_keep_loss():
            out = orig_update_fn()
            losses.update(out)
            return out

        alg.update = update_and_keep_loss
        runner.learn(1)
        alg.update = orig_update
        assert not zq, "every recorded draw was consumed"
        # rollout buffer: observations and dones are copies (bit-exact), actions = mean + std * z, teacher actions
        assert np.array_equal(snap["dones"], g[p + "storage.dones"])
        for k in meta["group_dims"]:
            assert np.array_equal(snap["obs." + k], g[p + "storage.obs." + k]), k
        assert np.array_equal(snap["rewards"], g[p + "storage.rewards"])
        assert rel_err(snap["actions"], g[p + "storage.actions"]) < tol
        assert rel_err(snap["privileged_actions"], g[p + "storage.privileged_actions"]) < tol
        # every optimiser step of the update (gradient_length steps of summed loss each, student-only clipping)
        ref_trace = g[p + "mb.params_after"]
        assert len(trace) == len(ref_trace), (len(trace), len(ref_trace))
        for i, (mine, ref) in enumerate(zip(trace, ref_trace)):
            assert rel_err(mine, ref) < 2.5 * tol, (it, i)
        assert abs(losses["behavior"] - float(g[p + "loss.behavior"])) < 5 * tol * max(1.0, abs(float(g[p + "loss.behavior"])))
        assert alg.num_updates == int(g[p + "num_updates"])
        for k, v in pol.state_dict().items():
            ref = g[p + "final." + k]
            if k.endswith("count"):
                assert int(v) == int(ref), k
            elif k.startswith(("teacher", "memory_t", "std", "log_std")):
                assert np.array_equal(v.cpu().numpy(), ref), k   # never trained: exactly as loaded
            else:
                assert rel_err(v.cpu().numpy(), ref) < 2.5 * tol, k
        hs = alg.last_hidden_states
        if pol.is_recurrent:
            for tag, h in zip(("s", "t"), hs):
                if h is None:   # the teacher's memory is reset at the start of update() and not advanced by it
                    assert p + f"hidden_{tag}0" not in g.files
                    continue
                for i, x in enumerate(h if isinstance(h, tuple) else (h,)):
                    assert not x.requires_grad
                    assert rel_err(x.cpu().numpy(), g[p + f"hidden_{tag}{i}"]) < 2.5 * tol, (tag, i)
        else:
            assert hs is None
    # optimiser checkpoint surface: a torch Adam state dict over policy.parameters()
    osd = alg.optimizer.state_dict()
    assert len(osd["param_groups"][0]["params"]) == len(meta["param_names"])
    return runner


def test_windowed_recurrent_distillation_matches_reference_golden_on_the_emulated_abi(monkeypatch):
    """RSLRL_DISTILL_WINDOWED=1 (the default): a recurrent student's update as one multi-step LSTM call per run of steps between two
    optimiser steps (trajectory split at dones, carried state for the trajectories that continue, state read back at the
    end of every environment's last trajectory) lands on the same golden as the step-by-step loop: windows of 3 over 8
    steps (the last two steps only advance the state), dones at 15 %, two iterations so the carried state matters."""
    from tests import host_emulation
    monkeypatch.setenv("RSLRL_DISTILL_WINDOWED", "1")
    lib = host_emulation.install(monkeypatch)
    runner = _check_distillation("distill_lstm_small", "cpu", 2e-5)
    assert runner.alg._windowed_recurrent and "trajectory_index" in lib.calls and "split_and_pad" in lib.calls
    # 2 iterations x (8 rollout steps x 2 memories + 3 runs in the update) multi-step calls instead of 2 x (16 + 8)
    assert lib.calls.count("lstm_step_forward") == 2 * (16 + 3)


@pytest.mark.parametrize("windowed", [False, True])
def test_recurrent_distillation_across_the_epoch_boundary_two_lstm_layers(windowed, monkeypatch):
    """Golden `distill_lstm2_cross`: a 2-layer LSTM student, a feed-forward teacher, windows of 3 over two epochs of 8
    steps -- the third window crosses the epoch boundary, where the reference re-installs the start-of-rollout hidden state
    while its loss sum and step counter carry on (distillation.py:110-112), and one trailing step only advances the state.
    Both update paths (step by step, and one multi-step LSTM call per run) land on it."""
    from tests import host_emulation
    monkeypatch.setenv("RSLRL_DISTILL_WINDOWED", "1" if windowed else "0")
    lib = host_emulation.install(monkeypatch)
    runner = _check_distillation("distill_lstm2_cross", "cpu", 2e-5)
    assert runner.alg._windowed_recurrent == windowed and runner.alg.policy.memory_s.rnn.num_layers == 2
    # per iteration: 8 rollout steps, then 16 stored steps = runs [0-2] [3-5] [6-7 | 0] [1-3] [4-6] [7]: 7 multi-step calls
    per_iter = (8 + 7) if windowed else (8 + 16)
    assert lib.calls.count("lstm_step_forward") == 2 * per_iter * 2      # (x 2 layers)


def _select_update_path(monkeypatch, batched):
    """batched: the defaults (a feed-forward student's window as one batch + grouped rollout launches, a recurrent
    student's window as one multi-step LSTM call); otherwise the step-by-step restatement of the reference's loop."""
    monkeypatch.setenv("RSLRL_DISTILL_BATCHED", "1" if batched else "0")
    monkeypatch.setenv("RSLRL_DISTILL_WINDOWED", "1" if batched else "0")


@pytest.mark.parametrize("name,batched", [("distill_mlp_small", True), ("distill_mlp_small", False),
                                          ("distill_lstm_small", True), ("distill_lstm_small", False),
                                          ("distill_mlp_adamw", True)])
def test_distillation_host_logic_matches_reference_golden_on_the_emulated_abi(name, batched, monkeypatch):
    """Both update paths of either student kind land on the reference's golden.  The feed-forward golden's windows
    (gradient_length 4 over 2 epochs of 10 steps) cross the epoch boundary; the recurrent golden has two trailing steps
    that only advance the hidden state."""
    from tests import host_emulation
    _select_update_path(monkeypatch, batched)
    lib = host_emulation.install(monkeypatch)
    runner = _check_distillation(name, "cpu", 2e-5)
    assert runner.alg._batched_update == (batched and "mlp" in name)
    assert runner.alg._windowed_recurrent == (batched and "lstm" in name)
    assert "store_transition" in lib.calls and "linear_backward_weight" in lib.calls
    # torch.optim.Adam is the fused kernel; AdamW / SGD / RMSprop are torch's optimisers on the flat-buffer views
    assert ("allreduce_clip_adam" in lib.calls) == (name != "distill_mlp_adamw")
    assert isinstance(runner.alg.optimizer, torch.optim.AdamW) == (name == "distill_mlp_adamw")
    if "lstm" in name:
        assert "lstm_step_backward" in lib.calls
    if name == "distill_mlp_small":
        # 2 iterations x (10 rollout steps + 2 epochs x 10 stored steps): layer launches per iteration
        fwd = lib.calls.count("linear_forward")
        assert fwd == (2 * (10 * 3 + 6 * 3) if batched else 2 * (10 * 6 + 20 * 3)), fwd


@pytest.mark.gpu
@pytest.mark.parametrize("name,batched", [("distill_mlp_small", True), ("distill_mlp_small", False),
                                          ("distill_lstm_small", True), ("distill_lstm_small", False)])
def test_distillation_matches_reference_golden(name, batched, monkeypatch):
    from rsl_rl_b200 import _native as nat
    _select_update_path(monkeypatch, batched)
    before = nat.launch_count()
    runner = _check_distillation(name, "cuda:0", 2e-5)
    assert runner.alg._batched_update == (batched and name == "distill_mlp_small")
    assert runner.alg._windowed_recurrent == (batched and name == "distill_lstm_small")
    assert nat.launch_count() > before   # the library's kernels ran (no silent framework path)


def test_distillation_resumes_from_a_reference_written_checkpoint(monkeypatch, tmp_path):
    """`tests/golden/distill_mlp_small_ref_checkpoint.pt` was written by the reference's `runner.save` after the golden's
    last regular iteration (its Adam state only has entries for the parameters that ever received a gradient: the student's);
    the reference then ran one more iteration (`resume.*`).  A fresh runner that loads the checkpoint reproduces that
    iteration step for step -- optimiser moments, bias-correction step count, normaliser statistics and the iteration
    counter all came back -- and its own checkpoint loads again with the same result."""
    import os
    from tests import host_emulation
    from tests.golden_util import GOLDEN
    host_emulation.install(monkeypatch)
    g = load("distill_mlp_small")
    meta = meta_of(g)
    ckpt = os.path.join(GOLDEN, "distill_mlp_small_ref_checkpoint.pt")
    ref = torch.load(ckpt, weights_only=False)
    assert sorted(ref["optimizer_state_dict"]["state"].keys()) == [1, 2, 3, 4, 5, 6]   # std (0) and the teacher: no state

    def resumed_runner(path):
        runner = _runner(g, meta, "cpu", parts=["resume."])
        runner.load(path)
        assert runner.alg.policy.loaded_teacher and runner.current_learning_iteration == int(g["resume.iter"])
        return runner

    def run_resumed_iteration(runner):
        alg, pol = runner.alg, runner.alg.policy
        zq = [torch.from_numpy(z) for z in g["resume.z_rollout"]]
        pol._noise_fn = lambda like, q=zq: q.pop(0)
        trace = []
        alg._mb_hook = lambda: trace.append(torch.cat([q.detach().reshape(-1) for q in pol.parameters()]).numpy().copy())
        runner.learn(1)
        assert len(trace) == len(g["resume.mb.params_after"])
        for i, (mine, want) in enumerate(zip(trace, g["resume.mb.params_after"])):
            assert rel_err(mine, want) < 5e-5, i
        for k, v in pol.state_dict().items():
            want = g["resume.final." + k]
            assert (int(v) == int(want)) if k.endswith("count") else rel_err(v.numpy(), want) < 5e-5, k

    first = resumed_runner(ckpt)
    for k, v in first.alg.policy.state_dict().items():
        assert np.array_equal(v.numpy(), g["it1.final." + k]), k          # the checkpoint IS the end of iteration 1
    opt = first.alg.optimizer
    assert opt._host_step == int(float(ref["optimizer_state_dict"]["state"][1]["step"])) == 10
    for i, p in enumerate(first.alg._group.params):
        m, v = first.alg._group.moment_views(i)
        if i in ref["optimizer_state_dict"]["state"]:
            assert torch.equal(m, ref["optimizer_state_dict"]["state"][i]["exp_avg"])
            assert torch.equal(v, ref["optimizer_state_dict"]["state"][i]["exp_avg_sq"])
        else:
            assert not m.any() and not v.any()
    mine = str(tmp_path / "model_0.pt")
    first.save(mine)                                   # our checkpoint of the same state, before the resumed iteration
    saved = torch.load(mine, weights_only=False)
    assert set(saved) == set(ref) and set(saved["model_state_dict"]) == set(ref["model_state_dict"])
    assert len(saved["optimizer_state_dict"]["param_groups"][0]["params"]) == 13
    run_resumed_iteration(first)
    run_resumed_iteration(resumed_runner(mine))
    # a PPO checkpoint through the same `load`: teacher only, nothing resumes, the optimiser is left alone
    ppo_ckpt = str(tmp_path / "ppo.pt")
    torch.save({"model_state_dict": {k[len("teacher_ckpt."):]: torch.from_numpy(g[k]) for k in g.files
                                     if k.startswith("teacher_ckpt.")},
                "optimizer_state_dict": {"state": {}, "param_groups": []}, "iter": 77, "infos": {"from": "ppo"}}, ppo_ckpt)
    fresh = _runner(g, meta, "cpu", parts=["resume."])
    assert fresh.load(ppo_ckpt) == {"from": "ppo"} and fresh.current_learning_iteration == 0
    assert fresh.alg.policy.loaded_teacher and fresh.alg.optimizer._host_step == 0
    assert np.array_equal(fresh.alg.policy.teacher[0].weight.detach().numpy(), g["teacher_ckpt.actor.0.weight"])


def test_gru_student_distillation_host_logic_on_the_emulated_abi(monkeypatch):
    """GRU memories on both sides (hidden state = one tensor instead of an (h, c) pair, reference memory.py:23,46-74): the
    step-by-step update -- the GRU recurrence is torch's (cuDNN on the GPU), everything around it is this library's."""
    from tests import host_emulation
    lib = host_emulation.install(monkeypatch)
    runner = _check_distillation("distill_gru_small", "cpu", 2e-5)
    alg = runner.alg
    assert not alg._windowed_recurrent and not alg._batched_update and not alg.policy.memory_s.native
    assert isinstance(alg.last_hidden_states[0], torch.Tensor) and "lstm_step_forward" not in lib.calls


def test_distillation_runner_logs_and_saves_like_the_on_policy_runner(monkeypatch, tmp_path, capsys):
    """`log_dir` set: the shared loop's logging path with a Distillation algorithm (loss dict `behavior`, no RND buffers),
    TensorBoard scalars, the console table, `model_<it>.pt` with the reference's checkpoint keys."""
    import os
    pytest.importorskip("tensorboard")
    from tests import host_emulation
    host_emulation.install(monkeypatch)
    g = load("distill_mlp_small")
    meta = meta_of(g)
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import DistillationRunner
    runner = _runner(g, meta, "cpu")
    cfg = dict(num_steps_per_env=meta["num_steps"], save_interval=1, obs_groups={k: list(v) for k, v in meta["obs_groups"].items()},
               policy=dict(class_name=meta["policy_class"], student_hidden_dims=list(meta["hidden"]),
                           teacher_hidden_dims=list(meta["hidden"]), activation=meta["activation"], init_noise_std=0.1,
                           **meta["policy_extra"]),
               algorithm=dict(meta["alg"], class_name="Distillation"))
    logged = DistillationRunner(runner.env, cfg, log_dir=str(tmp_path), device="cpu")
    logged.alg.policy.load_state_dict(_state(g, "init.", "cpu"))
    zq = [torch.from_numpy(z) for z in g["it0.z_rollout"]]
    logged.alg.policy._noise_fn = lambda like, q=zq: q.pop(0)
    logged.learn(1)
    out = capsys.readouterr().out
    assert "Learning iteration 0/1" in out and "behavior" in out and "Mean action noise std" in out
    ckpt = torch.load(os.path.join(str(tmp_path), "model_0.pt"), weights_only=False)
    assert set(ckpt) == {"model_state_dict", "optimizer_state_dict", "iter", "infos"} and ckpt["iter"] == 0
    assert any(f.startswith("events.out.tfevents") for f in os.listdir(str(tmp_path)))
    mine = torch.cat([q.detach().reshape(-1) for q in logged.alg.policy.parameters()]).numpy()
    assert rel_err(mine, g["it0.mb.params_after"][-1]) < 5e-5      # logging does not change the training


def test_distillation_storage_and_algorithm_argument_errors(monkeypatch):
    """Error behaviour of the reference: wrong generator for the training type (rollout_storage.py:153-154,161-162),
    unknown loss / optimiser names (distillation.py:71, utils.py:75), overflow of the rollout buffer (:79-80)."""
    from tests import host_emulation
    host_emulation.install(monkeypatch)
    from rsl_rl_b200.algorithms import Distillation
    from rsl_rl_b200.modules import StudentTeacher
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite
    obs = TensorDictLite({"policy": torch.zeros(4, 3)}, batch_size=[4])
    st = RolloutStorage("distillation", 4, 2, obs, [2], "cpu")
    assert not hasattr(st, "values") and st.privileged_actions.shape == (2, 4, 2)
    with pytest.raises(ValueError, match="only available for reinforcement learning"):
        next(st.mini_batch_generator(2))
    with pytest.raises(ValueError, match="only available for reinforcement learning"):
        next(st.recurrent_mini_batch_generator(2))
    rl = RolloutStorage("rl", 4, 2, obs, [2], "cpu")
    with pytest.raises(ValueError, match="only available for distillation"):
        next(rl.generator())
    with pytest.raises(ValueError, match="Unknown training_type"):
        RolloutStorage("other", 4, 2, obs, [2], "cpu")
    tr = RolloutStorage.Transition()
    tr.observations, tr.actions, tr.privileged_actions = obs, torch.ones(4, 2), 2 * torch.ones(4, 2)
    tr.rewards, tr.dones = torch.arange(4.0), torch.tensor([0, 1, 0, 1])
    st.add_transitions(tr)
    st.add_transitions(tr)
    with pytest.raises(OverflowError):
        st.add_transitions(tr)
    rows = list(st.generator())
    assert len(rows) == 2 and torch.equal(rows[1][2], 2 * torch.ones(4, 2)) and rows[1][3].dtype == torch.uint8
    assert rows[0][3].view(-1).tolist() == [0, 1, 0, 1] and rows[0][0]["policy"].shape == (4, 3)

    policy = StudentTeacher(obs, {"policy": ["policy"], "teacher": ["policy"]}, 2, student_hidden_dims=[8],
                            teacher_hidden_dims=[8])
    with pytest.raises(ValueError, match="Unknown loss type"):
        Distillation(policy, loss_type="l1", device="cpu")
    with pytest.raises(ValueError, match="Invalid optimizer"):
        Distillation(policy, optimizer="lion", device="cpu")
    alg = Distillation(policy, device="cpu", max_grad_norm=1.0)
    # the clipped span is exactly the student MLP's parameters inside the flat buffer
    sl = alg._student_slice
    flat_ids = torch.zeros(alg._group.numel)
    for p in policy.student.parameters():
        p.data.fill_(1.0)
    flat_ids[sl] = 1.0
    assert torch.equal((alg._group.values == 1.0).float() * flat_ids, (alg._group.values == 1.0).float())
    assert int((alg._group.values[sl] == 1.0).sum()) == sum(p.numel() for p in policy.student.parameters())


def test_distillation_has_no_cpu_path():
    from rsl_rl_b200.algorithms import Distillation
    from rsl_rl_b200.modules import StudentTeacher
    from rsl_rl_b200.utils import TensorDictLite
    obs = TensorDictLite({"policy": torch.zeros(4, 3)}, batch_size=[4])
    policy = StudentTeacher(obs, {"policy": ["policy"], "teacher": ["policy"]}, 2, student_hidden_dims=[8],
                            teacher_hidden_dims=[8])
    with pytest.raises(RuntimeError, match="no CPU"):
        Distillation(policy, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU|CUDA"):
        policy.act_inference(obs)
