"""CPU: host-side logic that needs no kernels -- obs container, config resolution, flat parameter
groups and optimiser state compatibility."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from rsl_rl_b200.algorithms.flat_params import FlatParamGroup, FusedAdam
from rsl_rl_b200.utils import TensorDictLite, resolve_nn_activation, resolve_obs_groups, string_to_callable


def test_tensordict_lite_surface():
    td = TensorDictLite({"a": torch.arange(24.).view(4, 3, 2), "b": torch.arange(12.).view(4, 3)}, batch_size=[4, 3])
    assert td[1]["a"].shape == (3, 2) and td[1].batch_size == torch.Size([3])
    assert td[:, 1:3].batch_size == torch.Size([4, 2])
    assert td[:, torch.tensor(1):torch.tensor(3)]["b"].shape == (4, 2)
    flat = td.flatten(0, 1)
    assert flat.batch_size == torch.Size([12]) and flat["a"].shape == (12, 2)
    idx = torch.tensor([5, 0, 7])
    assert torch.equal(flat[idx]["b"], td["b"].flatten()[idx])
    assert flat[idx].batch_size == torch.Size([3])
    assert "a" in td and set(td.keys()) == {"a", "b"}
    clone = td.clone()
    clone["a"].zero_()
    assert td["a"].sum() > 0
    td[0].copy_(TensorDictLite({"a": torch.ones(3, 2), "b": torch.ones(3)}, batch_size=[3]))
    assert td["a"][0].eq(1).all()


def test_resolve_obs_groups_defaults_and_errors():
    obs = TensorDictLite({"policy": torch.zeros(2, 3), "critic": torch.zeros(2, 4)}, batch_size=[2])
    with pytest.warns(UserWarning):
        out = resolve_obs_groups(obs, {"policy": ["policy"]}, ["critic"])
    assert out["critic"] == ["critic"]
    obs2 = TensorDictLite({"policy": torch.zeros(2, 3)}, batch_size=[2])
    with pytest.warns(UserWarning):
        out = resolve_obs_groups(obs2, {"policy": ["policy"]}, ["critic", "rnd_state"])
    assert out["critic"] == ["policy"] and out["rnd_state"] == ["policy"]
    with pytest.raises(ValueError):
        resolve_obs_groups(obs2, {"policy": ["missing"]}, [])
    with pytest.raises(ValueError):
        resolve_obs_groups(obs2, {"policy": []}, [])
    with pytest.raises(ValueError):
        resolve_obs_groups(TensorDictLite({"x": torch.zeros(2, 1)}, batch_size=[2]), {}, [])


def test_activation_resolution():
    assert isinstance(resolve_nn_activation("ELU"), nn.ELU)
    with pytest.raises(ValueError):
        resolve_nn_activation("nope")
    with pytest.raises(NotImplementedError):
        resolve_nn_activation("gelu")
    assert string_to_callable("math:sqrt")(4.0) == 2.0
    with pytest.raises(ValueError):
        string_to_callable("math:pi")


def _toy():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(5, 7), nn.ELU(), nn.Linear(7, 1))


def test_flat_param_group_views_and_rebinding():
    net = _toy()
    before = [p.detach().clone() for p in net.parameters()]
    g = FlatParamGroup(net.parameters(), "cpu")
    # 32-byte tensor starts: float4 / TMA alignment in fp32 and 16 bytes in the fp16 weight shadow of the same offsets
    assert g.numel % 8 == 0 and all(o % 8 == 0 for o in g.offsets)
    assert g.values16 is None and g.shadow16_stale       # the shadow only exists in gemm_precision="fp16" mode
    g.values16 = torch.zeros(g.numel, dtype=torch.float16)
    for q in net.parameters():
        v = g.shadow16(q)
        assert v.shape == q.shape and v.data_ptr() % 16 == g.values16.data_ptr() % 16
        assert (v.data_ptr() - g.values16.data_ptr()) // 2 == (q.data_ptr() - g.values.data_ptr()) // 4
    g.values16 = None
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p, b)
    assert g.is_bound()
    g.values.zero_()
    assert all(float(p.abs().sum()) == 0 for p in net.parameters())      # params are views
    g.grads.fill_(2.0)
    assert all(float(p.grad.mean()) == 2.0 for p in net.parameters())    # grads are views
    # load_state_dict copies in place -> stays bound; .to()/zero_grad(set_to_none) unbinds -> rebind works
    net.load_state_dict({k: torch.ones_like(v) for k, v in net.state_dict().items()})
    assert g.is_bound() and float(g.values[:35].min()) == 1.0
    net.zero_grad(set_to_none=True)
    assert not g.is_bound()
    g.ensure_bound()
    assert g.is_bound()


def test_fused_adam_is_a_torch_adam_with_compatible_state_dict():
    net = _toy()
    g = FlatParamGroup(net.parameters(), "cpu")
    opt = FusedAdam(g, lr=1e-3)
    assert isinstance(opt, torch.optim.Adam)
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    # a stock optimiser's checkpoint loads into the fused one and lands in the flat moment buffers
    ref_net = _toy()
    ref = torch.optim.Adam(ref_net.parameters(), lr=5e-4)
    for _ in range(3):
        ref.zero_grad()
        ref_net(torch.randn(8, 5)).pow(2).mean().backward()
        ref.step()
    opt.load_state_dict(ref.state_dict())
    assert opt._host_step == 3 and opt.param_groups[0]["lr"] == 5e-4
    flat_m = torch.cat([ref.state[p]["exp_avg"].reshape(-1) for p in ref_net.parameters()])
    mine = torch.cat([g.moment_views(i)[0].reshape(-1) for i in range(len(g.params))])
    assert torch.equal(flat_m, mine)
    # stock step() keeps working on the views
    net.load_state_dict(ref_net.state_dict())
    g.grads.normal_()
    opt.step()
    assert g.is_bound()


def test_fp16_storage_has_the_operand_precision_of_tf32():
    """within fp16's normal range a value stored as fp16 carries the same 10-bit mantissa the tensor core
    keeps of a tf32 operand (relative rounding error <= 2^-11 for both), and the power-of-two gradient scale of the
    fp16 path is exact."""
    import numpy as np
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-6, 6, 200000))).astype(np.float32)
    x = x[(np.abs(x) > 6.2e-5) & (np.abs(x) < 6.0e4)]                      # fp16 normal range
    f16 = x.astype(np.float16).astype(np.float32)
    u = x.view(np.uint32)
    tf32 = ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)   # round-to-nearest tf32 (ties away)
    rel16 = np.abs(f16 - x) / np.abs(x)
    rel32 = np.abs(tf32 - x) / np.abs(x)
    assert rel16.max() <= 2.0 ** -11 and rel32.max() <= 2.0 ** -11
    ties = (u & np.uint32(0x1FFF)) == np.uint32(0x1000)                    # the two roundings only differ on exact ties
    assert np.array_equal(f16[~ties], tf32[~ties])
    S = np.float32(2.0 ** 15)
    g = (rng.standard_normal(1000) * 1e-5).astype(np.float32)
    assert np.array_equal((g * S) / S, g)                                  # scaling by a power of two is exact


def test_policy_modules_export_with_torchscript_and_tracing():
    """Deployment surface: downstream play / export scripts deep-copy `policy.actor` (and the observation
    normaliser, and `memory_a.rnn` of a recurrent policy), move them to the CPU and `torch.jit.script` or trace them
    (ONNX export traces).  The kernel-backed modules script / trace to the plain layer chain; calling them eagerly on
    host tensors still raises (there is no CPU compute path)."""
    import copy

    import pytest
    import torch

    from rsl_rl_b200.modules import ActorCritic, ActorCriticRecurrent, StudentTeacher
    from rsl_rl_b200.utils import TensorDictLite

    obs = TensorDictLite({"policy": torch.randn(4, 10)}, batch_size=[4])
    groups = {"policy": ["policy"], "critic": ["policy"], "teacher": ["policy"]}
    torch.manual_seed(0)
    pol = ActorCritic(obs, groups, 3, actor_hidden_dims=[16, 8], critic_hidden_dims=[16, 8], actor_obs_normalization=True)
    with torch.no_grad():
        pol.actor_obs_normalizer._mean.normal_()
        pol.actor_obs_normalizer._std.uniform_(0.5, 2.0)

    class Exporter(torch.nn.Module):   # the shape of Isaac Lab's _TorchPolicyExporter
        def __init__(self, actor, normalizer):
            super().__init__()
            self.actor = copy.deepcopy(actor)
            self.normalizer = copy.deepcopy(normalizer)

        def forward(self, x):
            return self.actor(self.normalizer(x))

    x = torch.randn(5, 10)
    nrm = pol.actor_obs_normalizer

    def plain(mlp, inp):   # (iterating the Sequential, not .children(): one activation module serves every layer)
        for m in mlp:
            inp = m(inp)
        return inp

    want = plain(pol.actor, (x - nrm._mean) / (nrm._std + nrm.eps))
    exporter = Exporter(pol.actor, nrm).to("cpu")
    scripted = torch.jit.script(exporter)
    assert torch.equal(scripted(x), want)
    traced = torch.jit.trace(exporter, x, check_trace=False)   # (the check would re-run the module eagerly on the host)
    assert torch.equal(traced(x), want)
    with pytest.raises(RuntimeError, match="no CPU"):
        exporter(x)

    # a state-dependent std head (Unflatten at the end) and a student network export the same way
    sd = ActorCritic(obs, groups, 3, actor_hidden_dims=[8], critic_hidden_dims=[8], state_dependent_std=True)
    assert torch.jit.script(copy.deepcopy(sd.actor))(x).shape == (5, 2, 3)
    st = StudentTeacher(obs, groups, 3, student_hidden_dims=[8], teacher_hidden_dims=[8])
    assert torch.equal(torch.jit.script(copy.deepcopy(st.student))(x), plain(st.student, x))
    # recurrent policy: the memory's parameter container is a stock nn.LSTM
    rec = ActorCriticRecurrent(obs, groups, 3, actor_hidden_dims=[8], critic_hidden_dims=[8], rnn_hidden_dim=12)
    rnn = copy.deepcopy(rec.memory_a.rnn)
    assert type(rnn) is torch.nn.LSTM
    out, _ = rnn(x.unsqueeze(0))
    assert torch.equal(torch.jit.script(copy.deepcopy(rec.actor))(out.squeeze(0)), plain(rec.actor, out.squeeze(0)))


def test_error_behaviour_of_the_reference_surface(monkeypatch, tmp_path):
    """SURVEY 8(b) "Errors": the exceptions a caller of the reference can rely on -- buffer overflow, generators of the
    wrong training type, missing hidden states in batch mode, unknown noise parametrisation, a non-callable augmentation
    function, a device that does not match the local rank, an unknown logger, `forward()` of the policy classes."""
    import pytest
    import torch

    from tests import host_emulation
    host_emulation.install(monkeypatch)
    from rsl_rl_b200.algorithms import PPO
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.modules import ActorCritic, ActorCriticRecurrent, StudentTeacher
    from rsl_rl_b200.networks import Memory
    from rsl_rl_b200.runners import OnPolicyRunner
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite

    N, A = 8, 2
    obs = TensorDictLite({"policy": torch.randn(N, 5)}, batch_size=[N])
    groups = {"policy": ["policy"], "critic": ["policy"], "teacher": ["policy"]}
    for cls in (ActorCritic, ActorCriticRecurrent, StudentTeacher):
        with pytest.raises(ValueError, match="Unknown standard deviation type"):
            cls(obs, groups, A, noise_std_type="softplus")
    with pytest.raises(ValueError, match="Unknown standard deviation type"):
        ActorCritic(obs, groups, A, noise_std_type="softplus", state_dependent_std=True)
    policy = ActorCritic(obs, groups, A, actor_hidden_dims=[8], critic_hidden_dims=[8])
    with pytest.raises(NotImplementedError):
        policy.forward()
    with pytest.raises(NotImplementedError):
        ActorCriticRecurrent(obs, groups, A, actor_hidden_dims=[8], critic_hidden_dims=[8], rnn_hidden_dim=8).forward()
    with pytest.raises(AssertionError, match="only supports 1D observations"):
        ActorCritic(TensorDictLite({"policy": torch.zeros(N, 3, 3)}, batch_size=[N]), groups, A)
    with pytest.raises(ValueError, match="Hidden states not passed"):
        Memory(5, hidden_size=8)(torch.zeros(3, N, 5), masks=torch.ones(3, N, dtype=torch.bool), hidden_states=None)
    with pytest.raises(ValueError, match="not callable"):
        PPO(policy, device="cpu", symmetry_cfg=dict(use_data_augmentation=True, use_mirror_loss=False,
                                                    data_augmentation_func=42))

    st = RolloutStorage("rl", N, 1, obs, [A], "cpu")
    alg = PPO(policy, device="cpu", use_cuda_graph=False, rng_compat=False, gemm_precision="fp32")
    alg.init_storage("rl", N, 1, obs, [A])
    with torch.inference_mode():
        alg.act(obs)
        alg.process_env_step(obs, torch.zeros(N), torch.zeros(N, dtype=torch.long), {})
        with pytest.raises(OverflowError, match="Rollout buffer overflow"):
            alg.act(obs)
    with pytest.raises(ValueError, match="only available for distillation"):
        next(st.generator())
    with pytest.raises(ValueError, match="gemm_precision must be one of"):
        PPO(policy, device="cpu", gemm_precision="bf16")

    env = ReplayVecEnv({"policy": torch.zeros(2, N, 5)}, torch.zeros(1, N), torch.zeros(1, N, dtype=torch.long), None, A,
                       device="cpu")
    cfg = lambda **extra: dict(dict(num_steps_per_env=1, save_interval=1, obs_groups={"policy": ["policy"]},
                                    policy=dict(class_name="ActorCritic", actor_hidden_dims=[8], critic_hidden_dims=[8]),
                                    algorithm=dict(class_name="PPO", use_cuda_graph=False, rng_compat=False)), **extra)
    with pytest.raises(ValueError, match="Logger type not found"):
        OnPolicyRunner(env, cfg(logger="csv"), log_dir=str(tmp_path), device="cpu").learn(1)
    with pytest.raises(NotImplementedError, match="W&B"):
        OnPolicyRunner(env, cfg(logger="wandb"), log_dir=str(tmp_path), device="cpu").learn(1)
    with pytest.raises(ValueError, match="Unknown class_name"):
        OnPolicyRunner(env, cfg(policy=dict(class_name="Transformer")), device="cpu")
    with pytest.raises(ValueError, match="must contain the 'policy' key"):
        bad = ReplayVecEnv({"proprio": torch.zeros(2, N, 5)}, torch.zeros(1, N), torch.zeros(1, N), None, A, device="cpu")
        OnPolicyRunner(bad, cfg(obs_groups={}), device="cpu")
    # torchrun-style environment whose device does not match the local rank (reference on_policy_runner.py:378-390)
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("LOCAL_RANK", "1")
    with pytest.raises(ValueError, match="does not match expected device for local rank"):
        OnPolicyRunner(env, cfg(), device="cuda:0")
    monkeypatch.setenv("LOCAL_RANK", "3")
    with pytest.raises(ValueError, match="greater than or equal to world size"):
        OnPolicyRunner(env, cfg(), device="cuda:3")


def test_public_api_surface_matches_the_reference_snapshot():
    """The drop-in boundary: every public class / function of the reference's sub-packages exists here under
    the same name, and every method of its surface takes the reference's parameters -- same names, same order, same
    defaults (extra parameters here must be optional).  tests/golden/reference_api.json was recorded from the unmodified
    reference by oracle/make_api_snapshot.py."""
    import importlib
    import inspect
    import json
    import os

    from tests.golden_util import GOLDEN

    snap = {k: v for k, v in json.load(open(os.path.join(GOLDEN, "reference_api.json"))).items() if not k.startswith("__")}
    assert sum(len(v) for v in snap.values()) >= 24

    def sig(fn):
        return [[p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
                for p in inspect.signature(fn).parameters.values()]

    problems = []

    def compare(where, ref, mine):
        fixed = [r for r in ref if r[1] not in ("VAR_KEYWORD", "VAR_POSITIONAL")]
        for r in ref:
            if r[1] in ("VAR_KEYWORD", "VAR_POSITIONAL") and not any(m[1] == r[1] for m in mine):
                problems.append(f"{where}: no {r[1]} parameter")
        for i, r in enumerate(fixed):
            if i >= len(mine) or mine[i][0] != r[0]:
                problems.append(f"{where}: parameter {i} is {mine[i][0] if i < len(mine) else None!r}, reference {r[0]!r}")
            elif mine[i][2] != r[2]:
                problems.append(f"{where}: default of {r[0]!r} is {mine[i][2]}, reference {r[2]}")
        for m in mine[len(fixed):]:
            if m[1] not in ("VAR_KEYWORD", "VAR_POSITIONAL") and m[2] is None:
                problems.append(f"{where}: extra required parameter {m[0]!r}")

    def walk(prefix, cls, members):
        for name, m in members.items():
            if not hasattr(cls, name):
                problems.append(f"{prefix}.{name}: missing")
            elif isinstance(m, dict):
                walk(f"{prefix}.{name}", getattr(cls, name), m["members"])
            elif isinstance(m, list):
                compare(f"{prefix}.{name}", m, sig(getattr(cls, name)))

    for pkg, entries in snap.items():
        mod = importlib.import_module(f"rsl_rl_b200.{pkg}")
        for name, d in entries.items():
            if not hasattr(mod, name):
                problems.append(f"{pkg}.{name}: missing")
            elif d["kind"] == "function":
                compare(f"{pkg}.{name}", d["signature"], sig(getattr(mod, name)))
            else:
                walk(f"{pkg}.{name}", getattr(mod, name), d["members"])
    assert not problems, "\n".join(problems)


def test_instance_attributes_of_the_reference_are_present(monkeypatch):
    """What user code reads from live objects (`alg.policy`, `alg.learning_rate`, `alg.rnd_optimizer`,
    `storage.saved_hidden_states_a`, `runner.current_learning_iteration`, ...): every public instance attribute the
    reference's runner / algorithm / policy / storage / transition objects carry (recorded by oracle/make_api_snapshot.py
    from live reference objects) exists on this library's objects built from the same configurations."""
    import json
    import os

    import torch

    from oracle.symmetry_example import mirror_augmentation
    from tests import host_emulation
    from tests.golden_util import GOLDEN
    host_emulation.install(monkeypatch)
    from rsl_rl_b200.env import ReplayVecEnv
    from rsl_rl_b200.runners import DistillationRunner, OnPolicyRunner

    want = json.load(open(os.path.join(GOLDEN, "reference_api.json")))["__instance_attributes__"]
    N = 8
    env = ReplayVecEnv({"policy": torch.randn(3, N, 11)}, torch.zeros(2, N), torch.zeros(2, N, dtype=torch.long), None, 4,
                       device="cpu")
    cfg = dict(num_steps_per_env=4, save_interval=10, obs_groups={"policy": ["policy"], "critic": ["policy"]},
               policy=dict(class_name="ActorCriticRecurrent", actor_hidden_dims=[8], critic_hidden_dims=[8], rnn_hidden_dim=8),
               algorithm=dict(class_name="PPO", use_cuda_graph=False, rng_compat=False,
                              rnd_cfg=dict(weight=1.0, weight_schedule=None, reward_normalization=False,
                                           state_normalization=False, learning_rate=1e-3, num_outputs=1,
                                           predictor_hidden_dims=[-1], target_hidden_dims=[-1]),
                              symmetry_cfg=dict(use_data_augmentation=True, use_mirror_loss=True, mirror_loss_coeff=0.5,
                                                data_augmentation_func=mirror_augmentation)))
    runner = OnPolicyRunner(env, cfg, log_dir=None, device="cpu")
    alg = runner.alg
    with torch.inference_mode():
        alg.act(env.get_observations())
    objects = {"OnPolicyRunner": runner, "PPO": alg, "ActorCriticRecurrent": alg.policy, "Memory": alg.policy.memory_a,
               "RolloutStorage[rl]": alg.storage, "Transition": alg.transition, "RandomNetworkDistillation": alg.rnd}
    env = ReplayVecEnv({"policy": torch.randn(3, N, 6), "priv": torch.randn(3, N, 3)}, torch.zeros(2, N),
                       torch.zeros(2, N, dtype=torch.long), None, 2, device="cpu")
    cfg = dict(num_steps_per_env=4, save_interval=10, obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]},
               policy=dict(class_name="StudentTeacher", student_hidden_dims=[8], teacher_hidden_dims=[8],
                           student_obs_normalization=True),
               algorithm=dict(class_name="Distillation"))
    drunner = DistillationRunner(env, cfg, log_dir=None, device="cpu")
    objects.update({"DistillationRunner": drunner, "Distillation": drunner.alg, "StudentTeacher": drunner.alg.policy,
                    "EmpiricalNormalization": drunner.alg.policy.student_obs_normalizer,
                    "RolloutStorage[distillation]": drunner.alg.storage})
    assert set(objects) == set(want)
    missing = [f"{k}.{a}" for k, names in want.items() for a in names if not hasattr(objects[k], a)]
    assert not missing, missing
