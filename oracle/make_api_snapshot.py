#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- records the public API surface of the UNMODIFIED reference (the drop-in boundary is
its Python class API) into tests/golden/reference_api.json: for every public name of `rsl_rl.{runners, algorithms, modules,
storage, networks, env, utils}` its kind, and for classes the signature of `__init__` and of every public method /
property defined in the package (parameter names in order, kinds, defaults).  tests/test_host_logic.py compares this
library's surface with it.

    python oracle/make_api_snapshot.py [--ref /root/reference]      (build container only)
"""
import argparse
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PACKAGES = ("runners", "algorithms", "modules", "storage", "networks", "env", "utils")


def signature_of(fn):
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return None
    out = []
    for p in sig.parameters.values():
        default = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([p.name, p.kind.name, default])
    return out


def describe(obj, package_prefix):
    if inspect.isclass(obj):
        members = {}
        for name, member in inspect.getmembers(obj):
            if name.startswith("_") and name != "__init__":
                continue
            owner = next((c for c in obj.__mro__ if name in c.__dict__), None)
            if owner is None or not owner.__module__.startswith(package_prefix):
                continue                                   # inherited from torch / builtins: not this package's surface
            raw = owner.__dict__[name]
            if isinstance(raw, property):
                members[name] = "property"
            elif isinstance(raw, (staticmethod, classmethod)) or inspect.isfunction(raw):
                members[name] = signature_of(getattr(obj, name))
            elif inspect.isclass(raw):
                members[name] = describe(raw, package_prefix)
        return {"kind": "class", "members": members}
    if inspect.isfunction(obj):
        return {"kind": "function", "signature": signature_of(obj)}
    return None


def snapshot(root_module_name):
    import importlib
    out = {}
    for pkg in PACKAGES:
        mod = importlib.import_module(f"{root_module_name}.{pkg}")
        names = set(getattr(mod, "__all__", ())) | {n for n in dir(mod) if not n.startswith("_")}   # (`from .rnd import *`)
        entry = {}
        for n in sorted(names):
            obj = getattr(mod, n)
            if getattr(obj, "__module__", "").startswith(f"{root_module_name}.{pkg}"):   # defined in this sub-package
                d = describe(obj, root_module_name)
                if d is not None:
                    entry[n] = d
        out[pkg] = entry
    return out


def instance_attributes():
    """Public attribute names of live objects of the reference (what user code reads: `runner.alg.policy`,
    `alg.learning_rate`, `storage.saved_hidden_states_a`, ...): an OnPolicyRunner with a recurrent policy, RND and
    symmetry switched on, and a DistillationRunner, both on a tiny synthetic env on the CPU."""
    import contextlib
    import io

    sys.path.insert(0, os.path.join(HERE, ".."))
    import make_golden as mg
    from oracle.symmetry_example import mirror_augmentation
    from rsl_rl.runners import DistillationRunner, OnPolicyRunner

    Env = mg.make_env_class()
    public = lambda obj: sorted(k for k in vars(obj) if not k.startswith("_"))
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        env = Env(8, {"policy": 11}, 4, seed=1)
        cfg = dict(num_steps_per_env=4, save_interval=10, obs_groups={"policy": ["policy"], "critic": ["policy"]},
                   policy=dict(class_name="ActorCriticRecurrent", actor_hidden_dims=[8], critic_hidden_dims=[8],
                               rnn_hidden_dim=8),
                   algorithm=dict(class_name="PPO",
                                  rnd_cfg=dict(weight=1.0, weight_schedule=None, reward_normalization=False,
                                               state_normalization=False, learning_rate=1e-3, num_outputs=1,
                                               predictor_hidden_dims=[-1], target_hidden_dims=[-1]),
                                  symmetry_cfg=dict(use_data_augmentation=True, use_mirror_loss=True, mirror_loss_coeff=0.5,
                                                    data_augmentation_func=mirror_augmentation)))
        runner = OnPolicyRunner(env, cfg, log_dir=None, device="cpu")
        alg = runner.alg
        alg.act(env.get_observations())          # fills the transition, creates the hidden states
        out["OnPolicyRunner"] = public(runner)
        out["PPO"] = public(alg)
        out["ActorCriticRecurrent"] = public(alg.policy)
        out["Memory"] = public(alg.policy.memory_a)
        out["RolloutStorage[rl]"] = public(alg.storage)
        out["Transition"] = public(alg.transition)
        out["RandomNetworkDistillation"] = public(alg.rnd)
        env = Env(8, {"policy": 6, "priv": 3}, 2, seed=2)
        cfg = dict(num_steps_per_env=4, save_interval=10, obs_groups={"policy": ["policy"], "teacher": ["policy", "priv"]},
                   policy=dict(class_name="StudentTeacher", student_hidden_dims=[8], teacher_hidden_dims=[8],
                               student_obs_normalization=True),
                   algorithm=dict(class_name="Distillation"))
        runner = DistillationRunner(env, cfg, log_dir=None, device="cpu")
        out["DistillationRunner"] = public(runner)
        out["Distillation"] = public(runner.alg)
        out["StudentTeacher"] = public(runner.alg.policy)
        out["EmpiricalNormalization"] = public(runner.alg.policy.student_obs_normalizer)
        out["RolloutStorage[distillation]"] = public(runner.alg.storage)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden", "reference_api.json"))
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(HERE, "_shim"))
    sys.path.insert(0, args.ref)
    snap = snapshot("rsl_rl")
    snap["__instance_attributes__"] = instance_attributes()
    with open(args.out, "w") as f:
        json.dump(snap, f, indent=1, sort_keys=True)
    n = sum(len(v) for k, v in snap.items() if not k.startswith("__"))
    print("wrote", args.out, n, "public names")


if __name__ == "__main__":
    main()
