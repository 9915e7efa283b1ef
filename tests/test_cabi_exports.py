"""CPU: the C-ABI library builds/loads and exports every symbol include/rslrl_b200.h declares."""
import ctypes
import os
import re

from rsl_rl_b200 import _build, _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rslrl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rslrl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build_library()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    # the ctypes signature table covers exactly the header
    assert sorted(nat.exported_symbols()) == declared


def test_loader_checks_abi_and_struct_layout():
    lib = nat.lib()
    assert lib.rslrl_abi_version() == 3
    assert ctypes.sizeof(nat.OptState) == 64
    assert ctypes.sizeof(nat.LinearFwd) == 4 * 8 + 6 * 4 and ctypes.sizeof(nat.LinearBwdData) == 4 * 8 + 7 * 4 + 4 + 8
    assert ctypes.sizeof(nat.Mlp3Fwd) == 216   # static_assert'ed on the C side (mlp_fused.cu)
    assert lib.rslrl_gae_workspace_bytes(4096) >= 16
    assert lib.rslrl_loss_workspace_bytes() >= 8 * (4 + 64)


def test_no_cpu_fallback():
    import pytest
    import torch
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite

    obs = TensorDictLite({"policy": torch.zeros(4, 3)}, batch_size=[4])
    with pytest.raises(RuntimeError, match="no CPU|There is no CPU"):
        RolloutStorage("rl", 4, 2, obs, [2], device="cpu")
    with pytest.raises(RuntimeError):
        nat.require_cuda(torch.zeros(3))
