"""rsl_rl_b200 -- B200-native implementation of rsl_rl's PPO learning iteration.

The sub-packages mirror the reference's layout (`runners`, `algorithms`, `modules`, `storage`,
`networks`, `env`, `utils`) so that `rsl_rl_b200.install_as("rsl_rl")` makes existing
`from rsl_rl.runners import OnPolicyRunner` imports resolve to this implementation.
"""
import importlib
import sys

__version__ = "0.1.0"

_SUBMODULES = ("utils", "env", "networks", "storage", "modules", "algorithms", "runners")


def install_as(name: str = "rsl_rl"):
    """Registers this package (and its sub-packages) in sys.modules under `name`."""
    pkg = sys.modules[__name__]
    sys.modules[name] = pkg
    for sub in _SUBMODULES:
        sys.modules[f"{name}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
    return pkg
