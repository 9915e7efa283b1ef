"""Helper functions (mirror of `rsl_rl.utils`)."""
from .tensordict_lite import TensorDict, TensorDictLite, is_tensordict_like
from .utils import (TrajectoryIndex, resolve_nn_activation, resolve_obs_groups, resolve_optimizer,
                    split_and_pad_trajectories, store_code_state, string_to_callable, unpad_trajectories)

__all__ = ["TensorDict", "TensorDictLite", "is_tensordict_like", "TrajectoryIndex", "resolve_nn_activation",
           "resolve_obs_groups", "resolve_optimizer", "split_and_pad_trajectories", "store_code_state",
           "string_to_callable", "unpad_trajectories"]
