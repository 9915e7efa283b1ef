"""Helper functions mirroring `rsl_rl/utils/utils.py` of the reference (same names, arguments, errors).

`split_and_pad_trajectories` / `unpad_trajectories` run on the device through the segmented-scan
kernels of csrc/trajectory.cu instead of the reference's host-side split/pad_sequence loop.
"""
from __future__ import annotations

import importlib
import os
import pathlib
import warnings
from typing import Callable

import torch

from .. import _native as nat
from .tensordict_lite import TensorDict, is_tensordict_like

# activation name -> torch module factory.  The kernels fuse the activation into the GEMM epilogue and
# differentiate it from the stored OUTPUT, so only activations whose derivative is a function of their
# output are available natively (reference: rsl_rl/utils/utils.py:18-49 lists 12 names).
_NATIVE_ACTIVATIONS = {
    "elu": torch.nn.ELU, "selu": torch.nn.SELU, "relu": torch.nn.ReLU, "crelu": torch.nn.CELU,
    "lrelu": torch.nn.LeakyReLU, "tanh": torch.nn.Tanh, "sigmoid": torch.nn.Sigmoid,
    "softplus": torch.nn.Softplus, "identity": torch.nn.Identity,
}
_KNOWN_BUT_UNSUPPORTED = ("gelu", "swish", "mish")


def resolve_nn_activation(act_name: str) -> torch.nn.Module:
    """Activation module from its name (reference utils.py:18-49)."""
    name = act_name.lower()
    if name in _NATIVE_ACTIVATIONS:
        return _NATIVE_ACTIVATIONS[name]()
    if name in _KNOWN_BUT_UNSUPPORTED:
        raise NotImplementedError(
            f"Activation '{name}' has no sm_100a epilogue yet (its derivative is not a function of its output); "
            f"available: {list(_NATIVE_ACTIVATIONS)}")
    raise ValueError(f"Invalid activation function '{act_name}'. Valid activations are: "
                     f"{list(_NATIVE_ACTIVATIONS) + list(_KNOWN_BUT_UNSUPPORTED)}")


def resolve_optimizer(optimizer_name: str):
    """Optimizer class from its name (reference utils.py:52-75)."""
    table = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD,
             "rmsprop": torch.optim.RMSprop}
    name = optimizer_name.lower()
    if name not in table:
        raise ValueError(f"Invalid optimizer '{optimizer_name}'. Valid optimizers are: {list(table)}")
    return table[name]


# ---------------------------------------------------------------------------------------------------
# trajectories (reference utils.py:78-141)
# ---------------------------------------------------------------------------------------------------

class TrajectoryIndex:
    """Device-side (t, env) -> (trajectory, position) map of one rollout, built by one segmented scan."""

    def __init__(self, dones: torch.Tensor):
        nat.require_cuda(dones)
        T, N = dones.shape[0], dones.shape[1]
        d = dones.reshape(T, N)
        if d.dtype != torch.uint8:
            d = (d != 0).to(torch.uint8)
        d = d.contiguous()
        dev = d.device
        self.T, self.N = T, N
        i32 = dict(dtype=torch.int32, device=dev)
        self.traj_id = torch.empty(T, N, **i32)
        self.traj_pos = torch.empty(T, N, **i32)
        self.traj_len = torch.empty(T * N, **i32)
        self.traj_env = torch.empty(T * N, **i32)
        self.traj_start = torch.empty(T * N, **i32)
        self.env_traj_offset = torch.empty(N + 1, **i32)
        ws = torch.empty(int(nat.lib().rslrl_traj_workspace_bytes(T, N)), dtype=torch.uint8, device=dev)
        nat.check(nat.lib().rslrl_trajectory_index(d.data_ptr(), T, N, self.traj_id.data_ptr(), self.traj_pos.data_ptr(),
                                                   self.traj_len.data_ptr(), self.traj_env.data_ptr(),
                                                   self.traj_start.data_ptr(), self.env_traj_offset.data_ptr(),
                                                   ws.data_ptr(), nat.stream_ptr()))
        self._offsets_host = None

    def offsets_host(self):
        """env_traj_offset on the host (the single device->host read of the recurrent generator)."""
        if self._offsets_host is None:
            self._offsets_host = self.env_traj_offset.cpu().tolist()
        return self._offsets_host

    @property
    def n_traj(self) -> int:
        return self.offsets_host()[-1]

    def masks(self, first_traj: int, last_traj: int) -> torch.Tensor:
        cap = last_traj - first_traj
        out = torch.empty(self.T, cap, dtype=torch.bool, device=self.traj_len.device)
        nat.check(nat.lib().rslrl_trajectory_masks(self.traj_len.data_ptr(), out.data_ptr(), self.T, first_traj,
                                                   last_traj, cap, nat.stream_ptr()))
        return out

    def split_and_pad(self, x: torch.Tensor, first_traj: int, last_traj: int) -> torch.Tensor:
        T, N = self.T, self.N
        tail = x.shape[2:]
        D = int(torch.Size(tail).numel()) if len(tail) else 1
        xc = nat.f32c(x).reshape(T, N, D)
        cap = last_traj - first_traj
        out = torch.empty(T, cap, D, dtype=torch.float32, device=x.device)
        nat.check(nat.lib().rslrl_split_and_pad(xc.data_ptr(), self.traj_len.data_ptr(), self.traj_env.data_ptr(),
                                                self.traj_start.data_ptr(), out.data_ptr(), T, N, D, first_traj,
                                                last_traj, cap, nat.stream_ptr()))
        return out.reshape(T, cap, *tail)

    def pad_columns(self, x: torch.Tensor, env0: int, first_traj: int, last_traj: int) -> torch.Tensor:
        """split_and_pad of a tensor that holds only the env columns [env0, env0 + x.shape[1]) of the rollout
        ([T, ne, D] -> [T, last_traj - first_traj, D], zero padded): the inverse of `unpad` (gradients on their way back
        into a recurrent memory)."""
        T, ne, D = x.shape
        xc = nat.f32c(x)
        cap = last_traj - first_traj
        out = torch.empty(T, cap, D, dtype=torch.float32, device=x.device)
        # the kernel addresses x[(start + p), env, :] with the ABSOLUTE env index: shift the base by env0 columns
        nat.check(nat.lib().rslrl_split_and_pad(xc.data_ptr() - env0 * D * 4, self.traj_len.data_ptr(),
                                                self.traj_env.data_ptr(), self.traj_start.data_ptr(), out.data_ptr(), T, ne,
                                                D, first_traj, last_traj, cap, nat.stream_ptr()))
        return out

    def unpad(self, padded: torch.Tensor, env0: int, env1: int, first_traj: int) -> torch.Tensor:
        T = self.T
        cap = padded.shape[1]
        D = padded.shape[-1]
        pc = nat.f32c(padded)
        out = torch.empty(T, env1 - env0, D, dtype=torch.float32, device=padded.device)
        nat.check(nat.lib().rslrl_unpad(pc.data_ptr(), self.traj_id.data_ptr(), self.traj_pos.data_ptr(),
                                        out.data_ptr(), T, self.N, D, env0, env1, first_traj, cap, nat.stream_ptr()))
        return out

    def hidden_at_starts(self, saved: torch.Tensor, first_traj: int, last_traj: int) -> torch.Tensor:
        T, L, N, H = saved.shape
        cap = last_traj - first_traj
        out = torch.empty(L, cap, H, dtype=torch.float32, device=saved.device)
        nat.check(nat.lib().rslrl_hidden_at_starts(nat.f32c(saved).data_ptr(), self.traj_env.data_ptr(),
                                                   self.traj_start.data_ptr(), out.data_ptr(), T, L, N, H, first_traj,
                                                   last_traj, cap, nat.stream_ptr()))
        return out


def split_and_pad_trajectories(tensor, dones: torch.Tensor):
    """Splits trajectories at done indices, pads with zeros to the rollout length T and returns the
    validity masks -- same contract as the reference (utils.py:78-131):
    input [T, N, ...] (tensor or TensorDict) -> ([T, n_traj, ...], bool [T, n_traj])."""
    index = TrajectoryIndex(dones)
    n_traj = index.n_traj
    masks = index.masks(0, n_traj)
    if is_tensordict_like(tensor):
        padded = TensorDict({k: index.split_and_pad(v, 0, n_traj) for k, v in tensor.items()},
                            batch_size=[tensor.batch_size[0], n_traj])
    else:
        padded = index.split_and_pad(tensor, 0, n_traj)
    return padded, masks


def unpad_trajectories(trajectories: torch.Tensor, masks: torch.Tensor):
    """Inverse of split_and_pad_trajectories (reference utils.py:134-141): [T, n_traj, H] -> [T, N, H].

    The masks determine the map: trajectory j of length len_j fills T consecutive slots of the env-major
    flattened output.  Rebuilt here from the masks alone with the same scan kernels."""
    nat.require_cuda(trajectories, masks)
    T, n_traj = masks.shape
    lengths = masks.sum(dim=0, dtype=torch.int32)  # [n_traj]
    ends = torch.cumsum(lengths, 0)
    N = int(ends[-1].item()) // T if n_traj > 0 else 0
    # position of every valid (p, j) in the env-major flattened [N*T] layout
    starts = ends - lengths
    p = torch.arange(T, device=masks.device, dtype=torch.int32).unsqueeze(1)
    flat = (starts.unsqueeze(0) + p)[masks]  # row-major over (p, j) restricted to valid entries
    vals = trajectories[masks]
    out = torch.empty(N * T, trajectories.shape[-1], dtype=trajectories.dtype, device=trajectories.device)
    out[flat.long()] = vals
    return out.view(N, T, trajectories.shape[-1]).transpose(1, 0)


# ---------------------------------------------------------------------------------------------------
# logging / config glue (reference utils.py:144-304)
# ---------------------------------------------------------------------------------------------------

def store_code_state(logdir, repositories) -> list:
    """Dumps `git status` + `git diff` of every repository into <logdir>/git (reference utils.py:144-169)."""
    try:
        import git
    except Exception:  # GitPython is optional here
        return []
    target = os.path.join(logdir, "git")
    os.makedirs(target, exist_ok=True)
    written = []
    for repo_path in repositories:
        try:
            repo = git.Repo(repo_path, search_parent_directories=True)
            tree = repo.head.commit.tree
        except Exception:
            print(f"Could not find git repository in {repo_path}. Skipping.")
            continue
        name = pathlib.Path(repo.working_dir).name
        diff_file = os.path.join(target, f"{name}.diff")
        if os.path.isfile(diff_file):
            continue
        print(f"Storing git diff for '{name}' in: {diff_file}")
        with open(diff_file, "x", encoding="utf-8") as fh:
            fh.write(f"--- git status ---\n{repo.git.status()} \n\n\n--- git diff ---\n{repo.git.diff(tree)}")
        written.append(diff_file)
    return written


def string_to_callable(name: str) -> Callable:
    """'module:attribute' -> callable (reference utils.py:172-199)."""
    try:
        mod_name, attr_name = name.split(":")
        obj = getattr(importlib.import_module(mod_name), attr_name)
    except AttributeError as err:
        raise ValueError(
            "We could not interpret the entry as a callable object. The format of input should be"
            f" 'module:attribute_name'\nWhile processing input '{name}', received the error:\n {err}.") from err
    if not callable(obj):
        raise ValueError(f"The imported object is not callable: '{name}'")
    return obj


def resolve_obs_groups(obs, obs_groups: dict, default_sets: list) -> dict:
    """Validates the observation-set configuration and fills in missing default sets
    (behaviour of the reference utils.py:202-304: 'policy' is mandatory unless an observation group of
    that name exists; each default set falls back to a same-named group, else to the 'policy' set)."""
    if "policy" not in obs_groups:
        if "policy" not in obs:
            raise ValueError("The observation configuration dictionary 'obs_groups' must contain the 'policy' key."
                             f" Found keys: {list(obs_groups.keys())}")
        obs_groups["policy"] = ["policy"]
        warnings.warn("'obs_groups' has no 'policy' key; using the observation group named 'policy'."
                      " Add the key explicitly: this implicit behaviour will be removed.")
    for set_name, groups in obs_groups.items():
        if len(groups) == 0:
            hint = ""
            if set_name in default_sets:
                hint = (" Consider removing the key to default to the observations used for the 'policy' set."
                        if set_name not in obs else
                        f" Consider removing the key to default to the observation '{set_name}' from the environment.")
            raise ValueError(f"The '{set_name}' key in the 'obs_groups' dictionary can not be an empty list." + hint)
        for group in groups:
            if group not in obs:
                raise ValueError(f"Observation '{group}' in observation set '{set_name}' not found in the observations"
                                 f" from the environment. Available observations from the environment:"
                                 f" {list(obs.keys())}")
    for name in default_sets:
        if name in obs_groups:
            continue
        if name in obs:
            obs_groups[name] = [name]
            warnings.warn(f"'obs_groups' has no '{name}' key; using the observation group named '{name}'."
                          " Add the key explicitly: this implicit behaviour will be removed.")
        else:
            obs_groups[name] = list(obs_groups["policy"])
            warnings.warn(f"'obs_groups' has no '{name}' key; the observations of the 'policy' set are used."
                          " Add the key explicitly: this implicit behaviour will be removed.")
    print("-" * 80)
    print("Resolved observation sets: ")
    for set_name, groups in obs_groups.items():
        print("\t", set_name, ": ", groups)
    print("-" * 80)
    return obs_groups
