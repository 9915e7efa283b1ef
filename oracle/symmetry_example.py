"""TEST INFRASTRUCTURE -- a deterministic symmetry augmentation with the signature the reference expects
(`data_augmentation_func(obs, actions, env)`, rsl_rl/algorithms/ppo.py:226-234, 318-337): a fixed signed
permutation of every observation group's dims and of the action dims; returns cat([original, mirrored])."""
import numpy as np
import torch


def _signed_perm(dim: int, device):
    rng = np.random.default_rng(1000 + dim)
    perm = torch.as_tensor(rng.permutation(dim), device=device)
    sign = torch.as_tensor(rng.choice([-1.0, 1.0], size=dim), dtype=torch.float32, device=device)
    return perm, sign


def mirror(x: torch.Tensor) -> torch.Tensor:
    perm, sign = _signed_perm(x.shape[-1], x.device)
    return x[..., perm] * sign


def mirror_augmentation(obs=None, actions=None, env=None):
    obs_out = None
    if obs is not None:
        n = obs.batch_size[0]
        obs_out = type(obs)({k: torch.cat([v, mirror(v)], dim=0) for k, v in obs.items()}, batch_size=[2 * n])
    act_out = None if actions is None else torch.cat([actions, mirror(actions)], dim=0)
    return obs_out, act_out
