"""Random Network Distillation (mirror of `rsl_rl/modules/rnd.py`; Burda et al., arXiv:1810.12894).

A frozen random `target` MLP and a trained `predictor` MLP embed the `rnd_state` observation set; the
intrinsic reward is the L2 distance of the embeddings times a (scheduled) weight."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from ..networks import MLP, EmpiricalDiscountedVariationNormalization, EmpiricalNormalization


class RandomNetworkDistillation(nn.Module):
    def __init__(self, num_states: int, obs_groups: dict, num_outputs: int, predictor_hidden_dims: list,
                 target_hidden_dims: list, activation: str = "elu", weight: float = 0.0,
                 state_normalization: bool = False, reward_normalization: bool = False, device: str = "cpu",
                 weight_schedule: "dict | None" = None):
        super().__init__()
        self.num_states = num_states
        self.obs_groups = obs_groups
        self.num_outputs = num_outputs
        self.initial_weight = weight
        self.device = device
        self.state_normalization = state_normalization
        self.reward_normalization = reward_normalization
        self.state_normalizer = (EmpiricalNormalization(shape=[num_states], until=1.0e8).to(device)
                                 if state_normalization else torch.nn.Identity())
        self.reward_normalizer = (EmpiricalDiscountedVariationNormalization(shape=[], until=1.0e8).to(device)
                                  if reward_normalization else torch.nn.Identity())
        self.update_counter = 0
        if weight_schedule is not None:
            self.weight_scheduler_params = weight_schedule
            self.weight_scheduler = getattr(self, f"_{weight_schedule['mode']}_weight_schedule")
        else:
            self.weight_scheduler = None
        self.predictor = MLP(num_states, num_outputs, predictor_hidden_dims, activation).to(device)
        self.target = MLP(num_states, num_outputs, target_hidden_dims, activation).to(device)
        self.target.eval()

    def get_intrinsic_reward(self, obs) -> torch.Tensor:
        """Reference rnd.py:113-135 (the counter advances once per env step).  On CUDA the two embedding networks, the
        gap, its norm and (without a reward normaliser) the weight are ONE launch (`rslrl_rnd_reward`)."""
        self.update_counter += 1
        state = self.state_normalizer(self.get_rnd_state(obs))
        if self.weight_scheduler is not None:
            self.weight = self.weight_scheduler(step=self.update_counter, **self.weight_scheduler_params)
        else:
            self.weight = self.initial_weight
        plain = isinstance(self.reward_normalizer, torch.nn.Identity)
        norm = self._native_gap_norm(state, float(self.weight) if plain else 1.0)
        if norm is not None:
            return norm if plain else self.reward_normalizer(norm) * self.weight
        with torch.no_grad():
            gap = self.target(state) - self.predictor(state)
        reward = self.reward_normalizer(torch.linalg.norm(gap, dim=1))
        return reward * self.weight

    def _native_gap_norm(self, state: torch.Tensor, scale: float):
        """scale * ||target(state) - predictor(state)||_2 per row in one launch, or None when the fused kernel does not
        cover the networks (widths > 128, more than 4 layers, a tuple-shaped output, a last activation)."""
        from .. import _native as nat
        import ctypes

        if not nat.on_device(state) or state.dim() != 2 or os.environ.get("RSLRL_RND_FUSED", "1") == "0":
            return None
        plan = getattr(self, "_reward_plan", None)
        if plan is None:
            structs = []
            for mlp in (self.target, self.predictor):
                if any(isinstance(m, nn.Unflatten) for m in mlp):
                    structs = None
                    break
                layers, acts = mlp.linear_layers(), mlp.layer_activation_codes()
                dims = [layers[0].in_features] + [l.out_features for l in layers]
                n_par = sum(a * b + b for a, b in zip(dims[:-1], dims[1:]))
                # (both networks' parameters + three 32-row activation tiles must fit in 200 KB of shared memory)
                if len(layers) > 4 or max(dims) > 128 or acts[-1] != 0 or (2 * n_par + 96 * 129) * 4 > 200 * 1024:
                    structs = None
                    break
                q = nat.RndNet()
                for i, l in enumerate(layers):
                    q.w[i], q.b[i], q.act[i] = l.weight.data_ptr(), (l.bias.data_ptr() if l.bias is not None else None), acts[i]
                for i, d in enumerate(dims):
                    q.dims[i] = d
                q.num_layers = len(layers)
                structs.append((q, [l.weight.data_ptr() for l in layers]))
            plan = self._reward_plan = structs if structs is not None else False
        if plan is False:
            return None
        for (q, ptrs), mlp in zip(plan, (self.target, self.predictor)):   # parameters re-bound (flat buffers, load_state_dict)
            if ptrs != [l.weight.data_ptr() for l in mlp.linear_layers()]:
                self._reward_plan = None
                return self._native_gap_norm(state, scale)
        x = nat.f32c(state.detach())
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        nat.check(nat.lib().rslrl_rnd_reward(ctypes.byref(plan[0][0]), ctypes.byref(plan[1][0]), x.data_ptr(), x.stride(0),
                                             x.shape[0], out.data_ptr(), scale, nat.stream_ptr()))
        return out

    def forward(self, *args, **kwargs):
        raise RuntimeError("Forward method is not implemented. Use get_intrinsic_reward instead.")

    def train(self, mode: bool = True):
        self.predictor.train(mode)
        if self.state_normalization:
            self.state_normalizer.train(mode)
        if self.reward_normalization:
            self.reward_normalizer.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def get_rnd_state(self, obs):
        groups = self.obs_groups["rnd_state"]
        return obs[groups[0]] if len(groups) == 1 else torch.cat([obs[g] for g in groups], dim=-1)

    def update_normalization(self, obs):
        if self.state_normalization:
            self.state_normalizer.update(self.get_rnd_state(obs))

    # weight schedules (reference rnd.py:168-182)
    def _constant_weight_schedule(self, step: int, **kwargs):
        return self.initial_weight

    def _step_weight_schedule(self, step: int, final_step: int, final_value: float, **kwargs):
        return self.initial_weight if step < final_step else final_value

    def _linear_weight_schedule(self, step: int, initial_step: int, final_step: int, final_value: float, **kwargs):
        if step < initial_step:
            return self.initial_weight
        if step > final_step:
            return final_value
        frac = (step - initial_step) / (final_step - initial_step)
        return self.initial_weight + (final_value - self.initial_weight) * frac


def resolve_rnd_config(alg_cfg, obs, obs_groups, env):
    """Fills `num_states` / `obs_groups` and scales the weight by the env step (reference rnd.py:185-209)."""
    if "rnd_cfg" in alg_cfg and alg_cfg["rnd_cfg"] is not None:
        n = 0
        for g in obs_groups["rnd_state"]:
            assert len(obs[g].shape) == 2, "The RND module only supports 1D observations."
            n += obs[g].shape[-1]
        alg_cfg["rnd_cfg"]["num_states"] = n
        alg_cfg["rnd_cfg"]["obs_groups"] = obs_groups
        alg_cfg["rnd_cfg"]["weight"] *= env.unwrapped.step_dt
    return alg_cfg
