"""Proximal Policy Optimization on fused sm_100a kernels.

Mirror of `rsl_rl/algorithms/ppo.py` of the reference: same constructor arguments (plus three optional
keyword arguments at the end), same public methods and attributes
(`policy, optimizer, storage, transition, learning_rate, rnd, rnd_optimizer, symmetry, ...`), same
loss dictionary.  Differences are under the hood:

* the policy parameters live in one flat buffer (`FlatParamGroup`), `optimizer` is a `FusedAdam`
  (a real `torch.optim.Adam` subclass) -- clip_grad_norm_ + KL-adaptive LR + Adam are two launches;
* `act()` writes actions / mean / sigma / value / log-prob straight into row `step` of the rollout
  buffer, `process_env_step()` adds observations, rewards (with the time-out bootstrap) and dones
  with one launch;
* `update()` runs forward, the PPO objective with its hand-derived backward, the MLP backward and
  the optimiser step without autograd and without host synchronisation; the learning rate lives on
  the device and is read back once per update (so `learning_rate` is still a Python float);
* multi-GPU: ONE in-place all-reduce of the flat gradient buffer per minibatch; the minibatch KL
  rides in its tail and every rank applies the LR rule identically (reference: 3 collectives and
  2 host syncs per minibatch, ppo.py:272-290, 441-469).

Feed-forward policies without RND / symmetry take the fused engine; recurrent policies, RND and
symmetry augmentation take `_update_composite` (same kernels for GEMMs, storage, GAE, trajectories and
the optimiser, with the surrounding algebra expressed through autograd).
"""
from __future__ import annotations

import os
from itertools import chain

import torch
import torch.nn as nn

from .. import _native as nat
from ..modules.rnd import RandomNetworkDistillation
from ..storage import RolloutStorage
from ..utils import string_to_callable
from . import _engine
from .flat_params import FlatParamGroup, FusedAdam


_GRAPH_COLLECTIVES = os.environ.get("RSLRL_GRAPH_COLLECTIVES", "1") != "0"
# GEMM arithmetic a drop-in user gets without asking: the reference computes in fp32 (allow_tf32 is off by default in
# torch), so the default must be an fp32-CLASS mode.  "tf32x3" = three split-operand tcgen05 products per GEMM (gradients
# within 2e-6 of the fp32 oracle at the C1 shape, 3x faster than the CUDA-core "fp32" mode); shapes TMA cannot describe
# fall back to the CUDA-core kernels per launch.
DEFAULT_GEMM_PRECISION = "tf32x3"
_ACT_GRAPH = os.environ.get("RSLRL_ACT_GRAPH", "1") != "0"   # per-step act() graphs (forward + Gaussian head)


class PPO:
    """Proximal Policy Optimization algorithm (https://arxiv.org/abs/1707.06347)."""

    def __init__(
        self,
        policy,
        num_learning_epochs=5,
        num_mini_batches=4,
        clip_param=0.2,
        gamma=0.99,
        lam=0.95,
        value_loss_coef=1.0,
        entropy_coef=0.01,
        learning_rate=0.001,
        max_grad_norm=1.0,
        use_clipped_value_loss=True,
        schedule="adaptive",
        desired_kl=0.01,
        device="cpu",
        normalize_advantage_per_mini_batch=False,
        rnd_cfg: dict | None = None,
        symmetry_cfg: dict | None = None,
        multi_gpu_cfg: dict | None = None,
        # --- rsl_rl_b200 extensions (all optional) ---
        gemm_precision: str | None = None,   # None: $RSLRL_GEMM_PRECISION or DEFAULT_GEMM_PRECISION
                                         # "tf32x3": 3xTF32 split-operand tcgen05 products (fp32 class, 1e-5)
                                         # | "fp32": CUDA-core FMA (1e-5 class) | "tf32": tcgen05 tensor cores (1e-3 class)
                                         # | "fp16": tcgen05 kind::f16 on fp16-stored activations/gradients (1e-3 class)
        use_cuda_graph: bool = True,     # capture update() into one CUDA graph (single-GPU fused path)
        rng_compat: bool = True,         # consume the RNG exactly like the reference (discarded draws in update())
    ):
        nat.require_cuda_device(device, "rsl_rl_b200.PPO")
        if gemm_precision is None:
            gemm_precision = os.environ.get("RSLRL_GEMM_PRECISION", DEFAULT_GEMM_PRECISION)
        if gemm_precision not in nat.PREC_CODES:
            raise ValueError(f"gemm_precision must be one of {sorted(nat.PREC_CODES)}, got {gemm_precision!r}")
        nat.lib()
        self.device = device
        self.is_multi_gpu = multi_gpu_cfg is not None
        if multi_gpu_cfg is not None:
            self.gpu_global_rank = multi_gpu_cfg["global_rank"]
            self.gpu_world_size = multi_gpu_cfg["world_size"]
        else:
            self.gpu_global_rank = 0
            self.gpu_world_size = 1

        # RND (reference ppo.py:60-71)
        if rnd_cfg is not None:
            rnd_lr = rnd_cfg.pop("learning_rate", 1e-3)
            self.rnd = RandomNetworkDistillation(device=self.device, **rnd_cfg)
            self._rnd_group = FlatParamGroup(self.rnd.predictor.parameters(), self.device)
            self.rnd_optimizer = FusedAdam(self._rnd_group, lr=rnd_lr)
        else:
            self.rnd = None
            self.rnd_optimizer = None
            self._rnd_group = None

        # symmetry (reference ppo.py:73-92)
        if symmetry_cfg is not None:
            if not (symmetry_cfg["use_data_augmentation"] or symmetry_cfg["use_mirror_loss"]):
                print("Symmetry not used for learning. We will use it for logging instead.")
            if isinstance(symmetry_cfg["data_augmentation_func"], str):
                symmetry_cfg["data_augmentation_func"] = string_to_callable(symmetry_cfg["data_augmentation_func"])
            if symmetry_cfg["use_data_augmentation"] and not callable(symmetry_cfg["data_augmentation_func"]):
                raise ValueError("Data augmentation enabled but the function is not callable:"
                                 f" {symmetry_cfg['data_augmentation_func']}")
            self.symmetry = symmetry_cfg
        else:
            self.symmetry = None

        self.policy = policy
        self.policy.to(self.device)
        self.policy.set_gemm_precision(gemm_precision)
        if self.rnd is not None:
            # (MLPs called under autograd have no fp16-storage variant: tensor cores there mean tf32)
            self.rnd.predictor.precision = "tf32" if gemm_precision == "fp16" else gemm_precision
            self.rnd.target.precision = "tf32" if gemm_precision == "fp16" else gemm_precision
        self._group = FlatParamGroup(self.policy.parameters(), self.device)
        self.optimizer = FusedAdam(self._group, lr=learning_rate)
        if self.is_multi_gpu and os.environ.get("RSLRL_ALLREDUCE", "p2p") == "p2p":
            self._enable_peer_allreduce()
        self.storage: RolloutStorage = None  # type: ignore
        self.transition = RolloutStorage.Transition()

        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate
        self.normalize_advantage_per_mini_batch = normalize_advantage_per_mini_batch
        self.intrinsic_rewards = None

        self.gemm_precision = gemm_precision
        self.use_cuda_graph = use_cuda_graph
        self.rng_compat = rng_compat
        self._engine = None
        # the fused engine drives the ROLLOUT of every feed-forward ActorCritic; the UPDATE additionally needs the plain
        # PPO objective (RND / symmetry go through the composite update on the same parameters)
        self._engine_reason = _engine.supports(self.policy)
        self._obs_fast = None
        if self._engine_reason is None:
            pol = self.policy
            ga, gc = list(pol.obs_groups["policy"]), list(pol.obs_groups["critic"])
            if (len(ga) == 1 and len(gc) == 1 and isinstance(pol.actor_obs_normalizer, nn.Identity)
                    and isinstance(pol.critic_obs_normalizer, nn.Identity)):
                self._obs_fast = (ga[0], None if gc[0] == ga[0] else gc[0])
        self._update_reason = "RND" if self.rnd is not None else ("symmetry" if self.symmetry is not None else None)
        self._recurrent_fused = False
        self._rc = None   # per-update cache of the fused recurrent path
        self._in_place_actions = None
        self._host_actions, self._host_actions_id = None, 0   # pinned [2, N, A] ring the act kernel also writes (host envs)
        self.last_host_actions = None   # the ring slot the most recent act() wrote, valid after a stream synchronisation
        self._graph = None
        self._cgraph, self._cgraph_key, self._cgraph_failed, self._cgraph_launches = None, None, False, 0
        self._csums, self._cperm = None, None
        self._graph_failed = False
        self._graph_key = None
        self._graph_launches = 0
        self._eager_updates = 0
        self._act_plans = {}
        self._act_in = None
        self._lineup = None
        self._cf = None   # permuted rollout copy of the fused RND / symmetry update
        self._fp16_saturations = 0   # running count of saturated fp16 gradient launches (dynamic scale)
        self._draw_offset = None   # Philox offset consumed by one discarded torch.normal draw in update()
        self._mb_hook = None  # test hook: called after every minibatch backward in eager mode
        self.last_update_info = {}

    def _enable_peer_allreduce(self):
        """Gradient buffer into symmetric memory: the per-minibatch all-reduce then runs inside the optimiser kernel
        over NVLink peer loads (csrc/allreduce_adam.cu).  Every rank must take the same path, so the outcome of the
        set-up is agreed on with one NCCL all-reduce; on failure all ranks keep the NCCL all-reduce."""
        ok = 1
        try:
            self.optimizer.enable_peer_allreduce(self.gpu_global_rank, self.gpu_world_size)
        except Exception as exc:
            print(f"[rsl_rl_b200] peer-memory all-reduce unavailable ({type(exc).__name__}: {exc}); using NCCL")
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if int(flag.item()) == 0 and self.optimizer.peer_allreduce:
            self.optimizer.peer_world, self.optimizer.peer_rank = 1, 0
            self.optimizer.one_launch = False

    # ------------------------------------------------------------------------------------------
    def _recurrent_fusable(self) -> bool:
        """ActorCriticRecurrent with native LSTM memories and the plain PPO objective: rollout and update run on the
        fused engine (LSTM kernels + fused MLP forward / loss / hand-written backward), no autograd."""
        from ..modules.actor_critic_recurrent import ActorCriticRecurrent
        pol = self.policy
        if type(pol) is not ActorCriticRecurrent or self.rnd is not None or self.symmetry is not None:
            return False
        if os.environ.get("RSLRL_RECURRENT_FUSED", "1") == "0" or pol.state_dependent_std:
            return False
        ma, mc = pol.memory_a, pol.memory_c
        if not (ma.native and mc.native):
            return False
        ra, rc = ma.rnn, mc.rnn
        if (ra.num_layers, ra.hidden_size, ra.bias) != (rc.num_layers, rc.hidden_size, rc.bias) or ra.hidden_size % 4:
            return False
        for mlp in (pol.actor, pol.critic):
            if any(isinstance(m, nn.Unflatten) for m in mlp) or mlp.layer_activation_codes()[-1] != 0:
                return False
        return True

    def init_storage(self, training_type, num_envs, num_transitions_per_env, obs, actions_shape):
        self.storage = RolloutStorage(training_type, num_envs, num_transitions_per_env, obs, actions_shape, self.device)
        if self._engine_reason is not None and self._recurrent_fusable() and num_envs % self.num_mini_batches == 0:
            # recurrent policy: the engine drives the MLPs behind the memories (inputs = LSTM outputs)
            rows = num_transitions_per_env * (num_envs // self.num_mini_batches)
            precision = "tf32" if self.gemm_precision == "fp16" else self.gemm_precision
            self._engine = _engine.FeedForwardEngine(self.policy, rows, num_envs, rows, self.device, precision, self._group)
            self._update_reason = "recurrent"
            self._recurrent_fused = True
            self._last_values = torch.empty(num_envs, 1, dtype=torch.float32, device=self.device)
        if self._engine_reason is None:
            B = num_envs * num_transitions_per_env
            mb = B // self.num_mini_batches
            precision = self.gemm_precision
            if self._update_reason is not None and precision == "fp16" and not self._composite_fp16_ok():
                precision = "tf32"   # (shapes the fp16-storage chain does not cover, or the autograd composite path)
            self._engine = _engine.FeedForwardEngine(self.policy, B if self._update_reason is None else num_envs, num_envs,
                                                     mb if self._update_reason is None else num_envs, self.device,
                                                     precision, self._group,
                                                     scale_state=self.optimizer.dev_state.scale_state)
            self._last_values = torch.empty(num_envs, 1, dtype=torch.float32, device=self.device)
            if self._update_reason is None:
                self._alloc_update_buffers(num_envs, num_transitions_per_env)

    @property
    def fused(self) -> bool:
        """update() runs on the fused feed-forward engine."""
        return self._engine is not None and self._update_reason is None

    @property
    def rollout_fused(self) -> bool:
        """act() / process_env_step() / compute_returns() run on the fused engine (any feed-forward ActorCritic)."""
        return self._engine is not None

    # ------------------------------------------------------------------------------------------
    # rollout
    # ------------------------------------------------------------------------------------------
    def _inputs(self, obs):
        pol = self.policy
        x_a = pol.actor_obs_normalizer(pol.get_actor_obs(obs))
        if pol.critic_shares_actor_input:
            return nat.f32c(x_a), nat.f32c(x_a)
        x_c = pol.critic_obs_normalizer(pol.get_critic_obs(obs))
        return nat.f32c(x_a), nat.f32c(x_c)

    def act(self, obs):
        """Reference ppo.py:129-140."""
        if not self.rollout_fused:
            return self._act_composite(obs)
        st, eng, tr = self.storage, self._engine, self.transition
        s = st.step
        if s >= st.num_transitions_per_env:
            raise OverflowError("Rollout buffer overflow! You should call clear() before adding new transitions.")
        if s == 0 and eng.needs_refresh:
            eng.refresh_padded()  # parameters may have been loaded / modified since the last optimiser step
        fast = self._obs_fast
        if self._recurrent_fused:
            # both memories advance one step together (native LSTM kernels); their outputs feed the fused MLP forward
            from ..networks.memory import Memory
            pol = self.policy
            tr.hidden_states = pol.get_hidden_states()
            o_a = pol.actor_obs_normalizer(pol.get_actor_obs(obs))
            o_c = pol.critic_obs_normalizer(pol.get_critic_obs(obs))
            out_a, out_c = Memory.forward_pair(pol.memory_a, pol.memory_c, o_a, o_c)
            x_a, x_c = nat.f32c(out_a.squeeze(0)), nat.f32c(out_c.squeeze(0))
        elif fast is None:
            x_a, x_c = self._inputs(obs)
        else:   # one observation group per set, no normaliser: the env's tensors are the network inputs
            x_a = obs[fast[0]]
            x_c = x_a if fast[1] is None else obs[fast[1]]
            if x_a.dtype != torch.float32 or x_c.dtype != torch.float32 or not x_a.is_contiguous() \
                    or not x_c.is_contiguous():
                x_a, x_c = self._inputs(obs)
        key = (x_a.stride(0), x_c.stride(0), x_a.shape[1], x_c.shape[1], self._group.values.data_ptr(), self._host_actions_id)
        plan = self._act_plans.get(s)
        if plan is None or plan["key"] != key:
            plan = self._build_act_plan(s, x_a, x_c, key)
        mu_s, val_s, act_s, logp_s, sig_s = plan["views"]
        sa, sc, zbuf = self._act_in
        if self.policy._noise_fn is not None:
            zbuf.copy_(self.policy._noise_fn(mu_s))
        else:
            zbuf.normal_()  # same draw as torch.normal(mean, std): normal_(0,1) on a tensor of mu's shape
        one = plan["one_launch"]
        self.last_host_actions = plan["host"] if one is not None else None
        if one is not None:
            # ONE launch per env step: fused forward of both nets, Gaussian head in the actor's output epilogue, results
            # written straight into row s of the rollout buffer.  Observations are read in place when their rows are
            # 16-byte aligned, otherwise through the zero-padded staging buffers.
            if plan["direct"] and x_a.data_ptr() % 16 == 0 and x_c.data_ptr() % 16 == 0:
                one(x_a.data_ptr(), x_c.data_ptr())
            else:
                sa.copy_(x_a)
                if sc is not sa:
                    sc.copy_(x_c)
                plan["one_static"](sa.data_ptr(), sc.data_ptr())
        elif not (self.use_cuda_graph and _ACT_GRAPH):
            plan["fwd"](x_a.data_ptr(), x_c.data_ptr())
            eng.gaussian_head(mu_s, zbuf, act_s, logp_s, sig_s)
        else:
            # layer-by-layer forward (fp32 mode / shapes the fused forward does not cover) + Gaussian head of step s
            # replayed from a CUDA graph reading static input buffers
            sa.copy_(x_a)
            if sc is not sa:
                sc.copy_(x_c)
            if plan["graph"] is None:
                plan["visits"] += 1
                if plan["visits"] < 2:   # first visit runs eagerly (lazy initialisation), the second one captures
                    plan["fwd_static"](sa.data_ptr(), sc.data_ptr())
                    eng.gaussian_head(mu_s, zbuf, act_s, logp_s, sig_s)
                else:
                    torch.cuda.synchronize()
                    before = nat.launch_count()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        plan["fwd_static"](sa.data_ptr(), sc.data_ptr())
                        eng.gaussian_head(mu_s, zbuf, act_s, logp_s, sig_s)
                    plan["launches"] = nat.launch_count() - before
                    nat.lib().rslrl_add_launch_count(-plan["launches"] & 0xFFFFFFFFFFFFFFFF)
                    plan["graph"] = g
            if plan["graph"] is not None:
                plan["graph"].replay()
                nat.lib().rslrl_add_launch_count(plan["launches"])
        tr.actions, tr.values = act_s, val_s
        self._in_place_actions = act_s   # process_env_step recognises the in-place row by identity
        tr.actions_log_prob = logp_s.view(-1)
        tr.action_mean, tr.action_sigma = mu_s, sig_s
        tr.observations = obs
        pd = self.policy.__dict__   # plain attributes (set in _init_noise): skip nn.Module.__setattr__'s type checks
        pd["_mean"], pd["_std"] = mu_s, sig_s
        return act_s

    def enable_host_actions(self, ring: torch.Tensor):
        """ring: PINNED host tensor [2, num_envs, num_actions].  From the next act() on, the one-launch rollout step also
        writes its actions into ring[step & 1] (mapped pinned memory): a host-resident env reads them after one stream
        synchronisation instead of a device-to-host copy (`last_host_actions`)."""
        if not (ring.is_pinned() and ring.dtype == torch.float32 and ring.dim() == 3 and ring.shape[0] == 2):
            raise ValueError("enable_host_actions: a pinned float32 [2, num_envs, num_actions] tensor is required")
        self._host_actions, self._host_actions_id = ring, self._host_actions_id + 1

    def _build_act_plan(self, s, x_a, x_c, key):
        st, eng = self.storage, self._engine
        self._group.ensure_bound()
        nat.require_cuda(x_a, x_c)
        views = (st.mu[s], st.values[s], st.actions[s], st.actions_log_prob[s], st.sigma[s])
        if self._act_in is None or self._act_in[0].shape != x_a.shape or self._act_in[1].shape != x_c.shape:
            with torch.inference_mode(False):  # plain tensors: act() may also be called outside inference mode
                pad4 = lambda d: (d + 3) // 4 * 4
                z = lambda t: torch.zeros(t.shape[0], pad4(t.shape[1]), dtype=torch.float32,
                                          device=t.device)[:, :t.shape[1]]      # 16-byte row pitch, zero pad
                sa = z(x_a)
                sc = sa if x_c.data_ptr() == x_a.data_ptr() else z(x_c)
                self._act_in = (sa, sc, torch.empty(views[0].shape, dtype=torch.float32, device=x_a.device))
        sa, sc, zbuf = self._act_in
        N = st.num_envs
        direct = x_a.stride(0) % 4 == 0 and x_c.stride(0) % 4 == 0 and x_a.stride(1) == 1 and x_c.stride(1) == 1
        host = None if self._host_actions is None else self._host_actions[s & 1]
        one = eng.build_act_plan(N, x_a.stride(0), x_c.stride(0), *views[:2], zbuf, *views[2:], actions_host=host) if direct else None
        one_static = eng.build_act_plan(N, sa.stride(0), sc.stride(0), *views[:2], zbuf, *views[2:], actions_host=host)
        plan = dict(key=key, views=views, graph=None, launches=0, visits=0, direct=direct and one is not None, host=host,
                    one_launch=one if one is not None else one_static, one_static=one_static, fwd=None, fwd_static=None)
        if one_static is None:   # no fused forward for this policy / precision: layer-by-layer launch plans
            plan["fwd"] = eng.build_forward_plan(N, x_a.stride(0), x_c.stride(0), views[0], views[1])
            plan["fwd_static"] = eng.build_forward_plan(N, sa.stride(0), sc.stride(0), views[0], views[1])
        self._act_plans[s] = plan
        return plan

    def _act_composite(self, obs):
        tr = self.transition
        if self.policy.is_recurrent:
            tr.hidden_states = self.policy.get_hidden_states()
        if hasattr(self.policy, "act_evaluate"):   # recurrent policy: both memories in grouped launches
            actions, values = self.policy.act_evaluate(obs)
            tr.actions, tr.values = actions.detach(), values.detach()
        else:
            tr.actions = self.policy.act(obs).detach()
            tr.values = self.policy.evaluate(obs).detach()
        tr.actions_log_prob = self.policy.get_actions_log_prob(tr.actions).detach()
        tr.action_mean = self.policy.action_mean.detach()
        tr.action_sigma = self.policy.action_std.detach()
        tr.observations = obs
        return tr.actions

    def process_env_step(self, obs, rewards, dones, extras):
        """Reference ppo.py:142-169; the reward bootstrap and the buffer write are one kernel."""
        self.policy.update_normalization(obs)
        if self.rnd:
            self.rnd.update_normalization(obs)
        tr = self.transition
        tr.rewards = rewards
        tr.dones = dones
        intrinsic = None
        if self.rnd:
            self.intrinsic_rewards = self.rnd.get_intrinsic_reward(obs)
            intrinsic = self.intrinsic_rewards
        time_outs = extras["time_outs"] if "time_outs" in extras else None
        if tr.actions is not None and tr.actions is self._in_place_actions:   # act() wrote row `step` in place
            self._in_place_actions = None
            if tr.hidden_states is not None:
                self.storage._save_hidden_states(tr.hidden_states)
            self.storage._store_in_place(tr.observations, rewards, dones, time_outs, intrinsic, self.gamma)
        else:
            self.storage._store(tr, time_outs=time_outs, intrinsic=intrinsic, gamma=self.gamma)
        tr.clear()
        self.policy.reset(dones)

    def compute_returns(self, obs):
        """Reference ppo.py:171-176."""
        if self.rollout_fused and not self._recurrent_fused:
            _, x_c = self._inputs(obs)
            last_values = self._last_values
            self._engine.forward(x_c, x_c, None, last_values, actor=False)
        else:
            last_values = self.policy.evaluate(obs).detach()
        self.storage.compute_returns(last_values, self.gamma, self.lam,
                                     normalize_advantage=not self.normalize_advantage_per_mini_batch)

    # ------------------------------------------------------------------------------------------
    # update: fused feed-forward path
    # ------------------------------------------------------------------------------------------
    def _alloc_update_buffers(self, N, T):
        pol, dev = self.policy, self.device
        M = self.num_mini_batches
        self._mb = (N * T) // M
        Bp = self._mb * M
        f32 = dict(dtype=torch.float32, device=dev)
        pad4 = lambda d: (d + 3) // 4 * 4
        self._Oa, self._Oc = pol.num_actor_obs, pol.num_critic_obs
        self._xa_raw = torch.zeros(Bp, pad4(self._Oa), **f32)   # pad columns stay zero (padded first-layer weights)
        self._xc_raw = self._xa_raw if pol.critic_shares_actor_input else torch.zeros(Bp, pad4(self._Oc), **f32)
        A = self._engine.A
        self._p_actions = torch.empty(Bp, A, **f32)
        self._p_mu = torch.empty(Bp, A, **f32)
        self._p_sigma = torch.empty(Bp, A, **f32)
        self._p_values = torch.empty(Bp, **f32)
        self._p_adv = torch.empty(Bp, **f32)
        self._p_ret = torch.empty(Bp, **f32)
        self._p_logp = torch.empty(Bp, **f32)
        self._perm = torch.empty(Bp, dtype=torch.int64, device=dev)
        self._discard = torch.empty(self._mb, A, **f32)
        if self._engine.f16:   # fp16 copies of the permuted observations: the B operand of the first layer's wgrad
            pad8 = lambda d: (d + 7) // 8 * 8
            f16 = dict(dtype=torch.float16, device=dev)
            self._xa16 = torch.zeros(Bp, pad8(self._Oa), **f16)
            self._xc16 = self._xa16 if pol.critic_shares_actor_input else torch.zeros(Bp, pad8(self._Oc), **f16)

    def _hyper(self):
        return dict(clip_param=float(self.clip_param), value_loss_coef=float(self.value_loss_coef),
                    entropy_coef=float(self.entropy_coef), use_clipped_value_loss=bool(self.use_clipped_value_loss))

    def _gather_all(self):
        """Permute the whole rollout once (all minibatches of all epochs reuse it)."""
        st, pol = self.storage, self.policy
        fields, widths, outs, strides, cols = [], [], [], [], []

        def add_groups(groups, dst):
            col = 0
            for g in groups:
                src = st.observations[g]
                w = int(src[0, 0].numel())
                fields.append(src); widths.append(w); outs.append(dst); strides.append(dst.stride(0)); cols.append(col)
                col += w

        add_groups(pol.obs_groups["policy"], self._xa_raw)
        if not pol.critic_shares_actor_input:
            add_groups(pol.obs_groups["critic"], self._xc_raw)
        A = self._engine.A
        for src, dst, w in ((st.actions, self._p_actions, A), (st.mu, self._p_mu, A), (st.sigma, self._p_sigma, A),
                            (st.values, self._p_values, 1), (st.advantages, self._p_adv, 1),
                            (st.returns, self._p_ret, 1), (st.actions_log_prob, self._p_logp, 1)):
            fields.append(src); widths.append(w); outs.append(dst); strides.append(w); cols.append(0)
        # the gather's dst base pointer is per field; column offsets handle the concatenation
        st._gather(fields, widths, self._perm, outs, strides, cols)
        xa = self._xa_raw[:, :self._Oa]
        xc = self._xc_raw[:, :self._Oc]
        # observation normalisers: applied IN PLACE on the permuted copy (one launch each, 16-byte row pitch kept)
        if not isinstance(pol.actor_obs_normalizer, nn.Identity):
            pol.actor_obs_normalizer(xa, out=self._xa_raw)
        if pol.critic_shares_actor_input:
            xc = xa
        elif not isinstance(pol.critic_obs_normalizer, nn.Identity):
            pol.critic_obs_normalizer(xc, out=self._xc_raw)
        return xa, xc

    def _to_f16(self, x: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
        nat.check(nat.lib().rslrl_convert_f32_f16(x.data_ptr(), x.stride(0), dst.data_ptr(), dst.stride(0), x.shape[0],
                                                  x.shape[1], dst.shape[1], nat.stream_ptr()))
        return dst[:, :x.shape[1]]

    def _update_kernels(self):
        """Everything update() puts on the stream; capturable (static pointers, no host reads)."""
        eng, opt, hp = self._engine, self.optimizer, self._hyper()
        mb, M = self._mb, self.num_mini_batches
        xa, xc = self._gather_all()
        xa16 = xc16 = None
        if eng.f16:
            xa16, xc16 = self._to_f16(xa, self._xa16), None
            xc16 = xa16 if xc is xa else self._to_f16(xc, self._xc16)
        ds = opt.dev_state
        nat.check(nat.lib().rslrl_fill_zero(ds.loss_sums.data_ptr(), 24, nat.stream_ptr()))
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        for _ in range(self.num_learning_epochs):
            for i in range(M):
                r = slice(i * mb, (i + 1) * mb)
                mu, value = eng.mu[:mb], eng.value[:mb]
                eng.forward(xa[r], xc[r], mu, value)
                eng.loss(mb, mb, mu, value, self._p_actions[r], self._p_logp[r], self._p_adv[r], self._p_ret[r],
                         self._p_values[r], self._p_mu[r], self._p_sigma[r], self._group.kl_slot, ds.loss_sums, hp,
                         self.normalize_advantage_per_mini_batch)
                if eng.f16:
                    eng.backward(xa16[r], xc16[r], eng.dmu[:mb], eng.dvalue[:mb])
                else:
                    eng.backward(xa[r], xc[r], eng.dmu[:mb], eng.dvalue[:mb])
                if self._mb_hook is not None:
                    self._mb_hook(self)
                if self.is_multi_gpu and not opt.peer_allreduce:
                    torch.distributed.all_reduce(self._group.grads_ext, op=torch.distributed.ReduceOp.SUM)
                opt.fused_step(world_size=self.gpu_world_size, max_grad_norm=self.max_grad_norm, adaptive=adaptive,
                               desired_kl=self.desired_kl or 0.0, round_lr_fp32=self.is_multi_gpu)
                eng.refresh_padded(after_step=True)

    def _update_fused(self):
        st, eng, opt = self.storage, self._engine, self.optimizer
        self._group.ensure_bound()
        M, E = self.num_mini_batches, self.num_learning_epochs
        perm = st._randperm(M * self._mb)
        self._perm.copy_(perm)
        if self.rng_compat:  # the reference's policy.act() inside update() draws (and discards) a sample per minibatch
            self._skip_discarded_draws(M * E)
        opt.dev_state.set_lr(self.learning_rate)

        # multi-GPU: the per-minibatch NCCL all-reduce is captured with the kernels (NCCL collectives are graph
        # capturable); RSLRL_GRAPH_COLLECTIVES=0 keeps the multi-GPU update eager
        graph_ok = (self.use_cuda_graph and self._mb_hook is None and not self._graph_failed
                    and (not self.is_multi_gpu or _GRAPH_COLLECTIVES))
        if not graph_ok or self._eager_updates < (2 if self.is_multi_gpu else 1):
            # the first update always runs eagerly (module loading, lazy initialisation); later ones replay a graph
            self._eager_updates += 1
            self._update_kernels()
            return self._finish_update()
        key = (self._hyper_key(), st.actions.data_ptr(), self._group.values.data_ptr())
        if self._graph is None or self._graph_key != key:
            torch.cuda.synchronize()
            before, host_step = nat.launch_count(), opt._host_step
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._update_kernels()
            except Exception as exc:  # e.g. a communicator that cannot be captured: stay eager from now on
                if not self.is_multi_gpu:
                    raise
                print(f"[rsl_rl_b200] update() graph capture with the all-reduce failed ({exc}); running eagerly")
                self._graph_failed = True
                nat.lib().rslrl_add_launch_count(-(nat.launch_count() - before) & 0xFFFFFFFFFFFFFFFF)
                opt._host_step = host_step
                torch.cuda.synchronize()
                self._update_kernels()
                return self._finish_update()
            self._graph_launches = nat.launch_count() - before
            nat.lib().rslrl_add_launch_count(-self._graph_launches & 0xFFFFFFFFFFFFFFFF)  # capture launched nothing
            opt._host_step = host_step
            self._graph, self._graph_key = g, key
        self._graph.replay()
        nat.lib().rslrl_add_launch_count(self._graph_launches)
        opt._host_step += M * E
        return self._finish_update()

    def _skip_discarded_draws(self, count: int, rows: int = None):
        """Advances the CUDA generator exactly as `count` calls of `torch.normal(mean, std)` on a [rows, A] mean
        (default: one minibatch) would (reference ppo.py:249 -> actor_critic.py:140): the Philox offset one such call
        consumes is measured once per shape with a real `normal_()` launch, after which the stream is advanced
        arithmetically -- no launches."""
        idx = torch.device(self.device).index
        gen = torch.cuda.default_generators[torch.cuda.current_device() if idx is None else idx]
        discard = getattr(self, "_discard", None)
        if rows is None:
            rows = discard.shape[0]
        if self._draw_offset is None:
            self._draw_offset = {}
        per_draw = self._draw_offset.get(rows)
        if per_draw is None:
            if discard is not None and rows == discard.shape[0]:
                buf = discard
            else:
                A = discard.shape[1] if discard is not None else self._engine.A
                buf = torch.empty(rows, A, dtype=torch.float32, device=self.device)
            before = gen.get_offset()
            buf.normal_()
            per_draw = self._draw_offset[rows] = gen.get_offset() - before
            count -= 1
        if count > 0:
            gen.set_offset(gen.get_offset() + per_draw * count)

    def release_graphs(self):
        """Drops the captured CUDA graphs (update() and the per-step act() graphs).  Call it before
        `torch.distributed.destroy_process_group()`: the multi-GPU update graph holds NCCL work."""
        torch.cuda.synchronize()
        self._graph, self._graph_key = None, None
        self._cgraph, self._cgraph_key = None, None
        if self._rc is not None:
            self._rc["graphs"] = None
        for plan in self._act_plans.values():
            plan["graph"] = None
            plan["visits"] = 0

    def _hyper_key(self):
        return (self.clip_param, self.value_loss_coef, self.entropy_coef, self.use_clipped_value_loss, self.desired_kl,
                self.schedule, self.max_grad_norm, self.num_learning_epochs, self.num_mini_batches,
                self.normalize_advantage_per_mini_batch, self.gemm_precision)

    def _finish_update(self):
        info = self.optimizer.dev_state.read()  # the one device->host read of update()
        n = self.num_learning_epochs * self.num_mini_batches
        self.learning_rate = info["lr"]
        self.optimizer.sync_host_state(self.learning_rate)
        self.last_update_info = info
        self._fp16_saturations += info.get("fp16_saturations", 0)
        if info.get("fp16_saturations", 0):
            import warnings
            warnings.warn(f"rsl_rl_b200: {info['fp16_saturations']} minibatch(es) of this update had output gradients "
                          "beyond the fp16 range at the current dynamic scale (saturated, then rescaled); a persistent "
                          "warning means gemm_precision='tf32' is the safer choice for this reward scale")
        self.storage.clear()
        s = info["loss_sums"]
        return {"value_function": s[1] / n, "surrogate": s[0] / n, "entropy": s[2] / n}

    def update(self):
        """Reference ppo.py:178-422."""
        if self.is_multi_gpu and self.optimizer.peer_allreduce:
            # Line the ranks up on the stream before the first peer-memory step: rank 0 alone saves checkpoints and
            # logs between iterations, and the in-kernel flag wait should only ever see kernel-level skew (NCCL simply
            # waits; no host synchronisation here).
            if self._lineup is None:
                self._lineup = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.distributed.all_reduce(self._lineup)
        if self.fused:
            return self._update_fused()
        return self._update_composite()

    # ------------------------------------------------------------------------------------------
    # update: composite path (recurrent / RND / symmetry)
    # ------------------------------------------------------------------------------------------
    def _update_composite(self):
        """PPO.update for recurrent policies, RND and symmetry augmentation (reference ppo.py:178-422).

        Same kernels for the dense layers (through autograd), the rollout buffer, GAE, the trajectory
        split/pad/unpad and the flat-buffer clip + KL-adaptive LR + Adam; the objective itself is expressed with
        tensor algebra under autograd.  The KL goes to the device-side LR rule through the gradient tail slot, so
        this path has no per-minibatch host synchronisation either."""
        pol, st, opt = self.policy, self.storage, self.optimizer
        M, E = self.num_mini_batches, self.num_learning_epochs
        self._group.ensure_bound()
        if self.rnd:
            self._rnd_group.ensure_bound()
            self.rnd_optimizer.dev_state.set_lr(self.rnd_optimizer.param_groups[0]["lr"])
        opt.dev_state.set_lr(self.learning_rate)
        if self._csums is None:
            self._csums = torch.zeros(5, dtype=torch.float64, device=self.device)  # value, surrogate, entropy, rnd, symmetry
        sums = self._csums

        # Feed-forward policies (RND / symmetry configurations): every shape is static and nothing returns to the host
        # inside the minibatch loop, so the loop - gathers, forward, autograd backward, optimiser steps - is captured into
        # ONE CUDA graph after two eager updates and replayed; only the permutation is drawn outside (copied into a
        # static buffer).  A user augmentation function that cannot be captured (e.g. one that uploads host data on
        # every call) makes the capture fail once: the update then stays eager.  Recurrent policies have data-dependent
        # trajectory counts and stay eager.
        graph_ok = (self.use_cuda_graph and not pol.is_recurrent and not self.is_multi_gpu and self._mb_hook is None
                    and not self._cgraph_failed and os.environ.get("RSLRL_COMPOSITE_GRAPH", "1") != "0")
        if graph_ok and self._eager_updates >= 2:
            B = M * ((st.num_envs * st.num_transitions_per_env) // M)
            perm = st._randperm(B)
            if self._cperm is None:
                self._cperm = torch.empty(B, dtype=torch.int64, device=self.device)
            self._cperm.copy_(perm)
            key = (self._hyper_key(), st.actions.data_ptr(), self._group.values.data_ptr(), B)
            if self._cgraph is None or self._cgraph_key != key:
                torch.cuda.synchronize()
                steps = (opt._host_step, self.rnd_optimizer._host_step if self.rnd else 0)
                before = nat.launch_count()
                g = torch.cuda.CUDAGraph()
                orig_randperm = st._randperm
                st._randperm = lambda n: self._cperm
                try:
                    with torch.cuda.graph(g):
                        self._composite_minibatches(sums)
                    self._cgraph, self._cgraph_key = g, key
                    self._cgraph_launches = nat.launch_count() - before
                except Exception as exc:
                    print(f"[rsl_rl_b200] composite update() could not be captured into a CUDA graph "
                          f"({type(exc).__name__}: {exc}); running eagerly")
                    self._cgraph_failed, self._cgraph = True, None
                    torch.cuda.synchronize()
                finally:
                    st._randperm = orig_randperm
                    nat.lib().rslrl_add_launch_count(-(nat.launch_count() - before) & 0xFFFFFFFFFFFFFFFF)
                    opt._host_step = steps[0]
                    if self.rnd:
                        self.rnd_optimizer._host_step = steps[1]
            if self._cgraph is not None:
                self._cgraph.replay()
                nat.lib().rslrl_add_launch_count(self._cgraph_launches)
                opt._host_step += M * E
                if self.rnd:
                    self.rnd_optimizer._host_step += M * E
                return self._composite_finish(sums)
            st._randperm_once = perm   # capture failed: this update runs eagerly on the permutation already drawn
        self._eager_updates += 1
        self._composite_minibatches(sums)
        return self._composite_finish(sums)

    def _composite_fusable(self) -> bool:
        """RND and / or symmetry AUGMENTATION on a feed-forward ActorCritic: the policy objective runs on the fused
        engine (no autograd); only the RND predictor's small MSE stays under autograd."""
        return (self._engine is not None and not self.policy.is_recurrent
                and (self.symmetry is None or bool(self.symmetry["use_data_augmentation"]))
                and os.environ.get("RSLRL_COMPOSITE_FUSED", "1") != "0")

    def _composite_fp16_ok(self) -> bool:
        """gemm_precision='fp16' on the fused RND / symmetry update: the 16-bit storage chain needs the shapes the fused
        forward covers (3 hidden layers, widths multiples of 64, A % 4 == 0)."""
        if not self._composite_fusable() or os.environ.get("RSLRL_COMPOSITE_FP16", "1") == "0":
            return False
        pol = self.policy
        from ..networks.mlp import mlp3_eligible
        for net in (_engine._Net(pol.actor), _engine._Net(pol.critic)):
            if not mlp3_eligible(net.dims, net.acts) or any(d % 64 for d in net.dims[1:-1]) or net.dims[-1] > 12:
                return False
        return _engine._Net(pol.actor).dims[-1] % 4 == 0

    def _composite_engine(self, rows: int):
        """The engine with activation / gradient buffers for `rows` minibatch rows (k * minibatch with k-fold
        augmentation); the rollout keeps using the same object."""
        eng = self._engine
        if eng.rows < rows:
            precision = self.gemm_precision
            if precision == "fp16" and not self._composite_fp16_ok():
                precision = "tf32"
            eng = self._engine = _engine.FeedForwardEngine(self.policy, rows, self.storage.num_envs, rows, self.device,
                                                           precision, self._group,
                                                           scale_state=self.optimizer.dev_state.scale_state)
            self._act_plans.clear()
        if eng.f16 and self.symmetry and self.symmetry["use_mirror_loss"]:
            # the mirror-loss gradient is added to dL/dmu AFTER the loss launch: the fp16 copy of dL/dmu is rebuilt from
            # the sum and the output layer's data gradient runs as its own launch (not inside the loss kernel)
            eng.fused_head_dgrad = False
        return eng

    def _composite_minibatches_fused(self, sums):
        """Minibatch loop of the RND / symmetry configurations on the fused engine (reference ppo.py:199-395):
        one permuted gather of the rollout per update; per minibatch the user's augmentation function (torch ops),
        the grouped actor + critic forward on the k-fold batch, ONE loss launch (the per-sample scalars of the augmented
        rows are indexed i % minibatch instead of being repeat()-ed, ppo.py:240-244; entropy and KL on the original
        rows only, :255-257), the mirror-loss gradient added to dL/dmu of the augmented rows (the reference's third
        actor forward `act_inference(obs_aug)` recomputes exactly `mu`, :331-343), the hand-written backward and the
        fused optimiser step.  The RND predictor's MSE (:352-363) is two tiny MLPs under autograd on the native GEMMs."""
        pol, st, opt = self.policy, self.storage, self.optimizer
        M, E = self.num_mini_batches, self.num_learning_epochs
        N, T = st.num_envs, st.num_transitions_per_env
        mb = (N * T) // M
        B = mb * M
        once = getattr(st, "_randperm_once", None)
        st._randperm_once = None
        perm = once if once is not None else st._randperm(B)
        capturing = torch.cuda.is_current_stream_capturing()
        f32 = dict(dtype=torch.float32, device=self.device)
        if self._cf is None or self._cf["B"] != B:
            A = self._engine.A
            self._cf = dict(B=B, obs={k: torch.empty(B, int(v[0, 0].numel()), **f32) for k, v in st.observations.items()},
                            actions=torch.empty(B, A, **f32), mu=torch.empty(B, A, **f32), sigma=torch.empty(B, A, **f32),
                            values=torch.empty(B, **f32), adv=torch.empty(B, **f32), ret=torch.empty(B, **f32),
                            logp=torch.empty(B, **f32), discard={})
        cf = self._cf
        fields, widths, outs, strides, cols = [], [], [], [], []
        for k, src in st.observations.items():
            w = cf["obs"][k].shape[1]
            fields.append(src); widths.append(w); outs.append(cf["obs"][k]); strides.append(w); cols.append(0)
        A = self._engine.A
        for src, dst, w in ((st.actions, cf["actions"], A), (st.mu, cf["mu"], A), (st.sigma, cf["sigma"], A),
                            (st.values, cf["values"], 1), (st.advantages, cf["adv"], 1), (st.returns, cf["ret"], 1),
                            (st.actions_log_prob, cf["logp"], 1)):
            fields.append(src); widths.append(w); outs.append(dst); strides.append(w); cols.append(0)
        st._gather(fields, widths, perm, outs, strides, cols)
        ds = opt.dev_state
        nat.check(nat.lib().rslrl_fill_zero(ds.loss_sums.data_ptr(), 24, nat.stream_ptr()))
        sums.zero_()
        self._group.zero_grad()
        if self.rnd:
            self._rnd_group.zero_grad()
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        hp = self._hyper()
        sym = self.symmetry
        TD = type(st.observations)
        for _ in range(E):
            for i in range(M):
                r = slice(i * mb, (i + 1) * mb)
                obs_b = TD({k: v[r] for k, v in cf["obs"].items()}, batch_size=[mb], device=self.device)
                obs_in, actions_in, k_aug = obs_b, cf["actions"][r], 1
                if sym:
                    obs_in, actions_in = sym["data_augmentation_func"](obs=obs_b, actions=actions_in, env=sym["_env"])
                    k_aug = int(obs_in.batch_size[0] / mb)
                rows = k_aug * mb
                eng = self._composite_engine(rows)
                x_a = nat.f32c(pol.actor_obs_normalizer(pol.get_actor_obs(obs_in)))
                x_c = x_a if pol.critic_shares_actor_input else nat.f32c(pol.critic_obs_normalizer(pol.get_critic_obs(obs_in)))
                mu, value = eng.mu[:rows], eng.value[:rows]
                eng.forward(x_a, x_c, mu, value)
                if self.rng_compat:   # the sample the reference's policy.act() draws and discards (ppo.py:249)
                    if capturing:
                        d = cf["discard"].get(rows)
                        if d is None:
                            d = cf["discard"][rows] = torch.empty(rows, A, **f32)
                        d.normal_()
                    else:
                        self._skip_discarded_draws(1, rows)
                eng.loss(rows, mb, mu, value, nat.f32c(actions_in), cf["logp"][r], cf["adv"][r], cf["ret"][r],
                         cf["values"][r], cf["mu"][r], cf["sigma"][r], self._group.kl_slot, ds.loss_sums, hp,
                         self.normalize_advantage_per_mini_batch)
                if sym:
                    _, target = sym["data_augmentation_func"](obs=None, actions=mu[:mb], env=sym["_env"])
                    diff = mu[mb:rows] - target[mb:rows]
                    sums[4] += (diff * diff).mean()
                    if sym["use_mirror_loss"]:
                        eng.dmu[mb:rows].add_(diff, alpha=float(sym["mirror_loss_coeff"]) * 2.0 / diff.numel())
                        if eng.f16:   # S * (dL/dmu + mirror gradient) as fp16, S = the scale this loss launch used
                            S = eng.scale_state[0] if eng.scale_state is not None else eng.grad_scale
                            eng.dmu16[mb:rows].copy_((eng.dmu[mb:rows] * S).clamp_(-65504.0, 65504.0))
                if eng.f16:   # fp16 copies of the (augmented) minibatch observations: B operand of the first layer's wgrad
                    x16a = self._x16(cf, "a", x_a)
                    x16c = x16a if x_c is x_a else self._x16(cf, "c", x_c)
                    eng.backward(x16a, x16c, eng.dmu[:rows], eng.dvalue[:rows])
                else:
                    eng.backward(x_a, x_c, eng.dmu[:rows], eng.dvalue[:rows])
                if self.rnd:
                    with torch.no_grad():
                        rnd_state = self.rnd.state_normalizer(self.rnd.get_rnd_state(obs_b))
                    rnd_loss = torch.nn.functional.mse_loss(self.rnd.predictor(rnd_state),
                                                            self.rnd.target(rnd_state).detach())
                    rnd_loss.backward()
                    sums[3] += rnd_loss.detach()
                if self.is_multi_gpu:
                    if not opt.peer_allreduce:
                        torch.distributed.all_reduce(self._group.grads_ext, op=torch.distributed.ReduceOp.SUM)
                    if self.rnd:
                        torch.distributed.all_reduce(self._rnd_group.grads_ext, op=torch.distributed.ReduceOp.SUM)
                opt.fused_step(world_size=self.gpu_world_size, max_grad_norm=self.max_grad_norm, adaptive=adaptive,
                               desired_kl=self.desired_kl or 0.0, round_lr_fp32=self.is_multi_gpu, use_kl=adaptive)
                if self.rnd:
                    self.rnd_optimizer.fused_step(world_size=self.gpu_world_size, max_grad_norm=0.0, use_kl=False)
                eng.refresh_padded(after_step=True)
        # loss accumulators of the fused kernel: [surrogate, value, entropy] -> the composite order
        sums[0:1].copy_(ds.loss_sums[1:2]); sums[1:2].copy_(ds.loss_sums[0:1]); sums[2:3].copy_(ds.loss_sums[2:3])

    def _recurrent_minibatches_fused(self, sums):
        """Minibatch loop of a recurrent policy on the engine, without autograd (reference ppo.py:199-395 with
        rollout_storage.py:206-260): per minibatch (a block of env columns, all T steps) the padded trajectories go
        through both LSTMs together (native forward with saved gates), the outputs are un-padded into the [T * envs, H]
        input of the fused MLP forward, ONE loss launch, the hand-written MLP backward INCLUDING the gradient with respect
        to the LSTM outputs, which is padded back and sent through the native backward-through-time; every weight
        gradient accumulates straight into the flat gradient buffer; one fused optimiser step.  The trajectory index,
        the padded observations, the hidden states at the trajectory starts and the per-sample tensors are built once per
        update (recurrent minibatches are not shuffled)."""
        from ..networks.lstm import _layers_backward, _layers_forward
        from ..utils.utils import TrajectoryIndex
        pol, st, opt, eng = self.policy, self.storage, self.optimizer, self._engine
        M, E = self.num_mini_batches, self.num_learning_epochs
        T, N = st.num_transitions_per_env, st.num_envs
        mbN = N // M
        rows = T * mbN
        A = eng.A
        f32 = dict(dtype=torch.float32, device=self.device)
        TD = type(st.observations)
        rnns = [pol.memory_a.rnn, pol.memory_c.rnn]
        L, H, has_bias = rnns[0].num_layers, rnns[0].hidden_size, bool(rnns[0].bias)
        prec = nat.PREC_CODES["tf32" if pol.memory_a.precision == "fp16" else pol.memory_a.precision]
        index = TrajectoryIndex(st.dones)
        offs = index.offsets_host()           # the one host read of the update
        st._last_trajectory_index = index
        flat = lambda x, w, a, b: x[:, a:b].reshape(rows, w).contiguous()
        cached = []
        for i in range(M):
            a, b = i * mbN, (i + 1) * mbN
            f, l = offs[a], offs[b]
            obs_pad = TD({k: index.split_and_pad(st.observations[k], f, l) for k in st._obs_keys}, batch_size=[T, l - f],
                         device=self.device)
            x_a = nat.f32c(pol.actor_obs_normalizer(pol.get_actor_obs(obs_pad)))
            x_c = nat.f32c(pol.critic_obs_normalizer(pol.get_critic_obs(obs_pad)))
            hid = []
            for saved in (st.saved_hidden_states_a, st.saved_hidden_states_c):
                if saved is None:
                    z = torch.zeros(L, l - f, H, **f32)
                    hid.append((z, z))
                else:
                    hid.append(tuple(index.hidden_at_starts(s_, f, l) for s_ in saved))
            cached.append(dict(a=a, b=b, f=f, l=l, x=(x_a, x_c), hid=hid,
                               actions=flat(st.actions, A, a, b), logp=flat(st.actions_log_prob, 1, a, b).view(-1),
                               adv=flat(st.advantages, 1, a, b).view(-1), ret=flat(st.returns, 1, a, b).view(-1),
                               values=flat(st.values, 1, a, b).view(-1), mu=flat(st.mu, A, a, b),
                               sigma=flat(st.sigma, A, a, b)))
        ds = opt.dev_state
        nat.check(nat.lib().rslrl_fill_zero(ds.loss_sums.data_ptr(), 24, nat.stream_ptr()))
        sums.zero_()
        self._group.zero_grad()
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        hp = self._hyper()
        rc = self._rc
        if rc is None or rc["rows"] != rows:
            rc = self._rc = dict(rows=rows, dxa=torch.empty(rows, H, **f32), dxc=torch.empty(rows, H, **f32),
                                 db=[torch.zeros(4 * H, **f32) for _ in range(2)] if has_bias else [None, None])
        par = lambda r, name, l_: getattr(r, f"{name}_l{l_}")
        rc["updates"] = rc.get("updates", 0) + 1
        # Every minibatch has static shapes within an update (its trajectory count is fixed once the rollout is known), so
        # its ~150 launches (two C time loops of per-step projections + cells among them) are captured into ONE CUDA graph
        # at the first epoch and replayed in the others; the graphs die with the update (the trajectory counts of the
        # next rollout differ).  The first update of a run stays eager (lazy initialisations).
        use_graph = (self.use_cuda_graph and not self.is_multi_gpu and E > 1 and rc["updates"] > 1
                     and os.environ.get("RSLRL_RECURRENT_GRAPH", "1") != "0")
        graphs = [None] * len(cached)

        def minibatch(mb):
                a, b, f, l = mb["a"], mb["b"], mb["f"], mb["l"]
                # ---- both memories forward over the padded trajectories (all layers), gates saved
                inps, saved_layers = list(mb["x"]), []
                for l_ in range(L):
                    res = _layers_forward(inps, [mb["hid"][g][0][l_] for g in range(2)], [mb["hid"][g][1][l_] for g in range(2)],
                                          [par(r, "weight_ih", l_).detach() for r in rnns], [par(r, "weight_hh", l_).detach() for r in rnns],
                                          [par(r, "bias_ih", l_).detach() if has_bias else None for r in rnns],
                                          [par(r, "bias_hh", l_).detach() if has_bias else None for r in rnns], prec, save=True)
                    saved_layers.append(res)
                    inps = [r_[0][1:] for r_ in res]
                xa = index.unpad(inps[0], a, b, f).view(rows, H)
                xc = index.unpad(inps[1], a, b, f).view(rows, H)
                mu, value = eng.mu[:rows], eng.value[:rows]
                eng.forward(xa, xc, mu, value)
                eng.loss(rows, rows, mu, value, mb["actions"], mb["logp"], mb["adv"], mb["ret"], mb["values"], mb["mu"],
                         mb["sigma"], self._group.kl_slot, ds.loss_sums, hp, self.normalize_advantage_per_mini_batch)
                eng.backward(xa, xc, eng.dmu[:rows], eng.dvalue[:rows], dx_a=rc["dxa"], dx_c=rc["dxc"])
                # ---- gradients of the memory outputs back onto the padded trajectories, then backward through time
                douts = [index.pad_columns(rc["dxa"].view(T, mbN, H), a, f, l), index.pad_columns(rc["dxc"].view(T, mbN, H), a, f, l)]
                for l_ in range(L - 1, -1, -1):
                    res = saved_layers[l_]
                    per = [(res[g][3], res[g][0], res[g][1], res[g][2], par(rnns[g], "weight_ih", l_).detach(),
                            par(rnns[g], "weight_hh", l_).detach()) for g in range(2)]
                    for d_ in rc["db"]:
                        if d_ is not None:
                            d_.zero_()
                    targets = [(par(rnns[g], "weight_ih", l_).grad, par(rnns[g], "weight_hh", l_).grad, rc["db"][g]) for g in range(2)]
                    back = _layers_backward(per, [(douts[g], None, None) for g in range(2)], prec, has_bias,
                                            [l_ > 0, l_ > 0], False, targets)
                    if has_bias:
                        for g in range(2):   # b_ih and b_hh receive the same gradient
                            par(rnns[g], "bias_ih", l_).grad.add_(rc["db"][g])
                            par(rnns[g], "bias_hh", l_).grad.add_(rc["db"][g])
                    douts = [back[g][0] for g in range(2)]
                if self.is_multi_gpu and not opt.peer_allreduce:
                    torch.distributed.all_reduce(self._group.grads_ext, op=torch.distributed.ReduceOp.SUM)
                opt.fused_step(world_size=self.gpu_world_size, max_grad_norm=self.max_grad_norm, adaptive=adaptive,
                               desired_kl=self.desired_kl or 0.0, round_lr_fp32=self.is_multi_gpu, use_kl=adaptive)
                eng.refresh_padded(after_step=True)

        for _ in range(E):
            for i, mb in enumerate(cached):
                if self.rng_compat:   # the sample the reference's policy.act() draws and discards (ppo.py:249): host side
                    self._skip_discarded_draws(1, rows)
                if not use_graph:
                    minibatch(mb)
                    continue
                if graphs[i] is None:
                    host_step, before = opt._host_step, nat.launch_count()
                    g = self._capture_graph(lambda mb=mb: minibatch(mb))
                    mb["launches"] = nat.launch_count() - before
                    nat.lib().rslrl_add_launch_count(-mb["launches"] & 0xFFFFFFFFFFFFFFFF)   # capture launched nothing
                    opt._host_step = host_step
                    graphs[i] = g
                graphs[i].replay()
                nat.lib().rslrl_add_launch_count(mb["launches"])
                opt._host_step += 1
        # loss accumulators of the fused kernel: [surrogate, value, entropy] -> the composite order
        sums[0:1].copy_(ds.loss_sums[1:2]); sums[1:2].copy_(ds.loss_sums[0:1]); sums[2:3].copy_(ds.loss_sums[2:3])
        # the graphs of this update stay alive until the next update has captured its own in the same memory pool: the
        # pool then keeps its blocks and the new captures reuse them (no cudaMalloc / cudaFree per update)
        rc["graphs"] = graphs if use_graph else None

    def _capture_graph(self, fn):
        """Captures what `fn` puts on the stream into a CUDA graph in this object's persistent graph memory pool.
        (torch.cuda.graph() empties the caching allocator and runs the garbage collector before every capture -- tens of
        milliseconds, and every later allocation becomes a cudaMalloc; graphs that are re-captured on every update
        cannot afford that.)"""
        if getattr(self, "_graph_pool", None) is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
            self._cap_stream = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        cur = torch.cuda.current_stream()
        side = self._cap_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            g.capture_begin(pool=self._graph_pool)
            try:
                fn()
            finally:
                g.capture_end()
        cur.wait_stream(side)
        return g

    def _x16(self, cf, tag: str, x: torch.Tensor) -> torch.Tensor:
        key = ("x16", tag, x.shape[0], x.shape[1])
        buf = cf.get(key)
        if buf is None:
            buf = cf[key] = torch.zeros(x.shape[0], (x.shape[1] + 7) // 8 * 8, dtype=torch.float16, device=self.device)
        return self._to_f16(x, buf)

    def _composite_minibatches(self, sums):
        """The minibatch loop of the composite update (everything it puts on the stream is graph-capturable for
        feed-forward policies)."""
        if self._composite_fusable():
            return self._composite_minibatches_fused(sums)
        if self._recurrent_fused:
            return self._recurrent_minibatches_fused(sums)
        pol, st, opt = self.policy, self.storage, self.optimizer
        M, E = self.num_mini_batches, self.num_learning_epochs
        once = getattr(st, "_randperm_once", None)
        st._randperm_once = None
        orig_randperm = st._randperm
        if once is not None:   # reuse the permutation drawn before a failed capture (keeps the RNG stream aligned)
            st._randperm = lambda n: once
        gen = (st.recurrent_mini_batch_generator(M, E) if pol.is_recurrent else st.mini_batch_generator(M, E))
        sums.zero_()
        self._group.zero_grad()
        if self.rnd:
            self._rnd_group.zero_grad()
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        clip = self.clip_param

        for (obs_b, actions_b, target_values_b, adv_b, returns_b, old_logp_b, old_mu_b, old_sigma_b, hid_b,
             masks_b) in gen:
            st._randperm = orig_randperm   # the generator has drawn its permutation by now
            num_aug = 1
            original_batch_size = obs_b.batch_size[0]
            if self.normalize_advantage_per_mini_batch:
                with torch.no_grad():
                    adv_b = (adv_b - adv_b.mean()) / (adv_b.std() + 1e-8)
            if self.symmetry and self.symmetry["use_data_augmentation"]:
                aug = self.symmetry["data_augmentation_func"]
                obs_b, actions_b = aug(obs=obs_b, actions=actions_b, env=self.symmetry["_env"])
                num_aug = int(obs_b.batch_size[0] / original_batch_size)
                old_logp_b = old_logp_b.repeat(num_aug, 1)
                target_values_b = target_values_b.repeat(num_aug, 1)
                adv_b = adv_b.repeat(num_aug, 1)
                returns_b = returns_b.repeat(num_aug, 1)

            if pol.is_recurrent and hasattr(pol, "act_evaluate"):   # both memories advance in grouped launches
                _, value = pol.act_evaluate(obs_b, masks=masks_b, hidden_states=hid_b)
                logp = pol.get_actions_log_prob(actions_b)
            else:
                pol.act(obs_b, masks=masks_b, hidden_states=hid_b[0])  # also draws the sample the reference discards
                logp = pol.get_actions_log_prob(actions_b)
                value = pol.evaluate(obs_b, masks=masks_b, hidden_states=hid_b[1])
            mu_b = pol.action_mean[:original_batch_size]
            sigma_b = pol.action_std[:original_batch_size]
            entropy_b = pol.entropy[:original_batch_size]

            if adaptive:
                with torch.no_grad():
                    kl = torch.sum(torch.log(sigma_b / old_sigma_b + 1.0e-5)
                                   + (torch.square(old_sigma_b) + torch.square(old_mu_b - mu_b))
                                   / (2.0 * torch.square(sigma_b)) - 0.5, dim=-1)
                    self._group.kl_slot.copy_(kl.mean().reshape(1))

            adv = torch.squeeze(adv_b)
            ratio = torch.exp(logp - torch.squeeze(old_logp_b))
            surrogate_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
            if self.use_clipped_value_loss:
                value_clipped = target_values_b + (value - target_values_b).clamp(-clip, clip)
                value_loss = torch.max((value - returns_b).pow(2), (value_clipped - returns_b).pow(2)).mean()
            else:
                value_loss = (returns_b - value).pow(2).mean()
            entropy_mean = entropy_b.mean()
            loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_mean

            symmetry_loss = None
            if self.symmetry:
                aug = self.symmetry["data_augmentation_func"]
                if not self.symmetry["use_data_augmentation"]:
                    obs_b, _ = aug(obs=obs_b, actions=None, env=self.symmetry["_env"])
                    num_aug = int(obs_b.shape[0] / original_batch_size)
                mean_actions = pol.act_inference(obs_b.detach().clone())
                _, actions_mean_symm = aug(obs=None, actions=mean_actions[:original_batch_size],
                                           env=self.symmetry["_env"])
                symmetry_loss = torch.nn.functional.mse_loss(mean_actions[original_batch_size:],
                                                             actions_mean_symm.detach()[original_batch_size:])
                if self.symmetry["use_mirror_loss"]:
                    loss = loss + self.symmetry["mirror_loss_coeff"] * symmetry_loss
                else:
                    symmetry_loss = symmetry_loss.detach()

            rnd_loss = None
            if self.rnd:
                with torch.no_grad():
                    rnd_state = self.rnd.state_normalizer(self.rnd.get_rnd_state(obs_b[:original_batch_size]))
                rnd_loss = torch.nn.functional.mse_loss(self.rnd.predictor(rnd_state), self.rnd.target(rnd_state).detach())

            loss.backward()           # accumulates into the (zeroed) flat gradient views
            if self.rnd:
                rnd_loss.backward()
            if self.is_multi_gpu:
                if not opt.peer_allreduce:
                    torch.distributed.all_reduce(self._group.grads_ext, op=torch.distributed.ReduceOp.SUM)
                if self.rnd:
                    torch.distributed.all_reduce(self._rnd_group.grads_ext, op=torch.distributed.ReduceOp.SUM)
            opt.fused_step(world_size=self.gpu_world_size, max_grad_norm=self.max_grad_norm, adaptive=adaptive,
                           desired_kl=self.desired_kl or 0.0, round_lr_fp32=self.is_multi_gpu, use_kl=adaptive)
            if self.rnd:  # the reference neither clips the RND gradients nor adapts their learning rate
                self.rnd_optimizer.fused_step(world_size=self.gpu_world_size, max_grad_norm=0.0, use_kl=False)

            with torch.no_grad():
                sums[0] += value_loss.detach()
                sums[1] += surrogate_loss.detach()
                sums[2] += entropy_mean.detach()
                if rnd_loss is not None:
                    sums[3] += rnd_loss.detach()
                if symmetry_loss is not None:
                    sums[4] += symmetry_loss.detach()

    def _composite_finish(self, sums):
        opt, st = self.optimizer, self.storage
        n = self.num_learning_epochs * self.num_mini_batches
        info = opt.dev_state.read()
        host = (sums / n).cpu().tolist()
        self.learning_rate = info["lr"]
        opt.sync_host_state(self.learning_rate)
        if self.rnd:
            self.rnd_optimizer.sync_host_state(self.rnd_optimizer.param_groups[0]["lr"])
        self.last_update_info = info
        st.clear()
        loss_dict = {"value_function": host[0], "surrogate": host[1], "entropy": host[2]}
        if self.rnd:
            loss_dict["rnd"] = host[3]
        if self.symmetry:
            loss_dict["symmetry"] = host[4]
        return loss_dict

    # ------------------------------------------------------------------------------------------
    # multi-GPU helpers (reference ppo.py:428-469)
    # ------------------------------------------------------------------------------------------
    def broadcast_parameters(self):
        """Rank 0's parameters to every rank: one flat broadcast per parameter group (the reference
        pickles the state dicts, ppo.py:428-439)."""
        self._group.ensure_bound()
        torch.distributed.broadcast(self._group.values, src=0)
        buffers = [b for b in self.policy.buffers()]
        for b in buffers:
            torch.distributed.broadcast(b, src=0)
        if self.rnd:
            self._rnd_group.ensure_bound()
            torch.distributed.broadcast(self._rnd_group.values, src=0)

    def reduce_parameters(self):
        """Averages the gradients over all ranks in place (reference ppo.py:441-469)."""
        self._group.ensure_bound()
        groups = [self._group] + ([self._rnd_group] if self.rnd else [])
        for g in groups:
            torch.distributed.all_reduce(g.grads, op=torch.distributed.ReduceOp.SUM)
            g.grads /= self.gpu_world_size
