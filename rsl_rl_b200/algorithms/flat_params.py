"""Flat parameter / gradient / Adam-moment buffers with the parameters re-bound as views.

The reference keeps 17 separate parameter tensors, concatenates their gradients for every all-reduce
(`torch.cat` + 17 `copy_` back, rsl_rl/algorithms/ppo.py:441-469) and steps them with foreach-Adam.
Here a parameter group owns four flat fp32 buffers in HBM -- values, gradients, exp_avg, exp_avg_sq --
laid out in `module.parameters()` order (`[std, actor.0.weight, actor.0.bias, ...]`, the reference's
own flattening order), each tensor starting on a 32-byte boundary.  `param.data` and `param.grad`
are views, so state_dict()/load_state_dict()/checkpoints see ordinary tensors while the kernels see
one contiguous span: the gradient all-reduce is in place, clip+Adam is two launches.

One extra 16-byte tail after the gradients carries the minibatch KL so that it rides the same
all-reduce (reference: a separate 4-byte all-reduce plus a broadcast per minibatch, ppo.py:272-290).
"""
from __future__ import annotations

import os
from typing import Iterable, List

import torch
import torch.nn as nn

from .. import _native as nat

ALIGN = 8  # elements: 32 bytes in fp32 (float4 accesses, TMA base alignment) = 16 bytes in the fp16 weight shadow


class FlatParamGroup:
    def __init__(self, params: Iterable[nn.Parameter], device):
        self.params: List[nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FlatParamGroup needs at least one parameter")
        self.device = torch.device(device)
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("rsl_rl_b200 keeps parameters in fp32")
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off  # padded element count P
        f32 = dict(dtype=torch.float32, device=self.device)
        self.values = torch.zeros(off, **f32)
        self.grads_ext = torch.zeros(off + ALIGN, **f32)  # gradients + KL tail
        self.grads = self.grads_ext[:off]
        self.kl_slot = self.grads_ext[off:off + 1]
        self.exp_avg = torch.zeros(off, **f32)
        self.exp_avg_sq = torch.zeros(off, **f32)
        self.values16 = None  # fp16 shadow of `values` (same element offsets), see enable_shadow16()
        self.shadow16_stale = True
        self.bind()

    def _view(self, flat: torch.Tensor, i: int) -> torch.Tensor:
        p = self.params[i]
        return flat[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)

    def bind(self):
        """Copies the current parameter values into the flat buffer and re-points param.data/.grad."""
        with torch.no_grad():
            for i, p in enumerate(self.params):
                v = self._view(self.values, i)
                if p.data.data_ptr() != v.data_ptr():
                    v.copy_(p.data)
                    p.data = v
                g = self._view(self.grads, i)
                if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                    if p.grad is not None:
                        g.copy_(p.grad)
                    p.grad = g

    # ---- fp16 weight shadow (gemm_precision="fp16"): the kind::f16 MMAs read the weights from here
    def enable_shadow16(self):
        if self.values16 is None:
            self.values16 = torch.zeros(self.numel, dtype=torch.float16, device=self.device)
        self.refresh_shadow16()

    def refresh_shadow16(self, force: bool = True):
        """values16 = fp16(values): one launch over the flat buffer.  The one-launch optimiser step keeps the shadow
        current by itself (force=False then skips the launch); parameter loads / broadcasts need force=True."""
        if self.values16 is not None and (force or self.shadow16_stale):
            self.shadow16_stale = False
            nat.check(nat.lib().rslrl_convert_f32_f16(self.values.data_ptr(), self.numel, self.values16.data_ptr(),
                                                      self.numel, 1, self.numel, self.numel, nat.stream_ptr()))

    def shadow16(self, p: nn.Parameter) -> torch.Tensor:
        """The fp16 shadow view of parameter `p` (same shape)."""
        for off, q in zip(self.offsets, self.params):
            if q is p:
                return self.values16[off:off + p.numel()].view(p.shape)
        raise KeyError("parameter is not part of this group")

    def is_bound(self) -> bool:
        base = self.values.data_ptr()
        gbase = self.grads.data_ptr()
        for off, p in zip(self.offsets, self.params):
            if p.data.data_ptr() != base + 4 * off or p.grad is None or p.grad.data_ptr() != gbase + 4 * off:
                return False
        return True

    def ensure_bound(self):
        if not self.is_bound():
            self.bind()

    def moment_views(self, i: int):
        return self._view(self.exp_avg, i), self._view(self.exp_avg_sq, i)

    def make_symmetric(self, group=None, push: bool = True):
        """Peer-addressable (symmetric) buffers for the in-kernel all-reduce (collective call).  Returns
        (peer gradient pointers, peer flag pointers, peer result pointers, peer staging pointers).

        push=True (default): the exchange is pushed through a staging buffer, so the gradient buffer itself STAYS in
        ordinary device memory (the weight-gradient kernels keep reducing into plain cached memory) and only flags,
        staging and result buffers are symmetric.  push=False (round-1 pull protocol): the gradient buffer moves into
        symmetric memory so that the peers can read it over NVLink."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        world = dist.get_world_size(group)
        n = self.numel + ALIGN
        flags = symm_mem.empty(int(nat.lib().rslrl_allreduce_adam_flag_bytes()) // 4, dtype=torch.int32,
                               device=self.device)
        flags.zero_()
        res = symm_mem.empty(self.numel, dtype=torch.float32, device=self.device)   # reduced gradients (two-step mode)
        res.zero_()
        stage = symm_mem.empty(2 * world * (self.numel + ALIGN // 2), dtype=torch.float32, device=self.device)
        stage.zero_()
        g = None
        if not push:
            g = symm_mem.empty(n, dtype=torch.float32, device=self.device)
            g.zero_()
        torch.cuda.synchronize()
        hf = symm_mem.rendezvous(flags, group)
        hr = symm_mem.rendezvous(res, group)
        hs = symm_mem.rendezvous(stage, group)
        hg = None
        if not push:
            hg = symm_mem.rendezvous(g, group)
            with torch.no_grad():
                g.copy_(self.grads_ext)
            self.grads_ext = g
            self.grads = g[:self.numel]
            self.kl_slot = g[self.numel:self.numel + 1]
            self.bind()
        self._symm = (hg, hf, hr, hs, flags, res, stage, g)  # keep the handles (and the mappings) alive
        torch.cuda.synchronize()
        dist.barrier(group)
        grads = [int(p) for p in hg.buffer_ptrs] if hg is not None else [self.grads_ext.data_ptr()] * world
        return (grads, [int(p) for p in hf.buffer_ptrs], [int(p) for p in hr.buffer_ptrs],
                [int(p) for p in hs.buffer_ptrs])

    def zero_grad(self):
        nat.check(nat.lib().rslrl_fill_zero(self.grads_ext.data_ptr(), self.grads_ext.numel() * 4, nat.stream_ptr()))


class DeviceOptState:
    """`rslrl_opt_state_t` in device memory followed by the loss accumulators, read back with ONE copy.

    layout (8-byte words): [0] lr [1] clip_scale [2] total_norm [3] step_size [4] bc2_sqrt [5] kl_mean
                           [6] step (int64) [7] pad | [8..10] loss sums (surrogate, value, entropy) [11] spare
                           [12..15] = 8 floats: the dynamic fp16 gradient scale state of rslrl_ppo_loss_fwd_bwd
    """

    WORDS = 16

    def __init__(self, device, lr: float):
        self.buf = torch.zeros(self.WORDS, dtype=torch.float64, device=device)
        self.as_int = self.buf.view(torch.int64)
        self.buf[0] = lr
        self.loss_sums = self.buf[8:11]
        self.scale_state = self.buf[12:16].view(torch.float32)   # float[8], see include/rslrl_b200.h

    def ptr(self) -> int:
        return self.buf.data_ptr()

    def set_lr(self, lr: float):
        self.buf[0:1].fill_(lr)

    def set_step(self, step: int):
        self.as_int[6:7].fill_(int(step))

    def read(self):
        host = self.buf.cpu()
        step = int(host.view(torch.int64)[6])
        sat = int(host.view(torch.int32)[2 * 12 + 4])   # scale_state[4]: saturated fp16 gradient launches
        if sat:
            self.scale_state[4:5].zero_()
        return dict(lr=float(host[0]), clip_scale=float(host[1]), total_norm=float(host[2]), kl_mean=float(host[5]),
                    step=step, loss_sums=[float(host[8]), float(host[9]), float(host[10])], fp16_saturations=sat,
                    fp16_grad_scale=float(host.view(torch.float32)[2 * 12 + 2]))


class FusedAdam(torch.optim.Adam):
    """`torch.optim.Adam` whose state lives in a FlatParamGroup and whose step is the fused
    clip + KL-adaptive-LR + Adam kernel pair.

    It is a real `torch.optim.Adam` (default hyper-parameters; `state_dict()` has the usual `exp_avg`,
    `exp_avg_sq`, `step` entries per parameter, so reference checkpoints load and save unchanged,
    reference on_policy_runner.py:289-324).  `step()` without arguments behaves like the stock optimiser
    (it runs on the same flat-buffer views); the PPO update calls `fused_step()`.
    """

    def __init__(self, group: FlatParamGroup, lr: float):
        super().__init__(group.params, lr=lr, foreach=None, capturable=False, fused=None)
        self.group = group
        self.dev_state = DeviceOptState(group.device, lr)
        self._opt_ws = torch.zeros(int(nat.lib().rslrl_optim_workspace_bytes(group.numel)), dtype=torch.uint8,
                                   device=group.device)
        self._host_step = 0
        self._install_state()
        # one-launch step (csrc/allreduce_adam.cu): world 1 = the local buffer only; enable_peer_allreduce() adds the
        # peers' buffers (then the gradient all-reduce happens inside the same kernel)
        self._ar_ws = torch.zeros(int(nat.lib().rslrl_allreduce_adam_workspace_bytes()), dtype=torch.uint8,
                                  device=group.device)
        self.peer_world, self.peer_rank = 1, 0
        self._peer_grads = self._peer_flags = self._peer_results = self._peer_stage = None
        self.one_launch = os.environ.get("RSLRL_ONE_LAUNCH_STEP", "1") != "0"

    def enable_peer_allreduce(self, rank: int, world: int, group=None):
        """Gradients move to symmetric memory; from now on fused_step() all-reduces them itself (collective call)."""
        push = os.environ.get("RSLRL_ALLREDUCE_PUSH", "1") != "0"   # 0: pull protocol of round 1
        grads, flags, results, stage = self.group.make_symmetric(group, push=push)
        if len(grads) != world:
            raise RuntimeError("symmetric-memory rendezvous returned %d peers for world size %d" % (len(grads), world))
        self.peer_world, self.peer_rank = world, rank
        self._peer_grads, self._peer_flags = nat.void_array(grads), nat.void_array(flags)
        # two-step reduction (reduce-scatter + all-gather inside every slice) from RSLRL_TWO_STEP_MIN_WORLD ranks up
        two_step = world >= int(os.environ.get("RSLRL_TWO_STEP_MIN_WORLD", "4"))
        self._peer_results = nat.void_array(results) if two_step else None
        # pushed exchange (one NVLink hop per step) unless RSLRL_ALLREDUCE_PUSH=0 (pull protocol of round 1)
        self._peer_stage = nat.void_array(stage) if push else None

    @property
    def peer_allreduce(self) -> bool:
        return self.peer_world > 1

    def _install_state(self):
        for i, p in enumerate(self.group.params):
            m, v = self.group.moment_views(i)
            self.state[p] = {"step": torch.tensor(float(self._host_step)), "exp_avg": m, "exp_avg_sq": v}

    # ---- fused path
    def fused_step(self, *, world_size: int = 1, max_grad_norm: float = 0.0, adaptive: bool = False,
                   desired_kl: float = 0.0, round_lr_fp32: bool = False, use_kl: bool = True, zero_grad: bool = True):
        g = self.group
        pg = self.param_groups[0]
        beta1, beta2 = pg["betas"]
        L = nat.lib()
        if self.peer_allreduce or (self.one_launch and world_size == 1):
            if self.peer_allreduce and world_size != self.peer_world:
                raise RuntimeError("fused_step: world_size does not match the peer group")
            peers = self._peer_grads if self.peer_allreduce else nat.void_array([g.grads_ext.data_ptr()])
            nat.check(L.rslrl_allreduce_clip_adam(
                self.peer_world, self.peer_rank, peers, self._peer_flags,
                self._peer_results if self.peer_allreduce else None,
                self._peer_stage if self.peer_allreduce else None, g.values.data_ptr(), g.exp_avg.data_ptr(),
                g.exp_avg_sq.data_ptr(), g.numel, self.dev_state.ptr(), self._ar_ws.data_ptr(), float(max_grad_norm),
                int(adaptive), float(desired_kl or 0.0), int(round_lr_fp32), int(use_kl), beta1, beta2, pg["eps"],
                int(zero_grad), nat.ptr(g.values16), nat.stream_ptr()))
            g.shadow16_stale = False   # the Adam phase rewrote the fp16 shadow
            self._host_step += 1
            return
        nat.check(L.rslrl_grad_norm_lr(g.grads.data_ptr(), g.numel, g.kl_slot.data_ptr() if use_kl else None,
                                       self.dev_state.ptr(), self._opt_ws.data_ptr(), world_size, float(max_grad_norm),
                                       int(adaptive), float(desired_kl or 0.0), int(round_lr_fp32), beta1, beta2,
                                       nat.stream_ptr()))
        nat.check(L.rslrl_adam_step(g.values.data_ptr(), g.grads.data_ptr(), g.exp_avg.data_ptr(),
                                    g.exp_avg_sq.data_ptr(), g.numel, self.dev_state.ptr(), beta1, beta2, pg["eps"],
                                    int(zero_grad), nat.stream_ptr()))
        g.shadow16_stale = True
        self._host_step += 1

    def sync_host_state(self, lr: float):
        """Mirrors the device scalars into the torch-visible optimiser state (once per update())."""
        for pg in self.param_groups:
            pg["lr"] = lr
        for p in self.group.params:
            self.state[p]["step"] = torch.tensor(float(self._host_step))

    # ---- checkpoint compatibility
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        with torch.no_grad():
            step = 0
            for i, p in enumerate(self.group.params):
                st = self.state.get(p, None)
                m, v = self.group.moment_views(i)
                if st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    step = int(float(st["step"]))
            self._host_step = step
            self._install_state()
            self.dev_state.set_step(step)
            self.dev_state.set_lr(self.param_groups[0]["lr"])
