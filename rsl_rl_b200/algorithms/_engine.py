"""Launch plans for the feed-forward PPO path: rollout step, last-value evaluation and update().

`FeedForwardEngine` owns every scratch buffer (activations, gradients, permuted minibatch copy) and
turns one PPO phase into a fixed sequence of C-ABI kernel launches with static pointers:

  rollout step   L grouped Linear launches (actor + critic in one grid) -> gaussian head
                 (outputs land directly in row `step` of the rollout buffer)
  update()       1 gather launch for the WHOLE batch (the permutation is the same for every epoch,
                 reference rollout_storage.py:165, so the buffer is permuted once and minibatch i is
                 the contiguous row block i), then per minibatch:
                 L forward launches -> fused loss fwd+bwd -> L wgrad + (L-1) dgrad launches ->
                 [gradient all-reduce] -> grad-norm/KL/LR kernel -> Adam kernel
                 with every scalar (KL, lr, Adam step, loss sums) resident on the device.

Because pointers and shapes are static and nothing returns to the host, the whole update() is captured
once into a CUDA graph and replayed (single-GPU); the host reads back 128 bytes per update().
"""
from __future__ import annotations

from typing import List, Optional

import math
import os

import torch
import torch.nn as nn

from .. import _native as nat
from ..networks.mlp import (MLP, head_backward_ok, linear_backward_data, linear_backward_head,
                            linear_backward_weight, linear_forward, mlp3_eligible, mlp3_forward, mlp3_struct)
from .flat_params import FlatParamGroup, FusedAdam


class _Net:
    """One MLP (actor or critic) seen as lists of (weight, bias, activation code)."""

    def __init__(self, mlp: MLP):
        self.layers: List[nn.Linear] = mlp.linear_layers()
        self.acts = mlp.layer_activation_codes()
        self.dims = [self.layers[0].in_features] + [l.out_features for l in self.layers]
        self.L = len(self.layers)


class _Mlp3Plan:
    """Owner of a `rslrl_mlp3_plan_*` handle; call it with one input pointer per net."""

    def __init__(self, structs, n: int):
        import ctypes as C

        self._lib = nat.lib()
        self._handle = C.c_void_p()
        nat.check(self._lib.rslrl_mlp3_plan_create(n, structs, C.byref(self._handle)))
        self._xs = (C.c_void_p * n)()
        self._run = self._lib.rslrl_mlp3_plan_run

    def __call__(self, *x_ptrs: int):
        xs = self._xs
        for i, p in enumerate(x_ptrs):
            xs[i] = p
        rc = self._run(self._handle, xs, nat.stream_ptr())
        if rc != 0:
            nat.check(rc)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                self._lib.rslrl_mlp3_plan_destroy(h)
            except Exception:
                pass


def supports(policy) -> Optional[str]:
    """None if the fused feed-forward engine can drive `policy`, else the reason it cannot."""
    from ..modules.actor_critic import ActorCritic

    if type(policy) is not ActorCritic:
        return f"policy class {type(policy).__name__} is not the feed-forward ActorCritic"
    if policy.state_dependent_std:
        return "state_dependent_std"
    for mlp in (policy.actor, policy.critic):
        if any(isinstance(m, nn.Unflatten) for m in mlp):
            return "tuple-shaped MLP output"
        if mlp.layer_activation_codes()[-1] != 0:
            return "last_activation"
    return None


class FeedForwardEngine:
    def __init__(self, policy, storage_rows: int, num_envs: int, minibatch_rows: int, device, precision: str,
                 param_group: Optional[FlatParamGroup] = None, scale_state: Optional[torch.Tensor] = None):
        self.policy = policy
        self.device = torch.device(device)
        self.prec = nat.PREC_CODES[precision]
        self.f16 = precision == "fp16"
        self.group = param_group
        self.actor, self.critic = _Net(policy.actor), _Net(policy.critic)
        self.A = self.actor.dims[-1]
        self.std_param = policy.std if policy.noise_std_type == "scalar" else policy.log_std
        self.std_type = nat.STD_SCALAR if policy.noise_std_type == "scalar" else nat.STD_LOG
        self.rows = max(num_envs, minibatch_rows)
        f32 = dict(dtype=torch.float32, device=self.device)
        R = self.rows
        if self.f16:
            # 16-bit storage: saved activations and layer gradients are fp16 in HBM (half the bytes of the update's
            # dominant streams), the kind::f16 MMAs read fp16 weight shadows; layer gradients carry a power-of-two
            # scale (~ minibatch size, undoing the 1/mb of the mean losses) so that they sit in fp16's normal range
            why = self._fp16_ineligible()
            if why:
                raise ValueError(f"gemm_precision='fp16' is not available for this policy: {why}")
            if param_group is None:
                raise ValueError("gemm_precision='fp16' needs the flat parameter group (fp16 weight shadow)")
            param_group.enable_shadow16()
            self.grad_scale = float(2 ** round(math.log2(max(minibatch_rows, 1))))
        else:
            self.grad_scale = 1.0
        adt = dict(dtype=torch.float16 if self.f16 else torch.float32, device=self.device)
        # activations (outputs of hidden layers) and ping-pong gradient buffers
        self.h_a = [torch.empty(R, d, **adt) for d in self.actor.dims[1:-1]]
        self.h_c = [torch.empty(R, d, **adt) for d in self.critic.dims[1:-1]]
        wa, wc = max(self.actor.dims[1:-1] + [1]), max(self.critic.dims[1:-1] + [1])
        if self.f16:   # one gradient buffer per hidden layer: the weight gradients of all layers run as ONE launch at the end
            self.g_a = [torch.empty(R * d, **adt) for d in self.actor.dims[1:-1]]
            self.g_c = [torch.empty(R * d, **adt) for d in self.critic.dims[1:-1]]
        else:
            self.g_a = [torch.empty(R * wa, **adt) for _ in range(2)]
            self.g_c = [torch.empty(R * wc, **adt) for _ in range(2)]
        self.mu = torch.empty(R, self.A, **f32)
        self.value = torch.empty(R, 1, **f32)
        # gradient heads keep a 16-byte row pitch so that TMA can read them (dvalue: 1 used column of 4)
        self.dmu = torch.zeros(R, (self.A + 3) // 4 * 4, **f32)[:, :self.A]
        self.dvalue = torch.zeros(R, 4, **f32)[:, :1]
        # fp16 path, A % 4 == 0: the loss kernel also emits S * dmu, S * dvalue as fp16 (16-byte row pitch) and the output
        # layers run through the tensor-core dgrad / grouped wgrad like every other layer (no narrow-head kernel)
        self.head16 = self.f16 and self.A % 4 == 0 and os.environ.get("RSLRL_HEAD16", "1") != "0"
        self.scale_state = None
        self._scale_calibrated = False
        self._head_struct = None
        # the two middle layers' data gradients as one fused launch (fp16 chain, 3 hidden layers, equal shapes and
        # activations in both nets, N2 in {128, 256}, N1 in {128, 256, 512}); RSLRL_DGRAD_CHAIN=0: two launches
        a, c = self.actor, self.critic
        self.dgrad_chain = (self.f16 and a.L == 4 and c.L == 4 and a.dims[1:4] == c.dims[1:4] and a.acts[:3] == c.acts[:3]
                            and a.acts[0] == a.acts[1] and a.dims[3] in (64, 128)
                            and a.dims[2] in (128, 256) and a.dims[1] in (128, 256, 512)
                            and os.environ.get("RSLRL_DGRAD_CHAIN", "1") != "0")
        # output-layer data gradient fused into the loss kernel (same hidden width and activation in both nets): OPT-IN
        # (RSLRL_FUSED_HEAD_DGRAD=1).  Measured on the final round-2 build, 40-iteration runs, two repeats: the separate
        # K = 12 / K = 1 tensor-core launch 4.74 / 4.80 ms per iteration, the fused variant 4.81 / 5.01 ms -- inside the loss
        # kernel the product sits behind the objective's own latency chain in every block (29.6 us against 15 + 12 us).
        self.fused_head_dgrad = (self.head16 and self.actor.dims[-2] == self.critic.dims[-2]
                                 and self.actor.dims[-2] % 32 == 0 and self.actor.dims[-2] <= 128
                                 and self.actor.acts[-2] == self.critic.acts[-2]
                                 and os.environ.get("RSLRL_FUSED_HEAD_DGRAD", "0") == "1")
        if self.head16:
            f16 = dict(dtype=torch.float16, device=self.device)
            self.dmu16 = torch.zeros(R, (self.A + 7) // 8 * 8, **f16)[:, :self.A]
            self.dvalue16 = torch.zeros(R, 8, **f16)[:, :1]
            # Dynamic power-of-two gradient scale, resident on the device (include/rslrl_b200.h, scale_state): the loss
            # kernel scales its fp16 outputs with the value derived from the previous minibatch's largest gradient and
            # the weight-gradient kernel divides by the same value; [4] counts launches that saturated fp16.
            # It lives in the optimiser's device state block when one is given, so that the saturation counter comes
            # back with the one read of update().
            if os.environ.get("RSLRL_DYNAMIC_SCALE", "1") != "0":
                self.scale_state = (scale_state if scale_state is not None
                                    else torch.zeros(8, dtype=torch.float32, device=self.device))
                self.scale_state.zero_()
                self.scale_state[0], self.scale_state[1], self.scale_state[2] = self.grad_scale, 1.0 / self.grad_scale, self.grad_scale
        # First-layer weights whose row pitch is not a multiple of 16 bytes (e.g. 235 inputs) cannot be described to
        # TMA: keep a zero-padded shadow [N, pad4(K)] that the tensor-core forward reads (refreshed after every
        # optimiser step); inputs are then read at their padded width (pad columns are kept at zero).
        self.w0_pad = {}
        if precision != "fp32":
            for tag, net in (("a", self.actor), ("c", self.critic)):
                K = net.dims[0]
                if K % 4:
                    self.w0_pad[tag] = torch.zeros(net.dims[1], (K + 3) // 4 * 4, **f32)
        self._stage_bufs = {}
        self.refresh_padded()
        # whole-MLP forward in one launch (activations stay on chip between the layers): tf32 mode, eligible shapes
        self.fused_fwd = self.f16 or (precision == "tf32" and os.environ.get("RSLRL_FUSED_MLP", "1") != "0"
                                      and all(mlp3_eligible(n.dims, n.acts) for n in (self.actor, self.critic)))
        self.needs_refresh = bool(self.w0_pad) or self.f16
        self.loss_ws = torch.zeros(int(nat.lib().rslrl_loss_workspace_bytes()), dtype=torch.uint8, device=self.device)
        self.sum_ws = torch.zeros(64, dtype=torch.uint8, device=self.device)
        self.adv_stats = torch.zeros(2, dtype=torch.float64, device=self.device)

    # ---------------------------------------------------------------------------------------------
    def _fp16_ineligible(self) -> Optional[str]:
        for name, net in (("actor", self.actor), ("critic", self.critic)):
            if not mlp3_eligible(net.dims, net.acts):
                return f"{name} {net.dims} is not a 3-hidden-layer MLP the fused forward covers"
            if any(d % 64 for d in net.dims[1:-1]):
                return f"{name} hidden widths {net.dims[1:-1]} must be multiples of 64"
            if net.dims[-1] > 12:
                return f"{name} output width {net.dims[-1]} > 12 (the output layer's backward is the narrow-head kernel)"
        return None

    def refresh_padded(self, after_step: bool = False):
        """Derived copies of the parameters: zero-padded first-layer weights, fp16 weight shadow.
        after_step: called right after an optimiser step (which may have refreshed the shadow itself)."""
        if self.f16:
            self.group.refresh_shadow16(force=not after_step)
        with torch.no_grad():
            for tag, buf in self.w0_pad.items():
                net = self.actor if tag == "a" else self.critic
                buf[:, :net.dims[0]].copy_(net.layers[0].weight)

    def _layer0(self, tag: str, x: torch.Tensor, ld: int = None):
        """(input, weight) of the first layer: the padded pair when a shadow exists and x has the padded pitch."""
        net = self.actor if tag == "a" else self.critic
        buf = self.w0_pad.get(tag)
        ld = x.stride(0) if ld is None else ld
        if buf is not None and ld == buf.shape[1]:
            if x is None:
                return None, buf
            return x.as_strided((x.shape[0], buf.shape[1]), (ld, 1), x.storage_offset()), buf
        return x, net.layers[0].weight

    def forward(self, x_a: torch.Tensor, x_c: torch.Tensor, mu_out: torch.Tensor, v_out: torch.Tensor,
                actor: bool = True, critic: bool = True):
        """Grouped forward of actor and critic on M = x.shape[0] rows; hidden activations stay in
        self.h_a / self.h_c (rows [:M]) for the backward pass."""
        M = (x_a if actor else x_c).shape[0]
        if self.f16 and not self._fused_inputs_ok(x_a if actor else None, x_c if critic else None):
            # the fp16 path has no layer-by-layer forward: stage inputs whose rows are not 16-byte aligned (e.g. 235
            # concatenated observation columns) in a zero-padded buffer
            shared = x_c is x_a
            x_a = self._staged(x_a, "a") if actor else x_a
            x_c = x_a if (shared and actor) else (self._staged(x_c, "c") if critic else x_c)
        if self.fused_fwd and self._fused_inputs_ok(x_a if actor else None, x_c if critic else None):
            nets = []
            if actor:
                nets.append(self._mlp3("a", x_a.data_ptr(), x_a.stride(0), M, mu_out, store=True))
            if critic:
                nets.append(self._mlp3("c", x_c.data_ptr(), x_c.stride(0), M, v_out, store=True))
            mlp3_forward(nets)
            return
        La, Lc = self.actor.L if actor else 0, self.critic.L if critic else 0
        in_a, in_c = x_a, x_c
        for l in range(max(La, Lc)):
            probs = []
            if l < La:
                lin = self.actor.layers[l]
                out = mu_out if l == La - 1 else self.h_a[l][:M]
                xin, w = self._layer0("a", in_a) if l == 0 else (in_a, lin.weight)
                probs.append((xin, w, lin.bias, out, self.actor.acts[l]))
                in_a = out
            if l < Lc:
                lin = self.critic.layers[l]
                out = v_out if l == Lc - 1 else self.h_c[l][:M]
                xin, w = self._layer0("c", in_c) if l == 0 else (in_c, lin.weight)
                probs.append((xin, w, lin.bias, out, self.critic.acts[l]))
                in_c = out
            linear_forward(probs, self.prec)

    def _staged(self, x: torch.Tensor, tag: str) -> torch.Tensor:
        M, K = x.shape
        key = (tag, M, K)
        buf = self._stage_bufs.get(key)
        if buf is None:
            with torch.inference_mode(False):
                buf = torch.zeros(M, (K + 3) // 4 * 4, dtype=torch.float32, device=self.device)
            self._stage_bufs[key] = buf
        view = buf[:, :K]
        view.copy_(x)
        return view

    def _fused_inputs_ok(self, x_a, x_c) -> bool:
        return all(x is None or (x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and x.stride(1) == 1)
                   for x in (x_a, x_c))

    def _mlp3(self, tag: str, x_ptr: int, ldx: int, M: int, out: torch.Tensor, store: bool):
        net, hs = (self.actor, self.h_a) if tag == "a" else (self.critic, self.h_c)
        w0 = self._layer0(tag, None, ldx)[1]
        weights = [w0] + [self.group.shadow16(l.weight) if self.f16 else l.weight for l in net.layers[1:]]
        return mlp3_struct(x_ptr, ldx, M, weights, [l.bias for l in net.layers],
                           [h[:M] for h in hs] if store else None, out, net.acts[0],
                           nat.PREC_FP16 if self.f16 else nat.PREC_TF32)

    # ---- prebuilt launch plans for the rollout: everything static except the observation pointers ---------
    def build_forward_plan(self, M: int, ld_a: int, ld_c: int, mu_out: torch.Tensor, v_out: torch.Tensor,
                           actor: bool = True, critic: bool = True):
        """ctypes launch arrays for `forward` on M rows with outputs fixed; returns a callable(xa_ptr, xc_ptr)."""
        if self.fused_fwd and ld_a % 4 == 0 and ld_c % 4 == 0:
            # rollout: nothing is stored but the outputs (the update recomputes the activations of its minibatches)
            tags = (["a"] if actor else []) + (["c"] if critic else [])
            outs = {"a": mu_out, "c": v_out}
            lds = {"a": ld_a, "c": ld_c}
            structs = [self._mlp3(t, 0, lds[t], M, outs[t], store=False) for t in tags]
            arr = (nat.Mlp3Fwd * len(structs))(*structs)
            fn3, check3, n3 = nat.lib().rslrl_mlp3_forward, nat.check, len(structs)

            def run_fused(xa_ptr: int, xc_ptr: int):
                for i, t in enumerate(tags):
                    arr[i].x = xa_ptr if t == "a" else xc_ptr
                check3(fn3(n3, arr, nat.stream_ptr()))

            return run_fused
        La, Lc = self.actor.L if actor else 0, self.critic.L if critic else 0
        launches = []
        in_a = in_c = None  # layer-0 inputs are patched per call
        for l in range(max(La, Lc)):
            specs = []
            if l < La:
                lin = self.actor.layers[l]
                out = mu_out if l == La - 1 else self.h_a[l][:M]
                specs.append(("a", in_a, lin, out, self.actor.acts[l], ld_a))
                in_a = out
            if l < Lc:
                lin = self.critic.layers[l]
                out = v_out if l == Lc - 1 else self.h_c[l][:M]
                specs.append(("c", in_c, lin, out, self.critic.acts[l], ld_c))
                in_c = out
            arr = (nat.LinearFwd * len(specs))()
            patch = []
            for i, (tag, inp, lin, out, act, ld0) in enumerate(specs):
                xptr = 0 if inp is None else inp.data_ptr()
                ldx = ld0 if inp is None else inp.stride(0)
                w = self._layer0(tag, None, ld0)[1] if inp is None else lin.weight
                arr[i] = nat.LinearFwd(xptr, w.data_ptr(), lin.bias.data_ptr(), out.data_ptr(), M,
                                       w.shape[0], w.shape[1], ldx, out.stride(0), act)
                if inp is None:
                    patch.append((i, tag))
            launches.append((arr, len(specs), patch))
        fn, prec, check = nat.lib().rslrl_linear_forward, self.prec, nat.check

        def run(xa_ptr: int, xc_ptr: int):
            st = nat.stream_ptr()
            for arr, n, patch in launches:
                for i, tag in patch:
                    arr[i].x = xa_ptr if tag == "a" else xc_ptr
                check(fn(n, arr, prec, st))

        return run

    def build_act_plan(self, M: int, ld_a: int, ld_c: int, mu_out, v_out, z, actions_out, logp_out, sigma_out,
                       actions_host=None):
        """One-launch rollout step (fused forward of actor + critic with the Gaussian policy head in the actor's output
        epilogue): a C-side launch plan with every TMA descriptor pre-encoded; returns callable(xa_ptr, xc_ptr).  None
        when the fused forward does not cover this policy / input layout."""
        if not (self.fused_fwd and ld_a % 4 == 0 and ld_c % 4 == 0):
            return None
        sa = self._mlp3("a", 0, ld_a, M, mu_out, store=False)
        sc = self._mlp3("c", 0, ld_c, M, v_out, store=False)
        sa.head_z, sa.head_std, sa.head_std_type = z.data_ptr(), self.std_param.data_ptr(), self.std_type
        sa.head_actions, sa.head_logp, sa.head_sigma = actions_out.data_ptr(), logp_out.data_ptr(), nat.ptr(sigma_out)
        sa.head_actions_host = nat.ptr(actions_host)   # optional second copy of the actions in pinned host memory
        sa.x = sc.x = z.data_ptr()   # placeholder (any valid 16-byte aligned pointer): run() patches the real inputs
        arr = (nat.Mlp3Fwd * 2)(sa, sc)
        return _Mlp3Plan(arr, 2)

    def backward(self, x_a: torch.Tensor, x_c: torch.Tensor, dmu: torch.Tensor, dvalue: torch.Tensor,
                 dx_a: Optional[torch.Tensor] = None, dx_c: Optional[torch.Tensor] = None):
        """dL/dW, dL/db accumulated into the parameters' .grad views (flat gradient buffer).  fp16 mode: x_a / x_c are
        the fp16 copies of the minibatch observations.  dx_a / dx_c ([M, K0], optional, not in fp16 mode): also the
        gradient with respect to the networks' INPUTS (the outputs of a recurrent memory in front of the MLPs)."""
        if self.f16:
            if dx_a is not None or dx_c is not None:
                raise NotImplementedError("input gradients are not available on the fp16-storage chain")
            return self._backward16(x_a, x_c, dmu, dvalue)
        M = x_a.shape[0]
        nets = ((self.actor, self.h_a, self.g_a, x_a, dmu), (self.critic, self.h_c, self.g_c, x_c, dvalue))
        dx_in = (dx_a, dx_c)
        dy = [dmu, dvalue]
        for k in range(max(self.actor.L, self.critic.L)):
            wprobs, dprobs, hprobs, new_dy = [], [], [], list(dy)
            for ni, (net, hs, gbuf, x_in, _) in enumerate(nets):
                l = net.L - 1 - k
                if l < 0:
                    continue
                lin = net.layers[l]
                inp = x_in if l == 0 else hs[l - 1][:M]
                if k == 0 and l > 0 and head_backward_ok(lin.out_features, lin.in_features, inp):
                    # narrow output layer: one fused fp32 launch for dgrad + wgrad + bias gradient
                    dx = gbuf[0][:M * net.dims[l]].view(M, net.dims[l])
                    hprobs.append((dy[ni], lin.weight, inp, dx, lin.weight.grad, lin.bias.grad, net.acts[l - 1]))
                    new_dy[ni] = dx
                    continue
                wprobs.append((dy[ni], inp, lin.weight.grad, lin.bias.grad))
                if l > 0:
                    dx = gbuf[k % 2][:M * net.dims[l]].view(M, net.dims[l])
                    dprobs.append((dy[ni], lin.weight, inp, dx, net.acts[l - 1]))
                    new_dy[ni] = dx
                elif dx_in[ni] is not None:
                    # first layer: identity in front of it (the input tensor stands in for the saved activation)
                    dprobs.append((dy[ni], lin.weight, inp, dx_in[ni], 0))
            if hprobs:
                linear_backward_head(hprobs)
            if wprobs:
                linear_backward_weight(wprobs, self.prec)
            if dprobs:
                linear_backward_data(dprobs, self.prec)
            dy = new_dy

    def fp16_saturation_events(self) -> int:
        """Loss launches (since the last call) whose scaled fp16 output gradients left fp16's finite range (they were
        saturated, the step was still taken).  One device read: call it once per update() at most."""
        if self.scale_state is None:
            return 0
        n = int(self.scale_state[4:5].clone().view(torch.int32).item())
        if n:
            self.scale_state[4:5].zero_()
        return n

    def _backward16(self, x_a: torch.Tensor, x_c: torch.Tensor, dmu: torch.Tensor, dvalue: torch.Tensor):
        """fp16-storage backward: narrow-head kernel (fp32 math, writes S * dZ3 as fp16), one grouped dgrad launch per
        hidden layer down the chain (every dZ_l stays in its own buffer), then ONE wgrad launch for all layers of both
        nets (dW += dZ^T H / S, bias gradients from the same tiles; k-splits balanced across the problems)."""
        M = x_a.shape[0]
        S, prec = self.grad_scale, nat.PREC_FP16
        dyn = self.scale_state is not None
        inv_s = 1.0 if dyn else 1.0 / S                                         # static factor of the wgrad epilogue
        inv_dev = self.scale_state[1:2].data_ptr() if dyn else None            # device-resident 1 / (dynamic scale)
        nets = ((self.actor, self.h_a, self.g_a, x_a, dmu), (self.critic, self.h_c, self.g_c, x_c, dvalue))
        dz = lambda net, gbuf, l: gbuf[l - 1][:M * net.dims[l]].view(M, net.dims[l])   # gradient at the output of layer l-1
        hprobs, wprobs = [], []
        if self.head16:
            # output layers on the tensor cores: dY16 = S * dL/d(mu|V) was written by the loss kernel
            dprobs = []
            for (net, hs, gbuf, _, _), dy16 in zip(nets, (self.dmu16[:M], self.dvalue16[:M])):
                lin, l = net.layers[-1], net.L - 1
                wprobs.append((dy16, hs[l - 1][:M], lin.weight.grad, lin.bias.grad, inv_s, inv_dev))
                dprobs.append((dy16, self.group.shadow16(lin.weight), hs[l - 1][:M], dz(net, gbuf, l), net.acts[l - 1]))
            if not (self.fused_head_dgrad and M <= 148 * 1024):   # (else the loss kernel already wrote both dZ3)
                linear_backward_data(dprobs, prec)
        else:
            for net, hs, gbuf, _, d_out in nets:
                lin, l = net.layers[-1], net.L - 1
                hprobs.append((d_out, lin.weight, hs[l - 1][:M], dz(net, gbuf, l), lin.weight.grad, lin.bias.grad,
                               net.acts[l - 1], S))
            linear_backward_head(hprobs)
        chained = self._dgrad_chain(M, nets, dz)   # both middle layers' data gradients of both nets in ONE launch
        for l in range(self.actor.L - 2, -1, -1):
            dprobs = []
            for net, hs, gbuf, x_in, _ in nets:
                lin = net.layers[l]
                inp = x_in if l == 0 else hs[l - 1][:M]
                dy = dz(net, gbuf, l + 1)
                wprobs.append((dy, inp, lin.weight.grad, lin.bias.grad, inv_s, inv_dev))
                if l > 0 and not chained:
                    dprobs.append((dy, self.group.shadow16(lin.weight), inp, dz(net, gbuf, l), net.acts[l - 1]))
            if dprobs:
                linear_backward_data(dprobs, prec)
        wprobs.sort(key=lambda p: -(p[2].shape[0] * p[2].shape[1]))   # heaviest tiles first in the round-robin schedule
        linear_backward_weight(wprobs, prec)

    def _dgrad_chain(self, M: int, nets, dz) -> bool:
        """dZ2 = (dZ3 W3) * act'(H2), dZ1 = (dZ2 W2) * act'(H1) for actor and critic in one launch
        (`rslrl_dgrad_chain16`, csrc/dgrad_fused.cu).  False: not eligible / switched off (two layer-by-layer launches)."""
        if not self.dgrad_chain:
            return False
        arr = (nat.DgradChain * 2)()
        for i, (net, hs, gbuf, _, _) in enumerate(nets):
            l3, l2 = net.layers[2], net.layers[1]
            d3, d2, d1 = dz(net, gbuf, 3), dz(net, gbuf, 2), dz(net, gbuf, 1)
            h2, h1 = hs[1][:M], hs[0][:M]
            arr[i] = nat.DgradChain(d3.data_ptr(), self.group.shadow16(l3.weight).data_ptr(),
                                    self.group.shadow16(l2.weight).data_ptr(), h2.data_ptr(), h1.data_ptr(), d2.data_ptr(),
                                    d1.data_ptr(), M, net.dims[1], net.dims[2], net.dims[3], d3.stride(0), h2.stride(0),
                                    h1.stride(0), d2.stride(0), d1.stride(0), net.acts[1])
        nat.check(nat.lib().rslrl_dgrad_chain16(2, arr, nat.stream_ptr()))
        return True

    def _loss_head(self, rows: int):
        """rslrl_loss_head_t for the loss launch: the data gradient of the two output layers fused into the loss kernel
        (fp16 chain with fp16 output-gradient copies, four-threads-per-row kernel); None otherwise."""
        if not self.fused_head_dgrad or rows > 148 * 1024:
            return None
        if self._head_struct is None:
            la, lc = self.actor.layers[-1], self.critic.layers[-1]
            K3 = self.actor.dims[-2]
            h = nat.LossHead(la.weight.data_ptr(), lc.weight.data_ptr(), self.h_a[-1].data_ptr(), self.h_c[-1].data_ptr(),
                             self.g_a[-1].data_ptr(), self.g_c[-1].data_ptr(), K3, K3, self.actor.acts[-2])
            import ctypes
            self._head_struct = (h, ctypes.byref(h))
        return self._head_struct[1]

    def gaussian_head(self, mu, z, actions_out, logp_out, sigma_out):
        nat.check(nat.lib().rslrl_gaussian_act(mu.data_ptr(), self.std_param.data_ptr(), self.std_type, z.data_ptr(),
                                               actions_out.data_ptr(), logp_out.data_ptr(), nat.ptr(sigma_out),
                                               mu.shape[0], self.A, nat.stream_ptr()))

    def loss(self, rows: int, mb_orig: int, mu, value, actions, old_logp, adv, returns, old_values, old_mu, old_sigma,
             kl_out, loss_sums, hp, per_mb_norm: bool):
        L = nat.lib()
        stats_ptr = None
        if per_mb_norm:
            nat.check(L.rslrl_sum_sumsq(adv.data_ptr(), mb_orig, self.adv_stats.data_ptr(), self.sum_ws.data_ptr(),
                                        nat.stream_ptr()))
            stats_ptr = self.adv_stats.data_ptr()
        def launch(dstd_ptr, kl_ptr, sums_ptr):
            nat.check(L.rslrl_ppo_loss_fwd_bwd(
                mu.data_ptr(), self.std_param.data_ptr(), self.std_type, value.data_ptr(), actions.data_ptr(),
                old_logp.data_ptr(), adv.data_ptr(), returns.data_ptr(), old_values.data_ptr(), old_mu.data_ptr(),
                old_sigma.data_ptr(), stats_ptr, self.dmu.data_ptr(), self.dmu.stride(0), self.dvalue.data_ptr(),
                self.dvalue.stride(0), dstd_ptr, None, None, kl_ptr, sums_ptr,
                self.loss_ws.data_ptr(), rows,
                mb_orig, self.A, hp["clip_param"], hp["value_loss_coef"], hp["entropy_coef"],
                int(hp["use_clipped_value_loss"]),
                self.dmu16.data_ptr() if self.head16 else None, self.dmu16.stride(0) if self.head16 else 0,
                self.dvalue16.data_ptr() if self.head16 else None, self.dvalue16.stride(0) if self.head16 else 0,
                self.grad_scale, nat.ptr(self.scale_state), self._loss_head(rows), nat.stream_ptr()))

        if self.scale_state is not None and not self._scale_calibrated and not torch.cuda.is_current_stream_capturing():
            # The very first minibatch of a run: one extra launch whose only lasting effect is scale_state[2] -- the
            # dynamic scale derived from THIS minibatch's largest output gradient (its other outputs go to scratch or
            # are rewritten by the real launch below).  Every later launch scales with its predecessor's value.
            self._scale_calibrated = True
            scratch = torch.zeros(self.A, dtype=torch.float32, device=self.device)
            launch(scratch.data_ptr(), None, None)
            self.scale_state[4:5].zero_()   # (a saturation against the initial guess is not an event)
        launch(self.std_param.grad.data_ptr(), nat.ptr(kl_out), nat.ptr(loss_sums))
