"""Multi-layer perceptron whose forward/backward run in the library's fused GEMM kernels.

Mirror of `rsl_rl/networks/mlp.py` of the reference: an `nn.Sequential` of `nn.Linear` + activation
modules with identical construction rules (`-1` hidden dim = input dim, tuple output -> Unflatten,
optional last activation), hence identical `state_dict` keys (`0.weight`, `0.bias`, `2.weight`, ...).
The activation modules are kept for structure/printing only: `forward` never calls them, each
activation is fused into the epilogue of the GEMM that produces it (bias + activation in one pass) and
its derivative is fused into the data-gradient GEMM.
"""
from __future__ import annotations

from functools import reduce
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _native as nat
from ..utils.utils import resolve_nn_activation


def activation_code(module: nn.Module) -> int:
    table = {nn.ELU: "elu", nn.SELU: "selu", nn.ReLU: "relu", nn.CELU: "crelu", nn.LeakyReLU: "lrelu",
             nn.Tanh: "tanh", nn.Sigmoid: "sigmoid", nn.Softplus: "softplus", nn.Identity: "identity"}
    for cls, name in table.items():
        if type(module) is cls:
            return nat.ACT_CODES[name]
    raise NotImplementedError(f"no fused epilogue for activation module {module}")


def linear_forward(problems: Sequence[Tuple], precision: int):
    """problems: (x[M,K] (row stride ldx), weight[N,K], bias[N]|None, y[M,N], act_code).  One launch."""
    arr = (nat.LinearFwd * len(problems))()
    for i, (x, w, b, y, act) in enumerate(problems):
        arr[i] = nat.LinearFwd(x.data_ptr(), w.data_ptr(), nat.ptr(b), y.data_ptr(), x.shape[0], w.shape[0],
                               w.shape[1], x.stride(0), y.stride(0), act)
    nat.check(nat.lib().rslrl_linear_forward(len(problems), arr, precision, nat.stream_ptr()))


def mlp3_eligible(dims: Sequence[int], acts: Sequence[int]) -> bool:
    """Shapes the one-launch fused forward (csrc/mlp_fused.cu) covers: 3 hidden layers with one activation,
    N1 in 128..512 (step 128), N2 in {128, 256}, N3 <= 128 (step 32), N4 <= 16, linear output."""
    if len(dims) != 5 or len(acts) != 4:
        return False
    _, n1, n2, n3, n4 = dims
    return (n1 % 128 == 0 and 128 <= n1 <= 512 and n2 in (128, 256) and n3 % 32 == 0 and 32 <= n3 <= 128
            and 1 <= n4 <= 16 and acts[0] == acts[1] == acts[2] and acts[3] == 0)


def mlp3_struct(x_ptr: int, ldx: int, M: int, weights, biases, hs, y, act: int,
                precision: int = nat.PREC_TF32) -> "nat.Mlp3Fwd":
    """weights: 4 tensors [N_l, K_l] (weights[0] may be the zero-padded first-layer shadow); hs: 3 activation
    buffers [>=M, N_l] or None (nothing but y is stored).  precision = PREC_FP16: weights[1:] are the fp16 shadows
    and hs are fp16 buffers."""
    p = nat.Mlp3Fwd()
    p.x = x_ptr
    for l in range(4):
        p.w[l] = weights[l].data_ptr()
        p.b[l] = nat.ptr(biases[l])
    for l in range(3):
        p.h[l] = None if hs is None else hs[l].data_ptr()
        p.ldh[l] = 0 if hs is None else hs[l].stride(0)
    p.y = y.data_ptr()
    p.M, p.K0 = M, weights[0].shape[1]
    p.N1, p.N2, p.N3, p.N4 = (w.shape[0] for w in weights)
    p.ldx, p.ldw0, p.ldy, p.act = ldx, weights[0].stride(0), y.stride(0), act
    p.precision = precision
    return p


def mlp3_forward(structs: Sequence["nat.Mlp3Fwd"]):
    """One launch for all nets (actor and critic)."""
    arr = (nat.Mlp3Fwd * len(structs))(*structs)
    nat.check(nat.lib().rslrl_mlp3_forward(len(structs), arr, nat.stream_ptr()))


def linear_backward_data(problems: Sequence[Tuple], precision: int):
    """problems: (dy[M,N], weight[N,K], h[M,K]|None, dx[M,K], act_code_of_previous_layer[, db_prev[K]|None])."""
    arr = (nat.LinearBwdData * len(problems))()
    for i, prob in enumerate(problems):
        dy, w, h, dx, act = prob[:5]
        db_prev = prob[5] if len(prob) > 5 else None
        arr[i] = nat.LinearBwdData(dy.data_ptr(), w.data_ptr(), nat.ptr(h), dx.data_ptr(), dy.shape[0], w.shape[0],
                                   w.shape[1], dy.stride(0), 0 if h is None else h.stride(0), dx.stride(0), act,
                                   nat.ptr(db_prev))
    nat.check(nat.lib().rslrl_linear_backward_data(len(problems), arr, precision, nat.stream_ptr()))


def linear_backward_weight(problems: Sequence[Tuple], precision: int):
    """problems: (dy[M,N], x[M,K], dw[N,K] (accumulated), db[N]|None (accumulated)[, scale[, scale_dev_ptr]])."""
    arr = (nat.LinearBwdWeight * len(problems))()
    for i, prob in enumerate(problems):
        dy, x, dw, db = prob[:4]
        arr[i] = nat.LinearBwdWeight(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), nat.ptr(db), dy.shape[0],
                                     dw.shape[0], dw.shape[1], dy.stride(0), x.stride(0),
                                     prob[4] if len(prob) > 4 else 1.0, prob[5] if len(prob) > 5 else None)
    nat.check(nat.lib().rslrl_linear_backward_weight(len(problems), arr, precision, nat.stream_ptr()))


HEAD_MAX_N = 12


def head_backward_ok(N: int, K: int, *tensors) -> bool:
    return N <= HEAD_MAX_N and K % 4 == 0 and all(t is None or (t.data_ptr() % (8 if t.dtype == torch.float16 else 16) == 0 and t.stride(0) % 4 == 0)
                                                  for t in tensors)


def linear_backward_head(problems: Sequence[Tuple]):
    """Fused dgrad + wgrad + bias gradient of a narrow layer.
    problems: (dy[M,N], weight[N,K], h[M,K], dx[M,K]|None, dw[N,K], db[N]|None, act_code_of_previous_layer[, dx_scale]).
    h / dx may be fp16 tensors (the fp16-storage path); dx is then written as dx_scale * gradient."""
    arr = (nat.LinearBwdHead * len(problems))()
    for i, prob in enumerate(problems):
        dy, w, h, dx, dw, db, act = prob[:7]
        f16 = h.dtype == torch.float16
        if dx is not None and (dx.dtype == torch.float16) != f16:
            raise TypeError("linear_backward_head: h and dx must have the same dtype")
        arr[i] = nat.LinearBwdHead(dy.data_ptr(), w.data_ptr(), h.data_ptr(), nat.ptr(dx), dw.data_ptr(), nat.ptr(db),
                                   dy.shape[0], w.shape[0], w.shape[1], dy.stride(0), h.stride(0),
                                   0 if dx is None else dx.stride(0), act, int(f16),
                                   prob[7] if len(prob) > 7 else 1.0)
    nat.check(nat.lib().rslrl_linear_backward_head(len(problems), arr, nat.stream_ptr()))


class _MLPFunction(torch.autograd.Function):
    """Differentiable wrapper: forward = chain of fused Linear+activation launches, backward = fused
    dgrad (+activation derivative) and split-K wgrad launches.  Used whenever an MLP is called under
    autograd (RND, symmetry, recurrent policies); the feed-forward PPO update bypasses autograd."""

    @staticmethod
    def forward(ctx, x, acts: Tuple[int, ...], precision: int, *params):
        nat.require_cuda(x, *params)
        in_dim = params[0].shape[1]
        x2 = nat.f32c(x.reshape(-1, in_dim))
        M = x2.shape[0]
        saved = [x2]
        h = x2
        for li in range(len(acts)):
            w, b = params[2 * li], params[2 * li + 1]
            y = torch.empty(M, w.shape[0], dtype=torch.float32, device=x.device)
            if M > 0:
                linear_forward([(h, w, b, y, acts[li])], precision)
            saved.append(y)
            h = y
        ctx.acts, ctx.precision = acts, precision
        ctx.save_for_backward(*saved, *params)
        ctx.x_shape = x.shape
        return h.reshape(*x.shape[:-1], h.shape[-1])

    @staticmethod
    def backward(ctx, dy):
        acts, precision = ctx.acts, ctx.precision
        L = len(acts)
        saved = ctx.saved_tensors
        hs, params = saved[:L + 1], saved[L + 1:]
        M = hs[0].shape[0]
        d = nat.f32c(dy.reshape(M, -1))
        if acts[-1] != 0:  # a last activation: fold its derivative with an identity "layer"
            raise NotImplementedError("last_activation is not differentiable through the fused path yet")
        grads: List[torch.Tensor] = [None] * (2 * L)
        need_dx = ctx.needs_input_grad[0]
        for li in range(L - 1, -1, -1):
            w, b = params[2 * li], params[2 * li + 1]
            dw, db = torch.zeros_like(w), torch.zeros_like(b)
            if M > 0:
                linear_backward_weight([(d, hs[li], dw, db)], precision)
            grads[2 * li], grads[2 * li + 1] = dw, db
            if li > 0 or need_dx:
                dx = torch.empty(M, w.shape[1], dtype=torch.float32, device=d.device)
                if M > 0:
                    # (li == 0: identity in front of the first layer; the layer INPUT stands in for the saved activation so
                    #  that the tensor-core data-gradient kernel runs -- its factor is 1 whatever the tensor holds)
                    linear_backward_data([(d, w, hs[li], dx, acts[li - 1] if li > 0 else 0)], precision)
                d = dx
        dx_out = d.reshape(ctx.x_shape) if need_dx else None
        return (dx_out, None, None, *grads)


class MLP(nn.Sequential):
    """Multi-layer perceptron (reference rsl_rl/networks/mlp.py:15-120)."""

    def __init__(self, input_dim: int, output_dim, hidden_dims, activation: str = "elu",
                 last_activation: "str | None" = None):
        super().__init__()
        act = resolve_nn_activation(activation)
        last_act = resolve_nn_activation(last_activation) if last_activation is not None else None
        dims = [input_dim if d == -1 else d for d in hidden_dims]
        layers: List[nn.Module] = [nn.Linear(input_dim, dims[0]), act]
        for i in range(len(dims) - 1):
            layers += [nn.Linear(dims[i], dims[i + 1]), act]
        self._unflatten = None
        if isinstance(output_dim, int):
            layers.append(nn.Linear(dims[-1], output_dim))
        else:
            total = reduce(lambda a, b: a * b, output_dim)
            layers.append(nn.Linear(dims[-1], total))
            layers.append(nn.Unflatten(dim=-1, unflattened_size=output_dim))
        if last_act is not None:
            layers.append(last_act)
        for idx, layer in enumerate(layers):
            self.add_module(f"{idx}", layer)
        self.precision = "fp32"

    # ---- structure helpers used by the fused paths
    def linear_layers(self) -> List[nn.Linear]:
        return [m for m in self if isinstance(m, nn.Linear)]

    def layer_activation_codes(self) -> Tuple[int, ...]:
        """Activation code applied after each Linear layer (0 = none)."""
        mods = list(self)
        codes = []
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if isinstance(nxt, nn.Unflatten):
                    nxt = mods[i + 2] if i + 2 < len(mods) else None
                codes.append(0 if nxt is None or isinstance(nxt, (nn.Linear, nn.Unflatten)) else activation_code(nxt))
        return tuple(codes)

    def init_weights(self, scales):
        """Orthogonal init with per-layer gains (reference mlp.py:85-104)."""
        for idx, module in enumerate(self):
            if isinstance(module, nn.Linear):
                gain = scales[idx] if isinstance(scales, (list, tuple)) else scales
                nn.init.orthogonal_(module.weight, gain=gain)
                nn.init.zeros_(module.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.jit.is_scripting():
            # export surface (reference users script / trace `policy.actor` for deployment): the module IS an
            # nn.Sequential of Linear / activation / Unflatten modules, so its TorchScript is the plain layer chain
            for module in self:
                x = module(x)
            return x
        else:
            return self._forward_kernels(x)

    @torch.jit.unused
    def _forward_kernels(self, x: torch.Tensor) -> torch.Tensor:
        if torch.jit.is_tracing():   # torch.jit.trace / ONNX export record standard operators
            for module in self:
                x = module(x)
            return x
        params = []
        for lin in self.linear_layers():
            params += [lin.weight, lin.bias]
        out = _MLPFunction.apply(x, self.layer_activation_codes(), nat.PREC_CODES[self.precision], *params)
        for m in self:
            if isinstance(m, nn.Unflatten):
                out = m(out)
        return out

    def reset(self, dones=None, hidden_states=None):
        pass

    def detach_hidden_states(self, dones=None):
        pass
