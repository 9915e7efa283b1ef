"""LSTM layers on the library's kernels (no cuDNN): hoisted input projection, per-step recurrent projection on the GEMM
kernels, fused cell kernels (csrc/lstm.cu), backward-through-time with the weight gradients as two products over all
time steps.  Replaces what `nn.LSTM` does inside `rsl_rl/networks/memory.py:27-37`; the parameters stay those of an
`nn.LSTM` module (`weight_ih_l{k}`, `weight_hh_l{k}`, `bias_ih_l{k}`, `bias_hh_l{k}`), so state dicts are unchanged.

Several LSTMs of equal shape that run on the same time grid (the actor's and the critic's memory on the same padded
trajectories) advance TOGETHER: their per-step recurrent projections are one grouped GEMM launch and their cells one
launch (`lstm_forward_group`), which halves the launch count of the recurrent update.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .. import _native as nat
from .mlp import linear_backward_data, linear_backward_weight, linear_forward

_IDENT = 0
_MAX_GROUP = 2


def _layers_forward(xs, h0s, c0s, w_ihs, w_hhs, b_ihs, b_hhs, prec: int, save: bool):
    """One layer of G <= 2 LSTMs on the same [T, n] grid.  xs[g] [T,n,D_g], h0s/c0s[g] [n,H] ->
    per net (h_all [T+1,n,H], c_all [T+1,n,H], gates [T,n,4H] | None, x2d [T*n,D_g])."""
    G = len(xs)
    T, n = xs[0].shape[:2]
    H = w_hhs[0].shape[1]
    f32 = dict(dtype=torch.float32, device=xs[0].device)
    out, arr, keep = [], (nat.LstmFwd * G)(), []
    projections = []
    for g in range(G):
        x2d = nat.f32c(xs[g].reshape(T * n, xs[g].shape[2]))
        xg = torch.empty(T * n, 4 * H, **f32)
        projections.append((x2d, w_ihs[g], b_ihs[g], xg, _IDENT))
        h_all, c_all = torch.empty(T + 1, n, H, **f32), torch.empty(T + 1, n, H, **f32)
        h_all[0].copy_(h0s[g])
        c_all[0].copy_(c0s[g])
        gates = torch.empty(T, n, 4 * H, **f32) if save else None
        hg = torch.empty(n, 4 * H, **f32)
        arr[g] = nat.LstmFwd(xg.data_ptr(), w_hhs[g].data_ptr(), nat.ptr(b_hhs[g]), h_all.data_ptr(), c_all.data_ptr(),
                             nat.ptr(gates), hg.data_ptr())
        keep.append((xg, hg))
        out.append((h_all, c_all, gates, x2d))
    linear_forward(projections, prec)                 # input projections of ALL steps (and all nets): one launch
    # the time loop runs in C: per step one grouped recurrent projection (h_{t-1} W_hh^T) + one fused cell launch
    nat.check(nat.lib().rslrl_lstm_layers_forward(G, arr, T, n, H, prec, nat.stream_ptr()))
    if _state_tap is not None:
        _state_tap.append([(h_all, c_all) for h_all, c_all, _, _ in out])
    return out


_state_tap = None


class tap_states:
    """`with tap_states() as layers:` -- every LSTM layer evaluated inside the block appends its per-net
    `(h_all, c_all)`, the states before step 0 and after every step ([T+1, n, H] each; read-only: they are the buffers the
    backward pass reads).  For callers that need the cell state at a position inside a sequence (the windowed
    distillation update picks every environment's state at the end of its last trajectory); outputs and gradients of the
    LSTM functions are not affected."""

    def __enter__(self):
        global _state_tap
        self._outer, _state_tap = _state_tap, []
        return _state_tap

    def __exit__(self, *exc):
        global _state_tap
        _state_tap = self._outer
        return False


class _LSTMGroupFunction(torch.autograd.Function):
    """forward(prec, G, x_0.., h0_0.., c0_0.., w_ih_0.., w_hh_0.., b_ih_0.., b_hh_0..) -> (out_g, h_last_g, c_last_g) per net."""

    @staticmethod
    def forward(ctx, prec: int, G: int, has_bias: bool, *args):
        xs, h0s, c0s, w_ihs, w_hhs = (args[i * G:(i + 1) * G] for i in range(5))
        b_ihs = args[5 * G:6 * G] if has_bias else (None,) * G
        b_hhs = args[6 * G:7 * G] if has_bias else (None,) * G
        nat.require_cuda(*xs, *w_ihs, *w_hhs)
        res = _layers_forward(xs, [nat.f32c(h) for h in h0s], [nat.f32c(c) for c in c0s], w_ihs, w_hhs, b_ihs, b_hhs,
                              prec, save=True)
        ctx.prec, ctx.G, ctx.has_bias = prec, G, has_bias
        saved = []
        for g in range(G):
            h_all, c_all, gates, x2d = res[g]
            saved += [x2d, h_all, c_all, gates, w_ihs[g], w_hhs[g]]
        ctx.save_for_backward(*saved)
        outs = []
        for h_all, c_all, _, _ in res:
            outs += [h_all[1:], h_all[-1].clone(), c_all[-1].clone()]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        G = ctx.G
        saved = ctx.saved_tensors
        per = [saved[6 * g:6 * g + 6] for g in range(G)]
        need_dx = [ctx.needs_input_grad[3 + g] for g in range(G)]
        need_h0 = any(ctx.needs_input_grad[3 + G + g] for g in range(G))
        res = _layers_backward(per, [grads[3 * g:3 * g + 3] for g in range(G)], ctx.prec, ctx.has_bias, need_dx, need_h0)
        dxs, dh0s, dc0s, dw_ihs, dw_hhs, dbs = [], [], [], [], [], []
        for g, (dx, dh0, dc0, dw_ih, dw_hh, db) in enumerate(res):
            dxs.append(dx)
            dh0s.append(dh0 if ctx.needs_input_grad[3 + G + g] else None)
            dc0s.append(dc0 if ctx.needs_input_grad[3 + 2 * G + g] else None)
            dw_ihs.append(dw_ih); dw_hhs.append(dw_hh); dbs.append(db)
        out = [None, None, None] + dxs + dh0s + dc0s + dw_ihs + dw_hhs
        if ctx.has_bias:
            out += dbs + dbs                                      # b_ih and b_hh receive the same gradient
        return tuple(out)


def _layers_backward(per, grads, prec: int, has_bias: bool, need_dx, need_h0: bool, targets=None):
    """Backward through time of one layer of G <= 2 LSTMs advanced together.
    per[g] = (x2d, h_all, c_all, gates, w_ih, w_hh) as saved by `_layers_forward(save=True)`; grads[g] = (dout [T,n,H] | None,
    dh_T | None, dc_T | None).  targets[g] = (dw_ih, dw_hh, db | None) to ACCUMULATE into (e.g. the .grad views of the flat
    gradient buffer); fresh zero tensors otherwise.  Returns per net (dx [T,n,D] | None, dh0, dc0, dw_ih, dw_hh, db)."""
    G = len(per)
    arr, keep, per_net = (nat.LstmBwd * G)(), [], []
    for g in range(G):
        x2d, h_all, c_all, gates, w_ih, w_hh = per[g]
        T, n, H4 = gates.shape
        H = H4 // 4
        f32 = dict(dtype=torch.float32, device=gates.device)
        dout, dhT, dcT = (None if t is None else nat.f32c(t) for t in grads[g])
        dgates = torch.empty(T, n, H4, **f32)
        dh_bufs, dc_bufs = torch.empty(2, n, H, **f32), torch.empty(2, n, H, **f32)
        w_hh_t = w_hh.detach().t().contiguous()   # [H, 4H]: dh_{t-1} = dG_t W_hh as a forward product
        arr[g] = nat.LstmBwd(nat.ptr(dout), nat.ptr(dhT), nat.ptr(dcT), gates.data_ptr(), c_all.data_ptr(),
                             w_hh_t.data_ptr(), dgates.data_ptr(), dh_bufs.data_ptr(), dc_bufs.data_ptr())
        keep.append((dout, dhT, dcT, w_hh_t))
        per_net.append((x2d, h_all, w_ih, w_hh, dgates, dh_bufs, dc_bufs, T, n, H))
    T, n, H = per_net[0][7:]
    # backward through time in C: per step one fused cell-backward launch + the grouped dh_{t-1} product
    nat.check(nat.lib().rslrl_lstm_layers_backward(G, arr, int(need_h0), T, n, H, prec, nat.stream_ptr()))
    out, wprobs = [], []
    for g in range(G):
        x2d, h_all, w_ih, w_hh, dgates, dh_bufs, dc_bufs, T, n, H = per_net[g]
        dg2d = dgates.view(T * n, 4 * H)
        if targets is not None:
            dw_ih, dw_hh, db = targets[g]
        else:
            dw_ih, dw_hh = torch.zeros_like(w_ih), torch.zeros_like(w_hh)
            db = torch.zeros(4 * H, dtype=torch.float32, device=dgates.device) if has_bias else None
        # weight gradients of all steps as two products over T*n rows (h_all[:T] = the h_{t-1} of every step)
        wprobs += [(dg2d, x2d, dw_ih, db), (dg2d, h_all[:T].view(T * n, H), dw_hh, None)]
        dx = None
        if need_dx[g]:
            dx = torch.empty(T * n, w_ih.shape[1], dtype=torch.float32, device=dgates.device)
            # (x2d stands in for the saved activation of an identity layer: the tensor-core data-gradient kernel then runs)
            linear_backward_data([(dg2d, w_ih, x2d, dx, _IDENT)], prec)
            dx = dx.view(T, n, -1)
        out.append((dx, dh_bufs[0], dc_bufs[0], dw_ih, dw_hh, db))
    for i in range(0, len(wprobs), 2):                       # (the tensor-core wgrad groups two problems per launch)
        linear_backward_weight(wprobs[i:i + 2], prec)
    return out


def lstm_forward_group(rnns: Sequence[torch.nn.LSTM], xs: Sequence[torch.Tensor],
                       hiddens: Sequence[Optional[Tuple[torch.Tensor, torch.Tensor]]], precision: str = "fp32"):
    """`rnn_g(x_g, hidden_g)` for G <= 2 unidirectional, time-major nn.LSTMs of equal (num_layers, hidden_size, bias) on
    the same [T, n] grid, advanced together on the library's kernels.  Returns [(out_g [T,n,H], (h_n, c_n))]."""
    G = len(rnns)
    r0 = rnns[0]
    for rnn in rnns:
        if rnn.bidirectional or rnn.batch_first or getattr(rnn, "proj_size", 0) or (rnn.dropout and rnn.training):
            raise NotImplementedError("rsl_rl_b200 LSTM: unidirectional, time-major, no projection, no dropout")
        if (rnn.num_layers, rnn.hidden_size, rnn.bias) != (r0.num_layers, r0.hidden_size, r0.bias):
            raise ValueError("lstm_forward_group: the LSTMs must have equal num_layers / hidden_size / bias")
    if G > _MAX_GROUP or any(x.shape[:2] != xs[0].shape[:2] for x in xs):
        raise ValueError("lstm_forward_group: at most two nets on the same [T, n] grid")
    prec = nat.PREC_CODES["tf32" if precision == "fp16" else precision]
    T, n = xs[0].shape[:2]
    L, H = r0.num_layers, r0.hidden_size
    if H % 4:
        raise NotImplementedError("rsl_rl_b200 LSTM: hidden size must be a multiple of 4")
    hid = []
    for h in hiddens:
        if h is None:
            z = torch.zeros(L, n, H, dtype=torch.float32, device=xs[0].device)
            h = (z, z)
        hid.append(h)
    grad = torch.is_grad_enabled() and (any(x.requires_grad for x in xs)
                                        or any(p.requires_grad for rnn in rnns for p in rnn.parameters()))
    inps = list(xs)
    hn: List[List[torch.Tensor]] = [[] for _ in range(G)]
    cn: List[List[torch.Tensor]] = [[] for _ in range(G)]
    for l in range(L):
        w_ihs = [getattr(r, f"weight_ih_l{l}") for r in rnns]
        w_hhs = [getattr(r, f"weight_hh_l{l}") for r in rnns]
        b_ihs = [getattr(r, f"bias_ih_l{l}") if r.bias else None for r in rnns]
        b_hhs = [getattr(r, f"bias_hh_l{l}") if r.bias else None for r in rnns]
        h0s, c0s = [h[0][l] for h in hid], [h[1][l] for h in hid]
        if grad:
            args = list(inps) + h0s + c0s + w_ihs + w_hhs + ((b_ihs + b_hhs) if r0.bias else [])
            res = _LSTMGroupFunction.apply(prec, G, bool(r0.bias), *args)
            outs = [(res[3 * g], res[3 * g + 1], res[3 * g + 2]) for g in range(G)]
        else:
            nat.require_cuda(*inps, *w_ihs)
            det = lambda ts: [None if t is None else t.detach() for t in ts]
            r = _layers_forward(inps, [nat.f32c(h) for h in h0s], [nat.f32c(c) for c in c0s], det(w_ihs), det(w_hhs),
                                det(b_ihs), det(b_hhs), prec, save=False)
            outs = [(h_all[1:], h_all[-1], c_all[-1]) for h_all, c_all, _, _ in r]
        for g in range(G):
            hn[g].append(outs[g][1])
            cn[g].append(outs[g][2])
        inps = [o[0] for o in outs]
    return [(inps[g], (torch.stack(hn[g]), torch.stack(cn[g]))) for g in range(G)]


def lstm_forward(rnn: torch.nn.LSTM, x: torch.Tensor, hidden: Optional[Tuple[torch.Tensor, torch.Tensor]],
                 precision: str = "fp32"):
    """`rnn(x, hidden)` of a unidirectional, time-major nn.LSTM without projections, on the library's kernels.
    x [T,n,D]; hidden = (h0, c0), each [L,n,H], or None (zeros).  Returns (out [T,n,H], (h_n, c_n))."""
    return lstm_forward_group([rnn], [x], [hidden], precision)[0]
