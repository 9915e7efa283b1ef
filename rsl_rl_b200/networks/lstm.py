"""LSTM layers on the library's kernels (no cuDNN): hoisted input projection, per-step recurrent projection on the GEMM
kernels, fused cell kernels (csrc/lstm.cu), backward-through-time with the weight gradients as two products over all
time steps.  Replaces what `nn.LSTM` does inside `rsl_rl/networks/memory.py:27-37`; the parameters stay those of an
`nn.LSTM` module (`weight_ih_l{k}`, `weight_hh_l{k}`, `bias_ih_l{k}`, `bias_hh_l{k}`), so state dicts are unchanged.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _native as nat
from .mlp import linear_backward_data, linear_backward_weight, linear_forward

_IDENT = 0


def _cell_forward(xg, hg, b_hh, c_prev, gates, c_out, h_out):
    nat.check(nat.lib().rslrl_lstm_cell_forward(xg.data_ptr(), hg.data_ptr(), nat.ptr(b_hh), c_prev.data_ptr(),
                                                nat.ptr(gates), c_out.data_ptr(), h_out.data_ptr(), xg.shape[0],
                                                c_prev.shape[1], nat.stream_ptr()))


def _cell_backward(dh_out, dh_rec, dc_next, gates, c_prev, c_cur, dgates, dc_prev):
    nat.check(nat.lib().rslrl_lstm_cell_backward(nat.ptr(dh_out), nat.ptr(dh_rec), nat.ptr(dc_next), gates.data_ptr(),
                                                 c_prev.data_ptr(), c_cur.data_ptr(), dgates.data_ptr(),
                                                 dc_prev.data_ptr(), gates.shape[0], c_prev.shape[1], nat.stream_ptr()))


def _layer_forward(x, h0, c0, w_ih, w_hh, b_ih, b_hh, prec: int, save: bool):
    """x [T,n,D], h0/c0 [n,H] -> (h_all [T+1,n,H], c_all [T+1,n,H], gates [T,n,4H] | None, x2d)."""
    T, n, D = x.shape
    H = w_hh.shape[1]
    f32 = dict(dtype=torch.float32, device=x.device)
    x2d = nat.f32c(x.reshape(T * n, D))
    xg = torch.empty(T * n, 4 * H, **f32)
    linear_forward([(x2d, w_ih, b_ih, xg, _IDENT)], prec)              # input projection of ALL steps: one launch
    h_all, c_all = torch.empty(T + 1, n, H, **f32), torch.empty(T + 1, n, H, **f32)
    h_all[0].copy_(h0)
    c_all[0].copy_(c0)
    gates = torch.empty(T, n, 4 * H, **f32) if save else None
    hg = torch.empty(n, 4 * H, **f32)
    # the time loop runs in C: per step one recurrent projection (h_{t-1} W_hh^T) + one fused cell launch
    nat.check(nat.lib().rslrl_lstm_layer_forward(xg.data_ptr(), w_hh.data_ptr(), nat.ptr(b_hh), h_all.data_ptr(),
                                                 c_all.data_ptr(), nat.ptr(gates), hg.data_ptr(), T, n, H, prec,
                                                 nat.stream_ptr()))
    return h_all, c_all, gates, x2d


class _LSTMLayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, c0, w_ih, w_hh, b_ih, b_hh, prec: int):
        nat.require_cuda(x, h0, c0, w_ih, w_hh)
        h_all, c_all, gates, x2d = _layer_forward(x, nat.f32c(h0), nat.f32c(c0), w_ih, w_hh, b_ih, b_hh, prec, save=True)
        ctx.prec = prec
        ctx.has_bias = b_ih is not None
        ctx.save_for_backward(x2d, h_all, c_all, gates, w_ih, w_hh)
        return h_all[1:], h_all[-1].clone(), c_all[-1].clone()

    @staticmethod
    def backward(ctx, dout, dhT, dcT):
        x2d, h_all, c_all, gates, w_ih, w_hh = ctx.saved_tensors
        prec = ctx.prec
        T, n, H4 = gates.shape
        H = H4 // 4
        f32 = dict(dtype=torch.float32, device=gates.device)
        dout = None if dout is None else nat.f32c(dout)
        dgates = torch.empty(T, n, H4, **f32)
        dhT = None if dhT is None else nat.f32c(dhT)
        dcT = None if dcT is None else nat.f32c(dcT)
        dh_bufs, dc_bufs = torch.empty(2, n, H, **f32), torch.empty(2, n, H, **f32)
        # backward through time in C: per step one fused cell-backward launch + dh_{t-1} = dG_t W_hh
        w_hh_t = w_hh.detach().t().contiguous()   # [H, 4H]: dh_{t-1} = dG_t W_hh as a forward product
        nat.check(nat.lib().rslrl_lstm_layer_backward(nat.ptr(dout), nat.ptr(dhT), nat.ptr(dcT), gates.data_ptr(),
                                                      c_all.data_ptr(), w_hh_t.data_ptr(), dgates.data_ptr(),
                                                      dh_bufs.data_ptr(), dc_bufs.data_ptr(),
                                                      int(bool(ctx.needs_input_grad[1])), T, n, H, prec, nat.stream_ptr()))
        dh_rec, dc = dh_bufs[0], dc_bufs[0]
        dg2d = dgates.view(T * n, H4)
        dw_ih, dw_hh = torch.zeros_like(w_ih), torch.zeros_like(w_hh)
        db = torch.zeros(H4, **f32) if ctx.has_bias else None
        # weight gradients of all steps as two products over T*n rows (h_all[:T] = the h_{t-1} of every step)
        linear_backward_weight([(dg2d, x2d, dw_ih, db), (dg2d, h_all[:T].view(T * n, H), dw_hh, None)], prec)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(T * n, w_ih.shape[1], **f32)
            linear_backward_data([(dg2d, w_ih, None, dx, _IDENT)], prec)
            dx = dx.view(T, n, -1)
        dh0 = dh_rec if ctx.needs_input_grad[1] else None
        dc0 = dc if ctx.needs_input_grad[2] else None
        return dx, dh0, dc0, dw_ih, dw_hh, db, (db if ctx.has_bias else None), None


def lstm_forward(rnn: torch.nn.LSTM, x: torch.Tensor, hidden: Optional[Tuple[torch.Tensor, torch.Tensor]],
                 precision: str = "fp32"):
    """`rnn(x, hidden)` of a unidirectional, time-major nn.LSTM without projections, on the library's kernels.
    x [T,n,D]; hidden = (h0, c0), each [L,n,H], or None (zeros).  Returns (out [T,n,H], (h_n, c_n))."""
    if rnn.bidirectional or rnn.batch_first or getattr(rnn, "proj_size", 0) or (rnn.dropout and rnn.training):
        raise NotImplementedError("rsl_rl_b200 LSTM: unidirectional, time-major, no projection, no dropout")
    prec = nat.PREC_CODES["tf32" if precision == "fp16" else precision]
    T, n, _ = x.shape
    L, H = rnn.num_layers, rnn.hidden_size
    if H % 4:
        raise NotImplementedError("rsl_rl_b200 LSTM: hidden size must be a multiple of 4")
    if hidden is None:
        z = torch.zeros(L, n, H, dtype=torch.float32, device=x.device)
        hidden = (z, z)
    h0, c0 = hidden
    grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in rnn.parameters()))
    inp, hn, cn = x, [], []
    for l in range(L):
        w_ih, w_hh = getattr(rnn, f"weight_ih_l{l}"), getattr(rnn, f"weight_hh_l{l}")
        b_ih = getattr(rnn, f"bias_ih_l{l}") if rnn.bias else None
        b_hh = getattr(rnn, f"bias_hh_l{l}") if rnn.bias else None
        if grad:
            out, h_last, c_last = _LSTMLayerFunction.apply(inp, h0[l], c0[l], w_ih, w_hh, b_ih, b_hh, prec)
        else:
            nat.require_cuda(inp, w_ih)
            h_all, c_all, _, _ = _layer_forward(inp, nat.f32c(h0[l]), nat.f32c(c0[l]), w_ih.detach(), w_hh.detach(),
                                                None if b_ih is None else b_ih.detach(),
                                                None if b_hh is None else b_hh.detach(), prec, save=False)
            out, h_last, c_last = h_all[1:], h_all[-1], c_all[-1]
        hn.append(h_last)
        cn.append(c_last)
        inp = out
    return inp, (torch.stack(hn), torch.stack(cn))
