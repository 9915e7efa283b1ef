"""Recurrent memory in front of the MLPs (mirror of `rsl_rl/networks/memory.py`).

State-dict keys stay `rnn.weight_ih_l0`, ... (an `nn.LSTM` / `nn.GRU` module owns the parameters).  The padded
trajectory layout and its inverse come from the segmented-scan kernels (utils.TrajectoryIndex).  The LSTM recurrence
runs on this library's kernels (networks/lstm.py: hoisted input projection + per-step recurrent projection on the GEMM
kernels, fused cell kernels, hand-written backward through time) -- `self.rnn` is only the parameter container on the
CUDA path; it is called as a module for host tensors (export / inference off the training path) and for GRUs."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _native as nat
from ..utils.utils import unpad_trajectories
from .lstm import lstm_forward, lstm_forward_group

_GROUP = os.environ.get("RSLRL_LSTM_GROUP", "1") != "0"   # actor + critic memories in grouped launches (A/B switch)


class Memory(nn.Module):
    def __init__(self, input_size, type="lstm", num_layers=1, hidden_size=256):
        super().__init__()
        rnn_cls = nn.GRU if type.lower() == "gru" else nn.LSTM
        self.rnn = rnn_cls(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        self.hidden_states = None
        self._unpad = None  # optional fast unpad (set by the recurrent generator): callable(out) -> [T, N_mb, H]
        self.precision = "fp32"   # GEMM arithmetic of the native LSTM path ("fp32" | "tf32"), set by PPO
        self.native = isinstance(self.rnn, nn.LSTM)

    def _rnn(self, x, hidden):
        if self.native and nat.on_device(x):
            return lstm_forward(self.rnn, x, hidden, self.precision)
        return self.rnn(x, hidden)

    def forward(self, input, masks=None, hidden_states=None):
        batch_mode = masks is not None
        if batch_mode:
            if hidden_states is None:
                raise ValueError("Hidden states not passed to memory module during policy update")
            out, _ = self._rnn(input, hidden_states)
            out = unpad_trajectories(out, masks)
        else:
            out, self.hidden_states = self._rnn(input.unsqueeze(0), self.hidden_states)
        return out

    @staticmethod
    def forward_pair(mem_a: "Memory", mem_c: "Memory", input_a, input_c, masks=None, hidden_a=None, hidden_c=None):
        """`mem_a(input_a, masks, hidden_a)` and `mem_c(input_c, masks, hidden_c)` with both LSTMs advanced together
        (grouped launches) when they are native LSTMs of equal shape on CUDA; otherwise two ordinary calls."""
        ra, rc = mem_a.rnn, mem_c.rnn
        same = (_GROUP and mem_a.native and mem_c.native and nat.on_device(input_a) and nat.on_device(input_c)
                and (ra.num_layers, ra.hidden_size, ra.bias) == (rc.num_layers, rc.hidden_size, rc.bias)
                and input_a.shape[:-1] == input_c.shape[:-1])
        if not same:
            return mem_a(input_a, masks, hidden_a), mem_c(input_c, masks, hidden_c)
        if masks is not None:
            if hidden_a is None or hidden_c is None:
                raise ValueError("Hidden states not passed to memory module during policy update")
            (out_a, _), (out_c, _) = lstm_forward_group([ra, rc], [input_a, input_c], [hidden_a, hidden_c], mem_a.precision)
            return unpad_trajectories(out_a, masks), unpad_trajectories(out_c, masks)
        (out_a, mem_a.hidden_states), (out_c, mem_c.hidden_states) = lstm_forward_group(
            [ra, rc], [input_a.unsqueeze(0), input_c.unsqueeze(0)], [mem_a.hidden_states, mem_c.hidden_states],
            mem_a.precision)
        return out_a, out_c

    def reset(self, dones=None, hidden_states=None):
        if dones is None:
            self.hidden_states = hidden_states
        elif self.hidden_states is not None:
            if hidden_states is None:
                states = self.hidden_states if isinstance(self.hidden_states, tuple) else (self.hidden_states,)
                done = (dones == 1).view(1, -1, 1)  # masked_fill_: same result as h[..., dones == 1, :] = 0 without
                for h in states:                     # the host synchronisation of boolean-mask indexing
                    h.masked_fill_(done, 0.0)
            else:
                raise NotImplementedError(
                    "Resetting hidden states of done environments with custom hidden states is not implemented")

    def detach_hidden_states(self, dones=None):
        if self.hidden_states is None:
            return
        if dones is None:
            if isinstance(self.hidden_states, tuple):
                self.hidden_states = tuple(h.detach() for h in self.hidden_states)
            else:
                self.hidden_states = self.hidden_states.detach()
        else:
            # rows of done environments leave the autograd graph, the others stay attached (reference memory.py:69-74
            # assigns `h[..., dones == 1, :] = h[..., dones == 1, :].detach()`: boolean-mask indexing, one host
            # synchronisation per call; the same selection here is elementwise on the device)
            done = (dones == 1).view(1, -1, 1)
            cut = lambda h: torch.where(done, h.detach(), h) if h.requires_grad else h
            if isinstance(self.hidden_states, tuple):
                self.hidden_states = tuple(cut(h) for h in self.hidden_states)
            else:
                self.hidden_states = cut(self.hidden_states)
