"""Launches ONE grouped Linear kernel a few times (profiling target).
usage: python profiles/experiments/gemm_one.py {fwd|dgrad|wgrad} LAYER {fp32|tf32} [rows]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import linear_forward, linear_backward_data, linear_backward_weight

kind, l, prec = sys.argv[1], int(sys.argv[2]), nat.PREC_CODES[sys.argv[3]]
M = int(sys.argv[4]) if len(sys.argv) > 4 else 24576
dims_a, dims_c = [48, 512, 256, 128, 12], [48, 512, 256, 128, 1]
pad4 = lambda n: (n + 3) // 4 * 4
probs = []
for dims in (dims_a, dims_c):
    K, N = dims[l], dims[l + 1]
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    y = torch.empty(M, N, device="cuda"); dy = torch.randn(M, pad4(N), device="cuda")[:, :N]
    dx = torch.empty(M, K, device="cuda"); dw = torch.zeros(N, K, device="cuda"); dbp = torch.zeros(K, device="cuda")
    probs.append(dict(fwd=(x, w, b, y, 1 if l < 3 else 0), dgrad=(dy, w, x, dx, 1, dbp), wgrad=(dy, x, dw, None))[kind])
fn = dict(fwd=linear_forward, dgrad=linear_backward_data, wgrad=linear_backward_weight)[kind]
for _ in range(5):
    fn(probs, prec)
torch.cuda.synchronize()
print("ok")
