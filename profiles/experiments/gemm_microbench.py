"""Per-layer timing of the grouped (actor+critic) Linear launches at the C1/C2 minibatch shape.
usage: python profiles/experiments/gemm_microbench.py [fp32|tf32] [rows]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import linear_forward, linear_backward_data, linear_backward_weight, linear_backward_head

prec = nat.PREC_CODES[sys.argv[1] if len(sys.argv) > 1 else "tf32"]
M = int(sys.argv[2]) if len(sys.argv) > 2 else 24576
dev = "cuda"
KINDS = os.environ.get("MB_KINDS", "f,w,d,h").split(",")  # restrict the kinds timed (kernel experiments)
dims_a, dims_c = [48, 512, 256, 128, 12], [48, 512, 256, 128, 1]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ELU = nat.ACT_CODES["elu"]

def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

pad4 = lambda n: (n + 3) // 4 * 4
tot_us, tot_flop = 0.0, 0.0
for l in range(4):
    probs_f, probs_d, probs_w, probs_h = [], [], [], []
    flop = 0
    for dims in (dims_a, dims_c):
        K, N = dims[l], dims[l + 1]
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev)
        dy = torch.randn(M, pad4(N), device=dev)[:, :N]
        dx = torch.empty(M, K, device=dev); dw = torch.zeros(N, K, device=dev); dbp = torch.zeros(K, device=dev)
        probs_f.append((x, w, b, y, ELU if l < 3 else 0))
        probs_d.append((dy, w, x, dx, ELU, dbp))
        probs_w.append((dy, x, dw, None))
        probs_h.append((dy, w, x, dx, dw, dbp[:N].contiguous() if N <= K else None, ELU))
        flop += 2 * M * N * K
    tf = timeit(lambda: linear_forward(probs_f, prec)) if 'f' in KINDS else 0.0
    tw = timeit(lambda: linear_backward_weight(probs_w, prec)) if 'w' in KINDS else 0.0
    td = timeit(lambda: linear_backward_data(probs_d, prec)) if l > 0 and 'd' in KINDS else 0.0
    if l == 3 and 'h' in KINDS:
        th = timeit(lambda: linear_backward_head(probs_h))
        print(f"layer 3 fused head backward (dgrad+wgrad+db): {th:.1f} us")
    print(f"layer {l} K={dims_a[l]:4d} N={dims_a[l+1]:4d}  fwd {tf:8.1f} us {(flop/tf/1e6 if tf else 0):7.1f} TF | wgrad {tw:8.1f} us {(flop/tw/1e6 if tw else 0):7.1f} TF"
          f" | dgrad {td:8.1f} us {(flop/td/1e6 if td else 0):7.1f} TF")
    tot_us += tf + tw + td; tot_flop += flop * (3 if l > 0 else 2)
print(f"minibatch GEMM total {tot_us:.1f} us, {tot_flop/tot_us/1e6:.1f} TFLOP/s, x20 = {tot_us*20/1e3:.2f} ms/update")
