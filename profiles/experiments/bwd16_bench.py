"""Graph-replayed timing (L2 flushed) of the fp16-storage backward launches of one C2 minibatch:
head backward, wgrad L3/L2/L1, dgrad L3/L2.  usage: python profiles/experiments/bwd16_bench.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import linear_backward_data, linear_backward_weight, linear_backward_head
M, dev, ELU = 24576, "cuda", nat.ACT_CODES["elu"]
dims = [[48, 512, 256, 128, 12], [48, 512, 256, 128, 1]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
h16 = lambda *s: torch.randn(*s, device=dev).half()
def timed(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    ts = []
    for _ in range(10):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); print(f"{name:28s} median {ts[5]:7.1f} us  min {ts[0]:7.1f} us"); return ts[5]
tot = 0.0
hp = []
for d in dims:
    K, N = d[3], d[4]
    dy = torch.zeros(M, (N + 3) // 4 * 4, device=dev)[:, :N]; dy.normal_()
    hp.append((dy, torch.randn(N, K, device=dev), h16(M, K), torch.empty(M, K, device=dev, dtype=torch.float16),
               torch.zeros(N, K, device=dev), torch.zeros(N, device=dev), ELU, 32768.0))
tot += timed("head backward (2 nets)", lambda: linear_backward_head(hp))
allw = []
for l in (2, 1, 0):
    wp, dp = [], []
    for d in dims:
        K, N = d[l], d[l + 1]
        x, w, dy = h16(M, K), (torch.randn(N, K, device=dev) / K ** 0.5).half(), h16(M, N)
        if K == 48 and os.environ.get("XPITCH"):
            x = h16(M, int(os.environ["XPITCH"]))[:, :K]
        wp.append((dy, x, torch.zeros(N, K, device=dev), torch.zeros(N, device=dev), 1.0 / 32768))
        dp.append((dy, w, x, torch.empty(M, K, device=dev, dtype=torch.float16), ELU))
    allw += wp
    timed(f"wgrad layer {l} ({d[l]}->{d[l+1]})", lambda: linear_backward_weight(wp, nat.PREC_FP16))
    if l > 0:
        tot += timed(f"dgrad layer {l} ({d[l+1]}->{d[l]})", lambda: linear_backward_data(dp, nat.PREC_FP16))
allw.sort(key=lambda p: -(p[2].shape[0] * p[2].shape[1]))
tot += timed("wgrad all layers, one launch", lambda: linear_backward_weight(allw, nat.PREC_FP16))
nodb = [(p[0], p[1], p[2], None, p[4]) for p in allw]
timed("  (same without bias gradients)", lambda: linear_backward_weight(nodb, nat.PREC_FP16))
print(f"backward total (head + dgrads + merged wgrad) {tot:.1f} us")
