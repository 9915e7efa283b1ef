"""Timing + in-kernel timeline of the fused MLP forward (csrc/mlp_fused.cu) at the C2 minibatch / rollout shapes.
Timing: one launch replayed from a CUDA graph (no host-side tensor-map encoding in the timed region), L2 flushed before.
Timeline: needs a -DRSLRL_TC_TRACE build (RSLRL_B200_LIB=...).  usage: python profiles/experiments/mlp3_trace.py [M] [store] [tf32|fp16]"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import mlp3_forward, mlp3_struct

M = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
store = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
f16 = len(sys.argv) > 3 and sys.argv[3] == "fp16"
adt = torch.float16 if f16 else torch.float32
dev = "cuda"
dims = [[48, 512, 256, 128, 12], [48, 512, 256, 128, 1]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
keep, structs = [], []
for d in dims:
    x = torch.randn(M, d[0], device=dev)
    ws = [torch.randn(d[l + 1], d[l], device=dev) / d[l] ** 0.5 for l in range(4)]
    bs = [torch.randn(d[l + 1], device=dev) for l in range(4)]
    hs = [torch.empty(M, d[l + 1], device=dev, dtype=adt) for l in range(3)] if store else None
    if f16:
        ws = [ws[0]] + [w.half() for w in ws[1:]]
    y = torch.empty(M, d[4], device=dev)
    keep.append((x, ws, bs, hs, y))
    structs.append(mlp3_struct(x.data_ptr(), d[0], M, ws, bs, hs, y, nat.ACT_CODES["elu"], nat.PREC_FP16 if f16 else nat.PREC_TF32))
for _ in range(3):
    mlp3_forward(structs)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    mlp3_forward(structs)
ts = []
for _ in range(10):
    flush.fill_(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
flop = sum(2 * M * a * b for d in dims for a, b in zip(d[:-1], d[1:]))
print(f"mlp3 fused forward {sys.argv[3] if len(sys.argv) > 3 else 'tf32'} M={M} store={store}: median {ts[5]:.1f} us (min {ts[0]:.1f}), {flop / ts[5] / 1e6:.1f} TFLOP/s")
handle = ctypes.CDLL(nat.library_path())
if hasattr(handle, "rslrl_debug_read_trace_mlp3"):
    buf = np.zeros(148 * 32, dtype=np.uint64)
    assert handle.rslrl_debug_read_trace_mlp3(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(148, 32).astype(np.int64)
    t0 = t[:, 0].min()
    names = ["L1(0)mma", "E1(0)", "E1(last)", "acc2pub", "E2", "acc3pub", "E3", "acc4pub", "E4"]
    for cta in (0, 73, 147):
        r = (t[cta] - t0) / 1e3
        line = f"  cta {cta:3d}: start {r[0]:.2f} prologue {r[1]:.2f}"
        for j in range(3):
            seg = [f"{n} {r[2 + 9 * j + i]:.2f}" for i, n in enumerate(names) if t[cta][2 + 9 * j + i] > t0]
            if seg:
                line += f"\n      tile {j}: " + " | ".join(seg)
        print(line)
