"""In-kernel timeline (globaltimer) of one grouped forward launch per layer; needs a build with the TRACE hooks
(-DRSLRL_TC_TRACE: RSLRL_EXTRA_NVCC_FLAGS=-DRSLRL_TC_TRACE python -m rsl_rl_b200._build).  usage: python profiles/experiments/gemm_trace.py [f|w|d]"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import linear_forward, linear_backward_data, linear_backward_weight
kind = sys.argv[1] if len(sys.argv) > 1 else "f"
prec = nat.PREC_CODES["tf32"]; M = 24576; dev = "cuda"
dims_a, dims_c = [48, 512, 256, 128, 12], [48, 512, 256, 128, 1]
ELU = nat.ACT_CODES["elu"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pad4 = lambda n: (n + 3) // 4 * 4
handle = ctypes.CDLL(nat.library_path())
for l in range(4):
    probs = []
    for dims in (dims_a, dims_c):
        K, N = dims[l], dims[l + 1]
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev); dy = torch.randn(M, pad4(N), device=dev)[:, :N]
        dx = torch.empty(M, K, device=dev); dw = torch.zeros(N, K, device=dev)
        probs.append({"f": (x, w, b, y, ELU if l < 3 else 0), "d": (dy, w, x, dx, ELU, None), "w": (dy, x, dw, None)}[kind])
    if kind == "d" and l == 0: continue
    fn = {"f": linear_forward, "d": linear_backward_data, "w": linear_backward_weight}[kind]
    for _ in range(3): fn(probs, prec)
    flush.fill_(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(probs, prec); e1.record(); torch.cuda.synchronize()
    buf = np.zeros(148 * 32, dtype=np.uint64)
    assert handle.rslrl_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(148, 32).astype(np.int64)
    t0 = t[:, 0].min()
    rel = lambda a: (a - t0) / 1e3
    print(f"--- {kind} layer {l}: event time {e0.elapsed_time(e1)*1e3:.1f} us; CTA start spread {rel(t[:,0]).max():.2f} us, end min/max {rel(t[:,31]).min():.2f}/{rel(t[:,31]).max():.2f} us")
    for cta in (0, 73, 147):
        r = t[cta]
        ev = [f"start {rel(r[0]):.2f}", f"prologue {rel(r[1]):.2f}", f"tma0 {rel(r[2]):.2f}", f"full0 {rel(r[3]):.2f}"]
        for j in range(13):
            if r[4 + 2 * j] > t0: ev.append(f"mma{j} {rel(r[4+2*j]):.2f}")
            if r[5 + 2 * j] > t0: ev.append(f"epi{j} {rel(r[5+2*j]):.2f}")
        ev.append(f"end {rel(r[31]):.2f}")
        print(f"  cta {cta:3d}: " + " | ".join(ev))
