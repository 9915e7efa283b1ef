"""In-kernel timeline (globaltimer) of the fp16-storage dgrad / wgrad launches of one C2 minibatch; needs a build with
the TRACE hooks: python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_trace.so -DRSLRL_TC_TRACE, then
RSLRL_B200_LIB=rsl_rl_b200/lib/librslrl_trace.so python profiles/experiments/gemm_trace16.py [w|d]"""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import linear_backward_data, linear_backward_weight
kind = sys.argv[1] if len(sys.argv) > 1 else "d"
prec = nat.PREC_FP16; M = 24576; dev = "cuda"
dims_a, dims_c = [48, 512, 256, 128, 12], [48, 512, 256, 128, 1]
ELU = nat.ACT_CODES["elu"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
handle = ctypes.CDLL(nat.library_path())
h16 = lambda *s: torch.randn(*s, device=dev).half()
merged = []
for l in range(3):
    probs = []
    for dims in (dims_a, dims_c):
        K, N = dims[l], dims[l + 1]
        x = h16(M, K); w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        dy = h16(M, N); dx = torch.empty(M, K, device=dev, dtype=torch.float16); dw = torch.zeros(N, K, device=dev)
        db = torch.zeros(N, device=dev)
        probs.append({"d": (dy, w, x, dx, ELU), "w": (dy, x, dw, db, 1.0), "m": (dy, x, dw, db, 1.0)}[kind])
    if kind == "d" and l == 0: continue
    if kind == "m":
        merged += [(dy, x, dw, db, 1.0) for (dy, x, dw, db, _) in [(p[0], p[1], p[2], p[3], 0) for p in probs]]
        if l < 2: continue
        probs = sorted(merged, key=lambda p: -(p[2].shape[0] * p[2].shape[1]))
    fn = {"d": linear_backward_data, "w": linear_backward_weight, "m": linear_backward_weight}[kind]
    for _ in range(3): fn(probs, prec)
    flush.fill_(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(probs, prec); e1.record(); torch.cuda.synchronize()
    if not hasattr(handle, "rslrl_debug_read_trace"):
        print(f"--- fp16 {kind} layer {l}: event time {e0.elapsed_time(e1)*1e3:.1f} us (no trace hooks in this build)")
        continue
    buf = np.zeros(148 * 32, dtype=np.uint64)
    assert handle.rslrl_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(148, 32).astype(np.int64)
    t0 = t[:, 0].min()
    rel = lambda a: (a - t0) / 1e3
    print(f"--- fp16 {kind} layer {l}: event time {e0.elapsed_time(e1)*1e3:.1f} us; CTA start spread {rel(t[:,0]).max():.2f} us, end min/max {rel(t[:,31]).min():.2f}/{rel(t[:,31]).max():.2f} us")
    for cta in (0, 1, 73, 146, 147):
        r = t[cta]
        ev = [f"start {rel(r[0]):.2f}", f"prologue {rel(r[1]):.2f}", f"tma0 {rel(r[2]):.2f}", f"full0 {rel(r[3]):.2f}"]
        for j in range(13):
            if r[4 + 2 * j] > t0: ev.append(f"mma{j} {rel(r[4+2*j]):.2f}")
            if r[5 + 2 * j] > t0: ev.append(f"epi{j} {rel(r[5+2*j]):.2f}")
        ev.append(f"end {rel(r[31]):.2f}")
        print(f"  cta {cta:3d}: " + " | ".join(ev))
