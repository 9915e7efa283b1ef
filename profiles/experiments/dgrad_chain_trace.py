"""In-kernel timeline of dgrad_chain16_kernel at the C2 minibatch (trace build:
python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_trace.so -DRSLRL_DC_TRACE; run with RSLRL_B200_LIB=...)."""
import ctypes as C, os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200 import _native as nat

dev = "cuda:0"
lib = nat.lib()
M, N1, N2, N3 = 24576, 512, 256, 128
torch.manual_seed(0)
arr, keep = (nat.DgradChain * 2)(), []
for i in range(2):
    h1, h2 = torch.nn.functional.elu(torch.randn(M, N1, device=dev)).half(), torch.nn.functional.elu(torch.randn(M, N2, device=dev)).half()
    w3, w2 = (torch.randn(N3, N2, device=dev) / N3 ** 0.5).half(), (torch.randn(N2, N1, device=dev) / N2 ** 0.5).half()
    dz3 = torch.randn(M, N3, device=dev).half()
    dz2, dz1 = torch.empty(M, N2, device=dev, dtype=torch.float16), torch.empty(M, N1, device=dev, dtype=torch.float16)
    arr[i] = nat.DgradChain(dz3.data_ptr(), w3.data_ptr(), w2.data_ptr(), h2.data_ptr(), h1.data_ptr(), dz2.data_ptr(),
                            dz1.data_ptr(), M, N1, N2, N3, N3, N2, N1, N2, N1, nat.ACT_CODES["elu"])
    keep.append((h1, h2, w3, w2, dz3, dz2, dz1))
flush = torch.empty(64 * 2 ** 20, device=dev)
for _ in range(3):
    nat.check(lib.rslrl_dgrad_chain16(2, arr, nat.stream_ptr()))
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nat.check(lib.rslrl_dgrad_chain16(2, arr, nat.stream_ptr())); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"dgrad_chain16 M={M}: median {statistics.median(ts):.1f} us (min {min(ts):.1f})")
lib.rslrl_debug_read_trace_dgrad_chain.restype = C.c_int
buf = (C.c_ulonglong * (148 * 64))()
assert lib.rslrl_debug_read_trace_dgrad_chain(buf) == 0
t0 = min(buf[b * 64] for b in range(148) if buf[b * 64])
names = ["S1 operands", "S1 issued", "pass0 issued", "pass1 issued", "pass2 issued", "pass3 issued", "", "", "E1 done",
         "E pass0", "E pass1", "E pass2", "E pass3"]
for b in (0, 73, 147):
    r = lambda s: (buf[b * 64 + s] - t0) / 1e3 if buf[b * 64 + s] else float("nan")
    print(f"  cta {b:3d}: start {r(0):.2f} prologue {r(1):.2f}")
    for j in range(3):
        base = 2 + 20 * j
        print(f"      tile {j}: " + " | ".join(f"{n} {r(base + k):.2f}" for k, n in enumerate(names) if n))
