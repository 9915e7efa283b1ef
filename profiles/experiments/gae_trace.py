"""In-kernel timeline of gae_norm_reg_kernel (trace build: python -m rsl_rl_b200._build --out rsl_rl_b200/lib/librslrl_b200_trace.so
-DRSLRL_GAE_TRACE -DRSLRL_AR_TRACE; run with RSLRL_B200_LIB=... python profiles/experiments/gae_trace.py)."""
import ctypes as C, os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.storage import RolloutStorage
from rsl_rl_b200.utils import TensorDictLite

dev = "cuda:0"
lib = nat.lib()
lib.rslrl_debug_read_trace_gae.restype = C.c_int
names = ["entry->loads+scan", "block reduce+publish", "barrier wait", "stats", "stores"]
for N in (4096, 65536):
    T = 24
    sts = []
    for _ in range(12):
        st = RolloutStorage("rl", N, T, TensorDictLite({"policy": torch.zeros(N, 1, device=dev)}, batch_size=[N]), [1], dev)
        st.rewards.normal_(); st.values.normal_()
        st.dones.copy_((torch.rand(T, N, 1, device=dev) < 0.02).to(torch.uint8))
        sts.append(st)
    last = torch.randn(N, 1, device=dev)
    flush = torch.empty(64 * 2 ** 20, device=dev)
    rows = []
    for rep in range(12):
        st = sts[rep]
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.compute_returns(last, 0.99, 0.95); e1.record(); e1.synchronize()
        buf = (C.c_ulonglong * (148 * 8))()
        assert lib.rslrl_debug_read_trace_gae(buf) == 0
        threads = max(128, ((N + 147) // 148 + 31) // 32 * 32)
        blocks = (N + threads - 1) // threads
        stp = [[buf[b * 8 + s] for s in range(6)] for b in range(148)]
        newest = max(r[0] for r in stp)   # blocks of THIS launch: entry stamps within 1 ms of the newest one
        stp = [r for r in stp if r[0] and newest - r[0] < 1_000_000]
        blocks = len(stp)
        t0 = min(r[0] for r in stp)
        rows.append(dict(event_us=e0.elapsed_time(e1) * 1e3, kernel_us=(max(r[5] for r in stp) - t0) / 1e3,
                         entry_spread=(max(r[0] for r in stp) - t0) / 1e3,
                         phases=[statistics.median(r[k + 1] - r[k] for r in stp) / 1e3 for k in range(5)],
                         last_arrival=(max(r[2] for r in stp) - t0) / 1e3))
    rows = rows[2:]
    med = lambda f: statistics.median(f(r) for r in rows)
    print(f"N={N} blocks={blocks}x{threads}: event {med(lambda r: r['event_us']):.2f} us, first entry -> last exit "
          f"{med(lambda r: r['kernel_us']):.2f} us, entry spread {med(lambda r: r['entry_spread']):.2f}, last barrier arrival at "
          f"{med(lambda r: r['last_arrival']):.2f}; median per-block phases: "
          + ", ".join(f"{n} {med(lambda r, k=k: r['phases'][k]):.2f}" for k, n in enumerate(names)))
