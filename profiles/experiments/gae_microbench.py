"""GAE kernel timing with a rotating set of rollout buffers (> L2 in total), launches back to back."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200.storage import RolloutStorage
from rsl_rl_b200.utils import TensorDictLite

dev = "cuda:0"
def run(N, T=24, normalize=True, rounds=10):
    nbuf = max(2, int(400e6 // (17 * T * N)) + 1)
    nbuf = min(nbuf, 256)
    sts = []
    for _ in range(nbuf):
        st = RolloutStorage("rl", N, T, TensorDictLite({"p": torch.zeros(N, 1, device=dev)}, batch_size=[N]), [1], dev)
        st.rewards.normal_(); st.values.normal_(); st.dones.copy_((torch.rand(T, N, 1, device=dev) < 0.02).to(torch.uint8))
        sts.append(st)
    last = torch.randn(N, 1, device=dev)
    for st in sts: st.compute_returns(last, 0.99, 0.95, normalize)
    torch.cuda.synchronize()
    # the launches are captured in a CUDA graph: an eager loop is bounded by the ~8 us of Python per call
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for st in sts: st.compute_returns(last, 0.99, 0.95, normalize)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record(); e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (rounds * nbuf)
    nbytes = 17 * T * N + 4 * N
    print(f"N={N:7d} normalize={normalize} buffers={nbuf:3d}: {us:7.2f} us/launch  {nbytes/us/1e3:8.1f} GB/s")
for N in (4096, 65536, 262144):
    for norm in (True, False):
        run(N, normalize=norm)
