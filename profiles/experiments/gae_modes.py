"""Per-call time of rslrl_gae_norm replayed back to back from one CUDA graph (bench.py's method) under the launch
variants of gae.cu: RSLRL_GAE_TMA=0 (register kernel, cooperative grid barrier), RSLRL_GAE_CLUSTER_MAX=0 (TMA-staged,
cooperative grid barrier), =8 / =16 (TMA-staged, one cluster).  Also the floor: an empty-ish kernel (fill of 4 bytes)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench

dev = "cuda:0"
for N in (4096, 8192, 65536):
    r = bench.time_gae(dev, N)
    print(f"mode TMA={os.environ.get('RSLRL_GAE_TMA','1')} CLUSTER_MAX={os.environ.get('RSLRL_GAE_CLUSTER_MAX','16')} "
          f"N={N}: {r['us']:.2f} us per call ({r['gbs']:.0f} GB/s)")
from rsl_rl_b200.storage import RolloutStorage
from rsl_rl_b200.utils import TensorDictLite
for N in (4096, 65536):   # no normalisation: no barrier, an ordinary (non-cooperative, non-cluster) launch of the same kernel
    T = 24
    n = bench._n_sets(17 * T * N)
    sts = []
    for _ in range(n):
        st = RolloutStorage("rl", N, T, TensorDictLite({"policy": torch.zeros(N, 1, device=dev)}, batch_size=[N]), [1], dev)
        st.rewards.normal_(); st.values.normal_()
        sts.append(st)
    last = torch.randn(N, 1, device=dev)
    us = bench._rotating_time(lambda i: (lambda st=sts[i]: st.compute_returns(last, 0.99, 0.95, normalize_advantage=False)), n)
    print(f"mode TMA={os.environ.get('RSLRL_GAE_TMA','1')} no-normalisation N={N}: {us:.2f} us per call")
    us = bench._rotating_time(lambda i: (lambda st=sts[0]: st.compute_returns(last, 0.99, 0.95)), n)
    print(f"mode TMA={os.environ.get('RSLRL_GAE_TMA','1')} CLUSTER_MAX={os.environ.get('RSLRL_GAE_CLUSTER_MAX','16')} N={N} SAME buffers every call (warm L2 / TLB): {us:.2f} us per call")
x = [torch.zeros(1, device=dev) for _ in range(64)]
us = bench._rotating_time(lambda i: (lambda t=x[i]: t.fill_(1.0)), 64)
print(f"floor (1-element fill kernels back to back in the same harness): {us:.2f} us per call")
