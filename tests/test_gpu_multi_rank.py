"""GPU, >= 2 devices: the W > 1 branch of the one-launch optimiser step (peer-memory gradient all-reduce inside
`allreduce_clip_adam_kernel`) against the oracle on the mean of the rank gradients, and bit-identical replicas
(reference ppo.py:441-469).  Launched with torchrun as a subprocess; skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_clip_adam_matches_oracle_on_mean_gradient_two_ranks():
    r = _torchrun("mgpu_allreduce_parity.py", 2, 29517)
    assert r.returncode == 0 and "mgpu_allreduce_parity ok" in r.stdout, r.stdout[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_learn_keeps_replicas_bit_identical_graph_and_eager_two_ranks():
    r = _torchrun("mgpu_check.py", 2, 29518)
    assert r.returncode == 0 and "mgpu_check ok" in r.stdout, r.stdout[-4000:]
