"""Helpers shared by the CPU (oracle-vs-golden) and GPU (CUDA-vs-oracle) parity tests."""
from __future__ import annotations

import ast
import os
from collections import OrderedDict

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def meta_of(g):
    return ast.literal_eval(str(g["meta"]))


def rel_err(a, b):
    """max-abs error over max-abs reference value (the reading of 'relative')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


def params_from(g, prefix, names):
    return OrderedDict((k, g[f"{prefix}.{k}"].astype(np.float32).copy()) for k in names)


def flat(params):
    return np.concatenate([v.reshape(-1) for v in params.values()])


def stream_of(g, it, meta):
    p = f"it{it}.stream."
    groups = meta["obs_groups"]
    cat = lambda names: np.concatenate([g[p + "obs." + n] for n in names], axis=-1)
    return dict(obs_actor=cat(groups["policy"]), obs_critic=cat(groups["critic"]),
                rewards=g[p + "rewards"], dones=g[p + "dones"], time_outs=g[p + "time_outs"])


def cfg_of(meta):
    cfg = dict(meta["alg"])
    cfg["activation"] = meta["activation"]
    cfg["noise_std_type"] = meta["noise_std_type"]
    return cfg
