"""ctypes binding of `include/rslrl_b200.h` (the C ABI of the sm_100a kernels).

There is exactly one arithmetic back end.  If the shared library is missing, or a call is made with
tensors that are not on a CUDA device, this module raises -- it never computes anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _build

c_void_p, c_int, c_i64, c_double = C.c_void_p, C.c_int, C.c_int64, C.c_double

# enums of include/rslrl_b200.h
U8, I32, I64, F32 = 0, 1, 2, 3
ACT_CODES = {"identity": 0, "elu": 1, "relu": 2, "tanh": 3, "sigmoid": 4, "lrelu": 5, "selu": 6, "softplus": 7,
             "crelu": 1}  # nn.CELU(alpha=1) == nn.ELU(alpha=1)
PREC_FP32, PREC_TF32, PREC_FP16, PREC_TF32X3 = 0, 1, 2, 3
PREC_CODES = {"fp32": PREC_FP32, "tf32": PREC_TF32, "fp16": PREC_FP16, "tf32x3": PREC_TF32X3}
STD_SCALAR, STD_LOG = 0, 1
MAX_SEGMENTS, MAX_PROBLEMS = 16, 8
ABI_VERSION = 3   # RSLRL_ABI_VERSION of include/rslrl_b200.h


class LinearFwd(C.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("b", c_void_p), ("y", c_void_p),
                ("M", c_int), ("N", c_int), ("K", c_int), ("ldx", c_int), ("ldy", c_int), ("act", c_int)]


class LinearBwdData(C.Structure):
    _fields_ = [("dy", c_void_p), ("w", c_void_p), ("h", c_void_p), ("dx", c_void_p),
                ("M", c_int), ("N", c_int), ("K", c_int), ("ldy", c_int), ("ldh", c_int), ("ldx", c_int),
                ("act", c_int), ("db_prev", c_void_p)]


class LinearBwdWeight(C.Structure):
    _fields_ = [("dy", c_void_p), ("x", c_void_p), ("dw", c_void_p), ("db", c_void_p),
                ("M", c_int), ("N", c_int), ("K", c_int), ("ldy", c_int), ("ldx", c_int), ("scale", C.c_float),
                ("scale_dev", c_void_p)]


class LinearBwdHead(C.Structure):
    _fields_ = [("dy", c_void_p), ("w", c_void_p), ("h", c_void_p), ("dx", c_void_p), ("dw", c_void_p),
                ("db", c_void_p), ("M", c_int), ("N", c_int), ("K", c_int), ("ldy", c_int), ("ldh", c_int),
                ("ldx", c_int), ("act", c_int), ("f16", c_int), ("dx_scale", C.c_float)]


class Mlp3Fwd(C.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p * 4), ("b", c_void_p * 4), ("h", c_void_p * 3), ("y", c_void_p),
                ("M", c_int), ("K0", c_int), ("N1", c_int), ("N2", c_int), ("N3", c_int), ("N4", c_int),
                ("ldx", c_int), ("ldw0", c_int), ("ldh", c_int * 3), ("ldy", c_int), ("act", c_int), ("precision", c_int),
                # optional Gaussian policy head fused into the output epilogue (rollout)
                ("head_z", c_void_p), ("head_std", c_void_p), ("head_std_type", c_int), ("head_actions", c_void_p),
                ("head_logp", c_void_p), ("head_sigma", c_void_p), ("head_actions_host", c_void_p)]


class LstmFwd(C.Structure):
    _fields_ = [(k, c_void_p) for k in ("xg", "w_hh", "b_hh", "h_all", "c_all", "gates", "hg")]


class LstmBwd(C.Structure):
    _fields_ = [(k, c_void_p) for k in ("dout", "dhT", "dcT", "gates", "c_all", "w_hh_t", "dgates", "dh_scratch",
                                        "dc_scratch")]


class LossHead(C.Structure):
    _fields_ = [("w_actor", c_void_p), ("w_critic", c_void_p), ("h_actor", c_void_p), ("h_critic", c_void_p),
                ("dz_actor", c_void_p), ("dz_critic", c_void_p), ("K3", c_int), ("ldh", c_int), ("act", c_int)]


class DgradChain(C.Structure):
    _fields_ = [("dz3", c_void_p), ("w3", c_void_p), ("w2", c_void_p), ("h2", c_void_p), ("h1", c_void_p), ("dz2", c_void_p),
                ("dz1", c_void_p), ("M", c_int), ("N1", c_int), ("N2", c_int), ("N3", c_int), ("ld_dz3", c_int),
                ("ldh2", c_int), ("ldh1", c_int), ("ld_dz2", c_int), ("ld_dz1", c_int), ("act", c_int)]


class StorePlan(C.Structure):
    _fields_ = [("obs_dst", c_void_p * 16), ("obs_dim", c_int * 16), ("values_row", c_void_p), ("rewards_row", c_void_p),
                ("dones_row", c_void_p), ("gamma", c_double), ("n_groups", c_int), ("N", c_int), ("A", c_int)]


class RndNet(C.Structure):
    _fields_ = [("w", c_void_p * 4), ("b", c_void_p * 4), ("act", c_int * 4), ("dims", c_int * 5), ("num_layers", c_int)]


class OptState(C.Structure):
    _fields_ = [("lr", c_double), ("clip_scale", c_double), ("total_norm", c_double), ("step_size", c_double),
                ("bc2_sqrt", c_double), ("kl_mean", c_double), ("step", C.c_longlong), ("pad_", C.c_longlong)]


OPT_STATE_DOUBLES = C.sizeof(OptState) // 8  # the struct is viewed as 8 x 8-byte words on the device

_SIGNATURES = {
    "rslrl_abi_version": (c_int, []),
    "rslrl_last_error": (C.c_char_p, []),
    "rslrl_launch_count": (C.c_ulonglong, []),
    "rslrl_reset_launch_count": (None, []),
    "rslrl_add_launch_count": (None, [C.c_ulonglong]),
    "rslrl_gaussian_act": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int,
                                   c_void_p]),
    "rslrl_store_transition": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                       c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int, c_int, c_void_p]),
    "rslrl_gae_workspace_bytes": (c_i64, [c_int]),
    "rslrl_gae_norm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_double, c_double, c_int, c_void_p]),
    "rslrl_permute_gather": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                     c_void_p]),
    "rslrl_linear_forward": (c_int, [c_int, c_void_p, c_int, c_void_p]),
    "rslrl_linear_backward_data": (c_int, [c_int, c_void_p, c_int, c_void_p]),
    "rslrl_linear_backward_weight": (c_int, [c_int, c_void_p, c_int, c_void_p]),
    "rslrl_linear_backward_head": (c_int, [c_int, c_void_p, c_void_p]),
    "rslrl_mlp3_forward": (c_int, [c_int, c_void_p, c_void_p]),
    "rslrl_mlp3_plan_create": (c_int, [c_int, c_void_p, C.POINTER(c_void_p)]),
    "rslrl_mlp3_plan_run": (c_int, [c_void_p, c_void_p, c_void_p]),
    "rslrl_mlp3_plan_destroy": (None, [c_void_p]),
    "rslrl_loss_workspace_bytes": (c_i64, []),
    "rslrl_sum_sumsq": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p]),
    "rslrl_optim_workspace_bytes": (c_i64, [c_i64]),
    "rslrl_grad_norm_lr": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_int, c_double, c_int, c_double,
                                   c_int, c_double, c_double, c_void_p]),
    "rslrl_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_double, c_double, c_double,
                                c_int, c_void_p]),
    "rslrl_allreduce_adam_workspace_bytes": (c_i64, []),
    "rslrl_allreduce_adam_flag_bytes": (c_i64, []),
    "rslrl_allreduce_clip_adam": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_i64,
                                          c_void_p, c_void_p, c_double, c_int, c_double, c_int, c_int, c_double,
                                          c_double, c_double, c_int, c_void_p, c_void_p]),
    "rslrl_traj_workspace_bytes": (c_i64, [c_int, c_int]),
    "rslrl_trajectory_index": (c_int, [c_void_p, c_int, c_int] + [c_void_p] * 8),
    "rslrl_trajectory_masks": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rslrl_split_and_pad": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "rslrl_unpad": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p]),
    "rslrl_hidden_at_starts": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p]),
    "rslrl_lstm_cell_forward": (c_int, [c_void_p] * 7 + [c_i64, c_int, c_void_p]),
    "rslrl_lstm_cell_backward": (c_int, [c_void_p] * 8 + [c_i64, c_int, c_void_p]),
    "rslrl_lstm_layers_forward": (c_int, [c_int, c_void_p, c_int, c_i64, c_int, c_int, c_void_p]),
    "rslrl_lstm_layers_backward": (c_int, [c_int, c_void_p, c_int, c_int, c_i64, c_int, c_int, c_void_p]),
    "rslrl_empirical_norm_workspace_bytes": (c_i64, [c_int]),
    "rslrl_empirical_norm_update": (c_int, [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                            c_void_p, c_void_p]),
    "rslrl_empirical_normalize": (c_int, [c_void_p, c_i64, c_i64, c_int, c_void_p, c_void_p, c_double, c_void_p, c_i64,
                                          c_void_p]),
    "rslrl_fill_zero": (c_int, [c_void_p, c_i64, c_void_p]),
    "rslrl_dgrad_chain16": (c_int, [c_int, c_void_p, c_void_p]),
    "rslrl_store_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "rslrl_rnd_reward": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_double, c_void_p]),
    "rslrl_convert_f32_f16": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_int, c_int, c_void_p]),
}

# the ppo_loss signature above is easier to read spelled out:
_SIGNATURES["rslrl_ppo_loss_fwd_bwd"] = (c_int, [
    c_void_p, c_void_p, c_int, c_void_p,            # mu, std_param, std_type, value
    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,  # actions, old_logp, advantages, returns, old_values
    c_void_p, c_void_p, c_void_p,                   # old_mu, old_sigma, adv_stats
    c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,  # dmu, dmu_stride, dvalue, dvalue_stride, dstd, db_mu, db_value
    c_void_p, c_void_p, c_void_p,                   # kl_out, loss_sums, workspace
    c_i64, c_i64, c_int, c_double, c_double, c_double, c_int,
    c_void_p, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p])  # dmu16, stride, dvalue16, stride, grad_scale16, scale_state, head, stream

_lib = None


def library_path() -> str:
    # RSLRL_B200_LIB: load another build of the same ABI (A/B kernel experiments); never a fallback
    return os.environ.get("RSLRL_B200_LIB") or _build.LIBPATH


def lib():
    """The loaded shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"rsl_rl_b200: native library {path} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m rsl_rl_b200._build`). "
                "There is no PyTorch/CPU fallback for the PPO kernels.")
        handle = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.rslrl_abi_version() != ABI_VERSION:
            raise RuntimeError("rsl_rl_b200: ABI version mismatch between _native.py and the shared library")
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int):
    if rc != 0:
        raise RuntimeError("rsl_rl_b200 native call failed: " + lib().rslrl_last_error().decode())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (the fast C accessor when available)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def on_device(t: torch.Tensor) -> bool:
    """Whether `t` lives in memory the kernels can address (the one predicate every "is this a device tensor" test goes
    through, so that the host-logic tests can stand a host-memory emulation of the ABI in: tests/host_emulation.py)."""
    return t.is_cuda


def require_cuda_device(device, what: str):
    if torch.device(device).type != "cuda":
        raise RuntimeError(f"{what} runs on a CUDA device only (got device={device!r}); there is no CPU fallback.")


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not on_device(t):
            raise RuntimeError("rsl_rl_b200 kernels need CUDA tensors (got a tensor on %s); there is no CPU path"
                               % t.device)


def require_device_readable(*tensors: torch.Tensor):
    """CUDA tensors, or PINNED host tensors (page-locked memory is mapped into the device address space under UVA, so a
    kernel may read it in place over PCIe/NVLink-C2C: used for the few small per-step env outputs -- rewards, dones,
    time-outs -- of a host-resident env instead of one cudaMemcpyAsync each)."""
    for t in tensors:
        if t is not None and not t.is_cuda and not t.is_pinned():
            raise RuntimeError("rsl_rl_b200 kernels need CUDA (or pinned host) tensors, got a pageable tensor on %s; "
                               "there is no CPU path" % t.device)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous view of `t` (copies only when needed)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


_FLAG_DTYPES = {torch.uint8: U8, torch.bool: U8, torch.int32: I32, torch.int64: I64, torch.float32: F32}


def flag_tensor(t: torch.Tensor):
    """(tensor, dtype code) for dones / time_outs coming from the env in any integer/bool/float dtype."""
    if t.dtype not in _FLAG_DTYPES:
        t = t.to(torch.float32)
    if not t.is_contiguous():
        t = t.contiguous()
    return t, _FLAG_DTYPES[t.dtype]


def void_array(ptrs: Sequence[Optional[int]]):
    return (c_void_p * len(ptrs))(*ptrs)


def int_array(vals: Sequence[int]):
    return (c_int * len(vals))(*vals)


def launch_count() -> int:
    return int(lib().rslrl_launch_count())
