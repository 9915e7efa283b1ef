"""TEST INFRASTRUCTURE -- a host-memory stand-in for the C ABI of `include/rslrl_b200.h`, for `-m "not gpu"` tests only.

The product has ONE arithmetic back end, the sm_100a library, and refuses to run without it (tests/test_cabi_exports.py::
test_no_cpu_fallback).  The build container has no GPU, so the HOST LOGIC that sits above the ABI -- which launches are
issued in which order on which buffers: rollout store, the autograd wiring of the MLP / LSTM functions, hidden-state
carrying, flat gradient buffers, clip + Adam sequencing -- could otherwise only be checked on the GPU box.  `install()`
swaps the loaded library for `EmulatedLibrary`, a plain torch/numpy restatement of the entry points those host paths call
(each method cites the kernel it stands in for), reading and writing the callers' buffers through their raw addresses
exactly as the kernels do.  It says nothing about the kernels themselves: those are compared with the oracle and the
golden vectors by the `-m gpu` tests.  Nothing under `rsl_rl_b200/` imports this module.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from rsl_rl_b200 import _native as nat


def _view(ptr, count, dtype=np.float32):
    """numpy view of `count` elements at raw address `ptr`."""
    if not ptr or count == 0:
        return None
    nbytes = int(count) * np.dtype(dtype).itemsize
    return np.frombuffer((ctypes.c_char * nbytes).from_address(int(ptr)), dtype=dtype)


def _mat(ptr, rows, cols, ld):
    """[rows, cols] fp32 matrix with row stride `ld` at `ptr` (a writable strided numpy view)."""
    flat = _view(ptr, (rows - 1) * ld + cols)
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(4 * ld, 4))


class _HalfStore:
    """fp16 buffers of the 16-bit storage path, with an EXACT fp32 side copy of everything the emulation writes into them.
    The memory itself holds the rounded fp16 values (what the kernels store); reads prefer the side copy wherever it still
    agrees with memory after rounding (anything written by other means -- a torch op, a recycled allocation -- falls back to
    memory).  Storage rounding is thereby taken out of the host-logic tests: the fp16 chain's pointer wiring and power-of-two
    gradient scaling must reproduce the fp32 reference to fp32 accuracy, which a wiring mistake cannot do."""

    def __init__(self):
        self.regions = []      # [start address, element count, float32 array]

    def _find(self, ptr, span):
        for r in self.regions:
            if r[0] <= ptr and ptr + 2 * span <= r[0] + 2 * r[1]:
                return r, (ptr - r[0]) // 2
        return None

    @staticmethod
    def _strided(flat, off, rows, cols, ld):
        return np.lib.stride_tricks.as_strided(flat[off:], shape=(rows, cols), strides=(flat.itemsize * ld, flat.itemsize))

    def write(self, ptr, rows, cols, ld, values):
        ptr, span = int(ptr), (rows - 1) * ld + cols
        mem = _view(ptr, span, np.float16)
        self._strided(mem, 0, rows, cols, ld)[:] = np.clip(values, -65504.0, 65504.0).astype(np.float16)
        hit = self._find(ptr, span)
        if hit is None:
            end = ptr + 2 * span
            self.regions = [r for r in self.regions if r[0] + 2 * r[1] <= ptr or end <= r[0]]   # drop what it overlaps
            hit = ([ptr, span, mem.astype(np.float32)], 0)
            self.regions.append(hit[0])
        self._strided(hit[0][2], hit[1], rows, cols, ld)[:] = values

    def read(self, ptr, rows, cols, ld):
        ptr, span = int(ptr), (rows - 1) * ld + cols
        mem = self._strided(_view(ptr, span, np.float16), 0, rows, cols, ld)
        hit = self._find(ptr, span)
        if hit is None:
            return mem.astype(np.float32)
        exact = self._strided(hit[0][2], hit[1], rows, cols, ld)
        agrees = np.clip(exact, -65504.0, 65504.0).astype(np.float16) == mem
        return np.where(agrees, exact, mem.astype(np.float32)).astype(np.float32)


def _act(code, z):
    if code == 0:
        return z
    t = torch.from_numpy(np.ascontiguousarray(z))
    fn = {1: torch.nn.functional.elu, 2: torch.relu, 3: torch.tanh, 4: torch.sigmoid,
          5: torch.nn.functional.leaky_relu, 6: torch.selu, 7: torch.nn.functional.softplus}[code]
    return fn(t).numpy()


def _act_grad_from_output(code, h):
    """Derivative of the activation as a function of its OUTPUT h (how the data-gradient kernels evaluate it)."""
    if code == 0:
        return np.ones_like(h)
    if code == 1:
        return np.where(h > 0, 1.0, h + 1.0).astype(np.float32)
    if code == 2:
        return (h > 0).astype(np.float32)
    if code == 3:
        return (1.0 - h * h).astype(np.float32)
    if code == 4:
        return (h * (1.0 - h)).astype(np.float32)
    if code == 5:
        return np.where(h > 0, 1.0, 0.01).astype(np.float32)
    if code == 6:
        alpha, scale = 1.6732632423543772, 1.0507009873554805
        return np.where(h > 0, scale, h + scale * alpha).astype(np.float32)
    if code == 7:
        return (1.0 - np.exp(-h)).astype(np.float32)
    raise ValueError(code)


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)


class EmulatedLibrary:
    """The subset of the ABI the host paths under test call.  Unknown entry points raise AttributeError."""

    def __init__(self):
        self.launches = 0
        self.calls = []
        self.half = _HalfStore()

    def _launch(self, name):
        self.launches += 1
        self.calls.append(name)
        return 0

    # ---- plumbing
    def rslrl_abi_version(self):
        return nat.ABI_VERSION

    def rslrl_last_error(self):
        return b"emulated"

    def rslrl_launch_count(self):
        return self.launches

    def rslrl_reset_launch_count(self):
        self.launches = 0

    def rslrl_add_launch_count(self, n):
        self.launches = (self.launches + n) & 0xFFFFFFFFFFFFFFFF

    def rslrl_gae_workspace_bytes(self, N):
        return 256

    def rslrl_optim_workspace_bytes(self, numel):
        return 256

    def rslrl_allreduce_adam_workspace_bytes(self):
        return 256

    def rslrl_allreduce_adam_flag_bytes(self):
        return 256

    def rslrl_empirical_norm_workspace_bytes(self, D):
        return 64

    def rslrl_fill_zero(self, ptr, nbytes, stream):
        _view(ptr, nbytes, np.uint8)[:] = 0
        return self._launch("fill_zero")

    # ---- csrc/storage.cu: store_transition_kernel
    def rslrl_store_transition(self, n_groups, obs_src, obs_dst, obs_dim, actions, mu, sigma, sigma_stride, values,
                               log_prob, rewards, dones, dones_dtype, time_outs, time_outs_dtype, intrinsic, gamma,
                               actions_row, mu_row, sigma_row, values_row, log_prob_row, rewards_row, dones_row, N, A,
                               stream):
        def copy(src, dst, count):
            if src and dst:
                _view(dst, count)[:] = _view(src, count)

        for g in range(n_groups):
            copy(obs_src[g], obs_dst[g], N * obs_dim[g])
        copy(actions, actions_row, N * A)
        copy(mu, mu_row, N * A)
        copy(values, values_row, N)
        copy(log_prob, log_prob_row, N)
        if sigma and sigma_row:
            if sigma_stride == A:
                copy(sigma, sigma_row, N * A)
            else:
                _view(sigma_row, N * A).reshape(N, A)[:] = _view(sigma, A)
        flag_np = {nat.U8: np.uint8, nat.I32: np.int32, nat.I64: np.int64, nat.F32: np.float32}
        r = _view(rewards, N).copy()
        if intrinsic:
            r = r + _view(intrinsic, N)
        if time_outs:
            to = _view(time_outs, N, flag_np[time_outs_dtype]).astype(np.float32)
            r = r + np.float32(gamma) * (_view(values, N) * to)
        _view(rewards_row, N)[:] = r
        _view(dones_row, N, np.uint8)[:] = _view(dones, N, flag_np[dones_dtype]).astype(np.float32).astype(np.int32).astype(np.uint8)
        return self._launch("store_transition")

    # ---- csrc/storage.cu: the rollout fast path (row `step` fixed in a plan, only observations / rewards / dones move)
    def rslrl_store_step(self, plan_ref, obs_src, rewards, dones, dones_dtype, time_outs, time_outs_dtype, intrinsic,
                         stream):
        sp = plan_ref._obj                       # ctypes.byref(rslrl_store_plan_t)
        n, obs_dst, obs_dim = sp.n_groups, [sp.obs_dst[i] for i in range(sp.n_groups)], [sp.obs_dim[i] for i in range(sp.n_groups)]
        rc = self.rslrl_store_transition(n, obs_src, obs_dst, obs_dim, None, None, None, 0, sp.values_row, None, rewards,
                                         dones, dones_dtype, time_outs, time_outs_dtype, intrinsic, sp.gamma, None, None,
                                         None, None, None, sp.rewards_row, sp.dones_row, sp.N, sp.A, stream)
        self.calls[-1] = "store_step"
        return rc

    # ---- csrc/core.cu: Gaussian head of the rollout
    def rslrl_gaussian_act(self, mu, std_param, std_type, z, actions, logp, sigma_out, n, A, stream):
        f = np.float32
        m, zz = _view(mu, n * A).reshape(n, A), _view(z, n * A).reshape(n, A)
        sigma = _view(std_param, A) if std_type == nat.STD_SCALAR else np.exp(_view(std_param, A))
        a = (zz * sigma).astype(f) + m                     # two roundings, as normal_().mul_().add_()
        _view(actions, n * A).reshape(n, A)[:] = a
        lp = -((a - m) ** 2) / (f(2) * sigma * sigma) - np.log(sigma) - f(math.log(math.sqrt(2 * math.pi)))
        _view(logp, n)[:] = lp.sum(1)
        if sigma_out:
            _view(sigma_out, n * A).reshape(n, A)[:] = sigma
        return self._launch("gaussian_act")

    # ---- csrc/ppo_loss.cu (arithmetic: the numpy oracle's forward + hand-derived backward)
    def rslrl_loss_workspace_bytes(self):
        return 1024

    def rslrl_sum_sumsq(self, x, n, out, workspace, stream):
        v = _view(x, n).astype(np.float64)
        _view(out, 2, np.float64)[:] = (v.sum(), (v * v).sum())
        return self._launch("sum_sumsq")

    def rslrl_ppo_loss_fwd_bwd(self, mu, std_param, std_type, value, actions, old_logp, adv, returns, old_values, old_mu,
                               old_sigma, adv_stats, dmu, dmu_stride, dvalue, dvalue_stride, dstd, db_mu, db_value,
                               kl_out, loss_sums, workspace, rows, mb_orig, A, clip_param, value_loss_coef, entropy_coef,
                               use_clipped, dmu16, dmu16_stride, dvalue16, dvalue16_stride, grad_scale16, scale_state,
                               head, stream):
        from oracle import ppo_oracle as O
        assert not head, "the (opt-in) output-layer data gradient inside the loss kernel is not emulated"
        f = np.float32
        m = _view(mu, rows * A).reshape(rows, A)
        std = _view(std_param, A)
        sigma = np.broadcast_to(std if std_type == nat.STD_SCALAR else np.exp(std), (rows, A)).astype(f)
        a = _view(adv, mb_orig).copy()
        if adv_stats:                                       # per-minibatch normalisation (ppo.py:221-223), unbiased std
            s, ss = _view(adv_stats, 2, np.float64)
            mean = s / mb_orig
            a = ((a - f(mean)) / (f(math.sqrt(max(ss - mb_orig * mean * mean, 0.0) / (mb_orig - 1))) + f(1e-8))).astype(f)
        # symmetry augmentation: rows = k * mb_orig, the per-sample scalars repeat (row i reads element i % mb_orig)
        rep = lambda x: np.tile(x, rows // mb_orig)
        out = O.ppo_loss_fwd_bwd(m, sigma, _view(value, rows), _view(actions, rows * A).reshape(rows, A),
                                 rep(_view(old_logp, mb_orig)), rep(a), rep(_view(returns, mb_orig)),
                                 rep(_view(old_values, mb_orig)), clip_param, value_loss_coef, entropy_coef,
                                 bool(use_clipped), n_entropy=mb_orig)
        _mat(dmu, rows, A, dmu_stride)[:] = out["dmu"]
        _mat(dvalue, rows, 1, dvalue_stride)[:] = out["dvalue"]
        dsig = out["dsigma"].sum(0) if std_type == nat.STD_SCALAR else (out["dsigma"] * sigma).sum(0)
        if dstd:
            _view(dstd, A)[:] += dsig.astype(f)
        if db_mu:
            _view(db_mu, A)[:] += out["dmu"].sum(0)
        if db_value:
            _view(db_value, 1)[:] += out["dvalue"].sum()
        if kl_out:                                          # KL over the un-augmented block (ppo.py:255-269)
            _view(kl_out, 1)[0] = O.kl_mean(m[:mb_orig], sigma[:mb_orig], _view(old_mu, mb_orig * A).reshape(mb_orig, A),
                                            _view(old_sigma, mb_orig * A).reshape(mb_orig, A))
        if loss_sums:
            _view(loss_sums, 3, np.float64)[:] += (out["surrogate"], out["value_loss"], out["entropy"])
        if dmu16:
            # S * dL/dmu, S * dL/dV as fp16: S = scale_state[2] (dynamic, a power of two derived from the previous launch's
            # largest gradient) or the static argument; [0] / [1] publish S and 1 / S for the weight-gradient launches
            st = _view(scale_state, 8) if scale_state else None
            S = np.float32(st[2]) if st is not None else np.float32(grad_scale16)
            self.half.write(dmu16, rows, A, dmu16_stride, S * out["dmu"])
            self.half.write(dvalue16, rows, 1, dvalue16_stride, S * out["dvalue"].reshape(rows, 1))
            if st is not None:
                mx = max(float(np.abs(out["dmu"]).max()), float(np.abs(out["dvalue"]).max()))
                st[0], st[1] = S, np.float32(1.0) / S
                if mx * S > 60000.0:
                    _view(int(scale_state) + 16, 1, np.uint32)[0] += 1
                if mx > 0 and math.isfinite(mx):
                    st[2] = min(max(2.0 ** math.floor(math.log2(4096.0 / mx)), 2.0 ** -20), 2.0 ** 30)
        return self._launch("ppo_loss_fwd_bwd")

    # ---- csrc/rnd.cu: intrinsic reward in one launch
    def rslrl_rnd_reward(self, target, predictor, x, ldx, n, out, scale, stream):
        def embed(ref):
            q = ref._obj
            h = _mat(x, n, q.dims[0], ldx).astype(np.float32)
            for l in range(q.num_layers):
                z = h @ _mat(q.w[l], q.dims[l + 1], q.dims[l], q.dims[l]).T
                if q.b[l]:
                    z = z + _view(q.b[l], q.dims[l + 1])
                h = _act(q.act[l], z.astype(np.float32))
            return h
        gap = embed(target) - embed(predictor)
        _view(out, n)[:] = np.float32(scale) * np.sqrt((gap * gap).sum(1))
        return self._launch("rnd_reward")

    # ---- csrc/gae.cu (arithmetic: the numpy oracle, which is pinned bit-exactly to the reference's golden vectors)
    def rslrl_gae_norm(self, rewards, values, dones, last_values, returns, advantages, workspace, T, N, gamma, lam,
                       normalize, stream):
        from oracle import ppo_oracle as O
        r, v = _view(rewards, T * N).reshape(T, N), _view(values, T * N).reshape(T, N)
        d = _view(dones, T * N, np.uint8).reshape(T, N)
        ret, adv = O.compute_returns(r.copy(), v.copy(), d.copy(), _view(last_values, N).copy(), gamma, lam,
                                     normalize_advantage=bool(normalize))
        _view(returns, T * N)[:] = np.asarray(ret, np.float32).reshape(-1)
        _view(advantages, T * N)[:] = np.asarray(adv, np.float32).reshape(-1)
        return self._launch("gae_norm")

    # ---- csrc/storage.cu: permute_gather_kernel
    def rslrl_permute_gather(self, n_fields, src, dst, row_floats, dst_stride, dst_col, idx, n_rows, stream):
        rows = _view(idx, n_rows, np.int64)
        for f in range(n_fields):
            w, ld, col = row_floats[f], dst_stride[f], dst_col[f]
            total = int(rows.max()) + 1 if n_rows else 0
            s = _view(src[f], total * w).reshape(total, w)
            out = _view(dst[f], (n_rows - 1) * ld + col + w)
            np.lib.stride_tricks.as_strided(out[col:], shape=(n_rows, w), strides=(4 * ld, 4))[:] = s[rows]
        return self._launch("permute_gather")

    # ---- csrc/linear.cu -> gemm_simt.cu / gemm_tc.cu
    def rslrl_linear_forward(self, n, probs, precision, stream):
        for i in range(n):
            p = probs[i]
            x, w = _mat(p.x, p.M, p.K, p.ldx), _mat(p.w, p.N, p.K, p.K)
            z = x.astype(np.float32) @ w.T
            if p.b:
                z = z + _view(p.b, p.N)
            _mat(p.y, p.M, p.N, p.ldy)[:] = _act(p.act, z.astype(np.float32))
        return self._launch("linear_forward")

    def rslrl_linear_backward_data(self, n, probs, precision, stream):
        f16 = precision == nat.PREC_FP16     # 16-bit storage: dy, w (shadow), h and dx are fp16 tensors
        rd = (lambda ptr, r, c, ld: self.half.read(ptr, r, c, ld)) if f16 else _mat
        for i in range(n):
            p = probs[i]
            dy, w = rd(p.dy, p.M, p.N, p.ldy), rd(p.w, p.N, p.K, p.K)
            dx = dy @ w
            if p.act != 0:
                dx = dx * _act_grad_from_output(p.act, rd(p.h, p.M, p.K, p.ldh))
            dx = dx.astype(np.float32)
            if f16:
                self.half.write(p.dx, p.M, p.K, p.ldx, dx)
            else:
                _mat(p.dx, p.M, p.K, p.ldx)[:] = dx
            if p.db_prev:
                _view(p.db_prev, p.K)[:] += dx.sum(0)
        return self._launch("linear_backward_data")

    # ---- csrc/dgrad_fused.cu: data gradients of the two wide hidden layers of both nets in one launch (fp16 chain)
    def rslrl_dgrad_chain16(self, n, nets, stream):
        for i in range(n):
            q = nets[i]
            if not (q.N3 in (64, 128) and q.N2 in (128, 256) and q.N1 in (128, 256, 512)):
                return 4
            rd = self.half.read
            dz3 = rd(q.dz3, q.M, q.N3, q.ld_dz3)
            dz2 = (dz3 @ rd(q.w3, q.N3, q.N2, q.N2)) * _act_grad_from_output(q.act, rd(q.h2, q.M, q.N2, q.ldh2))
            self.half.write(q.dz2, q.M, q.N2, q.ld_dz2, dz2.astype(np.float32))
            dz1 = (dz2 @ rd(q.w2, q.N2, q.N1, q.N1)) * _act_grad_from_output(q.act, rd(q.h1, q.M, q.N1, q.ldh1))
            self.half.write(q.dz1, q.M, q.N1, q.ld_dz1, dz1.astype(np.float32))
        return self._launch("dgrad_chain16")

    # ---- csrc/core.cu: fp16 shadows (flat weights, permuted minibatch observations)
    def rslrl_convert_f32_f16(self, src, ld_src, dst, ld_dst, rows, cols, cols_pad, stream):
        vals = np.zeros((rows, cols_pad), np.float32)
        vals[:, :cols] = _mat(src, rows, cols, ld_src)
        self.half.write(dst, rows, cols_pad, ld_dst, vals)
        return self._launch("convert_f32_f16")

    def rslrl_linear_backward_head(self, n, probs, stream):
        """Narrow output layer: data gradient (x activation derivative of the previous layer), weight and bias gradient."""
        for i in range(n):
            p = probs[i]
            assert not p.f16
            dy, w, h = _mat(p.dy, p.M, p.N, p.ldy), _mat(p.w, p.N, p.K, p.K), _mat(p.h, p.M, p.K, p.ldh)
            if p.dx:
                _mat(p.dx, p.M, p.K, p.ldx)[:] = ((dy @ w) * _act_grad_from_output(p.act, h)).astype(np.float32)
            _mat(p.dw, p.N, p.K, p.K)[:] += dy.T @ h
            if p.db:
                _view(p.db, p.N)[:] += dy.sum(0)
        return self._launch("linear_backward_head")

    def rslrl_linear_backward_weight(self, n, probs, precision, stream):
        rd = (lambda ptr, r, c, ld: self.half.read(ptr, r, c, ld)) if precision == nat.PREC_FP16 else _mat
        for i in range(n):
            p = probs[i]
            dy, x = rd(p.dy, p.M, p.N, p.ldy), rd(p.x, p.M, p.K, p.ldx)
            scale = np.float32(p.scale if p.scale != 0 else 1.0)
            if p.scale_dev:
                scale = scale * _view(p.scale_dev, 1)[0]
            _mat(p.dw, p.N, p.K, p.K)[:] += scale * (dy.T @ x)
            if p.db:
                _view(p.db, p.N)[:] += scale * dy.sum(0)
        return self._launch("linear_backward_weight")

    # ---- csrc/lstm.cu: whole-layer drivers (time loop in C)
    def rslrl_lstm_layers_forward(self, n_nets, nets, T, n, H, precision, stream):
        for i in range(n_nets):
            q = nets[i]
            xg = _view(q.xg, T * n * 4 * H).reshape(T, n, 4 * H)
            w_hh = _view(q.w_hh, 4 * H * H).reshape(4 * H, H)
            b_hh = _view(q.b_hh, 4 * H) if q.b_hh else 0.0
            h_all = _view(q.h_all, (T + 1) * n * H).reshape(T + 1, n, H)
            c_all = _view(q.c_all, (T + 1) * n * H).reshape(T + 1, n, H)
            gates = _view(q.gates, T * n * 4 * H).reshape(T, n, 4 * H) if q.gates else None
            for t in range(T):
                pre = xg[t] + h_all[t] @ w_hh.T + b_hh
                gi, gf = _sigmoid(pre[:, :H]), _sigmoid(pre[:, H:2 * H])
                gg, go = np.tanh(pre[:, 2 * H:3 * H]), _sigmoid(pre[:, 3 * H:])
                c_all[t + 1] = gf * c_all[t] + gi * gg
                h_all[t + 1] = go * np.tanh(c_all[t + 1])
                if gates is not None:
                    gates[t] = np.concatenate([gi, gf, gg, go], axis=1)
            self._launch("lstm_step_forward")
        return 0

    def rslrl_lstm_layers_backward(self, n_nets, nets, need_dh0, T, n, H, precision, stream):
        for i in range(n_nets):
            q = nets[i]
            gates = _view(q.gates, T * n * 4 * H).reshape(T, n, 4 * H)
            c_all = _view(q.c_all, (T + 1) * n * H).reshape(T + 1, n, H)
            w_hh_t = _view(q.w_hh_t, H * 4 * H).reshape(H, 4 * H)
            dgates = _view(q.dgates, T * n * 4 * H).reshape(T, n, 4 * H)
            dh_s = _view(q.dh_scratch, 2 * n * H).reshape(2, n, H)
            dc_s = _view(q.dc_scratch, 2 * n * H).reshape(2, n, H)
            dout = _view(q.dout, T * n * H).reshape(T, n, H) if q.dout else None
            dh_rec = _view(q.dhT, n * H).reshape(n, H).copy() if q.dhT else None
            dc = _view(q.dcT, n * H).reshape(n, H).copy() if q.dcT else None
            for t in range(T - 1, -1, -1):
                gi, gf, gg, go = (gates[t][:, k * H:(k + 1) * H] for k in range(4))
                dh = np.zeros((n, H), np.float32)
                if dout is not None:
                    dh = dh + dout[t]
                if dh_rec is not None:
                    dh = dh + dh_rec
                tc = np.tanh(c_all[t + 1])
                d_o = dh * tc
                dcur = dh * go * (1.0 - tc * tc) + (dc if dc is not None else 0.0)
                dgates[t] = np.concatenate([dcur * gg * gi * (1 - gi), dcur * c_all[t] * gf * (1 - gf),
                                            dcur * gi * (1 - gg * gg), d_o * go * (1 - go)], axis=1)
                dc_s[t & 1] = dcur * gf
                dc = dc_s[t & 1].copy()
                if t > 0 or need_dh0:
                    dh_s[t & 1] = dgates[t] @ w_hh_t.T
                    dh_rec = dh_s[t & 1].copy()
            self._launch("lstm_step_backward")
        return 0

    # ---- csrc/trajectory.cu: segmented scan over dones, masks, split/pad, hidden states at trajectory starts
    def rslrl_traj_workspace_bytes(self, T, N):
        return 64

    def rslrl_trajectory_index(self, dones, T, N, traj_id, traj_pos, traj_len, traj_env, traj_start, env_traj_offset,
                               workspace, stream):
        d = _view(dones, T * N, np.uint8).reshape(T, N)
        i32 = lambda ptr, n: _view(ptr, n, np.int32)
        tid, tpos = i32(traj_id, T * N).reshape(T, N), i32(traj_pos, T * N).reshape(T, N)
        tlen, tenv, tstart, off = i32(traj_len, T * N), i32(traj_env, T * N), i32(traj_start, T * N), i32(env_traj_offset, N + 1)
        j = 0
        for env in range(N):
            off[env] = j
            start = 0
            for t in range(T):
                if d[t, env] or t == T - 1:
                    tlen[j], tenv[j], tstart[j] = t - start + 1, env, start
                    tid[start:t + 1, env] = j
                    tpos[start:t + 1, env] = np.arange(t - start + 1)
                    start, j = t + 1, j + 1
        off[N] = j
        return self._launch("trajectory_index")

    def rslrl_trajectory_masks(self, traj_len, masks, T, first_traj, n_traj_end, capacity, stream):
        tlen = _view(traj_len, n_traj_end, np.int32)
        m = _view(masks, T * capacity, np.uint8).reshape(T, capacity)
        for j in range(capacity):
            J = first_traj + j
            m[:, j] = (np.arange(T) < tlen[J]) if J < n_traj_end else 0
        return self._launch("trajectory_masks")

    def rslrl_split_and_pad(self, x, traj_len, traj_env, traj_start, padded, T, N, D, first_traj, n_traj_end, capacity,
                            stream):
        tlen, tenv, tstart = (_view(q, n_traj_end, np.int32) for q in (traj_len, traj_env, traj_start))
        out = _view(padded, T * capacity * D).reshape(T, capacity, D)
        out[:] = 0
        for j in range(capacity):
            J = first_traj + j
            if J < n_traj_end:
                for q in range(tlen[J]):   # (x may be addressed below its base by pad_columns: element-wise addresses)
                    out[q, j] = _view(int(x) + 4 * ((int(tstart[J]) + q) * N + int(tenv[J])) * D, D)
        return self._launch("split_and_pad")

    def rslrl_unpad(self, padded, traj_id, traj_pos, out, T, N, D, env0, env1, first_traj, capacity, stream):
        pd = _view(padded, T * capacity * D).reshape(T, capacity, D)
        tid, tpos = (_view(q, T * N, np.int32).reshape(T, N) for q in (traj_id, traj_pos))
        o = _view(out, T * (env1 - env0) * D).reshape(T, env1 - env0, D)
        for t in range(T):
            for env in range(env0, env1):
                o[t, env - env0] = pd[tpos[t, env], tid[t, env] - first_traj]
        return self._launch("unpad")

    def rslrl_hidden_at_starts(self, saved, traj_env, traj_start, out, T, L, N, H, first_traj, n_traj_end, capacity,
                               stream):
        sv = _view(saved, T * L * N * H).reshape(T, L, N, H)
        tenv, tstart = (_view(q, n_traj_end, np.int32) for q in (traj_env, traj_start))
        o = _view(out, L * capacity * H).reshape(L, capacity, H)
        for j in range(capacity):
            J = first_traj + j
            if J < n_traj_end:
                o[:, j] = sv[tstart[J], :, tenv[J]]
        return self._launch("hidden_at_starts")

    # ---- csrc/normalization.cu
    def rslrl_empirical_norm_update(self, x, ldx, N, D, mean, var, std, count, until, workspace, stream):
        cnt = _view(count, 1, np.int64)
        if until >= 0 and cnt[0] >= until:
            return self._launch("empirical_norm_update")
        xt = torch.from_numpy(np.ascontiguousarray(_mat(x, N, D, ldx)))
        m, v, s = (torch.from_numpy(_view(b, D)) for b in (mean, var, std))
        cnt[0] += N
        rate = N / int(cnt[0])
        var_x = torch.var(xt, dim=0, unbiased=False)
        mean_x = torch.mean(xt, dim=0)
        delta = mean_x - m
        m += rate * delta
        v += rate * (var_x - v + delta * (mean_x - m))
        s.copy_(torch.sqrt(v))
        return self._launch("empirical_norm_update")

    def rslrl_empirical_normalize(self, x, ldx, N, D, mean, std, eps, y, ldy, stream):
        _mat(y, N, D, ldy)[:] = (_mat(x, N, D, ldx) - _view(mean, D)) / (_view(std, D) + np.float32(eps))
        return self._launch("empirical_normalize")

    # ---- csrc/allreduce_adam.cu at world size 1: gradient norm, clip, (KL-adaptive LR,) Adam, re-zeroed gradients
    def rslrl_allreduce_clip_adam(self, world, rank, peers, peer_flags, peer_results, peer_stage, values, exp_avg,
                                  exp_avg_sq, numel, state, workspace, max_grad_norm, adaptive, desired_kl,
                                  round_lr_fp32, use_kl, beta1, beta2, eps, zero_grad, values16, stream):
        assert world == 1, "the emulation covers the single-rank step"
        g_ext = _view(peers[0], numel + 8)
        g = g_ext[:numel]
        st = ctypes.cast(int(state), ctypes.POINTER(nat.OptState)).contents
        norm = math.sqrt(float(np.sum(g.astype(np.float64) ** 2)))
        coef = min(max_grad_norm / (norm + 1e-6), 1.0) if max_grad_norm > 0 else 1.0
        lr = st.lr
        if adaptive:
            kl = np.float32(g_ext[numel]) if use_kl else np.float32(0)
            if kl > np.float32(desired_kl * 2.0):
                lr = max(1e-5, lr / 1.5)
            elif np.float32(0) < kl < np.float32(desired_kl / 2.0):
                lr = min(1e-2, lr * 1.5)
            if round_lr_fp32:
                lr = float(np.float32(lr))
            st.kl_mean = float(kl)
        t = st.step + 1
        bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
        p, m, v = _view(values, numel), _view(exp_avg, numel), _view(exp_avg_sq, numel)
        f = np.float32
        gc = g * f(coef)
        m[:] = m + (gc - m) * f(1.0 - beta1)
        v[:] = v * f(beta2) + f(1.0 - beta2) * gc * gc
        p[:] = p + f(-(lr / bc1)) * (m / (np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)))
        st.lr, st.clip_scale, st.total_norm, st.step_size, st.bc2_sqrt, st.step = lr, coef, norm, lr / bc1, math.sqrt(bc2), t
        if zero_grad:
            g_ext[:] = 0
        if values16:
            self.half.write(values16, 1, numel, numel, p.reshape(1, -1))
        return self._launch("allreduce_clip_adam")


def _fused_forward(cls):
    """csrc/mlp_fused.cu: the whole 3-hidden-layer MLP of one or two nets in one launch (tf32 mode), optionally with the
    Gaussian policy head in the output epilogue, and its rollout launch plans."""

    def _run_nets(self, n, nets, xs=None):
        for i in range(n):
            q = nets[i]
            x = int(xs[i]) if xs is not None else q.x
            dims = [q.K0, q.N1, q.N2, q.N3, q.N4]
            ok = (q.N1 % 128 == 0 and 128 <= q.N1 <= 512 and q.N2 in (128, 256) and q.N3 % 32 == 0 and 32 <= q.N3 <= 128
                  and 1 <= q.N4 <= 16 and x % 16 == 0 and q.ldx % 4 == 0)
            if not ok:
                return 4
            f16 = q.precision == nat.PREC_FP16     # layers 2-4 read the fp16 weight shadow, activations are stored as fp16
            h = _mat(x, q.M, q.K0, q.ldx).astype(np.float32)
            for l in range(4):
                if l == 0 or not f16:
                    w = _mat(q.w[l], dims[l + 1], dims[l], q.ldw0 if l == 0 else dims[l])
                else:
                    w = self.half.read(q.w[l], dims[l + 1], dims[l], dims[l])
                z = h @ w.T
                if q.b[l]:
                    z = z + _view(q.b[l], dims[l + 1])
                h = _act(q.act, z.astype(np.float32)) if l < 3 else z.astype(np.float32)
                if l < 3 and q.h[l]:
                    if f16:
                        self.half.write(q.h[l], q.M, dims[l + 1], q.ldh[l], h)
                    else:
                        _mat(q.h[l], q.M, dims[l + 1], q.ldh[l])[:] = h
            _mat(q.y, q.M, q.N4, q.ldy)[:] = h
            if q.head_z:
                A = q.N4
                assert q.ldy == A, "the head reads mu as a dense [M, A] matrix"
                self.rslrl_gaussian_act(q.y, q.head_std, q.head_std_type, q.head_z, q.head_actions, q.head_logp,
                                        q.head_sigma, q.M, A, 0)
                self.calls.pop()            # (same launch: the head is the forward's epilogue)
                self.launches -= 1
                if q.head_actions_host:
                    _view(q.head_actions_host, q.M * A)[:] = _view(q.head_actions, q.M * A)
        return 0

    def rslrl_mlp3_forward(self, n, nets, stream):
        rc = _run_nets(self, n, nets)
        return rc if rc else self._launch("mlp3_forward")

    def rslrl_mlp3_plan_create(self, n, nets, plan_out):
        copy = (nat.Mlp3Fwd * n)()
        for i in range(n):
            ctypes.memmove(ctypes.byref(copy[i]), ctypes.byref(nets[i]), ctypes.sizeof(nat.Mlp3Fwd))   # captured by value
        self._plans = getattr(self, "_plans", {})
        handle = len(self._plans) + 1
        self._plans[handle] = (n, copy)
        plan_out._obj.value = handle
        return 0

    def rslrl_mlp3_plan_run(self, plan, xs, stream):
        n, nets = self._plans[plan.value if hasattr(plan, "value") else int(plan)]
        rc = _run_nets(self, n, nets, xs)
        return rc if rc else self._launch("mlp3_plan_run")

    def rslrl_mlp3_plan_destroy(self, plan):
        self._plans.pop(plan.value if hasattr(plan, "value") else int(plan), None)

    cls.rslrl_mlp3_forward, cls.rslrl_mlp3_plan_create = rslrl_mlp3_forward, rslrl_mlp3_plan_create
    cls.rslrl_mlp3_plan_run, cls.rslrl_mlp3_plan_destroy = rslrl_mlp3_plan_run, rslrl_mlp3_plan_destroy
    return cls


def _two_launch_optimiser(cls):
    """csrc/optim.cu: the NCCL-path optimiser step as two launches (norm / clip scale / KL rule, then Adam)."""

    def rslrl_grad_norm_lr(self, g, P, kl_sum, state, workspace, world_size, max_grad_norm, adaptive, desired_kl,
                           round_lr_fp32, beta1, beta2, stream):
        st = ctypes.cast(int(state), ctypes.POINTER(nat.OptState)).contents
        w = float(world_size)
        norm = math.sqrt(float(np.sum(_view(g, P).astype(np.float64) ** 2))) / w      # || mean over ranks of g ||
        coef = min(max_grad_norm / (norm + 1e-6), 1.0) if max_grad_norm > 0 else 1.0
        lr = st.lr
        if adaptive:
            kl = np.float32(_view(kl_sum, 1)[0] / np.float32(w)) if kl_sum else np.float32(0)
            if kl > np.float32(desired_kl * 2.0):
                lr = max(1e-5, lr / 1.5)
            elif np.float32(0) < kl < np.float32(desired_kl / 2.0):
                lr = min(1e-2, lr * 1.5)
            if round_lr_fp32:
                lr = float(np.float32(lr))
            st.kl_mean = float(kl)
        t = st.step + 1
        bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
        st.lr, st.clip_scale, st.total_norm, st.step_size, st.bc2_sqrt, st.step = lr, coef / w, norm, lr / bc1, math.sqrt(bc2), t
        return self._launch("grad_norm_lr")

    def rslrl_adam_step(self, p, g, m, v, P, state, beta1, beta2, eps, zero_grad, stream):
        st = ctypes.cast(int(state), ctypes.POINTER(nat.OptState)).contents
        f = np.float32
        pp, gg, mm, vv = _view(p, P), _view(g, P), _view(m, P), _view(v, P)
        gc = gg * f(st.clip_scale)
        mm[:] = mm + (gc - mm) * f(1.0 - beta1)
        vv[:] = vv * f(beta2) + f(1.0 - beta2) * gc * gc
        pp[:] = pp + f(-st.step_size) * (mm / (np.sqrt(vv) / f(st.bc2_sqrt) + f(eps)))
        if zero_grad:
            gg[:] = 0
        return self._launch("adam_step")

    cls.rslrl_grad_norm_lr, cls.rslrl_adam_step = rslrl_grad_norm_lr, rslrl_adam_step
    return cls


_two_launch_optimiser(EmulatedLibrary)
_fused_forward(EmulatedLibrary)


def install(monkeypatch) -> EmulatedLibrary:
    """Routes `rsl_rl_b200._native` to the emulation for the duration of one test (pytest's monkeypatch undoes it)."""
    lib = EmulatedLibrary()
    monkeypatch.setattr(nat, "_lib", lib)
    monkeypatch.setattr(nat, "on_device", lambda t: True)
    monkeypatch.setattr(nat, "require_cuda_device", lambda device, what: None)
    monkeypatch.setattr(nat, "require_device_readable", lambda *tensors: None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)   # (no CUDA runtime here)
    from rsl_rl_b200.storage import rollout_storage
    monkeypatch.setattr(rollout_storage, "_store_step_fn", None)   # (a cached entry point of whichever library came first)
    monkeypatch.setattr(nat, "stream_ptr", lambda: 0)
    return lib
