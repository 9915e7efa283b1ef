"""GPU: the fp16-storage tensor-core path (gemm_precision="fp16", RSLRL_PREC_FP16).

Saved activations, layer gradients, minibatch observations and a weight shadow are fp16 in HBM and the MMAs are
tcgen05 kind::f16 with fp32 accumulation.  fp16 carries the same 10-bit mantissa tf32 gives those operands, so the
tolerance class is the north star's tensor-core class: 1e-3 (max-abs error / max-abs value) against fp64
restatements; the outputs that are themselves stored as fp16 carry one extra rounding (2^-11 relative)."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from tests.golden_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3


def _nat():
    from rsl_rl_b200 import _native as nat
    return nat


def _elu_grad_from_output(h):
    return np.where(h > 0, 1.0, h + 1.0)


@pytest.mark.parametrize("M,K,N", [(24576, 256, 128), (24576, 512, 256), (1000, 128, 64), (77, 64, 128), (4096, 512, 256)])
def test_fp16_backward_data(M, K, N):
    """dx = (dy W) * ELU'(h): dy, W, h fp16 in, dx fp16 out."""
    from rsl_rl_b200.networks.mlp import linear_backward_data
    nat = _nat()
    torch.manual_seed(M + K)
    h = torch.nn.functional.elu(torch.randn(M, K, device=DEV)).half()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).half()
    dy = torch.randn(M, N, device=DEV).half()
    dx = torch.full((M, K), float("nan"), device=DEV, dtype=torch.float16)
    linear_backward_data([(dy, w, h, dx, nat.ACT_CODES["elu"])], nat.PREC_FP16)
    torch.cuda.synchronize()
    ref = (dy.double() @ w.double()).cpu().numpy() * _elu_grad_from_output(h.double().cpu().numpy())
    assert rel_err(dx.float().cpu().numpy(), ref) < TOL


def test_fp16_backward_data_grouped_actor_and_critic():
    from rsl_rl_b200.networks.mlp import linear_backward_data
    nat = _nat()
    torch.manual_seed(5)
    M, probs, refs = 3000, [], []
    for K, N in ((512, 256), (256, 128)):
        h = torch.tanh(torch.randn(M, K, device=DEV)).half()
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).half()
        dy = torch.randn(M, N, device=DEV).half()
        dx = torch.full((M, K), float("nan"), device=DEV, dtype=torch.float16)
        probs.append((dy, w, h, dx, nat.ACT_CODES["tanh"]))
        refs.append((dy.double() @ w.double()).cpu().numpy() * (1 - h.double().cpu().numpy() ** 2))
    linear_backward_data(probs, nat.PREC_FP16)
    torch.cuda.synchronize()
    for p, ref in zip(probs, refs):
        assert rel_err(p[3].float().cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("M,K,N", [(24576, 48, 512), (24576, 512, 256), (24576, 256, 128), (1000, 64, 64), (333, 128, 192)])
def test_fp16_backward_weight(M, K, N):
    """dW += scale * dy^T x, db += scale * colsum(dy): dy, x fp16, dW / db fp32 accumulated."""
    from rsl_rl_b200.networks.mlp import linear_backward_weight
    nat = _nat()
    torch.manual_seed(M + K + N)
    Kp = (K + 7) // 8 * 8
    x = torch.zeros(M, Kp, device=DEV, dtype=torch.float16)[:, :K]
    x.copy_(torch.randn(M, K, device=DEV))
    dy = torch.randn(M, N, device=DEV).half()
    dw, db = torch.ones(N, K, device=DEV), torch.ones(N, device=DEV)
    scale = 1.0 / 64
    linear_backward_weight([(dy, x, dw, db, scale)], nat.PREC_FP16)
    torch.cuda.synchronize()
    ref_dw = 1 + scale * (dy.double().T @ x.double()).cpu().numpy()
    assert rel_err(dw.cpu().numpy(), ref_dw) < TOL
    db_err = (db.double() - 1 - scale * dy.double().sum(0)).abs().max() / (scale * dy.double().abs().sum(0).max())
    assert float(db_err) < 1e-5
    linear_backward_weight([(dy, x, dw, None, scale)], nat.PREC_FP16)       # accumulates; db optional
    assert rel_err(dw.cpu().numpy(), 2 * ref_dw - 1) < TOL


@pytest.mark.parametrize("M,K,N", [(24576, 128, 12), (4096, 128, 1), (1001, 64, 5)])
def test_fp16_head_backward(M, K, N):
    """Narrow output layer: h fp16 in, dx = S * (dy W) * ELU'(h) fp16 out, dW / db fp32 and unscaled."""
    from rsl_rl_b200.networks.mlp import linear_backward_head
    nat = _nat()
    torch.manual_seed(M + N)
    h = torch.nn.functional.elu(torch.randn(M, K, device=DEV)).half()
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    dy = torch.zeros(M, (N + 3) // 4 * 4, device=DEV)[:, :N]
    dy.normal_()
    dy.mul_(1.0 / M)
    S = float(2 ** round(np.log2(M)))
    dx = torch.full((M, K), float("nan"), device=DEV, dtype=torch.float16)
    dw, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    linear_backward_head([(dy, w, h, dx, dw, db, nat.ACT_CODES["elu"], S)])
    torch.cuda.synchronize()
    ref_dx = S * (dy.double() @ w.double()).cpu().numpy() * _elu_grad_from_output(h.double().cpu().numpy())
    assert rel_err(dx.float().cpu().numpy(), ref_dx) < TOL
    assert rel_err(dw.cpu().numpy(), (dy.double().T @ h.double()).cpu().numpy()) < 1e-5


def _mlp3_f16_case(M, dims_list, act, seed, store):
    from rsl_rl_b200.networks.mlp import mlp3_forward, mlp3_struct
    nat = _nat()
    torch.manual_seed(seed)
    structs, keep, refs = [], [], []
    for d in dims_list:
        x = torch.randn(M, d[0], device=DEV)
        ws = [torch.randn(d[l + 1], d[l], device=DEV) / d[l] ** 0.5 for l in range(4)]
        bs = [torch.randn(d[l + 1], device=DEV) * 0.3 for l in range(4)]
        w16 = [ws[0]] + [w.half() for w in ws[1:]]
        hs = [torch.full((M, d[l + 1]), float("nan"), device=DEV, dtype=torch.float16) for l in range(3)] if store else None
        y = torch.full((M, d[4]), float("nan"), device=DEV)
        structs.append(mlp3_struct(x.data_ptr(), d[0], M, w16, bs, hs, y, nat.ACT_CODES[act], nat.PREC_FP16))
        keep.append((x, w16, bs, hs, y))
        h, ref_h = x.double(), []
        for l in range(4):
            z = h @ ws[l].double().T + bs[l].double()
            h = torch.from_numpy(np.asarray(O.act_forward(act, z.cpu().numpy()), dtype=np.float64)).to(DEV) if l < 3 else z
            ref_h.append(h)
        refs.append(ref_h)
    mlp3_forward(structs)
    torch.cuda.synchronize()
    return keep, refs


@pytest.mark.parametrize("M", [24576, 4096, 1000, 77])
@pytest.mark.parametrize("store", [True, False])
def test_fp16_mlp3_forward_c2_shape(M, store):
    keep, refs = _mlp3_f16_case(M, [[48, 512, 256, 128, 12], [48, 512, 256, 128, 1]], "elu", M, store)
    for (x, w16, bs, hs, y), ref_h in zip(keep, refs):
        assert rel_err(y.cpu().numpy(), ref_h[3].cpu().numpy()) < TOL
        if store:
            for l in range(3):
                assert rel_err(hs[l].float().cpu().numpy(), ref_h[l].cpu().numpy()) < TOL, l


@pytest.mark.parametrize("dims,act", [([64, 384, 256, 64, 3], "tanh"), ([20, 256, 128, 128, 1], "elu"),
                                      ([48, 128, 128, 64, 12], "relu")])
def test_fp16_mlp3_forward_other_shapes(dims, act):
    keep, refs = _mlp3_f16_case(3000, [dims], act, sum(dims), True)
    (x, w16, bs, hs, y), ref_h = keep[0], refs[0]
    tol = 3e-3 if act == "tanh" else TOL
    assert rel_err(y.cpu().numpy(), ref_h[3].cpu().numpy()) < tol
    for l in range(3):
        assert rel_err(hs[l].float().cpu().numpy(), ref_h[l].cpu().numpy()) < tol, l


def test_fp16_convert_strided_and_flat():
    nat = _nat()
    torch.manual_seed(0)
    x = torch.randn(1000, 52, device=DEV)[:, :48]
    dst = torch.full((1000, 56), float("nan"), device=DEV, dtype=torch.float16)
    nat.check(nat.lib().rslrl_convert_f32_f16(x.data_ptr(), x.stride(0), dst.data_ptr(), dst.stride(0), 1000, 48, 56,
                                              nat.stream_ptr()))
    assert torch.equal(dst[:, :48], x.half()) and float(dst[:, 48:].abs().sum()) == 0.0
    flat = torch.randn(380321, device=DEV) * 100
    out = torch.empty(380321, device=DEV, dtype=torch.float16)
    nat.check(nat.lib().rslrl_convert_f32_f16(flat.data_ptr(), flat.numel(), out.data_ptr(), flat.numel(), 1,
                                              flat.numel(), flat.numel(), nat.stream_ptr()))
    assert torch.equal(out, flat.half())


@pytest.mark.parametrize("M,N1,N2,N3,act", [(24576, 512, 256, 128, "elu"), (1000, 512, 256, 128, "elu"), (4096, 256, 128, 64, "tanh"),
                                            (333, 128, 128, 64, "relu")])
def test_fp16_dgrad_chain_matches_two_layerwise_launches(M, N1, N2, N3, act):
    """rslrl_dgrad_chain16 (dZ2 and dZ1 of both nets in one launch, the dZ2 tile feeding the second product from shared
    memory) against two rslrl_linear_backward_data launches (same fp16 operands, fp32 accumulation) and against fp64."""
    import ctypes
    from rsl_rl_b200.networks.mlp import linear_backward_data
    nat = _nat()
    torch.manual_seed(M + N1)
    f = {"elu": torch.nn.functional.elu, "tanh": torch.tanh, "relu": torch.relu}[act]
    code = nat.ACT_CODES[act]
    arr, nets = (nat.DgradChain * 2)(), []
    for i in range(2):
        h1, h2 = f(torch.randn(M, N1, device=DEV)).half(), f(torch.randn(M, N2, device=DEV)).half()
        w3, w2 = (torch.randn(N3, N2, device=DEV) / N3 ** 0.5).half(), (torch.randn(N2, N1, device=DEV) / N2 ** 0.5).half()
        dz3 = torch.randn(M, N3, device=DEV).half()
        dz2 = torch.full((M, N2), float("nan"), device=DEV, dtype=torch.float16)
        dz1 = torch.full((M, N1), float("nan"), device=DEV, dtype=torch.float16)
        arr[i] = nat.DgradChain(dz3.data_ptr(), w3.data_ptr(), w2.data_ptr(), h2.data_ptr(), h1.data_ptr(), dz2.data_ptr(),
                                dz1.data_ptr(), M, N1, N2, N3, N3, N2, N1, N2, N1, code)
        nets.append((h1, h2, w3, w2, dz3, dz2, dz1))
    for _ in range(2):   # twice: the kernel must leave nothing behind
        nat.check(nat.lib().rslrl_dgrad_chain16(2, arr, nat.stream_ptr()))
    torch.cuda.synchronize()
    for h1, h2, w3, w2, dz3, dz2, dz1 in nets:
        r2 = torch.empty_like(dz2)
        r1 = torch.empty_like(dz1)
        linear_backward_data([(dz3, w3, h2, r2, code)], nat.PREC_FP16)
        linear_backward_data([(r2, w2, h1, r1, code)], nat.PREC_FP16)
        torch.cuda.synchronize()
        assert torch.equal(dz2, r2)
        assert rel_err(dz1.float().cpu().numpy(), r1.float().cpu().numpy()) < 1e-6
        grad = {"elu": lambda h: np.where(h > 0, 1.0, h + 1.0), "tanh": lambda h: 1 - h ** 2, "relu": lambda h: (h > 0) * 1.0}[act]
        ref2 = (dz3.double() @ w3.double()).cpu().numpy() * grad(h2.double().cpu().numpy())
        assert rel_err(dz2.float().cpu().numpy(), ref2) < TOL
        ref1 = (dz2.double() @ w2.double()).cpu().numpy() * grad(h1.double().cpu().numpy())
        assert rel_err(dz1.float().cpu().numpy(), ref1) < TOL
