"""GPU: every kernel behind include/rslrl_b200.h against the numpy oracle / the reference's golden
vectors.  Bit-exact for integer, byte, index and copy work and for the GAE scan; fp tolerances are the
north-star classes (1e-6 GAE, 1e-5 fp32 GEMM path) read as max-abs error / max-abs value."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from tests.golden_util import load, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def _nat():
    from rsl_rl_b200 import _native as nat
    return nat


# ----------------------------------------------------------------------------------------------- GAE
def _gae_cuda(rewards, values, dones, last_values, gamma, lam, normalize):
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite
    T, N = rewards.shape[:2]
    st = RolloutStorage("rl", N, T, TensorDictLite({"policy": torch.zeros(N, 1, device=DEV)}, batch_size=[N]), [1], DEV)
    st.rewards.copy_(cu(rewards)); st.values.copy_(cu(values)); st.dones.copy_(cu(dones))
    st.compute_returns(cu(last_values), gamma, lam, normalize_advantage=normalize)
    torch.cuda.synchronize()
    return st.returns.cpu().numpy(), st.advantages.cpu().numpy()


def test_gae_matches_reference_golden_vectors():
    g = load("gae_cases")
    names = sorted({k.rsplit(".", 1)[0] for k in g.files if k.endswith(".returns")})
    for n in names:
        gamma, lam = (0.9, 0.8) if n.startswith("odd") else (0.99, 0.95)
        ret, adv = _gae_cuda(g[n + ".rewards"], g[n + ".values"], g[n + ".dones"], g[n + ".last_values"], gamma, lam,
                             n.endswith("norm"))
        assert np.array_equal(ret, g[n + ".returns"]), n          # scan keeps the reference's fp32 op order
        if n.endswith("raw"):
            assert np.array_equal(adv, g[n + ".advantages"]), n
        else:
            assert rel_err(adv, g[n + ".advantages"]) < 1e-6, n


@pytest.mark.parametrize("T,N", [(24, 4096), (24, 65536), (40, 300), (7, 1), (24, 200000), (64, 5000)])
def test_gae_vs_oracle_all_kernel_variants(T, N):
    """Register kernels (T<=32), generic kernel (T>32 or more envs than resident threads)."""
    rng = np.random.default_rng(T * 1000 + N)
    rewards = rng.standard_normal((T, N, 1)).astype(np.float32)
    values = (2 * rng.standard_normal((T, N, 1))).astype(np.float32)
    dones = (rng.random((T, N, 1)) < 0.03).astype(np.uint8)
    last = rng.standard_normal((N, 1)).astype(np.float32)
    if T * N < 2:
        pytest.skip("std of a single element is undefined")
    ret_o, adv_o = O.compute_returns(rewards, values, dones, last, 0.99, 0.95, True)
    ret, adv = _gae_cuda(rewards, values, dones, last, 0.99, 0.95, True)
    assert np.array_equal(ret, ret_o)
    assert rel_err(adv, adv_o) < 1e-6
    # size-independent properties: normalised advantages have mean 0 / unbiased std 1
    a64 = adv.astype(np.float64)
    assert abs(a64.mean()) < 1e-5 and abs(a64.std(ddof=1) - 1) < 1e-5
    _, adv_raw = _gae_cuda(rewards, values, dones, last, 0.99, 0.95, False)
    assert np.array_equal(adv_raw, (ret_o - values))


# ------------------------------------------------------------------------------------ storage / gather
@pytest.mark.parametrize("dones_dtype", [torch.int64, torch.bool, torch.uint8, torch.float32, torch.int32])
@pytest.mark.parametrize("sigma_broadcast", [True, False])
def test_store_transition_bit_exact(dones_dtype, sigma_broadcast):
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite
    torch.manual_seed(1)
    N, A, T = 1000, 12, 3
    obs = TensorDictLite({"a": torch.randn(N, 48, device=DEV), "b": torch.randn(N, 7, device=DEV)}, batch_size=[N])
    st = RolloutStorage("rl", N, T, obs, [A], DEV)
    gamma = 0.99
    for step in range(T):
        tr = RolloutStorage.Transition()
        tr.observations = TensorDictLite({"a": torch.randn(N, 48, device=DEV), "b": torch.randn(N, 7, device=DEV)},
                                         batch_size=[N])
        tr.actions, tr.action_mean = torch.randn(N, A, device=DEV), torch.randn(N, A, device=DEV)
        std = torch.rand(A, device=DEV) + 0.5
        tr.action_sigma = std.expand(N, A) if sigma_broadcast else torch.rand(N, A, device=DEV)
        tr.values, tr.actions_log_prob = torch.randn(N, 1, device=DEV), torch.randn(N, device=DEV)
        tr.rewards = torch.randn(N, device=DEV)
        d = torch.rand(N, device=DEV) < 0.3
        tr.dones = d.to(dones_dtype)
        to = (torch.rand(N, device=DEV) < 0.5) & d if step != 1 else None
        exp_r = tr.rewards.clone()
        if to is not None:
            exp_r += gamma * torch.squeeze(tr.values * to.unsqueeze(1), 1)   # the reference's expression
        st._store(tr, time_outs=to, gamma=gamma)
        torch.cuda.synchronize()
        assert torch.equal(st.observations["a"][step], tr.observations["a"])
        assert torch.equal(st.observations["b"][step], tr.observations["b"])
        assert torch.equal(st.actions[step], tr.actions) and torch.equal(st.mu[step], tr.action_mean)
        assert torch.equal(st.sigma[step], tr.action_sigma.contiguous())
        assert torch.equal(st.values[step], tr.values)
        assert torch.equal(st.actions_log_prob[step].view(-1), tr.actions_log_prob)
        assert torch.equal(st.rewards[step].view(-1), exp_r)
        assert torch.equal(st.dones[step].view(-1), d.to(torch.uint8))
    with pytest.raises(OverflowError):
        st._store(tr)


@pytest.mark.parametrize("widths", [(48,), (235, 1, 12), (187, 48, 4, 3)])
def test_permute_gather_bit_exact_with_concat(widths):
    from rsl_rl_b200.storage import RolloutStorage
    torch.manual_seed(2)
    B, n = 5000, 1237
    idx = torch.randperm(B, device=DEV)[:n]
    srcs = [torch.randn(B, w, device=DEV) for w in widths]
    total = sum(widths)
    ld = (total + 3) // 4 * 4
    dst = torch.zeros(n, ld, device=DEV)
    st = RolloutStorage.__new__(RolloutStorage)
    cols = list(np.cumsum([0] + list(widths[:-1])))
    st._gather(srcs, list(widths), idx, [dst] * len(widths), [ld] * len(widths), [int(c) for c in cols])
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :total], torch.cat([s[idx] for s in srcs], dim=1))


def test_mini_batch_generator_matches_reference_protocol():
    """Same 10-tuple, bit-equal tensors for the same permutation (reference rollout_storage.py:160-203)."""
    from rsl_rl_b200.storage import RolloutStorage
    from rsl_rl_b200.utils import TensorDictLite
    torch.manual_seed(3)
    N, T, A, M, E = 64, 8, 4, 4, 2
    obs = TensorDictLite({"policy": torch.zeros(N, 12, device=DEV), "extra": torch.zeros(N, 5, device=DEV)}, batch_size=[N])
    st = RolloutStorage("rl", N, T, obs, [A], DEV)
    for t in (st.observations["policy"], st.observations["extra"], st.actions, st.values, st.advantages, st.returns,
              st.actions_log_prob, st.mu, st.sigma):
        t.normal_()
    perm = torch.randperm(N * T, device=DEV)
    st._randperm = lambda n: perm
    batches = list(st.mini_batch_generator(M, E))
    assert len(batches) == M * E
    mb = N * T // M
    for k, b in enumerate(batches):
        idx = perm[(k % M) * mb:(k % M + 1) * mb]
        obs_b, actions, values, adv, ret, logp, mu, sigma, hid, masks = b
        assert hid == (None, None) and masks is None and obs_b.batch_size[0] == mb
        assert torch.equal(obs_b["policy"], st.observations["policy"].flatten(0, 1)[idx])
        assert torch.equal(obs_b["extra"], st.observations["extra"].flatten(0, 1)[idx])
        for got, src in ((actions, st.actions), (values, st.values), (adv, st.advantages), (ret, st.returns),
                         (logp, st.actions_log_prob), (mu, st.mu), (sigma, st.sigma)):
            assert torch.equal(got, src.flatten(0, 1)[idx]) and got.is_contiguous()
    # default path: torch.randperm on the device, every index exactly once
    del st._randperm
    st._randperm = lambda n: torch.randperm(n, device=DEV)
    seen = torch.cat([b[1] for b in list(st.mini_batch_generator(M, 1))])
    assert seen.shape[0] == N * T


# -------------------------------------------------------------------------------------------- GEMMs
def _lin_ref(x, w, b, act):
    z = x.double() @ w.double().T + (0 if b is None else b.double())
    return torch.from_numpy(O.act_forward(act, z.cpu().numpy().astype(np.float32)).astype(np.float64)).to(DEV) \
        if act != "identity" else z


@pytest.mark.parametrize("M,K,N,act", [(1000, 48, 512, "elu"), (4096, 235, 512, "elu"), (777, 512, 256, "relu"),
                                        (2048, 128, 12, "identity"), (2048, 128, 1, "identity"), (130, 5, 3, "tanh"),
                                        (24576, 256, 128, "elu"), (64, 16, 200, "selu"), (300, 33, 65, "sigmoid"),
                                        (300, 33, 65, "lrelu"), (300, 33, 65, "softplus")])
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("tf32", 1e-3), ("tf32x3", 1e-5)])
def test_linear_forward(M, K, N, act, prec, tol):
    from rsl_rl_b200.networks.mlp import linear_forward
    nat = _nat()
    torch.manual_seed(M + K + N)
    x = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV) / K ** 0.5; b = torch.randn(N, device=DEV)
    y = torch.full((M, N), float("nan"), device=DEV)
    linear_forward([(x, w, b, y, nat.ACT_CODES[act])], nat.PREC_CODES[prec])
    ref = _lin_ref(x, w, b, act)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < tol


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("tf32", 1e-3), ("tf32x3", 1e-5)])
def test_linear_forward_grouped_and_strided_input(prec, tol):
    from rsl_rl_b200.networks.mlp import linear_forward
    nat = _nat()
    torch.manual_seed(5)
    xa = torch.randn(3000, 52, device=DEV)[:, :50]          # row stride 52, K = 50 (padded minibatch buffer)
    xc = torch.randn(3000, 48, device=DEV)
    wa, ba = torch.randn(96, 50, device=DEV), torch.randn(96, device=DEV)
    wc, bc = torch.randn(33, 48, device=DEV), None
    ya, yc = torch.empty(3000, 96, device=DEV), torch.empty(3000, 33, device=DEV)
    linear_forward([(xa, wa, ba, ya, nat.ACT_CODES["elu"]), (xc, wc, bc, yc, 0)], nat.PREC_CODES[prec])
    assert rel_err(ya.cpu().numpy(), _lin_ref(xa, wa, ba, "elu").cpu().numpy()) < tol
    assert rel_err(yc.cpu().numpy(), _lin_ref(xc, wc, bc, "identity").cpu().numpy()) < tol


@pytest.mark.parametrize("M,K,N,act", [(1000, 512, 256, "elu"), (4096, 128, 12, "elu"), (3000, 128, 1, "relu"),
                                        (24576, 256, 128, "elu"), (513, 37, 5, "tanh"), (600, 48, 512, "identity")])
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("tf32", 1e-3), ("tf32x3", 1e-5)])
def test_linear_backward(M, K, N, act, prec, tol):
    """dx = (dy W) * act'(h) (+ fused bias gradient of the previous layer), dW += dy^T x, db += colsum(dy)
    against an fp64 restatement."""
    from rsl_rl_b200.networks.mlp import linear_backward_data, linear_backward_weight
    nat = _nat()
    P = nat.PREC_CODES[prec]
    torch.manual_seed(M + N)
    z_prev = torch.randn(M, K, device=DEV)
    h = torch.from_numpy(O.act_forward(act, z_prev.cpu().numpy())).to(DEV)          # this layer's input
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    Np = (N + 3) // 4 * 4
    dy = torch.zeros(M, Np, device=DEV)[:, :N]                                      # 16-byte row pitch (TMA)
    dy.normal_()
    dx = torch.full((M, K), float("nan"), device=DEV)
    db_prev = torch.zeros(K, device=DEV)
    linear_backward_data([(dy, w, h, dx, nat.ACT_CODES[act], db_prev)], P)
    dh = (dy.double() @ w.double()).cpu().numpy()
    ref_dx = O.act_backward(act, z_prev.cpu().numpy(), h.cpu().numpy(), dh.astype(np.float32)) if act != "identity" else dh
    assert rel_err(dx.cpu().numpy(), ref_dx) < tol
    assert rel_err(db_prev.cpu().numpy(), np.asarray(ref_dx, dtype=np.float64).sum(0)) < 10 * tol
    dw = torch.zeros(N, K, device=DEV); db = torch.zeros(N, device=DEV)
    linear_backward_weight([(dy, h, dw, db)], P)
    assert rel_err(dw.cpu().numpy(), (dy.double().T @ h.double()).cpu().numpy()) < tol
    assert rel_err(db.cpu().numpy(), dy.double().sum(0).cpu().numpy()) < tol   # tf32: sums the rounded dY tiles
    linear_backward_weight([(dy, h, dw, None)], P)                                  # accumulates; db optional
    assert rel_err(dw.cpu().numpy(), 2 * (dy.double().T @ h.double()).cpu().numpy()) < tol


def test_mlp_autograd_function_matches_torch():
    from rsl_rl_b200.networks import MLP
    torch.manual_seed(7)
    mlp = MLP(20, 3, [32, -1, 16], "elu").to(DEV)
    ref = torch.nn.Sequential(*[torch.nn.Linear(l.in_features, l.out_features) if isinstance(l, torch.nn.Linear)
                                else torch.nn.ELU() for l in mlp]).to(DEV).double()
    ref.load_state_dict({k: v.double() for k, v in mlp.state_dict().items()})
    x = torch.randn(257, 20, device=DEV, requires_grad=True)
    xr = x.detach().double().requires_grad_(True)
    y, yr = mlp(x), ref(xr)
    assert rel_err(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < 1e-5
    y.square().sum().backward(); yr.square().sum().backward()
    assert rel_err(x.grad.cpu().numpy(), xr.grad.cpu().numpy()) < 1e-5
    for p, pr in zip(mlp.parameters(), ref.parameters()):
        assert rel_err(p.grad.cpu().numpy(), pr.grad.cpu().numpy()) < 1e-5
    assert list(mlp.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias",
                                             "6.weight", "6.bias"]


# ------------------------------------------------------------------------------- gaussian head / loss
@pytest.mark.parametrize("A,std_type", [(12, "scalar"), (5, "log"), (3, "scalar")])
def test_gaussian_act(A, std_type):
    nat = _nat()
    torch.manual_seed(A)
    n = 4096
    mu, z = torch.randn(n, A, device=DEV), torch.randn(n, A, device=DEV)
    param = (torch.rand(A, device=DEV) + 0.3) if std_type == "scalar" else torch.randn(A, device=DEV) * 0.3
    sigma = param if std_type == "scalar" else torch.exp(param)
    act, lp, sig = torch.empty(n, A, device=DEV), torch.empty(n, device=DEV), torch.empty(n, A, device=DEV)
    nat.check(nat.lib().rslrl_gaussian_act(mu.data_ptr(), param.data_ptr(), 0 if std_type == "scalar" else 1, z.data_ptr(),
                                           act.data_ptr(), lp.data_ptr(), sig.data_ptr(), n, A, nat.stream_ptr()))
    sg = sigma.expand(n, A).contiguous()
    if std_type == "scalar":
        assert torch.equal(act, z * sg + mu)       # bit-exact: mul then add, like Normal.sample()
        assert torch.equal(sig, sg)
    else:
        assert rel_err(act.cpu().numpy(), (z * sg + mu).cpu().numpy()) < 1e-6
    lp_o = O.gaussian_log_prob(mu.cpu().numpy(), sig.cpu().numpy(), act.cpu().numpy())
    assert rel_err(lp.cpu().numpy(), lp_o) < 2e-6


def _run_loss(mu, std_param, std_type, value, actions, old_logp, adv, ret, old_v, old_mu, old_sigma, clipped,
              mb_orig=None, per_mb_norm=False):
    nat = _nat()
    rows, A = mu.shape
    mb_orig = rows if mb_orig is None else mb_orig
    t = lambda a: cu(np.asarray(a, dtype=np.float32).reshape(a.shape))
    d = dict(mu=t(mu), sp=t(std_param), value=t(value.reshape(-1)), actions=t(actions), old_logp=t(old_logp.reshape(-1)),
             adv=t(adv.reshape(-1)), ret=t(ret.reshape(-1)), old_v=t(old_v.reshape(-1)), old_mu=t(old_mu), old_sigma=t(old_sigma))
    A4 = (A + 3) // 4 * 4
    dmu_buf, dv_buf = torch.zeros(rows, A4, device=DEV), torch.zeros(rows, 4, device=DEV)
    dmu, dv = dmu_buf[:, :A], dv_buf[:, 0]
    dstd, kl = torch.zeros(A, device=DEV), torch.zeros(1, device=DEV)
    db_mu, db_v = torch.zeros(A, device=DEV), torch.zeros(1, device=DEV)
    sums = torch.zeros(3, dtype=torch.float64, device=DEV)
    ws = torch.zeros(int(nat.lib().rslrl_loss_workspace_bytes()), dtype=torch.uint8, device=DEV)
    stats = None
    if per_mb_norm:
        stats = torch.zeros(2, dtype=torch.float64, device=DEV)
        sws = torch.zeros(64, dtype=torch.uint8, device=DEV)
        nat.check(nat.lib().rslrl_sum_sumsq(d["adv"].data_ptr(), mb_orig, stats.data_ptr(), sws.data_ptr(), nat.stream_ptr()))
    # optional scaled fp16 copies of the output gradients (four-threads-per-row kernel only: A % 4 == 0, A <= 16)
    f16out = A % 4 == 0 and A <= 16
    S16 = 4096.0
    dmu16 = torch.zeros(rows, (A + 7) // 8 * 8, device=DEV, dtype=torch.float16)
    dv16 = torch.zeros(rows, 8, device=DEV, dtype=torch.float16)
    # optional fused data gradient of the two output layers (same kernel): dZ3 = S * (dY W4) * ELU'(H3) as fp16
    head, head_keep = None, None
    if f16out and rows <= 148 * 1024:
        import ctypes
        K3 = 128
        g = torch.Generator(device=DEV).manual_seed(rows)
        h_a = torch.nn.functional.elu(torch.randn(rows, K3, device=DEV, generator=g)).half()
        h_c = torch.nn.functional.elu(torch.randn(rows, K3, device=DEV, generator=g)).half()
        w_a = torch.randn(A, K3, device=DEV, generator=g) / K3 ** 0.5
        w_c = torch.randn(1, K3, device=DEV, generator=g) / K3 ** 0.5
        dz_a = torch.full((rows, K3), float("nan"), device=DEV, dtype=torch.float16)
        dz_c = torch.full((rows, K3), float("nan"), device=DEV, dtype=torch.float16)
        hs = nat.LossHead(w_a.data_ptr(), w_c.data_ptr(), h_a.data_ptr(), h_c.data_ptr(), dz_a.data_ptr(), dz_c.data_ptr(),
                          K3, K3, nat.ACT_CODES["elu"])
        head, head_keep = ctypes.byref(hs), (hs, h_a, h_c, w_a, w_c, dz_a, dz_c)
    for _ in range(2):  # twice: the workspace must clean itself up
        dstd.zero_(); sums.zero_(); db_mu.zero_(); db_v.zero_()
        nat.check(nat.lib().rslrl_ppo_loss_fwd_bwd(
            d["mu"].data_ptr(), d["sp"].data_ptr(), std_type, d["value"].data_ptr(), d["actions"].data_ptr(),
            d["old_logp"].data_ptr(), d["adv"].data_ptr(), d["ret"].data_ptr(), d["old_v"].data_ptr(), d["old_mu"].data_ptr(),
            d["old_sigma"].data_ptr(), nat.ptr(stats), dmu.data_ptr(), A4, dv.data_ptr(), 4, dstd.data_ptr(),
            db_mu.data_ptr(), db_v.data_ptr(), kl.data_ptr(), sums.data_ptr(), ws.data_ptr(), rows, mb_orig, A, 0.2, 1.0,
            0.01, int(clipped), dmu16.data_ptr() if f16out else None, dmu16.stride(0), dv16.data_ptr() if f16out else None,
            dv16.stride(0), S16, None, head, nat.stream_ptr()))
    torch.cuda.synchronize()
    if f16out:
        sat = lambda t: (t * S16).clamp(-65504.0, 65504.0).half()   # saturating conversion (never inf)
        assert torch.equal(dmu16[:, :A], sat(dmu)) and torch.equal(dv16[:, 0], sat(dv))
        assert float(dmu16[:, A:].abs().sum()) == 0.0 and float(dv16[:, 1:].abs().sum()) == 0.0
    if head is not None:
        _, h_a, h_c, w_a, w_c, dz_a, dz_c = head_keep
        egrad = lambda h: torch.where(h > 0, torch.ones_like(h), h + 1.0)
        # the product consumes the (saturated) fp16 copies S * dmu, S * dV and the fp16-rounded weights with fp32
        # accumulation: the arithmetic of the tensor-core launch this fusion replaces
        ref_a = (dmu16[:, :A].double() @ w_a.half().double()) * egrad(h_a.double())
        ref_c = (dv16[:, :1].double() * w_c.half().double()) * egrad(h_c.double())
        assert rel_err(dz_a.double().clamp(-65504, 65504).cpu().numpy(), ref_a.clamp(-65504, 65504).cpu().numpy()) < 1e-3
        assert rel_err(dz_c.double().clamp(-65504, 65504).cpu().numpy(), ref_c.clamp(-65504, 65504).cpu().numpy()) < 1e-3
    # the fused bias gradients of the two output layers are the column sums of dmu / dvalue
    assert rel_err(db_mu.cpu().numpy(), dmu.double().sum(0).cpu().numpy()) < 1e-5
    assert abs(float(db_v) - float(dv.double().sum())) < 1e-5 * max(1e-6, float(dv.double().abs().sum()))
    assert float(dmu_buf[:, A:].abs().sum()) == 0.0 and float(dv_buf[:, 1:].abs().sum()) == 0.0
    return dmu.cpu().numpy(), dv.cpu().numpy(), dstd.cpu().numpy(), float(kl), sums.cpu().numpy()


@pytest.mark.parametrize("tag", ["clipped", "unclipped"])
def test_ppo_loss_against_reference_autograd(tag):
    g = load("loss_terms")
    f = lambda k: g[f"{tag}.{k}"]
    sigma = f("sigma")
    dmu, dv, dstd, kl, sums = _run_loss(f("mu"), sigma[0], 0, f("value"), f("actions"), f("old_logp"), f("advantages"),
                                        f("returns"), f("old_values"), f("old_mu"), f("old_sigma"), tag == "clipped")
    assert rel_err(dmu, f("dmu")) < 1e-5
    assert rel_err(dv, f("dvalue").reshape(-1)) < 1e-5
    assert rel_err(dstd, f("dsigma").sum(0)) < 1e-5
    assert abs(kl - float(f("kl"))) < 1e-5 * abs(float(f("kl")))
    assert abs(sums[0] - float(f("surrogate"))) < 1e-5 * max(1.0, abs(float(f("surrogate"))))
    assert abs(sums[1] - float(f("value_loss"))) < 1e-5 * abs(float(f("value_loss")))
    assert abs(sums[2] - float(f("entropy"))) < 1e-5 * abs(float(f("entropy")))


@pytest.mark.parametrize("A,std_type,rows,aug", [(12, 0, 24576, 1), (12, 1, 5000, 1), (7, 0, 999, 1), (12, 0, 4096, 2),
                                                 (12, 0, 160000, 1)])   # > 148 * 1024 rows: the row-per-thread kernel
def test_ppo_loss_against_oracle(A, std_type, rows, aug):
    rng = np.random.default_rng(rows + A)
    mb = rows
    R = rows * aug
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    mu, actions = f(R, A), f(R, A)
    param = (rng.random(A).astype(np.float32) + 0.5) if std_type == 0 else (0.2 * f(A))
    sigma_vec = param if std_type == 0 else np.exp(param)
    sigma = np.broadcast_to(sigma_vec, (R, A)).copy()
    value, old_v, ret, adv = f(R), f(mb), f(mb), f(mb)
    old_mu, old_sigma = mu[:mb] + 0.1 * f(mb, A), np.abs(f(mb, A)) + 0.5
    old_logp = O.gaussian_log_prob(mu[:mb], sigma[:mb], actions[:mb]) + 0.3 * f(mb)
    rep = lambda x: np.concatenate([x] * aug)
    o = O.ppo_loss_fwd_bwd(mu, sigma, value, actions, rep(old_logp), rep(adv), rep(ret), rep(old_v), 0.2, 1.0, 0.01, True,
                           n_entropy=mb)
    dmu, dv, dstd, kl, sums = _run_loss(mu, param, std_type, value, actions, old_logp, adv, ret, old_v, old_mu, old_sigma,
                                        True, mb_orig=mb)
    ds = o["dsigma"].sum(0, dtype=np.float64)
    if std_type == 1:
        ds = (o["dsigma"] * sigma).sum(0, dtype=np.float64)
    assert rel_err(dmu, o["dmu"]) < 1e-5 and rel_err(dv, o["dvalue"].reshape(-1)) < 1e-5
    assert rel_err(dstd, ds) < 1e-4
    assert abs(kl - O.kl_mean(mu[:mb], sigma[:mb], old_mu, old_sigma)) < 1e-5 * abs(kl)
    assert abs(sums[0] - o["surrogate"]) < 1e-5 * max(1, abs(o["surrogate"]))
    assert abs(sums[1] - o["value_loss"]) < 1e-5 * abs(o["value_loss"])
    assert abs(sums[2] - o["entropy"]) < 1e-5 * abs(o["entropy"])


def test_ppo_loss_per_minibatch_advantage_normalisation():
    rng = np.random.default_rng(11)
    mb, A = 3000, 12
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    mu, actions, value, old_v, ret, adv = f(mb, A), f(mb, A), f(mb), f(mb), f(mb), 3 * f(mb) + 1
    sig = np.ones(A, np.float32)
    sigma = np.broadcast_to(sig, (mb, A)).copy()
    old_logp = O.gaussian_log_prob(mu, sigma, actions) + 0.1 * f(mb)
    a64 = adv.astype(np.float64)
    adv_n = ((adv - np.float32(a64.mean())) / (np.float32(a64.std(ddof=1)) + np.float32(1e-8))).astype(np.float32)
    o = O.ppo_loss_fwd_bwd(mu, sigma, value, actions, old_logp, adv_n, ret, old_v, 0.2, 1.0, 0.01, True)
    dmu, dv, dstd, kl, sums = _run_loss(mu, sig, 0, value, actions, old_logp, adv, ret, old_v, mu, sigma, True,
                                        per_mb_norm=True)
    assert rel_err(dmu, o["dmu"]) < 1e-5 and abs(sums[0] - o["surrogate"]) < 1e-5 * max(1, abs(o["surrogate"]))


# --------------------------------------------------------------------------------------- optimiser
@pytest.mark.parametrize("world,adaptive", [(1, True), (2, True), (1, False)])
def test_clip_adam_kl_lr_against_oracle(world, adaptive):
    from collections import OrderedDict
    from rsl_rl_b200.algorithms.flat_params import FlatParamGroup, FusedAdam
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(48, 64), torch.nn.ELU(), torch.nn.Linear(64, 7)).to(DEV)
    names = [k for k, _ in net.named_parameters()]
    params = OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in net.named_parameters())
    group = FlatParamGroup(net.parameters(), DEV)
    opt = FusedAdam(group, lr=1e-3)
    adam = O.adam_init(params)
    lr = 1e-3
    rng = np.random.default_rng(1)
    kls = [0.05, 0.001, 0.03, 0.011, 0.0, 0.004, 0.1, 0.0049]
    for step, kl in enumerate(kls):
        grads = OrderedDict((k, (rng.standard_normal(v.shape) * (3.0 if step % 2 else 0.01)).astype(np.float32))
                            for k, v in params.items())
        for p, k in zip(net.parameters(), names):
            p.grad.copy_(cu(grads[k]) * world)                  # the buffer holds the SUM over ranks
        group.kl_slot.fill_(np.float32(kl) * world)
        opt.fused_step(world_size=world, max_grad_norm=1.0, adaptive=adaptive, desired_kl=0.01,
                       round_lr_fp32=world > 1)
        if adaptive:
            lr = O.adaptive_lr(lr, kl, 0.01)
            if world > 1:
                lr = float(np.float32(lr))
        total = O.clip_grad_norm(grads, 1.0)
        O.adam_step(params, grads, adam, lr)
        info = opt.dev_state.read()
        assert abs(info["lr"] - lr) < 1e-18 + 1e-12 * lr, (step, info["lr"], lr)
        assert abs(info["total_norm"] - total) < 1e-5 * total
        assert info["step"] == step + 1
        assert float(group.grads.abs().max()) == 0.0          # cleared for the next backward
        flat = np.concatenate([params[k].reshape(-1) for k in names])
        mine = np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in net.parameters()])
        assert rel_err(mine, flat) < 1e-6, step


# ------------------------------------------------------------------------------------ trajectories
def test_trajectory_kernels_bit_exact_vs_reference():
    from rsl_rl_b200.utils import TrajectoryIndex, split_and_pad_trajectories, unpad_trajectories
    g = load("trajectory_cases")
    for n in ("small", "none", "all", "c4like"):
        x, dones = cu(g[n + ".x"]), cu(g[n + ".dones"])
        padded, masks = split_and_pad_trajectories(x, dones)
        assert np.array_equal(masks.cpu().numpy(), g[n + ".masks"]), n
        assert np.array_equal(padded.cpu().numpy(), g[n + ".padded"]), n
        assert np.array_equal(unpad_trajectories(padded, masks).cpu().numpy(), g[n + ".unpadded"]), n
        index = TrajectoryIndex(dones)
        n_traj = index.n_traj
        assert n_traj == g[n + ".masks"].shape[1]
        hid = index.hidden_at_starts(cu(g[n + ".saved_hidden"]), 0, n_traj)
        assert np.array_equal(hid.cpu().numpy(), g[n + ".hidden_at_starts"]), n
        T, N = dones.shape[:2]
        # minibatch-style slices: env columns [e0,e1) <-> trajectories [off[e0], off[e1])
        offs = index.offsets_host()
        e0, e1 = N // 4, N // 2 if N >= 4 else N
        if e1 > e0:
            sub = index.split_and_pad(x, offs[e0], offs[e1])
            assert np.array_equal(sub.cpu().numpy(), g[n + ".padded"][:, offs[e0]:offs[e1]]), n
            un = index.unpad(sub, e0, e1, offs[e0])
            assert np.array_equal(un.cpu().numpy(), g[n + ".x"][:, e0:e1]), n


def test_trajectory_index_at_c4_size_properties():
    from rsl_rl_b200.utils import TrajectoryIndex
    torch.manual_seed(4)
    T, N = 24, 4096
    dones = (torch.rand(T, N, 1, device=DEV) < 0.02).to(torch.uint8)
    index = TrajectoryIndex(dones)
    n_traj = index.n_traj
    assert n_traj == N + int(dones[:-1].sum())
    masks = index.masks(0, n_traj)
    assert int(masks.sum()) == T * N                                 # every (t, env) appears exactly once
    x = torch.randn(T, N, 48, device=DEV)
    padded = index.split_and_pad(x, 0, n_traj)
    assert torch.equal(index.unpad(padded, 0, N, 0), x)              # round trip
    assert float(padded[~masks].abs().max()) == 0.0
    pad_o, mask_o = O.split_and_pad_trajectories(x[:, :64].cpu().numpy(), dones[:, :64].cpu().numpy())
    k = mask_o.shape[1]
    assert np.array_equal(padded[:, :k].cpu().numpy(), pad_o) and np.array_equal(masks[:, :k].cpu().numpy(), mask_o)


@pytest.mark.parametrize("M,K,N,act", [(24576, 128, 12, "elu"), (4096, 128, 1, "elu"), (3001, 256, 5, "tanh"),
                                        (2000, 48, 12, "identity")])
def test_linear_backward_head_fused(M, K, N, act):
    """Fused narrow-layer backward: dx, dW (accumulated), db (accumulated) against fp64."""
    from rsl_rl_b200.networks.mlp import head_backward_ok, linear_backward_head
    nat = _nat()
    torch.manual_seed(M + N)
    z_prev = torch.randn(M, K, device=DEV)
    h = torch.from_numpy(O.act_forward(act, z_prev.cpu().numpy())).to(DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    dy = torch.zeros(M, (N + 3) // 4 * 4, device=DEV)[:, :N]
    dy.normal_()
    dx = torch.full((M, K), float("nan"), device=DEV)
    dw, db = torch.ones(N, K, device=DEV), torch.ones(N, device=DEV)
    assert head_backward_ok(N, K, h, dx)
    linear_backward_head([(dy, w, h, dx, dw, db, nat.ACT_CODES[act])])
    dh = (dy.double() @ w.double()).cpu().numpy()
    ref_dx = O.act_backward(act, z_prev.cpu().numpy(), h.cpu().numpy(), dh.astype(np.float32)) if act != "identity" else dh
    assert rel_err(dx.cpu().numpy(), ref_dx) < 1e-5
    assert rel_err(dw.cpu().numpy(), 1 + (dy.double().T @ h.double()).cpu().numpy()) < 1e-5
    # a column sum of M zero-mean values cancels: measure the error against the sum of magnitudes
    db_err = (db.double() - 1 - dy.double().sum(0)).abs().max() / dy.double().abs().sum(0).max()
    assert float(db_err) < 1e-6


# ---------------------------------------------------------------------------------- fused MLP forward
def _mlp3_case(M, dims, act, seed, store, pad_k0=False):
    """Runs rslrl_mlp3_forward on one or two nets; returns (outputs, saved activations, fp64 references)."""
    from rsl_rl_b200.networks.mlp import mlp3_forward, mlp3_struct
    nat = _nat()
    torch.manual_seed(seed)
    structs, keep, refs = [], [], []
    for d in dims:
        K0 = d[0]
        ld = (K0 + 3) // 4 * 4 if pad_k0 else K0
        xbuf = torch.zeros(M, ld, device=DEV)
        xbuf[:, :K0] = torch.randn(M, K0, device=DEV)
        ws = [torch.randn(d[l + 1], d[l], device=DEV) / d[l] ** 0.5 for l in range(4)]
        bs = [torch.randn(d[l + 1], device=DEV) * 0.3 for l in range(4)]
        w0 = ws[0]
        if ld != K0:  # zero-padded first-layer shadow, as the engine keeps it
            w0 = torch.zeros(d[1], ld, device=DEV)
            w0[:, :K0] = ws[0]
        hs = [torch.full((M, d[l + 1]), float("nan"), device=DEV) for l in range(3)] if store else None
        y = torch.full((M, d[4]), float("nan"), device=DEV)
        structs.append(mlp3_struct(xbuf.data_ptr(), ld, M, [w0] + ws[1:], bs, hs, y, nat.ACT_CODES[act]))
        keep.append((xbuf, w0, ws, bs, hs, y))
        h = xbuf[:, :K0]
        ref_h = []
        for l in range(4):
            h = _lin_ref(h.float(), ws[l], bs[l], act if l < 3 else "identity").float()
            ref_h.append(h)
        refs.append(ref_h)
    mlp3_forward(structs)
    torch.cuda.synchronize()
    return keep, refs


@pytest.mark.parametrize("M", [24576, 4096, 1000, 77])
@pytest.mark.parametrize("store", [True, False])
def test_mlp3_fused_forward_c2_shape(M, store):
    """Actor + critic of the C2 config in one launch vs an fp64-accumulated reference (tf32 class: 1e-3)."""
    keep, refs = _mlp3_case(M, [[48, 512, 256, 128, 12], [48, 512, 256, 128, 1]], "elu", M, store)
    for (xbuf, w0, ws, bs, hs, y), ref_h in zip(keep, refs):
        assert rel_err(y.cpu().numpy(), ref_h[3].cpu().numpy()) < 1e-3
        if store:
            for l in range(3):
                assert rel_err(hs[l].cpu().numpy(), ref_h[l].cpu().numpy()) < 1e-3, l


@pytest.mark.parametrize("dims,act,pad", [([235, 512, 256, 128, 12], "elu", True), ([48, 128, 128, 32, 16], "relu", False),
                                          ([64, 384, 256, 96, 3], "tanh", False), ([20, 256, 128, 64, 1], "elu", False)])
def test_mlp3_fused_forward_other_shapes(dims, act, pad):
    keep, refs = _mlp3_case(3000, [dims], act, sum(dims), True, pad_k0=pad)
    (xbuf, w0, ws, bs, hs, y), ref_h = keep[0], refs[0]
    assert rel_err(y.cpu().numpy(), ref_h[3].cpu().numpy()) < 1e-3
    # a saturating activation bounds max|h| by 1, so the max-norm relative error of the 1.1M tf32 dot products is
    # larger than for ELU / ReLU layers (whose max|h| grows with the input scale)
    tol = 3e-3 if act == "tanh" else 1e-3
    for l in range(3):
        assert rel_err(hs[l].cpu().numpy(), ref_h[l].cpu().numpy()) < tol, l


def test_mlp3_fused_forward_matches_layerwise_tf32_path():
    """The fused kernel and the layer-by-layer tcgen05 path agree to tf32 noise (they differ in where the hidden
    activations are rounded to tf32: epilogue, ties away from zero, vs TMA on the next layer's load)."""
    from rsl_rl_b200.networks.mlp import linear_forward
    nat = _nat()
    M, d = 4096, [48, 512, 256, 128, 12]
    keep, _ = _mlp3_case(M, [d], "elu", 11, True)
    xbuf, w0, ws, bs, hs, y = keep[0]
    h = xbuf
    for l in range(4):
        out = torch.empty(M, d[l + 1], device=DEV)
        linear_forward([(h, ws[l], bs[l], out, nat.ACT_CODES["elu"] if l < 3 else 0)], nat.PREC_TF32)
        h = out
    assert rel_err(y.cpu().numpy(), h.cpu().numpy()) < 5e-4


@pytest.mark.parametrize("M,A,std_type,prec", [(4096, 12, "scalar", "tf32"), (1000, 12, "log", "fp16"), (77, 5, "scalar", "tf32")])
def test_mlp3_plan_with_fused_gaussian_head_bit_exact_vs_separate_kernels(M, A, std_type, prec):
    """The one-launch rollout step (rslrl_mlp3_plan_run with the Gaussian head in the actor's output epilogue) gives the
    bits of rslrl_mlp3_forward followed by rslrl_gaussian_act, also after the input pointer changes."""
    import ctypes as C
    from rsl_rl_b200.networks.mlp import mlp3_forward, mlp3_struct
    nat = _nat()
    torch.manual_seed(M + A)
    f16 = prec == "fp16"
    dims = [[48, 512, 256, 128, A], [48, 512, 256, 128, 1]]
    std = (torch.rand(A, device=DEV) + 0.3) if std_type == "scalar" else torch.randn(A, device=DEV) * 0.3
    code = nat.STD_SCALAR if std_type == "scalar" else nat.STD_LOG
    z = torch.randn(M, A, device=DEV)
    nets = []
    for d in dims:
        ws = [torch.randn(d[l + 1], d[l], device=DEV) / d[l] ** 0.5 for l in range(4)]
        bs = [torch.randn(d[l + 1], device=DEV) * 0.3 for l in range(4)]
        wk = [ws[0]] + [w.half() if f16 else w for w in ws[1:]]
        nets.append((ws, bs, wk))
    xs = [torch.randn(M, 48, device=DEV) for _ in range(3)]
    y_ref, y_new = [torch.empty(M, d[4], device=DEV) for d in dims], [torch.empty(M, d[4], device=DEV) for d in dims]
    act_ref, lp_ref, sig_ref = torch.empty(M, A, device=DEV), torch.empty(M, device=DEV), torch.empty(M, A, device=DEV)
    act_new, lp_new, sig_new = torch.empty(M, A, device=DEV), torch.empty(M, device=DEV), torch.empty(M, A, device=DEV)
    P = nat.PREC_FP16 if f16 else nat.PREC_TF32
    mk = lambda x, i, y: mlp3_struct(x.data_ptr(), 48, M, nets[i][2], nets[i][1], None, y, nat.ACT_CODES["elu"], P)
    sa, sc = mk(xs[0], 0, y_new[0]), mk(xs[0], 1, y_new[1])
    sa.head_z, sa.head_std, sa.head_std_type = z.data_ptr(), std.data_ptr(), code
    sa.head_actions, sa.head_logp, sa.head_sigma = act_new.data_ptr(), lp_new.data_ptr(), sig_new.data_ptr()
    arr = (nat.Mlp3Fwd * 2)(sa, sc)
    handle = C.c_void_p()
    nat.check(nat.lib().rslrl_mlp3_plan_create(2, arr, C.byref(handle)))
    try:
        for x in xs + [xs[0]]:
            mlp3_forward([mk(x, 0, y_ref[0]), mk(x, 1, y_ref[1])])
            nat.check(nat.lib().rslrl_gaussian_act(y_ref[0].data_ptr(), std.data_ptr(), code, z.data_ptr(), act_ref.data_ptr(),
                                                   lp_ref.data_ptr(), sig_ref.data_ptr(), M, A, nat.stream_ptr()))
            ptrs = (C.c_void_p * 2)(x.data_ptr(), x.data_ptr())
            nat.check(nat.lib().rslrl_mlp3_plan_run(handle, ptrs, nat.stream_ptr()))
            torch.cuda.synchronize()
            for a, b in ((y_ref[0], y_new[0]), (y_ref[1], y_new[1]), (act_ref, act_new), (lp_ref, lp_new), (sig_ref, sig_new)):
                assert torch.equal(a, b)
    finally:
        nat.lib().rslrl_mlp3_plan_destroy(handle)


def test_update_rng_skip_matches_the_discarded_draws():
    """rng_compat: update() advances the Philox offset instead of launching the 20 discarded torch.normal draws of the
    reference (ppo.py:249); the generator state afterwards must be the one real draws would leave."""
    from rsl_rl_b200.algorithms.ppo import PPO
    mb, A, n = 24576, 12, 20
    gen = torch.cuda.default_generators[0]
    buf = torch.empty(mb, A, device=DEV)

    class Stub:  # the two attributes _skip_discarded_draws touches
        device, _discard, _draw_offset = DEV, buf, None

    for trial in range(2):   # trial 0 measures the per-draw offset with one real launch, trial 1 is launch-free
        torch.manual_seed(123 + trial)
        for _ in range(n):
            torch.normal(torch.zeros(mb, A, device=DEV), torch.ones(mb, A, device=DEV))
        expect_state, expect_next = gen.get_state().clone(), torch.randperm(1000, device=DEV)
        torch.manual_seed(123 + trial)
        PPO._skip_discarded_draws(Stub, n)
        assert torch.equal(gen.get_state(), expect_state)
        assert torch.equal(torch.randperm(1000, device=DEV), expect_next)
    assert Stub._draw_offset and all(v > 0 for v in Stub._draw_offset.values())


# ------------------------------------------------------------------------------------ running normaliser
@pytest.mark.parametrize("N,D,strided", [(4096, 48, False), (65536, 235, False), (777, 5, True), (4096, 1, False), (33, 300, True)])
def test_empirical_normalization_kernels(N, D, strided):
    """EmpiricalNormalization.update / forward on the fused kernels vs the reference's tensor expressions
    (normalization.py:44-66) evaluated by torch in fp64 / fp32; forward is bit-exact."""
    from rsl_rl_b200.networks import EmpiricalNormalization
    torch.manual_seed(N + D)
    norm = EmpiricalNormalization(shape=[D] if D > 1 else [], until=None).to(DEV)
    mean, var, count = torch.zeros(1, D, dtype=torch.float64), torch.ones(1, D, dtype=torch.float64), 0
    ptrs = (norm._mean.data_ptr(), norm._var.data_ptr(), norm._std.data_ptr(), norm.count.data_ptr())
    for step in range(5):
        if strided:
            x = (torch.randn(N, D + 3, device=DEV) * (1 + step) + 3.0 * step)[:, :D]
        else:
            x = torch.randn(N, D, device=DEV) * (1 + step) + 3.0 * step
        xin = x[:, 0] if D == 1 and not strided else x
        norm.update(xin)
        xd = x.double().cpu().reshape(N, D)
        count += N
        rate = np.float32(N) / np.float32(count)
        mean_x, var_x = xd.mean(0, keepdim=True), xd.var(0, unbiased=False, keepdim=True)
        delta = mean_x - mean
        mean = mean + float(rate) * delta
        var = var + float(rate) * (var_x - var + delta * (mean_x - mean))
        assert int(norm.count) == count
        assert rel_err(norm._mean.cpu().numpy().reshape(1, D), mean.numpy()) < 2e-6, step
        assert rel_err(norm._var.cpu().numpy().reshape(1, D), var.numpy()) < 2e-6, step
        assert torch.equal(norm._std, torch.sqrt(norm._var))
        y = norm(xin)
        assert torch.equal(y, (xin - norm._mean) / (norm._std + norm.eps))      # bit-exact forward
        out = torch.zeros(N, D + 5, device=DEV)
        if xin.dim() == 2:
            norm(xin, out=out)
            assert torch.equal(out[:, :D], y) and float(out[:, D:].abs().sum()) == 0.0
    assert ptrs == (norm._mean.data_ptr(), norm._var.data_ptr(), norm._std.data_ptr(), norm.count.data_ptr()), \
        "buffers must be updated in place (CUDA graphs hold their addresses)"
    # `until`: frozen once count >= until (tested on the device), eval() mode never updates
    frozen = EmpiricalNormalization(shape=[D] if D > 1 else [], until=N).to(DEV)
    x = torch.randn(N, D, device=DEV)
    frozen.update(x if D > 1 else x[:, 0])
    snap = frozen._mean.clone()
    frozen.update((x if D > 1 else x[:, 0]) + 5.0)
    assert torch.equal(frozen._mean, snap) and int(frozen.count) == N
    norm.eval()
    snap = norm._mean.clone()
    norm.update(xin)
    assert torch.equal(norm._mean, snap)


# ------------------------------------------------------------------------------------------------ LSTM
@pytest.mark.parametrize("T,n,D,H,L,prec,tol", [(24, 1500, 48, 256, 1, "fp32", 2e-5), (8, 33, 10, 16, 1, "fp32", 2e-5),
                                                (6, 40, 12, 32, 2, "fp32", 2e-5), (24, 1024, 48, 256, 1, "tf32", 2e-3)])
def test_native_lstm_matches_torch_lstm(T, n, D, H, L, prec, tol):
    """networks/lstm.py (hoisted input GEMM, per-step recurrent GEMM + fused cell kernels, hand-written BPTT) against
    torch.nn.LSTM evaluated in fp64 on the host: outputs, final states and every parameter gradient."""
    from rsl_rl_b200.networks.lstm import lstm_forward
    torch.manual_seed(T + n + H)
    rnn = torch.nn.LSTM(D, H, num_layers=L).to(DEV)
    ref = torch.nn.LSTM(D, H, num_layers=L).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in rnn.state_dict().items()})
    x = torch.randn(T, n, D, device=DEV)
    h0, c0 = torch.randn(L, n, H, device=DEV) * 0.5, torch.randn(L, n, H, device=DEV) * 0.5
    w = torch.randn(T, n, H, device=DEV)                       # loss = sum(out * w) + sum(h_n) - sum(c_n)
    out, (hn, cn) = lstm_forward(rnn, x, (h0, c0), prec)
    (out * w).sum().add(hn.sum()).sub(cn.sum()).backward()
    xr = x.cpu().double().requires_grad_(True)
    out_r, (hn_r, cn_r) = ref(xr, (h0.cpu().double(), c0.cpu().double()))
    (out_r * w.cpu().double()).sum().add(hn_r.sum()).sub(cn_r.sum()).backward()
    assert rel_err(out.detach().cpu().numpy(), out_r.detach().numpy()) < tol
    assert rel_err(hn.detach().cpu().numpy(), hn_r.detach().numpy()) < tol
    assert rel_err(cn.detach().cpu().numpy(), cn_r.detach().numpy()) < tol
    for (k, p), pr in zip(rnn.named_parameters(), ref.parameters()):
        assert rel_err(p.grad.cpu().numpy(), pr.grad.numpy()) < 5 * tol, k
    # inference path (no autograd, hidden state None -> zeros) agrees with the training path
    with torch.inference_mode():
        out_i, (hn_i, cn_i) = lstm_forward(rnn, x, None, prec)
        out_z, _ = ref(xr.detach())
    assert rel_err(out_i.cpu().numpy(), out_z.detach().numpy()) < tol


def test_native_lstm_group_of_two_matches_separate_calls():
    """lstm_forward_group (actor and critic memories advanced together: grouped GEMM and cell launches) gives the bits of
    two separate lstm_forward calls, outputs and gradients."""
    from rsl_rl_b200.networks.lstm import lstm_forward, lstm_forward_group
    torch.manual_seed(3)
    T, n, H = 12, 700, 64
    ra, rc = torch.nn.LSTM(48, H).to(DEV), torch.nn.LSTM(20, H).to(DEV)
    xa, xc = torch.randn(T, n, 48, device=DEV), torch.randn(T, n, 20, device=DEV)
    ha = (torch.randn(1, n, H, device=DEV), torch.randn(1, n, H, device=DEV))
    hc = (torch.randn(1, n, H, device=DEV), torch.randn(1, n, H, device=DEV))
    w = torch.randn(T, n, H, device=DEV)
    (oa, (hna, cna)), (oc, (hnc, cnc)) = lstm_forward_group([ra, rc], [xa, xc], [ha, hc], "fp32")
    ((oa * w).sum() + (oc * w).sum() * 0.5 + hna.sum() - cnc.sum()).backward()
    grads = [p.grad.clone() for r in (ra, rc) for p in r.parameters()]
    for r in (ra, rc):
        r.zero_grad()
    oa2, (hna2, cna2) = lstm_forward(ra, xa, ha, "fp32")
    oc2, (hnc2, cnc2) = lstm_forward(rc, xc, hc, "fp32")
    ((oa2 * w).sum() + (oc2 * w).sum() * 0.5 + hna2.sum() - cnc2.sum()).backward()
    assert torch.equal(oa, oa2) and torch.equal(oc, oc2) and torch.equal(hna, hna2) and torch.equal(cnc, cnc2)
    for g, p in zip(grads, [p for r in (ra, rc) for p in r.parameters()]):
        assert rel_err(g.cpu().numpy(), p.grad.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("n,D,hidden,out,act", [(4096, 48, [48], 1, "elu"), (1000, 10, [64, 32], 8, "relu"), (33, 128, [96, 64, 7], 5, "tanh")])
def test_rnd_reward_kernel_matches_torch(n, D, hidden, out, act):
    """rslrl_rnd_reward (both embedding networks + gap + norm + weight in one launch) against the same expression on torch
    modules in fp64 (reference rnd.py:113-135)."""
    from rsl_rl_b200.modules.rnd import RandomNetworkDistillation
    torch.manual_seed(n + D)
    rnd = RandomNetworkDistillation(D, {"rnd_state": ["policy"]}, out, hidden, hidden, activation=act, weight=0.37, device=DEV)
    x = torch.randn(n, D, device=DEV)
    got = rnd._native_gap_norm(x, 0.37)
    assert got is not None
    f = {"elu": torch.nn.functional.elu, "relu": torch.relu, "tanh": torch.tanh}[act]

    def run(mlp, h):
        layers = mlp.linear_layers()
        for i, l in enumerate(layers):
            h = h @ l.weight.double().t() + l.bias.double()
            if i < len(layers) - 1:
                h = f(h)
        return h

    with torch.no_grad():
        ref = 0.37 * torch.linalg.norm(run(rnd.target, x.double()) - run(rnd.predictor, x.double()), dim=1)
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    obs = {"policy": x}
    r2 = rnd.get_intrinsic_reward(obs)
    assert torch.equal(r2, got)
