"""GPU parity tests of the individual kernels (through the C ABI) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import elementwise_np as EN
from oracle import mvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from multimodal_vae_public_b200 import ops as _ops
    return _ops


def _rel(a, ref):
    return (a.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


# tolerance stated: max|out - ref| / max|ref| against an fp64 reference.  TF32 = 10-bit mantissa operands.
# 3xTF32 = fp32-class products; the residual (measured 3e-7 at K=32 .. 7e-6 at K=784) is the tensor core's
# truncating fp32 accumulation over K/8 MMA steps -- the same error class as any fp32 GEMM with a different
# summation order (an fp32 FMA chain has ~1e-6 here).
TOL = {0: 3e-3, 1: 2e-5}


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (512, 512, 784), (300, 10, 512), (1000, 128, 512), (8, 512, 64),
                                   (4096, 784, 512), (1024, 48, 32), (640, 16, 64), (256, 176, 96)])
def test_linear_fwd(ops, prec, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g); w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    ld = (N + 3) // 4 * 4
    y = torch.full((M, ld), float("nan"), device="cuda")[:, :N]; h = torch.full((M, ld), float("nan"), device="cuda")[:, :N]
    ops.linear_fwd(x, w, b, y, h, precision=prec)
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(y, ref) < TOL[prec]
    assert _rel(h, ref * torch.sigmoid(ref)) < TOL[prec]
    y2 = torch.full((M, ld), float("nan"), device="cuda")[:, :N]
    ops.linear_fwd(x, w, None, y2, None, precision=prec)
    assert _rel(y2, x.double() @ w.double().t()) < TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (1024, 784, 512), (512, 10, 512), (512, 512, 64), (40, 128, 512)])
def test_linear_dgrad(ops, prec, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    ld = (N + 3) // 4 * 4
    dy = torch.randn(M, ld, device="cuda", generator=g)[:, :N]; w = torch.randn(N, K, device="cuda", generator=g) / N ** 0.5
    a = torch.randn(M, K, device="cuda", generator=g)
    dx = torch.full((M, K), float("nan"), device="cuda")
    ops.linear_dgrad(dy, w, dx, a_prev=a, precision=prec)
    s = torch.sigmoid(a.double())
    plain = dy.double() @ w.double()
    assert _rel(dx, plain * (s * (1 + a.double() * (1 - s)))) < TOL[prec]
    dx2 = torch.ones(M, K, device="cuda")
    ops.linear_dgrad(dy, w, dx2, accumulate=True, precision=prec)
    assert _rel(dx2, plain + 1.0) < TOL[prec]
    # fused bias-gradient column sums of the produced dA
    dx3 = torch.empty(M, K, device="cuda"); cs = torch.full((K,), 0.5, device="cuda")
    ops.gemm_batch([ops.gemm_desc(dy, w, dx3, M, K, N, b_mn=True, aux=a, epilogue=ops.EPI_MUL_DSWISH, colsum=cs)], prec)
    ref3 = plain * (s * (1 + a.double() * (1 - s)))
    assert _rel(dx3, ref3) < TOL[prec]
    assert _rel(cs, ref3.sum(0) + 0.5) < 5 * TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M,N,K,split", [(32, 128, 128, 1), (2048, 512, 784, 5), (1000, 10, 512, 3), (4096, 512, 64, 8),
                                         (4096, 256, 48, 4), (65536, 48, 32, 64),
                                         (8192, 128, 512, 16)])
def test_linear_wgrad(ops, prec, M, N, K, split):
    g = torch.Generator(device="cuda").manual_seed(M + N * 7 + K)
    ld = (N + 3) // 4 * 4
    dy = torch.randn(M, ld, device="cuda", generator=g)[:, :N]; x = torch.randn(M, K, device="cuda", generator=g)
    dw = torch.zeros(N, K, device="cuda")
    ops.linear_wgrad(dy, x, dw, split_k=split, precision=prec)
    ref = dy.double().t() @ x.double()
    assert _rel(dw, ref) < TOL[prec]
    ops.linear_wgrad(dy, x, dw, split_k=split, precision=prec)  # accumulates
    assert _rel(dw, 2 * ref) < TOL[prec]


def test_gemm_batch_mixed(ops):
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 700, 512, 512
    xs = [torch.randn(M, K, device="cuda", generator=g) for _ in range(3)]
    ws = [torch.randn(N, K, device="cuda", generator=g) / 23 for _ in range(3)]
    ys = [torch.empty(M, N, device="cuda") for _ in range(3)]
    dw = torch.zeros(N, K, device="cuda")
    descs = [ops.gemm_desc(xs[i], ws[i], ys[i], M, N, K) for i in range(3)]
    descs.append(ops.gemm_desc(ys[0], xs[0], dw, N, K, M, a_mn=True, b_mn=True, split_k=4, accumulate=True))
    ys[0].copy_(torch.randn(M, N, device="cuda", generator=g))  # the wgrad reads ys[0] while problem 0 writes it: use separate buffer
    y0_in = ys[0].clone()
    descs[3] = ops.gemm_desc(y0_in, xs[0], dw, N, K, M, a_mn=True, b_mn=True, split_k=4, accumulate=True)
    ops.gemm_batch(descs, 1)
    for i in range(3):
        assert _rel(ys[i], xs[i].double() @ ws[i].double().t()) < TOL[1]
    assert _rel(dw, y0_in.double().t() @ xs[0].double()) < TOL[1]


def test_colsum_swish_embedding(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    dy = torch.randn(1000, 16, device="cuda", generator=g)[:, :10]
    db = torch.ones(10, device="cuda")
    ops.colsum_accumulate(dy, db)
    assert _rel(db, dy.double().sum(0) + 1) < 1e-6
    dy = torch.randn(4100, 512, device="cuda", generator=g); db = torch.zeros(512, device="cuda")
    ops.colsum_accumulate(dy, db)
    assert _rel(db, dy.double().sum(0)) < 2e-6
    x = 3 * torch.randn(1003, device="cuda", generator=g); y = torch.empty_like(x); dx = torch.empty_like(x)
    ops.swish_fwd(x, y); ops.swish_bwd(x, torch.full_like(x, 2.0), dx)
    np.testing.assert_allclose(y.cpu().numpy(), EN.swish(x.cpu().double().numpy()), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(dx.cpu().numpy(), 2 * EN.swish_grad(x.cpu().double().numpy()), rtol=2e-6, atol=1e-6)
    table = torch.randn(10, 512, device="cuda", generator=g); idx = torch.randint(0, 10, (777,), device="cuda", generator=g)
    a = torch.empty(777, 512, device="cuda"); h = torch.empty(777, 512, device="cuda")
    ops.embedding_swish_fwd(table, idx, a, h)
    assert torch.equal(a, table[idx])
    np.testing.assert_allclose(h.cpu().numpy(), EN.swish(table[idx].cpu().double().numpy()), rtol=2e-6, atol=1e-7)
    dh = torch.randn(777, 512, device="cuda", generator=g); dt = torch.zeros(10, 512, device="cuda")
    ops.embedding_swish_bwd(table, idx, dh, dt)
    tt = table.double().cpu().requires_grad_(True)
    (O.swish(tt[idx.cpu()]) * dh.double().cpu()).sum().backward()
    assert _rel(dt.cpu(), tt.grad) < 5e-6


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("E,B,L,masks", [(2, 300, 64, [0b01, 0b11, 0b10]), (3, 17, 100, [0b111, 0b001, 0b110, 0b010]),
                                         (19, 9, 100, [(1 << 19) - 1, 1, 1 << 7, 0b1010101, 1 << 18])])
@pytest.mark.parametrize("training", [True, False])
def test_poe_fwd_bwd(ops, variant, E, B, L, masks, training):
    rs = np.random.RandomState(E + B)
    P = len(masks)
    enc = [torch.from_numpy(rs.standard_normal((B, 2 * L)).astype(np.float32) * np.float32(0.8)).cuda() for _ in range(E)]
    mu_e = [e[:, :L] for e in enc]; lv_e = [e[:, L:] for e in enc]
    noise = torch.from_numpy(rs.standard_normal((P * B, L)).astype(np.float32)).cuda()
    z = torch.empty(P * B, L, device="cuda"); mu_o = torch.empty(P * B, L, device="cuda"); lv_o = torch.empty(P * B, L, device="cuda")
    kl = torch.zeros(P, dtype=torch.float64, device="cuda")
    ops.poe_fwd(mu_e, lv_e, masks, B, L, z, variant=variant, training=training, noise=noise if training else None,
                mu_out=mu_o, lv_out=lv_o, kl_acc=kl)
    vn = "AB"[variant]
    mu_np = np.stack([m.cpu().double().numpy() for m in mu_e]); lv_np = np.stack([m.cpu().double().numpy() for m in lv_e])
    nz_np = noise.cpu().double().numpy().reshape(P, B, L) if training else None
    beta = 0.3
    rmu, rlv, rz, rkl = EN.poe_multipass_fwd(mu_np, lv_np, masks, nz_np, beta, vn)
    np.testing.assert_allclose(mu_o.cpu().numpy().reshape(P, B, L), rmu, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(lv_o.cpu().numpy().reshape(P, B, L), rlv, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(z.cpu().numpy().reshape(P, B, L), rz, rtol=2e-5, atol=3e-6)
    np.testing.assert_allclose(kl.cpu().numpy() * beta / B, rkl, rtol=1e-5, atol=1e-6)
    dz = torch.from_numpy(rs.standard_normal((P * B, L)).astype(np.float32)).cuda()
    d_enc = [torch.full((B, 2 * L), float("nan"), device="cuda") for _ in range(E)]
    bdev = torch.tensor([0.5], device="cuda")
    ops.poe_bwd(mu_e, lv_e, masks, B, L, dz, [d[:, :L] for d in d_enc], [d[:, L:] for d in d_enc], kl_scale=2 * beta / B,
                variant=variant, training=training, noise=noise if training else None, kl_scale_dev=bdev)
    rdmu, rdlv = EN.poe_multipass_bwd(mu_np, lv_np, masks, nz_np, beta, dz.cpu().double().numpy().reshape(P, B, L), vn)
    for e in range(E):
        np.testing.assert_allclose(d_enc[e][:, :L].cpu().numpy(), rdmu[e], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(d_enc[e][:, L:].cpu().numpy(), rdlv[e], rtol=1e-4, atol=2e-5)


def test_poe_philox_noise_is_standard_normal(ops):
    B, L = 4096, 64
    enc = [torch.zeros(B, 2 * L, device="cuda") for _ in range(2)]
    z = torch.empty(3 * B, L, device="cuda"); nz = torch.empty(3 * B, L, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    args = dict(variant=0, training=True, noise=None, noise_out=nz, seed=7, step_dev=step)
    ops.poe_fwd([e[:, :L] for e in enc], [e[:, L:] for e in enc], [1, 3, 2], B, L, z, **args)
    a = nz.clone()
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1) < 5e-3
    assert abs((a ** 4).mean().item() - 3) < 0.1
    ops.poe_fwd([e[:, :L] for e in enc], [e[:, L:] for e in enc], [1, 3, 2], B, L, z, **args)
    assert torch.equal(a, nz)  # deterministic for the same (seed, step)
    step += 1
    ops.poe_fwd([e[:, :L] for e in enc], [e[:, L:] for e in enc], [1, 3, 2], B, L, z, **args)
    assert not torch.equal(a, nz) and abs((a * nz).mean().item()) < 5e-3


@pytest.mark.parametrize("R,D,t_rows,seg", [(16, 784, 8, 8), (600, 784, 300, 300), (8, 12288, 8, 0), (5, 16, 5, 1)])
def test_bce(ops, R, D, t_rows, seg):
    rs = np.random.RandomState(R + D)
    x = torch.from_numpy((4 * rs.standard_normal((R, D))).astype(np.float32)).cuda()
    t = torch.from_numpy(rs.uniform(0, 1, (t_rows, D)).astype(np.float32)).cuda()
    nseg = 1 if seg == 0 else (R + seg - 1) // seg
    acc = torch.zeros(nseg, dtype=torch.float64, device="cuda")
    dx = torch.empty_like(x)
    ops.bce_logits_fwd_bwd(x, t, dx, 0.37, acc, seg_rows=seg)
    xn = x.cpu().double().numpy(); tn = np.tile(t.cpu().double().numpy(), (R // t_rows, 1))
    for s in range(nseg):
        rows = slice(0, R) if seg == 0 else slice(s * seg, min(R, (s + 1) * seg))
        loss, _ = EN.bce_logits_fwd_bwd(xn[rows], tn[rows], 1.0)
        assert abs(acc[s].item() - loss) <= 2e-6 * abs(loss)
    _, rdx = EN.bce_logits_fwd_bwd(xn, tn, 0.37)
    np.testing.assert_allclose(dx.cpu().numpy(), rdx, rtol=2e-5, atol=1e-7)
    ops.bce_logits_fwd_bwd(x, t, x, 0.37, None, seg_rows=seg)  # in place, no loss
    np.testing.assert_allclose(x.cpu().numpy(), rdx, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("R,t_rows,seg", [(16, 8, 8), (1000, 500, 500), (7, 7, 0)])
def test_ce(ops, R, t_rows, seg):
    rs = np.random.RandomState(R)
    xb = torch.from_numpy((3 * rs.standard_normal((R, 16))).astype(np.float32)).cuda()
    x = xb[:, :10]
    tg = torch.from_numpy(rs.randint(0, 10, t_rows)).cuda()
    nseg = 1 if seg == 0 else (R + seg - 1) // seg
    acc = torch.zeros(nseg, dtype=torch.float64, device="cuda")
    dxb = torch.zeros(R, 16, device="cuda")
    ops.ce_fwd_bwd(x, tg, dxb[:, :10], 10, 1.3, acc, seg_rows=seg)
    xn = x.cpu().double().numpy(); tn = np.tile(tg.cpu().numpy(), R // t_rows)
    for s in range(nseg):
        rows = slice(0, R) if seg == 0 else slice(s * seg, min(R, (s + 1) * seg))
        loss, _ = EN.ce_fwd_bwd(xn[rows], tn[rows], 1.0)
        assert abs(acc[s].item() - loss) <= 2e-6 * abs(loss)
    _, rdx = EN.ce_fwd_bwd(xn, tn, 1.3)
    np.testing.assert_allclose(dxb[:, :10].cpu().numpy(), rdx, rtol=2e-5, atol=1e-6)  # softmax - onehot cancels near 1
    assert torch.all(dxb[:, 10:] == 0)


def test_adam_matches_oracle(ops):
    rs = np.random.RandomState(0)
    n = 10007
    p0 = rs.standard_normal(n).astype(np.float32); p = torch.from_numpy(p0.copy()).cuda()
    n_pad = (n + 3) // 4 * 4
    buf = torch.zeros(4, n_pad, device="cuda")
    P, G, M, V = buf[0, :n], buf[1, :n], buf[2, :n], buf[3, :n]
    P.copy_(p)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ref = {"w": torch.from_numpy(p0.copy())}; st = {}
    for it in range(1, 4):
        g = rs.standard_normal(n).astype(np.float32) * 0.01
        G.copy_(torch.from_numpy(g))
        ops.adam_flat(P, G, M, V, step, lr=1e-3)
        O.adam_update(ref, {"w": torch.from_numpy(g)}, st, step=it, lr=1e-3)
        assert step.item() == it
        np.testing.assert_allclose(P.cpu().numpy(), ref["w"].numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("V,D,N3", [(10, 512, 128), (2, 512, 200), (16, 64, 8)])
def test_label_table_fwd_bwd(ops, V, D, N3):
    """Label encoder on its V-row class table (Embedding -> Swish -> Linear+Swish -> heads) against fp64 autograd."""
    rs = np.random.RandomState(V + D)
    t = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))  # noqa: E731
    emb, w2, b2, w3, b3, dtab = t(V, D), t(D, D) / D ** 0.5, t(D), t(N3, D) / D ** 0.5, t(N3), t(V, N3)
    dev = [x.cuda() for x in (emb, w2, b2, w3, b3, dtab)]
    a2 = torch.empty(V, D, device="cuda"); h2 = torch.empty(V, D, device="cuda"); tab = torch.empty(V, N3, device="cuda")
    ops.label_table_fwd(dev[0], dev[1], dev[2], dev[3], dev[4], a2, h2, tab)
    e64, w2_64, b2_64, w3_64, b3_64 = (x.double().requires_grad_(True) for x in (emb, w2, b2, w3, b3))
    ra2 = O.swish(e64) @ w2_64.t() + b2_64
    rtab = O.swish(ra2) @ w3_64.t() + b3_64
    assert _rel(a2.cpu(), ra2.detach()) < 2e-6 and _rel(h2.cpu(), O.swish(ra2).detach()) < 2e-6
    assert _rel(tab.cpu(), rtab.detach()) < 2e-6
    (rtab * dtab.double()).sum().backward()
    # gradients ACCUMULATE: start from a non-zero value to check that
    base = 0.25
    g = {k: torch.full(s, base, device="cuda") for k, s in (("emb", (V, D)), ("w2", (D, D)), ("b2", (D,)), ("w3", (N3, D)), ("b3", (N3,)))}
    d_a2 = torch.zeros(V, D, device="cuda")        # accumulation target: zero-initialised by the caller
    ops.label_table_bwd(dev[0], dev[1], dev[3], a2, h2, dev[5], d_a2, g["emb"], g["w2"], g["b2"], g["w3"], g["b3"])
    for k, ref in (("emb", e64.grad), ("w2", w2_64.grad), ("b2", b2_64.grad), ("w3", w3_64.grad), ("b3", b3_64.grad)):
        assert _rel(g[k].cpu() - base, ref) < 5e-6, k


@pytest.mark.parametrize("variant", [0, 1])
def test_poe_table_expert_equals_per_sample_expert(ops, variant):
    """An expert given as a V-row table + row index per sample (mvae_poe_*_g) == the same expert expanded to [B, 2L]; its
    backward is the class-wise sum of the per-sample gradients."""
    rs = np.random.RandomState(3)
    B, L, V = 700, 64, 10
    masks = [0b01, 0b11, 0b10]
    P = len(masks)
    enc_i = torch.from_numpy((0.8 * rs.standard_normal((B, 2 * L))).astype(np.float32)).cuda()
    tab = torch.from_numpy((0.8 * rs.standard_normal((V, 2 * L))).astype(np.float32)).cuda()
    idx = torch.from_numpy(rs.randint(0, V, B)).cuda()
    enc_t = tab[idx].contiguous()
    noise = torch.from_numpy(rs.standard_normal((P * B, L)).astype(np.float32)).cuda()
    dz = torch.from_numpy(rs.standard_normal((P * B, L)).astype(np.float32)).cuda()
    out = {}
    for mode in ("table", "rows"):
        et = tab if mode == "table" else enc_t
        gather = [None, idx] if mode == "table" else None
        z = torch.empty(P * B, L, device="cuda"); kl = torch.zeros(P, dtype=torch.float64, device="cuda")
        ops.poe_fwd([enc_i[:, :L], et[:, :L]], [enc_i[:, L:], et[:, L:]], masks, B, L, z, variant=variant, training=True,
                    noise=noise, kl_acc=kl, gather=gather)
        d_i = torch.empty(B, 2 * L, device="cuda"); d_t = torch.zeros_like(et)
        ops.poe_bwd([enc_i[:, :L], et[:, :L]], [enc_i[:, L:], et[:, L:]], masks, B, L, dz, [d_i[:, :L], d_t[:, :L]],
                    [d_i[:, L:], d_t[:, L:]], kl_scale=0.3 / B, variant=variant, training=True, noise=noise, gather=gather)
        out[mode] = (z, kl, d_i, d_t)
    assert torch.equal(out["table"][0], out["rows"][0]) and torch.equal(out["table"][2], out["rows"][2])
    np.testing.assert_allclose(out["table"][1].cpu().numpy(), out["rows"][1].cpu().numpy(), rtol=1e-12)
    seg = torch.zeros(V, 2 * L, dtype=torch.float64, device="cuda").index_add_(0, idx, out["rows"][3].double())
    assert _rel(out["table"][3].cpu(), seg.cpu()) < 2e-6
