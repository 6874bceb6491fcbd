"""GPU parity of the CelebA-flavour building blocks (segment-wise train-mode BatchNorm fwd/bwd + running stats, Dropout,
k4 conv data movement with stride/pad, NCHW->NHWC, narrow-row BCE) against torch CPU (fp64)."""
import numpy as np
import pytest
import torch

from oracle import mvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from multimodal_vae_public_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("S,seg_rows,C,act", [(3, 200, 64, True), (1, 37, 512, True), (2, 1000, 32, False)])
def test_batchnorm_train_forward_backward(ops, S, seg_rows, C, act):
    rs = np.random.RandomState(S * 7 + C)
    R = S * seg_rows
    x = torch.from_numpy((rs.standard_normal((R, C)) * 2 + 0.5).astype(np.float32))
    gamma = torch.from_numpy(rs.uniform(0.5, 1.5, C).astype(np.float32)); beta = torch.from_numpy(rs.standard_normal(C).astype(np.float32))
    rm0 = torch.from_numpy(rs.standard_normal(C).astype(np.float32)); rv0 = torch.from_numpy(rs.uniform(0.5, 2, C).astype(np.float32))
    dh = torch.from_numpy(rs.standard_normal((R, C)).astype(np.float32))
    order = list(range(S))[::-1] + [0]           # arbitrary call order, segment 0 "called" twice
    # ---- reference: one nn.functional.batch_norm call per segment, in `order`, fp64
    xr = x.double().requires_grad_(True); g64 = gamma.double().requires_grad_(True); b64 = beta.double().requires_grad_(True)
    rm, rv = rm0.double().clone(), rv0.double().clone()
    outs = {}
    for s in order:
        seg = xr[s * seg_rows:(s + 1) * seg_rows]
        y = torch.nn.functional.batch_norm(seg, rm, rv, g64, b64, True, 0.1, 1e-5)
        outs[s] = O.swish(y) if act else y
    live0, nlive = (1, S - 1) if S > 1 else (0, 1)            # backward only through the "live" segments
    loss = sum((outs[s] * dh[s * seg_rows:(s + 1) * seg_rows].double()).sum() for s in range(live0, live0 + nlive))
    loss.backward()
    # ---- kernels
    dev = "cuda"
    xd, hd = x.to(dev), torch.empty(R, C, device=dev)
    mean, invstd = torch.empty(S, C, device=dev), torch.empty(S, C, device=dev)
    acc = torch.empty(S, C, 2, dtype=torch.float64, device=dev)
    rmd, rvd = rm0.to(dev).clone(), rv0.to(dev).clone()
    ops.bn_forward(xd, hd, S, seg_rows, gamma.to(dev), beta.to(dev), mean, invstd, acc, rmd, rvd, update_order=order,
                   training=True, act=act)
    for s in range(S):
        np.testing.assert_allclose(hd[s * seg_rows:(s + 1) * seg_rows].cpu().numpy(), outs[s].detach().numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(rmd.cpu().numpy(), rm.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rvd.cpu().numpy(), rv.numpy(), rtol=1e-5, atol=1e-6)
    dx = torch.zeros(R, C, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    ops.bn_backward(xd, dh.to(dev), dx, S, seg_rows, live0, nlive, gamma.to(dev), beta.to(dev), mean, invstd, acc, dg, db, act=act)
    lo, hi = live0 * seg_rows, (live0 + nlive) * seg_rows
    ref_dx = xr.grad[lo:hi]
    scale = ref_dx.abs().max().item()
    assert (dx[lo:hi].cpu().double() - ref_dx).abs().max().item() <= 3e-4 * scale
    np.testing.assert_allclose(dg.cpu().numpy(), g64.grad.numpy(), rtol=3e-4, atol=3e-4 * g64.grad.abs().max().item())
    np.testing.assert_allclose(db.cpu().numpy(), b64.grad.numpy(), rtol=3e-4, atol=3e-4 * b64.grad.abs().max().item())
    # ---- eval mode uses the running statistics
    he = torch.empty(R, C, device=dev)
    ops.bn_forward(xd, he, S, seg_rows, gamma.to(dev), beta.to(dev), mean, invstd, acc, rmd, rvd, training=False, act=act)
    ye = torch.nn.functional.batch_norm(x.double(), rm, rv, gamma.double(), beta.double(), False, 0.1, 1e-5)
    np.testing.assert_allclose(he.cpu().numpy(), (O.swish(ye) if act else ye).numpy(), rtol=2e-4, atol=2e-5)


def test_dropout_masks_and_backward(ops):
    B, D, p = 512, 512, 0.1
    x = torch.randn(B, D, device="cuda")
    y = torch.empty(2 * B, D, device="cuda"); mask = torch.empty(2 * B, D, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.dropout_fwd(x, y, 2, p, mask_out=mask, seed=5, step_dev=step)
    assert set(mask.unique().tolist()) <= {0.0, 1.0}
    assert abs(mask.mean().item() - (1 - p)) < 5e-3
    assert not torch.equal(mask[:B], mask[B:])                      # fresh mask per stacked call
    assert torch.allclose(y, torch.cat([x, x]) * mask / (1 - p))
    y2 = torch.empty_like(y)
    ops.dropout_fwd(x, y2, 2, p, mask_in=mask)                       # injected mask (parity tests)
    assert torch.equal(y, y2)
    dy = torch.randn(2 * B, D, device="cuda"); dx = torch.empty(B, D, device="cuda")
    ops.dropout_bwd(dy, mask, dx, 2, p)
    assert torch.allclose(dx, ((dy * mask)[:B] + (dy * mask)[B:]) / (1 - p), atol=1e-6)
    step += 1
    mask2 = torch.empty_like(mask)
    ops.dropout_fwd(x, y2, 2, p, mask_out=mask2, seed=5, step_dev=step)
    assert not torch.equal(mask, mask2)


@pytest.mark.parametrize("stride,pad,H,C", [(1, 0, 8, 128), (2, 1, 16, 32), (2, 1, 64, 3)])
def test_general_k4_conv_data_movement(ops, stride, pad, H, C):
    B = 2
    g = torch.Generator().manual_seed(stride * 10 + H)
    x = torch.randn(B, H, H, C, generator=g)
    OH = (H + 2 * pad - 4) // stride + 1
    cols = torch.empty(B * OH * OH, 16 * C, device="cuda")
    ops.im2col_k4(x.cuda().contiguous(), cols, B, H, H, C, stride, pad)
    Co = 6
    w = torch.randn(Co, C, 4, 4, generator=g)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, stride, pad)
    got = (cols.cpu().double() @ w.permute(0, 2, 3, 1).reshape(Co, -1).double().t()).reshape(B, OH, OH, Co).permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() < 1e-9
    # transpose: [B,OH,OH,Cin] -> [B,H,H,C]
    Cin = 5
    x2 = torch.randn(B, OH, OH, Cin, generator=g); wt = torch.randn(Cin, C, 4, 4, generator=g)
    reft = torch.nn.functional.conv_transpose2d(x2.permute(0, 3, 1, 2).double(), wt.double(), None, stride, pad)
    assert reft.shape[-1] == H
    colsT = (x2.reshape(-1, Cin).double() @ wt.permute(2, 3, 1, 0).reshape(16 * C, Cin).double().t()).float().cuda().contiguous()
    out = torch.empty(B, H, H, C, device="cuda")
    ops.col2im_k4(colsT, out, B, OH, OH, C, stride, pad)
    assert (out.cpu().double().permute(0, 3, 1, 2) - reft).abs().max().item() < 1e-4


def test_nchw_to_nhwc_and_narrow_bce(ops):
    x = torch.randn(3, 3, 64 * 64, device="cuda"); y = torch.empty(3, 64 * 64, 3, device="cuda")
    ops.nchw_to_nhwc(x, y, 3, 3, 64 * 64)
    assert torch.equal(y, x.permute(0, 2, 1).contiguous())
    rs = np.random.RandomState(1)
    lg = torch.from_numpy((2 * rs.standard_normal((200, 20))).astype(np.float32)).cuda()[:, :18]
    t = torch.from_numpy(rs.randint(0, 2, (100, 18)).astype(np.float32)).cuda()
    acc = torch.zeros(2, dtype=torch.float64, device="cuda"); dx = torch.zeros(200, 20, device="cuda")
    ops.bce_logits_fwd_bwd(lg, t, dx[:, :18], 0.7, acc, seg_rows=100)
    for s in range(2):
        ref = O.bce_with_logits(lg[s * 100:(s + 1) * 100].cpu().double(), t.cpu().double()).sum().item()
        assert abs(acc[s].item() - ref) <= 2e-6 * abs(ref)
    ref_dx = 0.7 * (torch.sigmoid(lg.cpu().double()) - torch.cat([t, t]).cpu().double())
    np.testing.assert_allclose(dx[:, :18].cpu().numpy(), ref_dx.numpy(), rtol=2e-5, atol=1e-6)
    assert torch.all(dx[:, 18:] == 0)
