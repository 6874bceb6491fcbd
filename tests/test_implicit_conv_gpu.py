"""Implicit-GEMM convolution operands (mvae_conv_view: TMA im2col-mode loads feed the tcgen05 GEMM straight from the NHWC
activation, no cols buffer) against torch's fp64 conv2d / autograd on the CPU -- the ops behind nn.Conv2d /
nn.ConvTranspose2d of fashionmnist/model.py:80-82,112-113 and celeba/model.py:78-87,117-125."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = {0: 3e-3, 1: 2e-5}


@pytest.fixture(scope="module")
def ops():
    from multimodal_vae_public_b200 import ops as _ops
    return _ops


def _rel(a, ref):
    return (a.double().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def _case(rs, B, H, C, Cout):
    x = torch.from_numpy(rs.standard_normal((B, C, H, H)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((Cout, C, 4, 4)) / (16 * C) ** 0.5).astype(np.float32))
    return x, w


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("B,H,C,Cout", [(3, 14, 64, 128), (37, 14, 64, 128), (2, 32, 32, 64), (5, 16, 64, 128), (300, 14, 64, 128)])
def test_conv_forward_and_weight_gradient(ops, prec, B, H, C, Cout):
    """Conv2d(k4 s2 p1): y = im2col(x) W^T with A = the im2col view (K-major); dW = dy^T im2col(x) with B = the view
    (MN-major).  OW = 7 makes every 128-pixel tile straddle rows and images; the last tile runs past the last image."""
    rs = np.random.RandomState(B + H + C)
    x, w = _case(rs, B, H, C, Cout)
    OH = H // 2
    M, K = B * OH * OH, 16 * C
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wc = w.permute(0, 2, 3, 1).reshape(Cout, K).contiguous().cuda()              # [Cout][(kh,kw,ci)]
    view = ops.conv_view(B, H, H, C)
    y = torch.full((M, Cout), float("nan"), device="cuda"); h = torch.full((M, Cout), float("nan"), device="cuda")
    ops.gemm_batch([ops.gemm_desc(x_nhwc, wc, y, M, Cout, K, out2=h, epilogue=ops.EPI_BIAS_SWISH, a_view=view)], prec)
    w64 = w.double().requires_grad_(True)
    ref = F.conv2d(x.double(), w64, stride=2, padding=1).permute(0, 2, 3, 1).reshape(M, Cout)
    assert _rel(y, ref.detach()) < TOL[prec]
    assert _rel(h, (ref * torch.sigmoid(ref)).detach()) < TOL[prec]
    dy = torch.from_numpy(rs.standard_normal((M, Cout)).astype(np.float32))
    (ref * dy.double()).sum().backward()
    dw = torch.zeros(Cout, K, device="cuda")
    split = max(1, min(M // 512, 64))
    ops.gemm_batch([ops.gemm_desc(dy.cuda(), x_nhwc, dw, Cout, K, M, a_mn=True, b_mn=True, split_k=split, accumulate=True,
                                  b_view=view)], prec)
    assert _rel(dw, w64.grad.permute(0, 2, 3, 1).reshape(Cout, K)) < TOL[prec] * 2


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("B,IH,Cin,Cout", [(3, 7, 128, 64), (29, 7, 128, 64), (2, 16, 64, 32), (260, 7, 128, 64)])
def test_conv_transpose_backward(ops, prec, B, IH, Cin, Cout):
    """ConvTranspose2d(k4 s2 p1) backward: d x = im2col(dy) Wt (A = the view of dy, K-major, B MN-major) with the fused
    Swish' epilogue, dWt = im2col(dy)^T x (A = the view, MN-major)."""
    rs = np.random.RandomState(B + IH)
    a = torch.from_numpy(rs.standard_normal((B, Cin, IH, IH)).astype(np.float32))          # pre-activation below the layer
    w = torch.from_numpy((rs.standard_normal((Cin, Cout, 4, 4)) / (16 * Cin) ** 0.5).astype(np.float32))
    a64 = a.double().requires_grad_(True); w64 = w.double().requires_grad_(True)
    y = F.conv_transpose2d(a64 * torch.sigmoid(a64), w64, stride=2, padding=1)             # [B,Cout,2IH,2IH]
    dy = torch.from_numpy(rs.standard_normal((B, 2 * IH, 2 * IH, Cout)).astype(np.float32))
    (y.permute(0, 2, 3, 1) * dy.double()).sum().backward()
    P_in = B * IH * IH
    K = 16 * Cout
    wt = w.permute(2, 3, 1, 0).reshape(K, Cin).contiguous().cuda()                         # [(kh,kw,co)][ci]
    a_rows = a.permute(0, 2, 3, 1).reshape(P_in, Cin).contiguous().cuda()
    h_rows = (a_rows.double() * torch.sigmoid(a_rows.double())).float()
    view = ops.conv_view(B, 2 * IH, 2 * IH, Cout)
    dyd = dy.cuda()
    dx = torch.full((P_in, Cin), float("nan"), device="cuda")
    dwt = torch.zeros(K, Cin, device="cuda")
    split = max(1, min(P_in // 512, 64))
    ops.gemm_batch([ops.gemm_desc(dyd, wt, dx, P_in, Cin, K, b_mn=True, aux=a_rows, epilogue=ops.EPI_MUL_DSWISH, a_view=view),
                    ops.gemm_desc(dyd, h_rows, dwt, K, Cin, P_in, a_mn=True, b_mn=True, split_k=split, accumulate=True,
                                  a_view=view)], prec)
    assert _rel(dx, a64.grad.permute(0, 2, 3, 1).reshape(P_in, Cin)) < TOL[prec] * 2
    assert _rel(dwt, w64.grad.permute(2, 3, 1, 0).reshape(K, Cin)) < TOL[prec] * 2


def test_view_must_match_the_problem(ops):
    from multimodal_vae_public_b200 import _lib
    x = torch.randn(2, 14, 14, 64, device="cuda"); w = torch.randn(128, 1024, device="cuda"); y = torch.empty(98, 128, device="cuda")
    with pytest.raises(_lib.MvaeError):       # M does not equal the number of output pixels
        ops.gemm_batch([ops.gemm_desc(x, w, y, 97, 128, 1024, a_view=ops.conv_view(2, 14, 14, 64))], 1)
    with pytest.raises(_lib.MvaeError):       # C % 32 != 0
        ops.gemm_batch([ops.gemm_desc(x, w, y, 98, 128, 16 * 48, a_view=ops.conv_view(2, 14, 14, 48))], 1)


def test_implicit_operands_equal_materialised_im2col(monkeypatch):
    """The FashionMNIST trainer with implicit conv operands (default) == with im2col buffers in HBM."""
    from multimodal_vae_public_b200.trainer_fashion import FashionMVAETrainer
    B, L = 80, 64
    rs = np.random.RandomState(6)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32))
    a = FashionMVAETrainer(n_latents=L, batch_size=B, use_graph=False)
    monkeypatch.setenv("MVAE_IMPLICIT_CONV", "0")
    b = FashionMVAETrainer(n_latents=L, batch_size=B, use_graph=False)
    assert a.implicit_conv and not b.implicit_conv
    b.load_state_dict(a.state_dict())
    la = a.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    lb = b.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    assert abs(la - lb) <= 2e-6 * abs(la)
    for k in a.grads:
        err = (a.grads[k] - b.grads[k]).abs().max().item() / max(b.grads[k].abs().max().item(), 1e-12)
        assert err <= 1e-4, (k, err)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("B,IH,Cin,Cout", [(3, 7, 128, 64), (40, 7, 128, 64), (2, 16, 64, 32), (300, 7, 128, 64)])
def test_subpixel_conv_transpose_forward(ops, prec, B, IH, Cin, Cout):
    """ConvTranspose2d(k4 s2 p1) + Swish as four sub-pixel implicit GEMMs (2x2 stride-1 views, tap-split B, row-mapped
    stores of pre-activation and activation): no cols buffer, no col2im."""
    rs = np.random.RandomState(B + IH + 1)
    x = torch.from_numpy(rs.standard_normal((B, Cin, IH, IH)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((Cin, Cout, 4, 4)) / (4 * Cin) ** 0.5).astype(np.float32))
    ref = F.conv_transpose2d(x.double(), w.double(), stride=2, padding=1).permute(0, 2, 3, 1)      # [B,2IH,2IH,Cout]
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wt = w.permute(2, 3, 1, 0).reshape(16 * Cout, Cin).contiguous().cuda()                       # [(kh,kw,co)][ci]
    P_out = B * 4 * IH * IH
    a = torch.full((P_out, Cout), float("nan"), device="cuda"); h = torch.full((P_out, Cout), float("nan"), device="cuda")
    descs = ops.subpixel_k4s2p1(x_nhwc, wt, a, B, IH, IH, Cin, Cout, out2=h, epilogue=ops.EPI_BIAS_SWISH)
    ops.gemm_chain(descs, [-1] * 4, ops.chain_workspace("cuda"), prec)
    assert _rel(a, ref.reshape(P_out, Cout)) < TOL[prec]
    assert _rel(h, (ref * torch.sigmoid(ref)).reshape(P_out, Cout)) < TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("B,H,C,Cout", [(3, 14, 64, 128), (33, 14, 64, 128), (2, 32, 32, 64), (280, 14, 64, 128)])
def test_subpixel_conv_data_gradient(ops, prec, B, H, C, Cout):
    """Conv2d(k4 s2 p1) data gradient through the Swish below: d x = ConvT(d y; W) * swish'(a) as four sub-pixel implicit
    GEMMs (B = the conv weight [co][(kh,kw,ci)] read MN-major, tap-split), aux and C row-mapped."""
    rs = np.random.RandomState(B + H + 2)
    a = torch.from_numpy(rs.standard_normal((B, C, H, H)).astype(np.float32))                     # pre-activation of the layer below
    w = torch.from_numpy((rs.standard_normal((Cout, C, 4, 4)) / (16 * C) ** 0.5).astype(np.float32))
    a64 = a.double().requires_grad_(True)
    y = F.conv2d(a64 * torch.sigmoid(a64), w.double(), stride=2, padding=1)                         # [B,Cout,H/2,H/2]
    OH = H // 2
    dy = torch.from_numpy(rs.standard_normal((B, OH, OH, Cout)).astype(np.float32))
    (y.permute(0, 2, 3, 1) * dy.double()).sum().backward()
    wc = w.permute(0, 2, 3, 1).reshape(Cout, 16 * C).contiguous().cuda()                           # [co][(kh,kw,ci)]
    a_rows = a.permute(0, 2, 3, 1).reshape(B * H * H, C).contiguous().cuda()
    dx = torch.full((B * H * H, C), float("nan"), device="cuda")
    dyd = dy.cuda()          # (descriptors hold raw pointers: the operand must outlive the launch)
    descs = ops.subpixel_k4s2p1(dyd, wc, dx, B, OH, OH, Cout, C, w_is_conv=True, aux=a_rows, epilogue=ops.EPI_MUL_DSWISH)
    ops.gemm_chain(descs, [-1] * 4, ops.chain_workspace("cuda"), prec)
    torch.cuda.synchronize()
    assert _rel(dx, a64.grad.permute(0, 2, 3, 1).reshape(B * H * H, C)) < TOL[prec] * 2


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("B,IH,Cin,Cout", [(3, 5, 256, 128), (50, 5, 256, 128), (4, 3, 64, 32)])
def test_full_correlation_conv_transpose_k4s1p0(ops, prec, B, IH, Cin, Cout):
    """ConvTranspose2d(k4 s1 p0) forward (celeba/model.py:117, 5x5 -> 8x8) and the data gradient of Conv2d(k4 s1 p0)
    (:85, 8x8 <- 5x5) as ONE implicit GEMM each: 4x4 stride-1 view with lower corner -3 and the filter taps reversed."""
    rs = np.random.RandomState(B + IH + 3)
    OH = IH + 3
    x = torch.from_numpy(rs.standard_normal((B, Cin, IH, IH)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((Cin, Cout, 4, 4)) / (16 * Cin) ** 0.5).astype(np.float32))
    ref = F.conv_transpose2d(x.double(), w.double(), stride=1, padding=0).permute(0, 2, 3, 1).reshape(B * OH * OH, Cout)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wt = w.permute(2, 3, 1, 0).reshape(16 * Cout, Cin).contiguous().cuda()                       # [(kh,kw,co)][ci]
    y = torch.full((B * OH * OH, Cout), float("nan"), device="cuda")
    ops.gemm_chain(ops.full_k4s1p0(x_nhwc, wt, y, B, IH, IH, Cin, Cout), [-1], ops.chain_workspace("cuda"), prec)
    assert _rel(y, ref) < TOL[prec] * 2
    # data gradient of a k4 s1 p0 convolution: d in [B,OH,OH,Cout'] from d out [B,IH,IH,Cin'] with Wc [Cin'][(kh,kw,Cout')]
    a = torch.from_numpy(rs.standard_normal((B, Cout, OH, OH)).astype(np.float32)).double().requires_grad_(True)
    wc = torch.from_numpy((rs.standard_normal((Cin, Cout, 4, 4)) / (16 * Cout) ** 0.5).astype(np.float32))   # Conv2d(Cout -> Cin)
    out = F.conv2d(a, wc.double(), stride=1, padding=0)                                           # [B,Cin,IH,IH]
    dy = torch.from_numpy(rs.standard_normal((B, IH, IH, Cin)).astype(np.float32))
    (out.permute(0, 2, 3, 1) * dy.double()).sum().backward()
    wcm = wc.permute(0, 2, 3, 1).reshape(Cin, 16 * Cout).contiguous().cuda()                      # [co][(kh,kw,ci)]
    dx = torch.full((B * OH * OH, Cout), float("nan"), device="cuda")
    dyd = dy.cuda()
    ops.gemm_chain(ops.full_k4s1p0(dyd, wcm, dx, B, IH, IH, Cin, Cout, w_is_conv=True), [-1], ops.chain_workspace("cuda"), prec)
    torch.cuda.synchronize()
    assert _rel(dx, a.grad.permute(0, 2, 3, 1).reshape(B * OH * OH, Cout)) < TOL[prec] * 2
