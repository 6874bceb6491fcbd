"""Direct 4x4 / stride-2 / pad-1 convolutions of the image-side layers (csrc/conv_small.cu) against torch's fp64
conv2d / conv_transpose2d + autograd on the CPU (the ops the reference's nn.Conv2d / nn.ConvTranspose2d layers issue:
fashionmnist/model.py:79,114; celeba/model.py:77,126)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from multimodal_vae_public_b200 import ops as _ops
    return _ops


def _rel(a, ref):
    return (a.double().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def _swish(x):
    return x * torch.sigmoid(x)


@pytest.mark.parametrize("B,H,Cin,Cout", [(3, 28, 1, 64), (37, 28, 1, 64), (2, 64, 3, 32), (5, 8, 3, 32), (600, 28, 1, 64)])
def test_conv_cin_fwd_and_wgrad(ops, B, H, Cin, Cout):
    rs = np.random.RandomState(B + H)
    x = torch.from_numpy(rs.uniform(0, 1, (B, Cin, H, H)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((Cout, Cin, 4, 4)) / (16 * Cin) ** 0.5).astype(np.float32))
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    wc = w.permute(0, 2, 3, 1).reshape(Cout, 16 * Cin).contiguous().cuda()          # [Cout][(kh,kw,ci)]
    OH = H // 2
    a = torch.full((B * OH * OH, Cout), float("nan"), device="cuda"); h = torch.full_like(a, float("nan"))
    ops.conv_cin_fwd(x_nhwc, wc, a, h, B, H, H, Cin, Cout)
    w64 = w.double().requires_grad_(True)
    ra = F.conv2d(x.double(), w64, stride=2, padding=1)                              # [B,Cout,OH,OW]
    ra_rows = ra.permute(0, 2, 3, 1).reshape(B * OH * OH, Cout)
    assert _rel(a, ra_rows.detach()) < 2e-6
    assert _rel(h, _swish(ra_rows.detach())) < 3e-6
    da = torch.from_numpy(rs.standard_normal((B * OH * OH, Cout)).astype(np.float32))
    (ra_rows * da.double()).sum().backward()
    dwc = torch.full((Cout, 16 * Cin), 0.5, device="cuda")                           # accumulates
    ops.conv_cin_wgrad(x_nhwc, da.cuda(), dwc, B, H, H, Cin, Cout)
    ref = w64.grad.permute(0, 2, 3, 1).reshape(Cout, 16 * Cin)
    assert _rel(dwc - 0.5, ref) < 5e-6


@pytest.mark.parametrize("B,IH,Cin,Cout", [(3, 14, 64, 1), (41, 14, 64, 1), (2, 32, 32, 3), (3, 4, 32, 3), (700, 14, 64, 1)])
def test_convt_cout_fwd_and_bwd(ops, B, IH, Cin, Cout):
    rs = np.random.RandomState(B + IH)
    ain = torch.from_numpy(rs.standard_normal((B, Cin, IH, IH)).astype(np.float32))
    w = torch.from_numpy((rs.standard_normal((Cin, Cout, 4, 4)) / (16 * Cin) ** 0.5).astype(np.float32))   # ConvTranspose2d layout
    a64 = ain.double().requires_grad_(True); w64 = w.double().requires_grad_(True)
    rout = F.conv_transpose2d(_swish(a64), w64, stride=2, padding=1)                # [B,Cout,2IH,2IW]
    a_rows = ain.permute(0, 2, 3, 1).reshape(B * IH * IH, Cin).contiguous().cuda()
    h_rows = _swish(a_rows.double()).float()
    wt = w.permute(2, 3, 1, 0).reshape(16 * Cout, Cin).contiguous().cuda()           # [(kh,kw,co)][ci]
    out = torch.full((B, 2 * IH, 2 * IH, Cout), float("nan"), device="cuda")
    ops.convT_cout_fwd(h_rows, wt, out, B, IH, IH, Cin, Cout)
    assert _rel(out, rout.detach().permute(0, 2, 3, 1)) < 3e-6
    dout = torch.from_numpy(rs.standard_normal((B, 2 * IH, 2 * IH, Cout)).astype(np.float32))
    (rout.permute(0, 2, 3, 1) * dout.double()).sum().backward()
    dhin = torch.full((B * IH * IH, Cin), float("nan"), device="cuda")
    dwt = torch.full((16 * Cout, Cin), -0.25, device="cuda")
    ops.convT_cout_bwd(dout.cuda(), h_rows, a_rows, wt, dhin, dwt, B, IH, IH, Cin, Cout)
    assert _rel(dhin, a64.grad.permute(0, 2, 3, 1).reshape(B * IH * IH, Cin)) < 5e-6    # through the Swish below the layer
    assert _rel(dwt + 0.25, w64.grad.permute(2, 3, 1, 0).reshape(16 * Cout, Cin)) < 5e-6


def test_direct_layers_equal_im2col_gemm_path(monkeypatch):
    """The FashionMNIST trainer with the direct 1-channel kernels (default) == with im2col + tensor-core GEMM."""
    from multimodal_vae_public_b200.trainer_fashion import FashionMVAETrainer
    B, L = 96, 64
    rs = np.random.RandomState(5)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32))
    a = FashionMVAETrainer(n_latents=L, batch_size=B, use_graph=False)
    monkeypatch.setenv("MVAE_DIRECT_C1", "0")
    b = FashionMVAETrainer(n_latents=L, batch_size=B, use_graph=False)
    assert a.direct_c1 and not b.direct_c1
    b.load_state_dict(a.state_dict())
    la = a.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    lb = b.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    assert abs(la - lb) <= 2e-6 * abs(la)
    for k in a.grads:
        err = (a.grads[k] - b.grads[k]).abs().max().item() / max(b.grads[k].abs().max().item(), 1e-12)
        assert err <= 1e-4, (k, err)
