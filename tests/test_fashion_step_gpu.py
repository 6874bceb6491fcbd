"""GPU parity of the conv (FashionMNIST-flavour) path: im2col/col2im kernels and the fused trainer against the
oracle and the golden fixture produced by the unmodified reference (fashionmnist/model.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import mvae_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("B,H,C", [(3, 28, 1), (2, 14, 64), (2, 6, 20)])
def test_im2col_col2im_are_conv_data_movement(B, H, C):
    """im2col + matmul == conv2d(k4,s2,p1); matmul + col2im == conv_transpose2d(k4,s2,p1) (NHWC vs torch NCHW)."""
    from multimodal_vae_public_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + H + C)
    x = torch.randn(B, H, H, C, generator=g)
    cols = torch.full((B * (H // 2) ** 2, 16 * C), float("nan")).cuda()
    ops.im2col_k4s2p1(x.cuda().contiguous(), cols, B, H, H, C)
    Cout = 5
    w = torch.randn(Cout, C, 4, 4, generator=g)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, 2, 1)       # [B,Cout,OH,OW]
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 16 * C).double()
    got = (cols.cpu().double() @ wp.t()).reshape(B, H // 2, H // 2, Cout).permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() < 1e-9
    # transpose conv: x2 [B,IH,IW,Cin] -> [B,2IH,2IW,Cout]
    IH, Cin, Co = H // 2, 7, C
    x2 = torch.randn(B, IH, IH, Cin, generator=g)
    wt = torch.randn(Cin, Co, 4, 4, generator=g)
    reft = torch.nn.functional.conv_transpose2d(x2.permute(0, 3, 1, 2).double(), wt.double(), None, 2, 1)
    wtp = wt.permute(2, 3, 1, 0).reshape(16 * Co, Cin).double()
    colsT = (x2.reshape(-1, Cin).double() @ wtp.t()).float().cuda().contiguous()
    out = torch.empty(B, 2 * IH, 2 * IH, Co).cuda(); act = torch.empty_like(out)
    ops.col2im_k4s2p1(colsT, out, B, IH, IH, Co, out_act=act)
    assert (out.cpu().double().permute(0, 3, 1, 2) - reft).abs().max().item() < 5e-5
    assert (act.cpu().double() - O.swish(out.cpu().double())).abs().max().item() < 1e-5
    aux = torch.randn(B, 2 * IH, 2 * IH, Co, generator=g).cuda()
    out2 = torch.empty_like(out)
    ops.col2im_k4s2p1(colsT, out2, B, IH, IH, Co, aux=aux)
    s = torch.sigmoid(aux.cpu().double())
    assert (out2.cpu().double() - out.cpu().double() * s * (1 + aux.cpu().double() * (1 - s))).abs().max().item() < 1e-4


def _trainer(B, prec=1, **kw):
    from multimodal_vae_public_b200.trainer_fashion import FashionMVAETrainer
    return FashionMVAETrainer(n_latents=64, batch_size=B, precision=prec, use_graph=False, **kw)


def test_state_dict_round_trip():
    tr = _trainer(4)
    p32 = O.make_params(O.fashion_param_shapes(64), seed=9)
    tr.load_state_dict(p32)
    sd = tr.state_dict()
    assert list(sd.keys()) == [k for k, _ in O.fashion_param_shapes(64)]
    for k, v in p32.items():
        assert torch.equal(sd[k].cpu(), v), k


@pytest.mark.parametrize("prec", [1, 0])
def test_fashion_step_matches_reference_golden(prec):
    fa = dict(np.load(os.path.join(G, "fashion_golden.npz")))
    tr = _trainer(4, prec)
    tr.load_state_dict(O.make_params(O.fashion_param_shapes(64), seed=0))
    image = torch.from_numpy(fa["image"]); text = torch.from_numpy(fa["text"]); noise = torch.from_numpy(fa["noises"])
    tr.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    ls = tr.losses()
    rt = {1: 3e-6, 0: 2e-4}[prec]
    for name, ref in zip(("joint", "image", "text"), fa["terms"]):
        assert abs(ls[name] - ref) <= rt * abs(ref) + 1e-5, (name, ls[name], ref)
    gt = {1: 3e-4, 0: 4e-2}[prec]
    grads = tr.export_grads()
    for k, g in grads.items():
        g = g.cpu()
        head = fa[f"grad_head/{k}"]
        scale = max(np.abs(head).max(), fa[f"grad_digest/{k}"][2] / np.sqrt(g.numel()))
        assert np.abs(g.reshape(-1)[:64].numpy() - head).max() <= gt * scale + 1e-7, k
        d = fa[f"grad_digest/{k}"]
        assert abs(g.double().norm().item() - d[2]) <= 5 * gt * d[2] + 1e-7, k


@pytest.mark.parametrize("B", [32, 130, 512, 4096])
def test_fashion_step_matches_oracle_fp64(B):
    # B = 4096: BASELINE.json configs[2]; B = 512: its per-GPU batch at 8 GPUs (the fp64 oracle step takes ~15 s of CPU)
    L = 64
    rs = np.random.RandomState(B)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32))
    p32 = O.make_params(O.fashion_param_shapes(L), seed=3)
    loss, terms, grads, _ = O.fashion_step_grads({k: v.double() for k, v in p32.items()}, image.double(), text, L,
                                                 [n.double() for n in noise], 1.0, 10.0, 0.5)
    tr = _trainer(B)
    tr.load_state_dict(p32)
    got = tr.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    assert abs(got - loss.item()) <= 3e-6 * abs(loss.item())
    mine = tr.export_grads()
    for k, gref in grads.items():
        err = (mine[k].cpu().double() - gref).abs().max().item() / max(gref.abs().max().item(), 1e-12)
        assert err <= 3e-4, (k, err)
