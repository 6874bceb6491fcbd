"""The implicit-GEMM geometry of conv_views.py is the reference convolution's im2col (pure host arithmetic, numpy):
the rank-5 overlapping-stride view of the zero-bordered NHWC input equals the explicit [B*OH*OW, 16*C] patch matrix of
Conv2d(k=4, s=2, p=1) (fashionmnist/model.py:79-82, celeba/model.py:76-92), and every 128x32 operand tile is one box."""
import numpy as np
import pytest
import torch

from multimodal_vae_public_b200.conv_views import conv_k4s2p1_view


def _explicit_im2col(x):   # x: [B, H, W, C] -> [B*OH*OW, (kh, kw, ci)]
    B, H, W, C = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    OH, OW = H // 2, W // 2
    out = np.empty((B, OH, OW, 4, 4, C), x.dtype)
    for kh in range(4):
        for kw in range(4):
            out[:, :, :, kh, kw, :] = xp[:, kh:kh + 2 * OH:2, kw:kw + 2 * OW:2, :]
    return out.reshape(B * OH * OW, 16 * C)


@pytest.mark.parametrize("B,H,C", [(4, 64, 8), (2, 32, 32), (4, 16, 64), (3, 32, 16)])
def test_view_equals_im2col_and_conv2d(B, H, C):
    rs = np.random.RandomState(B + H + C)
    x = rs.standard_normal((B, H, H, C)).astype(np.float32)
    xp = np.ascontiguousarray(np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0))))
    v = conv_k4s2p1_view(B, H, H, C)
    # numpy wants outermost-first shapes / byte strides
    view = np.lib.stride_tricks.as_strided(xp, shape=v.dims[::-1], strides=tuple(4 * s for s in v.strides[::-1]))
    # (b, oh, kh, ow, r) -> rows (b, oh, ow), columns (kh, r)
    mat = view.transpose(0, 1, 3, 2, 4).reshape(v.rows, v.cols)
    ref = _explicit_im2col(x)
    assert np.array_equal(mat, ref)
    # and it is the reference convolution: im2col @ W^T with W in (kh, kw, ci) column order
    w = rs.standard_normal((5, C, 4, 4)).astype(np.float32)
    y = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(w).double(), None, 2, 1)
    got = (mat.astype(np.float64) @ w.transpose(0, 2, 3, 1).reshape(5, -1).astype(np.float64).T)
    got = got.reshape(B, H // 2, H // 2, 5).transpose(0, 3, 1, 2)
    assert np.abs(got - y.numpy()).max() < 1e-9
    # every 128 x 32 operand tile is one box of the view
    bx = v.box
    assert bx[1] * bx[3] * bx[4] == 128 or v.rows < 128
    if v.rows % 128 == 0:
        for m_blk in (0, v.rows // 128 - 1):
            for kb in (0, v.cols // 32 - 1):
                c = v.tile_coords(m_blk, kb)
                sl = view[c[4]:c[4] + bx[4], c[3]:c[3] + bx[3], c[2]:c[2] + 1, c[1]:c[1] + bx[1], c[0]:c[0] + 32]
                tile = sl.transpose(0, 1, 3, 2, 4).reshape(128, 32)     # smem order: b, oh, ow rows; r fastest
                assert np.array_equal(tile, ref[m_blk * 128:(m_blk + 1) * 128, kb * 32:(kb + 1) * 32])


def test_unsupported_geometries_are_rejected():
    with pytest.raises(ValueError):
        conv_k4s2p1_view(2, 28, 28, 64)       # OW = 14 does not divide 128 (FashionMNIST: stays on the materialised path)
    with pytest.raises(ValueError):
        conv_k4s2p1_view(2, 64, 64, 3)        # RGB input layer: 4*C = 12 floats per kh run


@pytest.mark.parametrize("B,IH,C,Co", [(2, 8, 16, 6), (3, 16, 8, 5), (1, 5, 4, 3)])
def test_conv_transpose_is_four_subpixel_convolutions(B, IH, C, Co):
    """ConvTranspose2d(k=4, s=2, p=1) (celeba/model.py:116-126) == four implicit 2x2 GEMMs over the zero-bordered input,
    one per output parity, each storing a quarter of the output pixels."""
    from multimodal_vae_public_b200.conv_views import convt_k4s2p1_subpixel_view, convt_subpixel_weight
    rs = np.random.RandomState(IH + C)
    x = rs.standard_normal((B, IH, IH, C)).astype(np.float32)
    wt = torch.from_numpy(rs.standard_normal((C, Co, 4, 4)).astype(np.float32))
    ref = torch.nn.functional.conv_transpose2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), wt.double(), None, 2, 1)
    xp = np.ascontiguousarray(np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0))))
    y = np.zeros((B, 2 * IH, 2 * IH, Co))
    for ph in range(2):
        for pw in range(2):
            dims, strides, base = convt_k4s2p1_subpixel_view(B, IH, IH, C, ph, pw)
            view = np.lib.stride_tricks.as_strided(xp.reshape(-1)[base:], shape=dims[::-1],
                                                   strides=tuple(4 * s for s in strides[::-1]))
            mat = view.transpose(0, 1, 3, 2, 4).reshape(B * IH * IH, 4 * C).astype(np.float64)   # rows (b, j, i), cols (dh, r)
            w = convt_subpixel_weight(wt, ph, pw).double().numpy()
            y[:, ph::2, pw::2, :] = (mat @ w.T).reshape(B, IH, IH, Co)
    assert np.abs(y.transpose(0, 3, 1, 2) - ref.numpy()).max() < 1e-9


def test_wgrad_operand_tile_is_one_box():
    from multimodal_vae_public_b200.conv_views import wgrad_box
    B, H, C = 2, 32, 16
    rs = np.random.RandomState(0)
    x = rs.standard_normal((B, H, H, C)).astype(np.float32)
    xp = np.ascontiguousarray(np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0))))
    v = conv_k4s2p1_view(B, H, H, C)
    view = np.lib.stride_tricks.as_strided(xp, shape=v.dims[::-1], strides=tuple(4 * s for s in v.strides[::-1]))
    ref = _explicit_im2col(x)
    bx = wgrad_box(v)
    assert bx[1] * bx[3] * bx[4] == 32
    OW = H // 2
    for p0 in (0, 96, 480):                      # first pixel (reduction row) of a 32-row k-block
        b0, rem = divmod(p0, OW * OW)
        oh0, ow0 = divmod(rem, OW)
        assert ow0 % bx[1] == 0 and oh0 % bx[3] == 0
        for kb in (0, 5, v.cols // 32 - 1):
            kh, rb = divmod(kb, v.kblocks_per_kh)
            sl = view[b0:b0 + 1, oh0:oh0 + bx[3], kh:kh + 1, ow0:ow0 + bx[1], rb * 32:rb * 32 + 32]
            tile = sl.transpose(0, 1, 3, 2, 4).reshape(32, 32)          # [32 pixels (k rows)][32 columns (n)]
            assert np.array_equal(tile, ref[p0:p0 + 32, kb * 32:kb * 32 + 32])


def test_conv_transpose_backward_is_convolution_of_dy():
    """What the decoder backward needs (celeba/model.py:116-126 under autograd): d x of ConvTranspose2d(k4,s2,p1) is
    Conv2d(k4,s2,p1) of dy with the same weight tensor, and d Wt = im2col(dy)^T x -- both implicit GEMMs over the
    zero-bordered dy, i.e. the same machinery as the encoder's forward / wgrad."""
    rs = np.random.RandomState(3)
    B, IH, Ci, Co = 2, 8, 6, 5
    x = torch.from_numpy(rs.standard_normal((B, Ci, IH, IH))).requires_grad_(True)
    wt = torch.from_numpy(rs.standard_normal((Ci, Co, 4, 4))).requires_grad_(True)
    y = torch.nn.functional.conv_transpose2d(x, wt, None, 2, 1)
    dy = torch.from_numpy(rs.standard_normal(tuple(y.shape)))
    y.backward(dy)
    dx = torch.nn.functional.conv2d(dy, wt, None, 2, 1)                 # wt read as a Conv2d weight [out=Ci, in=Co, 4, 4]
    assert (dx - x.grad).abs().max().item() < 1e-10
    cols = torch.from_numpy(_explicit_im2col(dy.permute(0, 2, 3, 1).numpy()))          # [B*IH*IW, (kh, kw, co)]
    xm = x.detach().permute(0, 2, 3, 1).reshape(-1, Ci)                                  # [B*IH*IW, Ci]
    dwt = (cols.t() @ xm).reshape(4, 4, Co, Ci).permute(3, 2, 0, 1)                      # -> [Ci, Co, kh, kw]
    assert (dwt - wt.grad).abs().max().item() < 1e-9
