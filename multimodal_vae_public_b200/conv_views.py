"""Tensor-map geometry for implicit-GEMM 4x4 / stride-2 / pad-1 convolutions (groundwork for DESIGN.md section 8.1).

The convolution layers currently materialise im2col matrices in HBM (16x the activation bytes).  With activations
stored NHWC **with a one-pixel zero border**, the im2col matrix never has to exist: row m = (b, oh, ow), column
k = (kh, kw, ci) of it is the element

    xp[b, 2*oh + kh, 2*ow + kw, ci]          xp: [B, H+2, W+2, C]  (padded input)

and for a fixed kh the 4*C values over (kw, ci) are CONTIGUOUS in memory.  The matrix is therefore a rank-5 strided
view of xp with overlapping strides -- (r, ow, kh, oh, b), r = kw*C + ci -- which is exactly what a tiled TMA tensor
map describes (no im2col-mode map needed), and a 128-row x 32-column operand tile of the GEMM is ONE box of that map
whenever OW divides 128 (CelebA: OW = 32, 16, 8).  `conv_k4s2p1_view` returns the dims / strides / box of that map;
`tests/test_conv_views_cpu.py` proves with numpy's as_strided that the view equals the reference im2col
(fashionmnist/model.py:79-82, celeba/model.py:76-92 are the layers it will serve).  The weight matrices already use the
matching column order [Cout, (kh, kw, ci)].
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class ConvView:
    dims: Tuple[int, int, int, int, int]        # (r, ow, kh, oh, b) sizes, innermost first (cuTensorMapEncodeTiled order)
    strides: Tuple[int, int, int, int, int]     # element strides of the same dims (strides[0] == 1)
    box: Tuple[int, int, int, int, int]         # box of one 128-row x 32-column operand tile
    rows: int                                   # M = B * OH * OW
    cols: int                                   # K = 16 * C
    kblocks_per_kh: int                         # 4*C / 32 k-blocks before kh advances

    def tile_coords(self, m_blk: int, kb: int) -> Tuple[int, int, int, int, int]:
        """TMA coordinates (r, ow, kh, oh, b) of operand tile (128-row block m_blk, 32-column block kb)."""
        r, ow_n, _, oh_n, _ = self.dims
        _, bow, _, boh, bb = self.box
        kh, rb = divmod(kb, self.kblocks_per_kh)
        row0 = m_blk * 128
        b0, rem = divmod(row0, oh_n * ow_n)
        oh0, ow0 = divmod(rem, ow_n)
        assert ow0 == 0 and oh0 % boh == 0 and b0 % bb == 0, "tile does not start on a box boundary"
        return (rb * 32, 0, kh, oh0, b0)


def conv_k4s2p1_view(B: int, H: int, W: int, C: int) -> ConvView:
    """Geometry of the implicit im2col matrix of Conv2d(C -> *, 4, 2, 1) over a padded NHWC input [B, H+2, W+2, C]."""
    if H % 2 or W % 2:
        raise ValueError("even H, W required")
    OH, OW = H // 2, W // 2
    if (4 * C) % 32:
        raise ValueError("4*C must be a multiple of the 32-float k-block (C % 8 == 0)")
    if OW > 128 or 128 % OW:
        raise ValueError("OW must divide the 128-row tile (use the materialised path otherwise)")
    boh = min(OH, 128 // OW)
    if OH % boh:
        raise ValueError("OH must be a multiple of the rows of one tile")
    bb = 128 // (OW * boh)
    Wp, Hp = W + 2, H + 2
    dims = (4 * C, OW, 4, OH, B)
    strides = (1, 2 * C, Wp * C, 2 * Wp * C, Hp * Wp * C)
    if any((s * 4) % 16 for s in strides[1:]):
        raise ValueError("TMA strides must be multiples of 16 bytes")
    return ConvView(dims, strides, (32, OW, 1, boh, bb), B * OH * OW, 16 * C, (4 * C) // 32)


# ---------------------------------------------------------------------------------------------------- ConvTranspose
# ConvTranspose2d(Cin -> Cout, 4, 2, 1) (fashionmnist/model.py:112-114, celeba/model.py:116-126) as FOUR stride-1 2x2
# convolutions, one per output parity class (ph, pw): y[b, 2j+ph, 2i+pw, :] only receives the taps kh = KH[ph][dh],
# kw = KH[pw][dw] from the input pixels (j + ph + dh - 1, i + pw + dw - 1).  On the zero-bordered input that is again a
# rank-5 strided view -- (r = (dw, ci), i, dh, j, b) with a base offset of (ph, pw) pixels -- times a [Cout, 4*Cin] slice
# of the weights: no cols matrix, no col2im, one quarter of the output pixels per GEMM with a strided epilogue store.
KH = ((3, 1), (2, 0))       # KH[parity][d]: kernel index met by input offset d of that output parity


def convt_k4s2p1_subpixel_view(B: int, IH: int, IW: int, C: int, ph: int, pw: int):
    """(dims, strides, base element offset) of the implicit patch matrix [B*IH*IW, 4*C] of output parity (ph, pw) over
    the zero-bordered NHWC input [B, IH+2, IW+2, C]; dims innermost first: (r, i, dh, j, b), r = dw*C + ci."""
    Wp, Hp = IW + 2, IH + 2
    dims = (2 * C, IW, 2, IH, B)
    strides = (1, C, Wp * C, Wp * C, Hp * Wp * C)
    return dims, strides, (ph * Wp + pw) * C


def convt_subpixel_weight(wt, ph: int, pw: int):
    """[Cout, (dh, dw, ci)] operand of parity (ph, pw) from ConvTranspose2d's weight ``wt`` [Cin, Cout, 4, 4]."""
    sub = wt[:, :, list(KH[ph]), :][:, :, :, list(KH[pw])]          # [Cin, Cout, dh, dw]
    return sub.permute(1, 2, 3, 0).reshape(wt.shape[1], -1)


def wgrad_box(view: ConvView) -> Tuple[int, int, int, int, int]:
    """Box of one MN-major wgrad operand tile of the same map: 32 reduction rows (output pixels) x 32 columns -- the B
    operand of dW[Cout, (kh,kw,ci)] = dY^T[Cout, pixels] * im2col[pixels, (kh,kw,ci)] (and, for a ConvTranspose, of
    dWt = im2col(dY)^T * X): 32 consecutive pixels are one row of OW = 32, two rows of 16 or four rows of 8."""
    ow = view.dims[1]
    if 32 % ow and ow % 32:
        raise ValueError("OW must divide 32 or be a multiple of it")
    bow = min(ow, 32)
    return (32, bow, 1, 32 // bow, 1)
