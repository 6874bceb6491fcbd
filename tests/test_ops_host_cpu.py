"""Host-side logic of ops.py on the CPU (no CUDA library call): the GEMM descriptors it builds mean what the reference's
layers compute.  A small numpy EMULATOR of mvae_gemm_desc (include/mvae_b200.h: a_view = im2col view of an NHWC tensor,
b_tap_* = tap-split B, rowmap_* = output row map) executes the descriptors of

  * ops.subpixel_k4s2p1   -- ConvTranspose2d(k4, s2, p1) forward (fashionmnist/model.py:112-114, celeba/model.py:120-126)
                             and the data gradient of Conv2d(k4, s2, p1) (fashionmnist/model.py:79-82)
  * ops.full_k4s1p0       -- ConvTranspose2d(k4, s1, p0) (celeba/model.py:117)
  * an implicit Conv2d(k4, s2, p1) forward descriptor (ops.conv_view defaults)

and the results are compared with torch.nn.functional on the same data; plus the pre-split-weights registry
(register_lo_arena: pointer lookup, the N <= max_n policy, purge of stale entries)."""
import ctypes

import numpy as np
import pytest
import torch

from multimodal_vae_public_b200 import ops


def _np(ptr_owner_map, ptr):
    return ptr_owner_map[ptr]


def _emulate(d, tensors, out_shape):
    """Execute one GemmDesc on numpy arrays.  `tensors`: data_ptr -> numpy array (the arrays ops.gemm_desc saw).
    Returns (C written densely or through the row map) for plain-store descriptors (no epilogue math)."""
    M, N, K = d.M, d.N, d.K
    # ---- A
    v = d.a_view
    assert v.C > 0, "this emulator covers implicit A operands"
    x = tensors[d.A].reshape(v.N, v.H, v.W, v.C)
    s = v.stride
    OH = (v.H + v.upper_h - v.lower_h - 1) // s + 1        # window origins lower, lower + s, ... while origin <= H - 1 + upper
    OW = (v.W + v.upper_w - v.lower_w - 1) // s + 1
    assert M == v.N * OH * OW and K == v.taps_h * v.taps_w * v.C
    A = np.zeros((v.N, OH, OW, v.taps_h, v.taps_w, v.C), np.float64)
    for th in range(v.taps_h):
        for tw in range(v.taps_w):
            for oh in range(OH):
                h = v.lower_h + oh * s + th
                if not 0 <= h < v.H:
                    continue
                for ow in range(OW):
                    w_ = v.lower_w + ow * s + tw
                    if 0 <= w_ < v.W:
                        A[:, oh, ow, th, tw, :] = x[:, h, w_, :]
    A = A.reshape(M, K)
    # ---- B  (K-major: B[n][k] at B + n ldb + k ; MN-major: B[k][n] at B + k ldb + n)
    wmat = tensors[d.B]
    ldb = d.ldb
    flat = wmat.reshape(-1)
    Bm = np.zeros((N, K), np.float64)
    if d.b_tap_slots:
        kper, mnper = d.b_tap_k, d.b_tap_mn
        assert K == d.b_tap_slots * kper
        for t in range(d.b_tap_slots):
            base = d.b_tap_table[t] * mnper
            for n in range(N):
                for k in range(kper):
                    Bm[n, t * kper + k] = flat[(base + n) * ldb + k] if not d.b_mn_major else flat[k * ldb + base + n]
    else:
        for n in range(N):
            for k in range(K):
                Bm[n, k] = flat[n * ldb + k] if not d.b_mn_major else flat[k * ldb + n]
    Cm = A @ Bm.T
    # ---- store
    out = np.zeros(out_shape, np.float64)
    if d.rowmap_IH:
        IH, IW, st, py, px = d.rowmap_IH, d.rowmap_IW, d.rowmap_s, d.rowmap_py, d.rowmap_px
        o = out.reshape(-1, st * IH, st * IW, N)
        o[:, py::st, px::st, :] = Cm.reshape(-1, IH, IW, N)
        return out, (slice(None), slice(py, None, st), slice(px, None, st))
    out.reshape(M, N)[:] = Cm
    return out, None


def _registry(*arrs):
    return {a.data_ptr(): a.numpy() for a in arrs}


@pytest.mark.parametrize("n,IH,Cx,Cy", [(2, 7, 8, 4), (3, 4, 4, 8), (1, 5, 12, 4)])
def test_subpixel_descriptors_are_conv_transpose(n, IH, Cx, Cy):
    rs = np.random.RandomState(n + IH + Cx)
    x = torch.from_numpy(rs.standard_normal((n, IH, IH, Cx)).astype(np.float32))                 # NHWC
    wt = torch.from_numpy(rs.standard_normal((Cx, Cy, 4, 4)).astype(np.float32))                 # torch ConvT layout
    w = wt.permute(2, 3, 1, 0).reshape(16 * Cy, Cx).contiguous()                                  # Wt [(kh,kw,cy)][cx]
    out = torch.zeros(n * 4 * IH * IH, Cy)
    descs = ops.subpixel_k4s2p1(x, w, out, n, IH, IH, Cx, Cy)
    assert len(descs) == 4 and sorted((d.rowmap_py, d.rowmap_px) for d in descs) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    reg = _registry(x, w, out)
    got = np.zeros((n, 2 * IH, 2 * IH, Cy))
    for d in descs:
        part, sl = _emulate(d, reg, (n, 2 * IH, 2 * IH, Cy))
        got[sl[0], sl[1], sl[2]] = part[sl[0], sl[1], sl[2]]
    ref = torch.nn.functional.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), None, 2, 1)
    assert np.abs(got - ref.permute(0, 2, 3, 1).numpy()).max() < 1e-9


@pytest.mark.parametrize("n,IH,Cout,Cin", [(2, 7, 8, 4), (2, 4, 4, 8)])
def test_subpixel_descriptors_are_conv_data_gradient(n, IH, Cout, Cin):
    """w_is_conv: x = d out [n,IH,IW,Cout], w = Wc [Cout][(kh,kw,ci)] (MN-major B), y = d in [n,2IH,2IW,Cin]."""
    rs = np.random.RandomState(7 * n + IH)
    dy = torch.from_numpy(rs.standard_normal((n, IH, IH, Cout)).astype(np.float32))
    wc4 = torch.from_numpy(rs.standard_normal((Cout, Cin, 4, 4)).astype(np.float32))            # torch Conv2d layout
    wc = wc4.permute(0, 2, 3, 1).reshape(Cout, 16 * Cin).contiguous()
    out = torch.zeros(n * 4 * IH * IH, Cin)
    descs = ops.subpixel_k4s2p1(dy, wc, out, n, IH, IH, Cout, Cin, w_is_conv=True)
    assert all(d.b_mn_major == 1 for d in descs)
    reg = _registry(dy, wc, out)
    got = np.zeros((n, 2 * IH, 2 * IH, Cin))
    for d in descs:
        part, sl = _emulate(d, reg, (n, 2 * IH, 2 * IH, Cin))
        got[sl[0], sl[1], sl[2]] = part[sl[0], sl[1], sl[2]]
    xin = torch.zeros(n, Cin, 2 * IH, 2 * IH, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(xin, wc4.double(), None, 2, 1)
    y.backward(dy.permute(0, 3, 1, 2).double())
    assert np.abs(got - xin.grad.permute(0, 2, 3, 1).numpy()).max() < 1e-9


def test_full_correlation_descriptor_is_stride1_conv_transpose():
    n, IH, Cx, Cy = 2, 5, 8, 4
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.standard_normal((n, IH, IH, Cx)).astype(np.float32))
    wt = torch.from_numpy(rs.standard_normal((Cx, Cy, 4, 4)).astype(np.float32))
    w = wt.permute(2, 3, 1, 0).reshape(16 * Cy, Cx).contiguous()
    out = torch.zeros(n * (IH + 3) * (IH + 3), Cy)
    (d,) = ops.full_k4s1p0(x, w, out, n, IH, IH, Cx, Cy)
    got, _ = _emulate(d, _registry(x, w, out), (n, IH + 3, IH + 3, Cy))
    ref = torch.nn.functional.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), None, 1, 0)
    assert np.abs(got - ref.permute(0, 2, 3, 1).numpy()).max() < 1e-9


def test_implicit_conv_forward_descriptor_is_conv2d():
    n, H, Cin, Cout = 2, 8, 4, 8
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.standard_normal((n, H, H, Cin)).astype(np.float32))
    w4 = torch.from_numpy(rs.standard_normal((Cout, Cin, 4, 4)).astype(np.float32))
    w = w4.permute(0, 2, 3, 1).reshape(Cout, 16 * Cin).contiguous()                               # [Cout][(kh,kw,ci)]
    out = torch.zeros(n * (H // 2) ** 2, Cout)
    v = ops.conv_view(n, H, H, Cin)
    assert (v.lower_h, v.upper_h, v.stride, v.taps_h) == (-1, -2, 2, 4)
    d = ops.gemm_desc(x, w, out, n * (H // 2) ** 2, Cout, 16 * Cin, a_view=v)
    assert d.lda == 0 and d.a_view.C == Cin
    got, _ = _emulate(d, _registry(x, w, out), (n, H // 2, H // 2, Cout))
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w4.double(), None, 2, 1)
    assert np.abs(got - ref.permute(0, 2, 3, 1).numpy()).max() < 1e-9


def test_lo_arena_registry_policy_and_purge():
    saved = list(ops._LO_ARENAS)
    try:
        ops._LO_ARENAS.clear()
        params, lo = torch.zeros(4096), torch.zeros(4096)
        ops.register_lo_arena(params, lo, max_n=64)
        w_narrow, w_wide = params[:64 * 32].view(64, 32), params[2048:2048 + 128 * 16].view(128, 16)
        A, Cn, Cw = torch.zeros(8, 32), torch.zeros(8, 64), torch.zeros(8, 128)
        d = ops.gemm_desc(A, w_narrow, Cn, 8, 64, 32)
        assert d.B_lo == lo.data_ptr()                                   # N = 64 <= max_n: the twin at the same offset
        d = ops.gemm_desc(A[:, :16], w_wide, Cw, 8, 128, 16)
        assert not d.B_lo                                                # N = 128 > max_n: generated in the loop instead
        d = ops.gemm_desc(A, torch.zeros(64, 32), Cn, 8, 64, 32)
        assert not d.B_lo                                                # not a registered parameter
        ops.register_lo_arena(params, lo)                                # re-registration replaces the overlapping entry
        assert len(ops._LO_ARENAS) == 1 and ops._LO_ARENAS[0][3] == 0
        d = ops.gemm_desc(A[:, :16], w_wide, Cw, 8, 128, 16)
        assert d.B_lo == lo.data_ptr() + 2048 * 4
        with pytest.raises(Exception):
            ops.register_lo_arena(params, torch.zeros(10))
        ops.unregister_lo_arena(params)
        assert ops._LO_ARENAS == []
    finally:
        ops._LO_ARENAS[:] = saved


def test_descriptor_struct_matches_header_size():
    # ctypes mirror of mvae_gemm_desc: 2 views of 11 int32, 16-entry tap table, 5 row-map ints (include/mvae_b200.h)
    from multimodal_vae_public_b200 import _lib
    assert ctypes.sizeof(_lib.ConvView) == 44
    assert ctypes.sizeof(_lib.GemmDesc) % 8 == 0 and ctypes.sizeof(_lib.GemmDesc) >= 13 * 8 + 2 * 44 + 19 * 4 + 5 * 4
