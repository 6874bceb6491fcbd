"""torch.Tensor-level wrappers over the C ABI (device pointers + sizes in, status out).

torch is used for device memory and streams only.  Every function enqueues on the current CUDA
stream and raises ``MvaeError`` on failure; there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import EPI_BIAS_SWISH, EPI_MUL_DSWISH, EPI_STORE, GemmDesc, PREC_3XTF32, PREC_TF32  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk2d(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.dtype != torch.float32 or not t.is_cuda or t.stride(1) != 1:
        raise _lib.MvaeError(f"{name}: expected a CUDA fp32 2-D tensor with unit inner stride, got "
                             f"{tuple(t.shape)} {t.dtype} strides={t.stride()} device={t.device}")


def conv_view(N, H, W, Cch, taps=4, stride=2, pad=1, lower=None, upper=None) -> _lib.ConvView:
    """im2col view of an NHWC tensor [N,H,W,C] for mvae_gemm_desc.a_view / b_view (include/mvae_b200.h): a taps x taps
    filter walking with `stride`; by default the padding is symmetric (`pad`), `lower` / `upper` (h, w) override it."""
    v = _lib.ConvView()
    v.N, v.H, v.W, v.C = N, H, W, Cch
    lo = (-pad, -pad) if lower is None else lower
    up = (pad - (taps - 1), pad - (taps - 1)) if upper is None else upper
    v.lower_h, v.lower_w, v.upper_h, v.upper_w = lo[0], lo[1], up[0], up[1]
    v.stride, v.taps_h, v.taps_w = stride, taps, taps
    return v


# Pre-split weights (3xTF32): arenas registered here hold parameters whose low halves (x - trunc_tf32(x)) live at the same
# offset in a twin buffer, refreshed once per step by split_lo(); gemm_desc() then hands the GEMM the twin of any B operand
# that points into a registered arena (weights are the B operand of every forward / dgrad problem).  ``max_n`` restricts
# this to problems with N <= max_n: narrow tiles (N <= 64) are paced by the operand-splitter warps, whose B pass costs as
# much as for a 128-wide tile while the TMA moves only half the bytes, so fetching B_lo pays there -- and only there
# (profiles/r02_notes.md section 6).
_LO_ARENAS = []      # (base_ptr, end_ptr, lo_base_ptr, max_n)


def register_lo_arena(params: torch.Tensor, lo: torch.Tensor, max_n: int = 0) -> None:
    if params.numel() != lo.numel() or params.dtype != torch.float32 or lo.dtype != torch.float32:
        raise _lib.MvaeError("register_lo_arena: params / lo must be fp32 buffers of the same size")
    b0, e0 = params.data_ptr(), params.data_ptr() + params.numel() * 4
    l0, l1 = lo.data_ptr(), lo.data_ptr() + lo.numel() * 4
    # drop every older entry that overlaps the new buffers: its owner is gone and the allocator reused the memory (a stale
    # entry would hand a GEMM a B_lo pointer into somebody else's tensor)
    _LO_ARENAS[:] = [a for a in _LO_ARENAS
                     if not (a[0] < e0 and b0 < a[1]) and not (a[2] < l1 and l0 < a[2] + (a[1] - a[0]))
                     and not (a[0] < l1 and l0 < a[1]) and not (a[2] < e0 and b0 < a[2] + (a[1] - a[0]))]
    _LO_ARENAS.append((b0, e0, l0, int(max_n)))


def unregister_lo_arena(params: torch.Tensor) -> None:
    _LO_ARENAS[:] = [a for a in _LO_ARENAS if a[0] != params.data_ptr()]


def split_lo(x: torch.Tensor, lo: torch.Tensor) -> None:
    """lo = x - trunc_tf32(x), element-wise (the low halves of 3xTF32 operands)."""
    _lib.check(_lib.load().mvae_split_lo(x.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "mvae_split_lo")


def gemm_desc(A, B, Cmat, M, N, K, a_mn=False, b_mn=False, bias=None, aux=None, out2=None, epilogue=EPI_STORE,
              split_k=1, accumulate=False, colsum=None, split_ws=None, a_view=None, b_view=None, b_taps=None,
              rowmap=None) -> GemmDesc:
    """b_taps = (table, k_per_slot, mn_per_slot): tap-split B; rowmap = (IH, IW, s, py, px): output row map (sub-pixel
    transposed convolutions, see include/mvae_b200.h)."""
    d = GemmDesc()
    if b_taps is not None:
        table, k_per, mn_per = b_taps
        d.b_tap_slots, d.b_tap_k, d.b_tap_mn = len(table), k_per, mn_per
        for i, t in enumerate(table):
            d.b_tap_table[i] = int(t)
    if rowmap is not None:
        d.rowmap_IH, d.rowmap_IW, d.rowmap_s, d.rowmap_py, d.rowmap_px = rowmap
    bp = B.data_ptr()
    for base, end, lo_base, max_n in _LO_ARENAS:
        if base <= bp < end and b_view is None and (max_n == 0 or N <= max_n):
            d.B_lo = lo_base + (bp - base)
            break
    d.A, d.lda, d.a_mn_major = A.data_ptr(), (A.stride(0) if a_view is None else 0), int(a_mn)
    d.B, d.ldb, d.b_mn_major = B.data_ptr(), (B.stride(0) if b_view is None else 0), int(b_mn)
    if a_view is not None:
        d.a_view = a_view
    if b_view is not None:
        d.b_view = b_view
    d.M, d.N, d.K = M, N, K
    d.C, d.ldc = Cmat.data_ptr(), Cmat.stride(0)
    d.bias = _p(bias)
    d.aux, d.ldaux = _p(aux), (aux.stride(0) if aux is not None else 0)
    d.out2, d.ldout2 = _p(out2), (out2.stride(0) if out2 is not None else 0)
    d.colsum = _p(colsum)
    d.epilogue, d.split_k, d.accumulate = epilogue, split_k, int(accumulate)
    d.split_ws = _p(split_ws)      # fused split-K scratch [M, ceil4(N)] (zeroed once; see include/mvae_b200.h)
    return d


def subpixel_k4s2p1(x, w, out, n, IH, IW, Cx, Cy, w_is_conv=False, out2=None, aux=None, epilogue=EPI_STORE):
    """The four sub-pixel problems of a 4x4 / stride-2 / pad-1 TRANSPOSED convolution  y [n,2IH,2IW,Cy] = ConvT(x [n,IH,IW,Cx])
    -- ConvTranspose2d forward (w = Wt [(kh,kw,cy)][cx], K-major B) or Conv2d data gradient (w_is_conv: w = Wc [cx][(kh,kw,cy)],
    x = d out, y = d in; MN-major B) -- as implicit GEMMs: output pixel (2j+py, 2i+px) only ever sees the 2 x 2 taps
    kh = 3-2a / 2-2a (py = 0 / 1), kw likewise, of input pixels (j+ly+a, i+lx+b), ly = -1 / 0.  No cols matrix, no col2im:
    A = 2x2 stride-1 im2col view of x, B = the class's taps picked out of the full weight matrix (tap-split), rows stored
    straight at their output pixels (row map); `out2` / `aux` ([n,2IH,2IW,Cy] like out) follow the same map."""
    descs = []
    M = n * IH * IW
    for py in (0, 1):
        for px in (0, 1):
            ly, lx = (-1 if py == 0 else 0), (-1 if px == 0 else 0)
            table = []
            for a in (0, 1):
                for b in (0, 1):
                    kh = 3 - 2 * a if py == 0 else 2 - 2 * a
                    kw = 3 - 2 * b if px == 0 else 2 - 2 * b
                    table.append(kh * 4 + kw)
            view = conv_view(n, IH, IW, Cx, taps=2, stride=1, lower=(ly, lx), upper=(ly, lx))
            descs.append(gemm_desc(x, w, out, M, Cy, 4 * Cx, b_mn=w_is_conv, aux=aux, out2=out2, epilogue=epilogue, a_view=view,
                                   b_taps=(table, Cx, Cy), rowmap=(IH, IW, 2, py, px)))
    return descs


def full_k4s1p0(x, w, out, n, IH, IW, Cx, Cy, w_is_conv=False, out2=None, aux=None, epilogue=EPI_STORE):
    """4x4 / stride-1 / pad-0 TRANSPOSED convolution y [n,IH+3,IW+3,Cy] = ConvT(x [n,IH,IW,Cx]) (celeba/model.py:117 forward,
    :85 data gradient) as ONE implicit GEMM: output (oy, ox) sums x[oy - kh, ox - kw] W[kh, kw], i.e. a 4x4 stride-1 view
    with lower corner -3 whose tap (th, tw) meets filter tap (3 - th, 3 - tw): tap-split B with the table reversed."""
    view = conv_view(n, IH, IW, Cx, taps=4, stride=1, lower=(-3, -3), upper=(0, 0))
    table = [15 - t for t in range(16)]
    return [gemm_desc(x, w, out, n * (IH + 3) * (IW + 3), Cy, 16 * Cx, b_mn=w_is_conv, aux=aux, out2=out2, epilogue=epilogue,
                      a_view=view, b_taps=(table, Cx, Cy))]


def gemm_batch(descs: Sequence[GemmDesc], precision: int = PREC_3XTF32) -> None:
    arr = (GemmDesc * len(descs))(*descs)
    _lib.check(_lib.load().mvae_gemm_batch(arr, len(descs), precision, _stream()), "mvae_gemm_batch")


def gemm_chain(descs: Sequence[GemmDesc], deps: Sequence[int], ws: torch.Tensor, precision: int = PREC_3XTF32) -> None:
    """Up to 16 problems in one launch; problem i reads, as its A operand, the output of problem deps[i] < i (or -1).
    ``ws``: int32 CUDA workspace, zero-initialised once by the caller (see mvae_gemm_chain in include/mvae_b200.h)."""
    if len(deps) != len(descs):
        raise _lib.MvaeError("gemm_chain: one dependency entry per problem")
    if ws.dtype != torch.int32 or not ws.is_cuda or not ws.is_contiguous():
        raise _lib.MvaeError("gemm_chain: ws must be a contiguous CUDA int32 tensor")
    arr = (GemmDesc * len(descs))(*descs)
    dep = (C.c_int32 * len(deps))(*[int(d) for d in deps])
    _lib.check(_lib.load().mvae_gemm_chain(arr, dep, len(descs), ws.data_ptr(), ws.numel(), precision, _stream()),
               "mvae_gemm_chain")


def chain_workspace(device, ints: int = 1 << 18) -> torch.Tensor:
    """Counter workspace of mvae_gemm_chain: one int per 128-row block of every problem of a launch (+ 8 per output tile of a
    fused split-K problem); 1 MiB covers the largest launches of the conv flavours (4 sub-pixel problems of 3 B x 32 x 32
    rows at B = 1024: 24.6 K counters)."""
    return torch.zeros(ints, dtype=torch.int32, device=device)


def linear_fwd(x, w, bias, y, h=None, precision=PREC_3XTF32):
    """y = x @ w.T + bias ; optionally h = swish(y)."""
    _chk2d(x, "x"); _chk2d(w, "w"); _chk2d(y, "y")
    M, K = x.shape
    N = w.shape[0]
    _lib.check(_lib.load().mvae_linear_fwd(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _p(bias), y.data_ptr(),
                                           y.stride(0), _p(h), h.stride(0) if h is not None else 0, M, N, K, precision,
                                           _stream()), "mvae_linear_fwd")


def linear_dgrad(dy, w, dx, a_prev=None, accumulate=False, precision=PREC_3XTF32):
    """dx = dy @ w  (optionally * swish'(a_prev))."""
    _chk2d(dy, "dy"); _chk2d(w, "w"); _chk2d(dx, "dx")
    M, N = dy.shape
    K = w.shape[1]
    _lib.check(_lib.load().mvae_linear_dgrad(dy.data_ptr(), dy.stride(0), w.data_ptr(), w.stride(0), _p(a_prev),
                                             a_prev.stride(0) if a_prev is not None else 0, dx.data_ptr(), dx.stride(0),
                                             M, N, K, int(accumulate), precision, _stream()), "mvae_linear_dgrad")


def linear_wgrad(dy, x, dw, split_k=1, precision=PREC_3XTF32):
    """dw += dy.T @ x."""
    _chk2d(dy, "dy"); _chk2d(x, "x"); _chk2d(dw, "dw")
    M, N = dy.shape
    K = x.shape[1]
    _lib.check(_lib.load().mvae_linear_wgrad(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw.data_ptr(),
                                             dw.stride(0), M, N, K, split_k, precision, _stream()), "mvae_linear_wgrad")


def colsum_accumulate(dy, db):
    _chk2d(dy, "dy")
    M, N = dy.shape
    _lib.check(_lib.load().mvae_colsum_accumulate(dy.data_ptr(), dy.stride(0), db.data_ptr(), M, N, _stream()),
               "mvae_colsum_accumulate")


def swish_fwd(x, y):
    _lib.check(_lib.load().mvae_swish_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "mvae_swish_fwd")


def swish_bwd(x, dy, dx):
    _lib.check(_lib.load().mvae_swish_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _stream()),
               "mvae_swish_bwd")


def embedding_swish_fwd(table, idx, a, h):
    B = idx.numel()
    V, D = table.shape
    _lib.check(_lib.load().mvae_embedding_swish_fwd(table.data_ptr(), idx.data_ptr(), _p(a), h.data_ptr(), B, D, V,
                                                    _stream()), "mvae_embedding_swish_fwd")


def embedding_swish_bwd(table, idx, dh, dtable):
    B = idx.numel()
    V, D = table.shape
    _lib.check(_lib.load().mvae_embedding_swish_bwd(table.data_ptr(), idx.data_ptr(), dh.data_ptr(), dh.stride(0),
                                                    dtable.data_ptr(), B, D, V, _stream()), "mvae_embedding_swish_bwd")


def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _gather_array(gather, E):
    """gather: None or a list of E entries (int64 [B] CUDA tensor or None) -> (ctypes array or None)."""
    if gather is None or all(g is None for g in gather):
        return None
    if len(gather) != E:
        raise _lib.MvaeError("poe: one gather entry (index tensor or None) per expert")
    for g in gather:
        if g is not None and (g.dtype != torch.int64 or not g.is_cuda or not g.is_contiguous()):
            raise _lib.MvaeError("poe: gather indices must be contiguous CUDA int64 tensors")
    return (C.c_void_p * E)(*[None if g is None else g.data_ptr() for g in gather])


def poe_fwd(mu_e, lv_e, masks, B, L, z, variant=0, training=True, noise=None, noise_out=None, seed=0, offset=0,
            step_dev=None, mu_out=None, lv_out=None, kl_acc=None, gather=None):
    """``gather[e]`` (optional int64 [B]): expert e is a V-row table, sample b uses row gather[e][b] (label tables)."""
    E, P = len(mu_e), len(masks)
    m = (C.c_uint32 * P)(*masks)
    _lib.check(_lib.load().mvae_poe_fwd_g(_ptr_array(mu_e), _ptr_array(lv_e), mu_e[0].stride(0), E, _gather_array(gather, E),
                                          m, P, B, L, variant, int(training), _p(noise), _p(noise_out), seed, offset,
                                          _p(step_dev), z.data_ptr(), z.stride(0), _p(mu_out), _p(lv_out), _p(kl_acc),
                                          _stream()), "mvae_poe_fwd")


def poe_bwd(mu_e, lv_e, masks, B, L, dz, dmu_e, dlv_e, kl_scale, variant=0, training=True, noise=None,
            kl_scale_dev=None, dmu_up=None, dlv_up=None, gather=None):
    """With ``gather[e]`` the gradient of table expert e is ADDED into rows gather[e][b] of dmu_e[e] / dlv_e[e] (zero them)."""
    E, P = len(mu_e), len(masks)
    m = (C.c_uint32 * P)(*masks)
    _lib.check(_lib.load().mvae_poe_bwd_g(_ptr_array(mu_e), _ptr_array(lv_e), mu_e[0].stride(0), E, _gather_array(gather, E),
                                          m, P, B, L, variant, int(training), _p(noise), dz.data_ptr(), dz.stride(0),
                                          _p(dmu_up), _p(dlv_up), float(kl_scale), _p(kl_scale_dev), _ptr_array(dmu_e),
                                          _ptr_array(dlv_e), dmu_e[0].stride(0), _stream()), "mvae_poe_bwd")


def label_table_fwd(emb, w2, b2, w3, b3, a2, h2, tab):
    """Label encoder on its V-row table: a2 = swish(emb) w2^T + b2, h2 = swish(a2), tab = h2 w3^T + b3."""
    V, D = emb.shape
    N3 = w3.shape[0]
    _lib.check(_lib.load().mvae_label_table_fwd(emb.data_ptr(), w2.data_ptr(), _p(b2), w3.data_ptr(), _p(b3), a2.data_ptr(),
                                                h2.data_ptr(), tab.data_ptr(), V, D, N3, _stream()), "mvae_label_table_fwd")


def label_table_bwd(emb, w2, w3, a2, h2, dtab, d_a2, d_emb, dw2, db2, dw3, db3):
    V, D = emb.shape
    N3 = w3.shape[0]
    _lib.check(_lib.load().mvae_label_table_bwd(emb.data_ptr(), w2.data_ptr(), w3.data_ptr(), a2.data_ptr(), h2.data_ptr(),
                                                dtab.data_ptr(), d_a2.data_ptr(), d_emb.data_ptr(), dw2.data_ptr(), _p(db2),
                                                dw3.data_ptr(), _p(db3), V, D, N3, _stream()), "mvae_label_table_bwd")


def reparam_fwd(mu, logvar, z, noise=None, noise_out=None, seed=0, offset=0):
    _lib.check(_lib.load().mvae_reparam_fwd(mu.data_ptr(), logvar.data_ptr(), _p(noise), _p(noise_out), seed, offset,
                                            z.data_ptr(), mu.numel(), _stream()), "mvae_reparam_fwd")


def reparam_bwd(logvar, noise, dz, dlogvar):
    _lib.check(_lib.load().mvae_reparam_bwd(logvar.data_ptr(), noise.data_ptr(), dz.data_ptr(), dlogvar.data_ptr(),
                                            logvar.numel(), _stream()), "mvae_reparam_bwd")


def kl_fwd_bwd(mu, logvar, dmu, dlogvar, scale, kl_acc=None):
    _lib.check(_lib.load().mvae_kl_fwd_bwd(mu.data_ptr(), logvar.data_ptr(), _p(dmu), _p(dlogvar), mu.numel(),
                                           float(scale), _p(kl_acc), _stream()), "mvae_kl_fwd_bwd")


def bce_logits_fwd_bwd(x, t, dx, scale, loss_acc=None, seg_rows=0, loss_elem=None):
    """x [R,D] logits, t [t_rows,D] targets (row r uses t[r % t_rows])."""
    R, D = x.shape
    _lib.check(_lib.load().mvae_bce_logits_fwd_bwd(x.data_ptr(), x.stride(0), t.data_ptr(), t.stride(0), t.shape[0],
                                                   _p(dx), dx.stride(0) if dx is not None else 0, R, D, float(scale),
                                                   _p(loss_acc), seg_rows, _p(loss_elem),
                                                   loss_elem.stride(0) if loss_elem is not None else 0, _stream()),
               "mvae_bce_logits_fwd_bwd")


def ce_fwd_bwd(x, target, dx, K, scale, loss_acc=None, seg_rows=0, loss_rows=None):
    R = x.shape[0]
    _lib.check(_lib.load().mvae_ce_fwd_bwd(x.data_ptr(), x.stride(0), target.data_ptr(), target.numel(), _p(dx),
                                           dx.stride(0) if dx is not None else 0, R, K, float(scale), _p(loss_acc),
                                           seg_rows, _p(loss_rows), loss_rows.stride(0) if loss_rows is not None else 0,
                                           _stream()), "mvae_ce_fwd_bwd")


def im2col_k4s2p1(x, cols, B, H, W, Cch):
    """x NHWC [B,H,W,C] (contiguous) -> cols [B*(H/2)*(W/2), 16*C]."""
    _lib.check(_lib.load().mvae_im2col_k4s2p1(x.data_ptr(), cols.data_ptr(), cols.stride(0), B, H, W, Cch, _stream()),
               "mvae_im2col_k4s2p1")


def col2im_k4s2p1(cols, out, B, IH, IW, Cch, out_act=None, aux=None):
    """cols [B*IH*IW, 16*C] -> out NHWC [B,2IH,2IW,C]; optional out_act = swish(out) or out *= swish'(aux)."""
    _lib.check(_lib.load().mvae_col2im_k4s2p1(cols.data_ptr(), cols.stride(0), out.data_ptr(), _p(out_act), _p(aux), B,
                                              IH, IW, Cch, _stream()), "mvae_col2im_k4s2p1")


def im2col_k4(x, cols, B, H, W, Cch, stride, pad):
    _lib.check(_lib.load().mvae_im2col_k4(x.data_ptr(), cols.data_ptr(), cols.stride(0), B, H, W, Cch, stride, pad,
                                          _stream()), "mvae_im2col_k4")


def col2im_k4(cols, out, B, IH, IW, Cch, stride, pad, out_act=None, aux=None):
    _lib.check(_lib.load().mvae_col2im_k4(cols.data_ptr(), cols.stride(0), out.data_ptr(), _p(out_act), _p(aux), B, IH, IW,
                                          Cch, stride, pad, _stream()), "mvae_col2im_k4")


def conv_cin_fwd(x, wc, a, h, B, H, W, Cin, Cout):
    """Direct Conv2d(k4 s2 p1, Cin in {1,3} -> Cout) + Swish: x NHWC -> a (pre-activation), h = swish(a) [B*H/2*W/2, Cout]."""
    _lib.check(_lib.load().mvae_conv_k4s2p1_cin_fwd(x.data_ptr(), wc.data_ptr(), a.data_ptr(), h.data_ptr(), B, H, W, Cin, Cout,
                                                    _stream()), "mvae_conv_k4s2p1_cin_fwd")


def conv_cin_wgrad(x, da, dwc, B, H, W, Cin, Cout):
    """dwc[Cout, 16 Cin] += da^T im2col(x) without materialising im2col."""
    _lib.check(_lib.load().mvae_conv_k4s2p1_cin_wgrad(x.data_ptr(), da.data_ptr(), dwc.data_ptr(), B, H, W, Cin, Cout,
                                                      _stream()), "mvae_conv_k4s2p1_cin_wgrad")


def convT_cout_fwd(hin, wt, out, B, IH, IW, Cin, Cout):
    """Direct ConvTranspose2d(k4 s2 p1, Cin -> Cout in {1,3}): hin NHWC [B,IH,IW,Cin] -> out [B,2IH,2IW,Cout]."""
    _lib.check(_lib.load().mvae_convt_k4s2p1_cout_fwd(hin.data_ptr(), wt.data_ptr(), out.data_ptr(), B, IH, IW, Cin, Cout,
                                                      _stream()), "mvae_convt_k4s2p1_cout_fwd")


def convT_cout_bwd(dout, hin, ain, wt, dhin, dwt, B, IH, IW, Cin, Cout):
    """Backward of the above through the Swish below it: dhin = (ConvT^T dout) * swish'(ain), dwt += dcols^T hin."""
    _lib.check(_lib.load().mvae_convt_k4s2p1_cout_bwd(dout.data_ptr(), hin.data_ptr(), ain.data_ptr(), wt.data_ptr(),
                                                      dhin.data_ptr(), dwt.data_ptr(), B, IH, IW, Cin, Cout, _stream()),
               "mvae_convt_k4s2p1_cout_bwd")


def bn_forward(x, h, S, seg_rows, gamma, beta, mean, invstd, acc, running_mean=None, running_var=None, update_order=(),
               training=True, act=True, eps=1e-5, momentum=0.1):
    """Train/eval BatchNorm (+Swish) over x [S*seg_rows, C] -> h; fills mean/invstd [S, C]."""
    lib = _lib.load()
    Cch = x.shape[1]
    if training:
        _lib.check(lib.mvae_bn_stats(x.data_ptr(), x.stride(0), S, seg_rows, Cch, acc.data_ptr(), _stream()), "mvae_bn_stats")
        order = (C.c_int32 * max(len(update_order), 1))(*update_order)
        _lib.check(lib.mvae_bn_finalize(acc.data_ptr(), S, seg_rows, Cch, eps, momentum, mean.data_ptr(), invstd.data_ptr(),
                                        _p(running_mean), _p(running_var), order, len(update_order), _stream()),
                   "mvae_bn_finalize")
    else:
        _lib.check(lib.mvae_bn_eval_stats(running_mean.data_ptr(), running_var.data_ptr(), S, Cch, eps, mean.data_ptr(),
                                          invstd.data_ptr(), _stream()), "mvae_bn_eval_stats")
    _lib.check(lib.mvae_bn_apply(x.data_ptr(), x.stride(0), h.data_ptr(), h.stride(0), S * seg_rows, seg_rows, Cch,
                                 mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), int(act), _stream()),
               "mvae_bn_apply")


def bn_backward(x, dh, dx, S, seg_rows, seg0, nseg, gamma, beta, mean, invstd, acc2, dgamma, dbeta, act=True,
                training=True):
    Cch = x.shape[1]
    _lib.check(_lib.load().mvae_bn_bwd(x.data_ptr(), x.stride(0), dh.data_ptr(), dh.stride(0), dx.data_ptr(), dx.stride(0), S,
                                       seg_rows, Cch, seg0, nseg, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                       beta.data_ptr(), int(act), int(training), acc2.data_ptr(), dgamma.data_ptr(),
                                       dbeta.data_ptr(),
                                       _stream()), "mvae_bn_bwd")


def dropout_fwd(x, y, copies, p, mask_out=None, mask_in=None, seed=0, step_dev=None):
    x_rows, D = x.shape
    _lib.check(_lib.load().mvae_dropout_fwd(x.data_ptr(), x_rows, y.data_ptr(), _p(mask_out), _p(mask_in), copies, D, p, seed,
                                            _p(step_dev), _stream()), "mvae_dropout_fwd")


def dropout_bwd(dy, mask, dx, copies, p):
    x_rows, D = dx.shape
    _lib.check(_lib.load().mvae_dropout_bwd(dy.data_ptr(), mask.data_ptr(), dx.data_ptr(), x_rows, copies, D, p, _stream()),
               "mvae_dropout_bwd")


def nchw_to_nhwc(x, y, B, Cch, HW):
    _lib.check(_lib.load().mvae_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), B, Cch, HW, _stream()), "mvae_nchw_to_nhwc")


def gather_batch_u8(data_u8, labels, idx, out, labels_out=None):
    """out[b, :] = data_u8[idx[b], :] / 255 ; labels_out[b] = labels[idx[b]] (device-resident dataset -> batch)."""
    if data_u8.dtype != torch.uint8 or data_u8.dim() != 2 or not data_u8.is_contiguous() or idx.dtype != torch.int64:
        raise _lib.MvaeError("gather_batch_u8: data must be a contiguous [N, D] uint8 tensor and idx int64")
    _lib.check(_lib.load().mvae_gather_batch_u8(data_u8.data_ptr(), data_u8.shape[1], _p(labels), idx.data_ptr(), idx.numel(),
                                                out.data_ptr(), out.stride(0), _p(labels_out), _stream()),
               "mvae_gather_batch_u8")


def adam_flat(p, g, m, v, step_count, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, lr_mult_dev=None):
    _lib.check(_lib.load().mvae_adam_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr,
                                          _p(lr_mult_dev), beta1, beta2, eps, grad_scale, step_count.data_ptr(),
                                          _stream()), "mvae_adam_flat")


def elbo_finalize(recon_img, recon_txt, kl, P, lambda_image, lambda_text, beta, inv_batch, out, beta_dev=None):
    _lib.check(_lib.load().mvae_elbo_finalize(_p(recon_img), _p(recon_txt), _p(kl), P, lambda_image, lambda_text, beta,
                                              _p(beta_dev), inv_batch, out.data_ptr(), _stream()), "mvae_elbo_finalize")


def allreduce_adam_p2p(grad_ptrs, param_ptrs, flag_ptrs, m, v, n, tail, tail_out, rank, world, step_count, lr,
                       lr_mult_dev=None, beta1=0.9, beta2=0.999, eps=1e-8):
    """Fused NVLink reduce-scatter + Adam + all-gather (one launch per rank); the pointer lists are the peer-mapped base
    addresses of every rank's gradient / parameter / flag buffers in rank order (see include/mvae_b200.h)."""
    arr = lambda ptrs: (C.c_void_p * len(ptrs))(*[int(x) for x in ptrs])  # noqa: E731
    _lib.check(_lib.load().mvae_allreduce_adam_p2p(arr(grad_ptrs), arr(param_ptrs), arr(flag_ptrs), m.data_ptr(), v.data_ptr(),
                                                   n, tail, _p(tail_out), rank, world, lr, _p(lr_mult_dev), beta1, beta2,
                                                   eps, step_count.data_ptr(), _stream()), "mvae_allreduce_adam_p2p")
