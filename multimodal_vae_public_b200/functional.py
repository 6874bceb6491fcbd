"""torch.autograd.Function wrappers over the C-ABI kernels: the building blocks of the drop-in
``MVAE`` modules (``mnist/model.py``-style surface).  Forward AND backward run in libmvae_b200.so;
there is no eager fallback -- CPU tensors raise ``MvaeError``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch.autograd import Function

from . import _lib, ops

DEFAULT_PRECISION = ops.PREC_3XTF32


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MvaeError("mvae_b200 modules run on CUDA tensors only (no CPU fallback for the hot path)")


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _padded(rows: int, cols: int, like: torch.Tensor, zero: bool = False) -> torch.Tensor:
    """[rows, cols] fp32 view whose row stride is a multiple of 4 elements (TMA requirement)."""
    mk = torch.zeros if zero else torch.empty
    return mk(rows, _pad4(cols), dtype=torch.float32, device=like.device)[:, :cols]


def _as_operand(t: torch.Tensor) -> torch.Tensor:
    """fp32, unit inner stride, row stride % 4 == 0, 16-byte aligned -- copy into a padded buffer if not."""
    t = t.to(torch.float32)
    if t.dim() != 2:
        t = t.reshape(t.shape[0], -1)
    if t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    out = _padded(t.shape[0], t.shape[1], t)
    out.copy_(t)
    return out


def _split_k(m_rows: int) -> int:
    return max(1, min(m_rows // 256, 16))


class _LinearFn(Function):
    """y = x W^T + b  (nn.Linear.forward, mnist/model.py:81-84) and, with ``act``, Swish fused in the epilogue."""

    @staticmethod
    def forward(ctx, x, w, b, act: bool, prec: int):
        _need_cuda(x, w, b)
        x = _as_operand(x.detach()); w2 = _as_operand(w.detach())
        M, N = x.shape[0], w2.shape[0]
        a = _padded(M, N, x)
        h = _padded(M, N, x) if act else None
        ops.linear_fwd(x, w2, b.detach() if b is not None else None, a, h, prec)
        ctx.act, ctx.prec, ctx.has_bias = act, prec, b is not None
        ctx.save_for_backward(x, w2, a if act else None)
        return h if act else a

    @staticmethod
    def backward(ctx, dy):
        x, w, a = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        dy = _as_operand(dy)
        if ctx.act:
            da = _padded(M, N, x)
            if da.stride(0) == N and dy.stride(0) == N and a.stride(0) == N:
                ops.swish_bwd(a, dy, da)
            else:  # ragged widths never carry an activation in the MVAE nets; keep a general path anyway
                da.copy_(dy * torch.sigmoid(a) * (1 + a * (1 - torch.sigmoid(a))))
            dy = da
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _padded(M, K, x)
            ops.linear_dgrad(dy, w, dx, precision=ctx.prec)
        if ctx.needs_input_grad[1]:
            dwp = _padded(N, K, x, zero=True)
            ops.linear_wgrad(dy, x, dwp, split_k=_split_k(M), precision=ctx.prec)
            dw = dwp
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, dtype=torch.float32, device=x.device)
            ops.colsum_accumulate(dy, db)
        return dx, dw, db, None, None


def linear(x, w, b=None, precision: int = DEFAULT_PRECISION):
    return _LinearFn.apply(x, w, b, False, precision)


def linear_swish(x, w, b, precision: int = DEFAULT_PRECISION):
    return _LinearFn.apply(x, w, b, True, precision)


def _conv_weight_cols(w: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [Cout,Cin,4,4] -> GEMM operand [Cout, (kh,kw,ci)]."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _convT_weight_cols(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight [Cin,Cout,4,4] -> GEMM operand [(kh,kw,co), Cin]."""
    return w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous()


class _Conv4x4s2Fn(Function):
    """nn.Conv2d(Cin, Cout, 4, 2, 1, bias=False) on NCHW tensors (fashionmnist/model.py:79-82), optional fused Swish.
    NHWC internally: im2col kernel + tcgen05 GEMM; backward = GEMM dgrad + col2im gather, GEMM wgrad."""

    @staticmethod
    def forward(ctx, x, w, act: bool, prec: int, stride: int = 2, pad: int = 1, nhwc: bool = False):
        _need_cuda(x, w)
        if nhwc:
            B, H, W, Cin = x.shape
            xh = x.detach().to(torch.float32).contiguous()
        else:
            B, Cin, H, W = x.shape
            xh = x.detach().to(torch.float32).permute(0, 2, 3, 1).contiguous()
        Cout = w.shape[0]
        OH, OW = (H + 2 * pad - 4) // stride + 1, (W + 2 * pad - 4) // stride + 1
        cols = _padded(B * OH * OW, 16 * Cin, xh)
        ops.im2col_k4(xh, cols, B, H, W, Cin, stride, pad)
        wp = _as_operand(_conv_weight_cols(w.detach().to(torch.float32)))
        a = _padded(B * OH * OW, Cout, xh)
        h = _padded(B * OH * OW, Cout, xh) if act else None
        ops.gemm_batch([ops.gemm_desc(cols, wp, a, B * OH * OW, Cout, 16 * Cin, out2=h,
                                      epilogue=ops.EPI_BIAS_SWISH if act else ops.EPI_STORE)], prec)
        ctx.save_for_backward(cols, wp, a if act else None)
        ctx.meta = (B, Cin, H, W, Cout, act, prec, stride, pad, nhwc, OH, OW)
        out = (h if act else a).reshape(B, OH, OW, Cout)
        return out if nhwc else out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        cols, wp, a = ctx.saved_tensors
        B, Cin, H, W, Cout, act, prec, stride, pad, nhwc, OH, OW = ctx.meta
        M = B * OH * OW
        dyh = dy.to(torch.float32) if nhwc else dy.to(torch.float32).permute(0, 2, 3, 1)
        d = _as_operand(dyh.reshape(M, Cout))
        if act:
            da = _padded(M, Cout, d)
            if d.stride(0) == Cout and a.stride(0) == Cout:
                ops.swish_bwd(a, d.contiguous(), da)
            else:
                s = torch.sigmoid(a)
                da.copy_(d * s * (1 + a * (1 - s)))
            d = da
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dwp = _padded(Cout, 16 * Cin, d, zero=True)
            ops.gemm_batch([ops.gemm_desc(d, cols, dwp, Cout, 16 * Cin, M, a_mn=True, b_mn=True, split_k=_split_k(M),
                                          accumulate=True)], prec)
            dw = dwp.reshape(Cout, 4, 4, Cin).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[0]:
            dcols = _padded(M, 16 * Cin, d)
            ops.gemm_batch([ops.gemm_desc(d, wp, dcols, M, 16 * Cin, Cout, b_mn=True)], prec)
            dxh = torch.empty(B, H, W, Cin, dtype=torch.float32, device=d.device)
            ops.col2im_k4(dcols, dxh, B, OH, OW, Cin, stride, pad)
            dx = dxh if nhwc else dxh.permute(0, 3, 1, 2)
        return dx, dw, None, None, None, None, None


class _ConvT4x4s2Fn(Function):
    """nn.ConvTranspose2d(Cin, Cout, 4, 2, 1, bias=False) on NCHW tensors (fashionmnist/model.py:112-114), optional
    fused Swish: tcgen05 GEMM + col2im gather; backward = im2col + GEMM dgrad / wgrad."""

    @staticmethod
    def forward(ctx, x, w, act: bool, prec: int, stride: int = 2, pad: int = 1, nhwc: bool = False):
        _need_cuda(x, w)
        if nhwc:
            B, IH, IW, Cin = x.shape
            xh = _as_operand(x.detach().to(torch.float32).reshape(B * IH * IW, Cin))
        else:
            B, Cin, IH, IW = x.shape
            xh = _as_operand(x.detach().to(torch.float32).permute(0, 2, 3, 1).reshape(B * IH * IW, Cin))
        Cout = w.shape[1]
        OH, OW = (IH - 1) * stride - 2 * pad + 4, (IW - 1) * stride - 2 * pad + 4
        wp = _as_operand(_convT_weight_cols(w.detach().to(torch.float32)))
        cols = _padded(B * IH * IW, 16 * Cout, xh)
        ops.gemm_batch([ops.gemm_desc(xh, wp, cols, B * IH * IW, 16 * Cout, Cin)], prec)
        a = torch.empty(B, OH, OW, Cout, dtype=torch.float32, device=xh.device)
        h = torch.empty_like(a) if act else None
        ops.col2im_k4(cols, a, B, IH, IW, Cout, stride, pad, out_act=h)
        ctx.save_for_backward(xh, wp, a if act else None)
        ctx.meta = (B, Cin, IH, IW, Cout, act, prec, stride, pad, nhwc, OH, OW)
        out = h if act else a
        return out if nhwc else out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, wp, a = ctx.saved_tensors
        B, Cin, IH, IW, Cout, act, prec, stride, pad, nhwc, OH, OW = ctx.meta
        M = B * IH * IW
        d = (dy.to(torch.float32) if nhwc else dy.to(torch.float32).permute(0, 2, 3, 1)).contiguous()
        if act:
            da = torch.empty_like(d)
            ops.swish_bwd(a, d, da)
            d = da
        dcols = _padded(M, 16 * Cout, d)
        ops.im2col_k4(d, dcols, B, OH, OW, Cout, stride, pad)
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dwp = _padded(16 * Cout, Cin, d, zero=True)
            ops.gemm_batch([ops.gemm_desc(dcols, xh, dwp, 16 * Cout, Cin, M, a_mn=True, b_mn=True, split_k=_split_k(M),
                                          accumulate=True)], prec)
            dw = dwp.reshape(4, 4, Cout, Cin).permute(3, 2, 0, 1)
        if ctx.needs_input_grad[0]:
            dxh = _padded(M, Cin, d)
            ops.gemm_batch([ops.gemm_desc(dcols, wp, dxh, M, Cin, 16 * Cout, b_mn=True)], prec)
            dx = dxh.reshape(B, IH, IW, Cin)
            dx = dx if nhwc else dx.permute(0, 3, 1, 2)
        return dx, dw, None, None, None, None, None


def conv4x4s2(x, w, swish_act: bool = False, precision: int = DEFAULT_PRECISION):
    return _Conv4x4s2Fn.apply(x, w, swish_act, precision)


def conv_transpose4x4s2(x, w, swish_act: bool = False, precision: int = DEFAULT_PRECISION):
    return _ConvT4x4s2Fn.apply(x, w, swish_act, precision)


def conv4x4(x, w, stride: int, pad: int, swish_act: bool = False, nhwc: bool = False, precision: int = DEFAULT_PRECISION):
    """nn.Conv2d(Cin, Cout, 4, stride, pad, bias=False); NCHW in/out, or NHWC in/out with ``nhwc=True``."""
    return _Conv4x4s2Fn.apply(x, w, swish_act, precision, stride, pad, nhwc)


def conv_transpose4x4(x, w, stride: int, pad: int, swish_act: bool = False, nhwc: bool = False,
                      precision: int = DEFAULT_PRECISION):
    """nn.ConvTranspose2d(Cin, Cout, 4, stride, pad, bias=False); NCHW in/out, or NHWC in/out with ``nhwc=True``."""
    return _ConvT4x4s2Fn.apply(x, w, swish_act, precision, stride, pad, nhwc)


class _BatchNormActFn(Function):
    """BatchNorm{1,2}d (train: batch statistics + running-stat update, eval: running statistics) + optional Swish over
    channels-last activations [..., C] (celeba/model.py:80,149, ...)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training: bool, act: bool):
        _need_cuda(x, weight, bias)
        Cc = x.shape[-1]
        x2 = x.detach().to(torch.float32).reshape(-1, Cc).contiguous()
        R = x2.shape[0]
        h = torch.empty_like(x2)
        mean = torch.empty(1, Cc, dtype=torch.float32, device=x2.device); invstd = torch.empty_like(mean)
        acc = torch.empty(2 * Cc, dtype=torch.float64, device=x2.device)
        ops.bn_forward(x2, h, 1, R, weight.detach(), bias.detach(), mean, invstd, acc, running_mean, running_var,
                       update_order=(0,), training=training, act=act)
        ctx.save_for_backward(x2, weight.detach(), bias.detach(), mean, invstd)
        ctx.meta = (x.shape, act, training)
        return h.reshape(x.shape)

    @staticmethod
    def backward(ctx, dh):
        x2, w, b, mean, invstd = ctx.saved_tensors
        shape, act, training = ctx.meta
        Cc = x2.shape[1]
        d = dh.to(torch.float32).reshape(-1, Cc).contiguous()
        dx = torch.empty_like(x2)
        dg = torch.zeros(Cc, dtype=torch.float32, device=x2.device); db = torch.zeros_like(dg)
        acc = torch.empty(2 * Cc, dtype=torch.float64, device=x2.device)
        ops.bn_backward(x2, d, dx, 1, x2.shape[0], 0, 1, w, b, mean, invstd, acc, dg, db, act=act, training=training)
        return dx.reshape(shape), dg, db, None, None, None, None


def batch_norm_act(x, weight, bias, running_mean, running_var, training: bool, swish_act: bool = True):
    """x channels-last [..., C]."""
    return _BatchNormActFn.apply(x, weight, bias, running_mean, running_var, training, swish_act)


class _DropoutFn(Function):
    """nn.Dropout(p) in training mode (celeba/model.py:91)."""

    @staticmethod
    def forward(ctx, x, p: float, seed: int):
        _need_cuda(x)
        x2 = x.detach().to(torch.float32).reshape(x.shape[0], -1).contiguous()
        y = torch.empty_like(x2); mask = torch.empty_like(x2)
        ops.dropout_fwd(x2, y, 1, p, mask_out=mask, seed=seed)
        ctx.save_for_backward(mask)
        ctx.p = p
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        d = dy.to(torch.float32).reshape(mask.shape).contiguous()
        dx = torch.empty_like(d)
        ops.dropout_bwd(d, mask, dx, 1, ctx.p)
        return dx.reshape(dy.shape), None, None


_dropout_counter = [0]


def dropout(x, p: float, training: bool):
    if not training or p == 0.0:
        return x
    _dropout_counter[0] += 1
    seed = (int(torch.initial_seed()) * 1000003 + _dropout_counter[0]) & 0x7FFFFFFF
    return _DropoutFn.apply(x, p, seed)


class _SwishFn(Function):
    """x * sigmoid(x)  (Swish, mnist/model.py:166-169)."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        xc = x.detach().to(torch.float32).contiguous()
        y = torch.empty_like(xc)
        ops.swish_fwd(xc, y)
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        dx = torch.empty_like(xc)
        ops.swish_bwd(xc, dy.to(torch.float32).contiguous(), dx)
        return dx


def swish(x):
    return _SwishFn.apply(x)


class _EmbeddingSwishFn(Function):
    """swish(table[idx])  (nn.Embedding + Swish, mnist/model.py:116,123)."""

    @staticmethod
    def forward(ctx, idx, table):
        _need_cuda(idx, table)
        t = table.detach().to(torch.float32).contiguous()
        idx = idx.detach().to(torch.int64).contiguous().view(-1)
        h = torch.empty(idx.numel(), t.shape[1], dtype=torch.float32, device=t.device)
        ops.embedding_swish_fwd(t, idx, None, h)
        ctx.save_for_backward(idx, t)
        return h

    @staticmethod
    def backward(ctx, dh):
        idx, t = ctx.saved_tensors
        dt = torch.zeros_like(t)
        ops.embedding_swish_bwd(t, idx, _as_operand(dh), dt)
        return None, dt


def embedding_swish(idx, table):
    return _EmbeddingSwishFn.apply(idx, table)


class _PoEFn(Function):
    """Product of Gaussian experts -> (mu, logvar).  ``with_prior``: the N(0,1) prior expert is implicit
    (MVAE.infer, mnist/model.py:46-64); otherwise the given experts are the whole product
    (ProductOfExperts.forward on an explicit stack, mnist/model.py:156-163)."""

    @staticmethod
    def forward(ctx, variant: int, with_prior: bool, n_experts: int, *tensors):
        mus = [t.detach().to(torch.float32).contiguous() for t in tensors[:n_experts]]
        lvs = [t.detach().to(torch.float32).contiguous() for t in tensors[n_experts:]]
        _need_cuda(*mus, *lvs)
        B, L = mus[0].shape
        code = variant | (0 if with_prior else 2)
        z = torch.empty(B, L, dtype=torch.float32, device=mus[0].device)
        mu = torch.empty_like(z); lv = torch.empty_like(z)
        mask = (1 << n_experts) - 1
        ops.poe_fwd(mus, lvs, [mask], B, L, z, variant=code, training=False, mu_out=mu, lv_out=lv)
        ctx.code, ctx.n = code, n_experts
        ctx.save_for_backward(*mus, *lvs)
        return mu, lv

    @staticmethod
    def backward(ctx, dmu, dlv):
        n = ctx.n
        saved = ctx.saved_tensors
        mus, lvs = list(saved[:n]), list(saved[n:])
        B, L = mus[0].shape
        gm = [torch.empty_like(m) for m in mus]; gl = [torch.empty_like(m) for m in mus]
        dz = torch.zeros(B, L, dtype=torch.float32, device=mus[0].device)
        ops.poe_bwd(mus, lvs, [(1 << n) - 1], B, L, dz, gm, gl, kl_scale=0.0, variant=ctx.code, training=False,
                    dmu_up=dmu.to(torch.float32).contiguous() if dmu is not None else None,
                    dlv_up=dlv.to(torch.float32).contiguous() if dlv is not None else None)
        return (None, None, None, *gm, *gl)


def product_of_experts(mus: Sequence[torch.Tensor], logvars: Sequence[torch.Tensor], variant: int = 0,
                       with_prior: bool = True):
    if len(mus) > 20:
        raise _lib.MvaeError("at most 20 experts are supported")
    return _PoEFn.apply(variant, with_prior, len(mus), *mus, *logvars)


class _ReparamFn(Function):
    """z = eps * exp(0.5 logvar) + mu with eps ~ N(0,1) (MVAE.reparametrize training branch, mnist/model.py:29-33)."""

    @staticmethod
    def forward(ctx, mu, logvar, noise, seed, offset):
        _need_cuda(mu, logvar)
        m = mu.detach().to(torch.float32).contiguous(); lv = logvar.detach().to(torch.float32).contiguous()
        z = torch.empty_like(m)
        if noise is None:
            noise = torch.empty_like(m)
            ops.reparam_fwd(m, lv, z, noise=None, noise_out=noise, seed=seed, offset=offset)
        else:
            noise = noise.detach().to(torch.float32).contiguous()
            ops.reparam_fwd(m, lv, z, noise=noise)
        ctx.save_for_backward(lv, noise)
        return z

    @staticmethod
    def backward(ctx, dz):
        lv, noise = ctx.saved_tensors
        dz = dz.to(torch.float32).contiguous()
        dlv = torch.empty_like(lv)
        ops.reparam_bwd(lv, noise, dz, dlv)
        return dz, dlv, None, None, None


_noise_counter = [0]


def reparametrize(mu, logvar, noise: Optional[torch.Tensor] = None, seed: Optional[int] = None):
    if seed is None:
        seed = int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)
    off = _noise_counter[0]
    _noise_counter[0] += mu.numel()
    return _ReparamFn.apply(mu, logvar, noise, seed, off)


class _BceSumFn(Function):
    """sum over ALL elements of binary_cross_entropy_with_logits (mnist/train.py:62-74); gradient
    sigmoid(x) - t computed in the same pass."""

    @staticmethod
    def forward(ctx, x, t):
        _need_cuda(x, t)
        if t.size() != x.size():
            raise ValueError("Target size ({}) must be the same as input size ({})".format(t.size(), x.size()))
        x2 = _as_operand(x.detach().reshape(x.shape[0], -1) if x.dim() > 1 else x.detach().reshape(1, -1))
        t2 = _as_operand(t.detach().reshape(x2.shape))
        acc = torch.zeros(1, dtype=torch.float64, device=x.device)
        dx = torch.empty(x2.shape, dtype=torch.float32, device=x.device)
        ops.bce_logits_fwd_bwd(x2, t2, dx, 1.0, acc)
        ctx.save_for_backward(dx)
        ctx.shape = x.shape
        return acc[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return (dx * g).reshape(ctx.shape), None


class _BceElemFn(Function):
    """Element-wise binary_cross_entropy_with_logits (same-shape output, mnist/train.py:62-74)."""

    @staticmethod
    def forward(ctx, x, t):
        _need_cuda(x, t)
        if t.size() != x.size():
            raise ValueError("Target size ({}) must be the same as input size ({})".format(t.size(), x.size()))
        n = x.numel()
        pad = (-n) % 4
        xf = torch.zeros(n + pad, dtype=torch.float32, device=x.device); xf[:n] = x.detach().reshape(-1)
        tf = torch.zeros(n + pad, dtype=torch.float32, device=x.device); tf[:n] = t.detach().reshape(-1)
        le = torch.empty_like(xf); dx = torch.empty_like(xf)
        ops.bce_logits_fwd_bwd(xf.view(1, -1), tf.view(1, -1), dx.view(1, -1), 1.0, None, loss_elem=le.view(1, -1))
        ctx.save_for_backward(dx[:n].view(x.shape))
        return le[:n].view(x.shape)

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None


class _CeRowsFn(Function):
    """cross_entropy(input, target, eps=1e-6) -> [N, K] = -onehot * log_softmax (mnist/train.py:77-94)."""

    @staticmethod
    def forward(ctx, x, target):
        _need_cuda(x, target)
        if target.size(0) != x.size(0):
            raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(0), x.size(0)))
        x2 = x.detach().to(torch.float32).contiguous()
        R, K = x2.shape
        out = torch.empty(R, K, dtype=torch.float32, device=x.device); dx = torch.empty_like(out)
        ops.ce_fwd_bwd(x2, target.detach().to(torch.int64).contiguous(), dx, K, 1.0, None, loss_rows=out)
        ctx.save_for_backward(dx, target.detach().to(torch.int64))
        return out

    @staticmethod
    def backward(ctx, g):
        dx, tg = ctx.saved_tensors
        # d/dx sum_k g[r,k]*out[r,k] = g[r,target_r] * (softmax - onehot)
        return dx * g.gather(1, tg.view(-1, 1)), None


class _CeSumFn(Function):
    """sum over rows of cross_entropy(input, target, eps=1e-6) (mnist/train.py:77-94)."""

    @staticmethod
    def forward(ctx, x, target):
        _need_cuda(x, target)
        if target.size(0) != x.size(0):
            raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(0), x.size(0)))
        x2 = x.detach().to(torch.float32)
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        R, K = x2.shape
        acc = torch.zeros(1, dtype=torch.float64, device=x.device)
        dx = torch.empty(R, K, dtype=torch.float32, device=x.device)
        ops.ce_fwd_bwd(x2, target.detach().to(torch.int64).contiguous(), dx, K, 1.0, acc)
        ctx.save_for_backward(dx)
        return acc[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None


class _KlSumFn(Function):
    """sum over the batch of KLD = -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) (mnist/train.py:56)."""

    @staticmethod
    def forward(ctx, mu, logvar):
        _need_cuda(mu, logvar)
        m = mu.detach().to(torch.float32).contiguous(); lv = logvar.detach().to(torch.float32).contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=m.device)
        dm = torch.empty_like(m); dl = torch.empty_like(m)
        ops.kl_fwd_bwd(m, lv, dm, dl, 1.0, acc)
        ctx.save_for_backward(dm, dl)
        return acc[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        dm, dl = ctx.saved_tensors
        return dm * g, dl * g


def bce_with_logits_sum(x, t):
    return _BceSumFn.apply(x, t)


def bce_with_logits(x, t):
    return _BceElemFn.apply(x, t)


def cross_entropy_rows(x, target):
    return _CeRowsFn.apply(x, target)


def cross_entropy_sum(x, target):
    return _CeSumFn.apply(x, target)


def kl_sum(mu, logvar):
    return _KlSumFn.apply(mu, logvar)
