"""Fused training step of the FashionMNIST-flavour MVAE (conv image encoder/decoder, fashionmnist/model.py:70-121;
label nets and objective as in mnist; loop body fashionmnist/train.py:196-219).

Same structure as ``trainer.MnistMVAETrainer`` (pass stacking, flat arenas, CUDA graph, one NCCL all-reduce); the four
4x4/stride-2 convolutions run as im2col / col2im data movement around the tcgen05 GEMM, activations are NHWC:

  Conv2d(Cin->Cout)           cols = im2col(x) [B*OH*OW, 16 Cin] ; y = cols * Wp^T  (Swish fused in the GEMM epilogue)
  ConvTranspose2d(Cin->Cout)  cols = x * WTp^T [B*IH*IW, 16 Cout] ; y = col2im(cols) (Swish fused in col2im)
  backward                    d cols(ConvT) = im2col(dy) ; dx(Conv) = col2im(d cols) * Swish'(a)

Parameters are stored in GEMM-operand order (permutations of the reference tensors, applied in
``load_state_dict`` / ``state_dict`` / ``export_grads``; Adam is element-wise, so it is layout-agnostic):
  features.2.weight   [128,64,4,4] -> [128, (kh,kw,ci)]       classifier.0.weight [512,(c,h,w)] -> [512,(h,w,c)]
  upsampler.2.weight  [(c,h,w),512] -> [(h,w,c),512] (+bias)   hallucinate.k.weight [ci,co,kh,kw] -> [(kh,kw,co), ci]
This first version materialises the im2col matrices in HBM (explicit GEMM operands); an implicit-GEMM (TMA im2col)
main loop is the planned replacement.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from . import ops
from .trainer import MnistMVAETrainer, _PASS_MASKS

_TXT_ENC = ("text_encoder.net.0", "text_encoder.net.2", "text_encoder.net.4")
_TXT_DEC = ("text_decoder.net.0", "text_decoder.net.2", "text_decoder.net.4", "text_decoder.net.6")


def fashion_reference_shapes(L: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Reference state_dict names/shapes in registration order (fashionmnist/model.py:24-31,76-87,104-116,131-161)."""
    return [
        ("image_encoder.features.0.weight", (64, 1, 4, 4)),
        ("image_encoder.features.2.weight", (128, 64, 4, 4)),
        ("image_encoder.classifier.0.weight", (512, 6272)), ("image_encoder.classifier.0.bias", (512,)),
        ("image_encoder.classifier.2.weight", (2 * L, 512)), ("image_encoder.classifier.2.bias", (2 * L,)),
        ("image_decoder.upsampler.0.weight", (512, L)), ("image_decoder.upsampler.0.bias", (512,)),
        ("image_decoder.upsampler.2.weight", (6272, 512)), ("image_decoder.upsampler.2.bias", (6272,)),
        ("image_decoder.hallucinate.0.weight", (128, 64, 4, 4)),
        ("image_decoder.hallucinate.2.weight", (64, 1, 4, 4)),
        ("text_encoder.net.0.weight", (10, 512)),
        ("text_encoder.net.2.weight", (512, 512)), ("text_encoder.net.2.bias", (512,)),
        ("text_encoder.net.4.weight", (2 * L, 512)), ("text_encoder.net.4.bias", (2 * L,)),
        ("text_decoder.net.0.weight", (512, L)), ("text_decoder.net.0.bias", (512,)),
        ("text_decoder.net.2.weight", (512, 512)), ("text_decoder.net.2.bias", (512,)),
        ("text_decoder.net.4.weight", (512, 512)), ("text_decoder.net.4.bias", (512,)),
        ("text_decoder.net.6.weight", (10, 512)), ("text_decoder.net.6.bias", (10,)),
    ]


# reference tensor -> internal (GEMM operand) tensor and back
def _to_internal(name: str, t: torch.Tensor) -> torch.Tensor:
    if name == "image_encoder.features.0.weight":
        return t.reshape(64, 16)
    if name == "image_encoder.features.2.weight":
        return t.permute(0, 2, 3, 1).reshape(128, 1024)
    if name == "image_encoder.classifier.0.weight":
        return t.reshape(512, 128, 7, 7).permute(0, 2, 3, 1).reshape(512, 6272)
    if name == "image_decoder.upsampler.2.weight":
        return t.reshape(128, 7, 7, 512).permute(1, 2, 0, 3).reshape(6272, 512)
    if name == "image_decoder.upsampler.2.bias":
        return t.reshape(128, 7, 7).permute(1, 2, 0).reshape(6272)
    if name == "image_decoder.hallucinate.0.weight":
        return t.permute(2, 3, 1, 0).reshape(1024, 128)
    if name == "image_decoder.hallucinate.2.weight":
        return t.permute(2, 3, 1, 0).reshape(16, 64)
    return t


def _to_reference(name: str, t: torch.Tensor) -> torch.Tensor:
    if name == "image_encoder.features.0.weight":
        return t.reshape(64, 1, 4, 4)
    if name == "image_encoder.features.2.weight":
        return t.reshape(128, 4, 4, 64).permute(0, 3, 1, 2)
    if name == "image_encoder.classifier.0.weight":
        return t.reshape(512, 7, 7, 128).permute(0, 3, 1, 2).reshape(512, 6272)
    if name == "image_decoder.upsampler.2.weight":
        return t.reshape(7, 7, 128, 512).permute(2, 0, 1, 3).reshape(6272, 512)
    if name == "image_decoder.upsampler.2.bias":
        return t.reshape(7, 7, 128).permute(2, 0, 1).reshape(6272)
    if name == "image_decoder.hallucinate.0.weight":
        return t.reshape(4, 4, 64, 128).permute(3, 2, 0, 1)
    if name == "image_decoder.hallucinate.2.weight":
        return t.reshape(4, 4, 1, 64).permute(3, 2, 0, 1)
    return t


_INTERNAL_SHAPE = {
    "image_encoder.features.0.weight": (64, 16), "image_encoder.features.2.weight": (128, 1024),
    "image_decoder.hallucinate.0.weight": (1024, 128), "image_decoder.hallucinate.2.weight": (16, 64),
}


class FashionMVAETrainer(MnistMVAETrainer):
    """Whole-step trainer, FashionMNIST flavour.  Same public API as MnistMVAETrainer."""

    def _make_layout(self, L: int):
        return [(k, _INTERNAL_SHAPE.get(k, shp)) for k, shp in fashion_reference_shapes(L)]

    def _label_encoder(self, buf: int):
        a = self.arena
        return (a.view(buf, "text_encoder.net.0.weight"), a.view(buf, "text_encoder.net.2.weight"),
                a.view(buf, "text_encoder.net.2.bias"), a.view(buf, "text_encoder.net.4.weight"),
                a.view(buf, "text_encoder.net.4.bias"))

    def _alloc_activations(self, f) -> None:
        import os
        B = self.B
        # direct kernels for the 1-channel conv layers (default); MVAE_DIRECT_C1=0 restores im2col + tensor-core GEMM
        self.direct_c1 = os.environ.get("MVAE_DIRECT_C1", "1") != "0"
        # ... of which the last transposed conv's BACKWARD only on request (MVAE_DIRECT_C1_BWD=1): measured at B = 4096 the
        # direct forward kernels win (conv1 133 vs 217 us, last transposed conv 235 vs 322 us) but the CUDA-core backward of the
        # transposed conv does not (741 vs 562 us: 2,048 FMAs per pixel are too many for the FMA pipes;
        # profiles/r02_conv_small_ncu.txt)
        self.direct_c1_bwd = self.direct_c1 and os.environ.get("MVAE_DIRECT_C1_BWD", "0") == "1"
        # implicit-GEMM operands for the 64-channel conv layers (default); MVAE_IMPLICIT_CONV=0 materialises im2col in HBM
        self.implicit_conv = os.environ.get("MVAE_IMPLICIT_CONV", "1") != "0"
        # transposed convolutions (decoder forward, encoder data gradient) as four sub-pixel implicit GEMMs: no cols
        # buffers, no col2im passes (ops.subpixel_k4s2p1); MVAE_SUBPIXEL=0 restores GEMM -> cols -> col2im
        self.subpixel = self.implicit_conv and os.environ.get("MVAE_SUBPIXEL", "1") != "0"
        # image encoder (B rows), NHWC
        self.cols1 = f(B * 196, 16) if not self.direct_c1 else None
        self.c1_a, self.c1_h = f(B * 196, 64), f(B * 196, 64)
        self.cols2 = f(B * 49, 1024) if not self.implicit_conv else None
        self.c2_a, self.c2_h = f(B * 49, 128), f(B * 49, 128)           # == [B, 6272] in (h,w,c) order
        self.fc_a, self.fc_h = f(B, 512), f(B, 512)
        # text encoder
        self.te_h1, self.te_a2, self.te_h2 = f(B, 512), f(B, 512), f(B, 512)
        # image decoder (2B rows)
        self.u1_a, self.u1_h = f(2 * B, 512), f(2 * B, 512)
        self.u2_a, self.u2_h = f(2 * B, 6272), f(2 * B, 6272)           # == [2B*49, 128]
        self.colsT1 = f(2 * B * 49, 1024)
        self.t1_a, self.t1_h = f(2 * B * 196, 64), f(2 * B * 196, 64)
        self.colsT2 = f(2 * B * 196, 16) if not self.direct_c1 else None
        # text decoder (2B rows)
        self.td_a = [f(2 * B, 512) for _ in range(3)]; self.td_h = [f(2 * B, 512) for _ in range(3)]
        # backward scratch
        self.dcolsT2 = f(2 * B * 196, 16) if not self.direct_c1_bwd else None
        self.d_t1 = f(2 * B * 196, 64)
        self.dcolsT1 = f(2 * B * 49, 1024) if not self.implicit_conv else None
        self.d_u2 = f(2 * B, 6272)
        self.d_u1 = f(2 * B, 512)
        self.td_dA = [f(2 * B, 512) for _ in range(2)]
        self.d_fc = f(B, 512)
        self.d_c2 = f(B * 49, 128)
        self.dcols2 = f(B * 49, 1024)
        self.d_c1 = f(B * 196, 64)
        self.te_dA = [f(B, 512) for _ in range(2)]

    # ------------------------------------------------------------------ parameters (reference <-> internal order)
    def init_parameters(self, seed: int = 0) -> None:
        g = torch.Generator(device="cpu").manual_seed(seed)
        sd = {}
        shapes = dict(fashion_reference_shapes(self.L))
        for name, shape in shapes.items():
            if name == "text_encoder.net.0.weight":
                sd[name] = torch.randn(shape, generator=g)
            else:
                wshape = shapes[name.replace(".bias", ".weight")]
                fan_in = int(math.prod(wshape[1:]))
                if name.startswith("image_decoder.hallucinate"):
                    fan_in = wshape[1] * 16   # ConvTranspose2d: fan_in counts dim 1
                sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        self.load_state_dict(sd)
        self.adam_m.zero_(); self.adam_v.zero_(); self.step_count.zero_()

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k, _ in fashion_reference_shapes(self.L):
            if k not in sd:
                raise KeyError(f"missing key {k}")
            self.params[k].copy_(_to_internal(k, sd[k].to(torch.float32)).contiguous())

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: _to_reference(k, self.params[k].detach()).contiguous().clone() for k, _ in fashion_reference_shapes(self.L)}

    def export_grads(self) -> Dict[str, torch.Tensor]:
        """Gradients in the reference's tensor layouts (tests / interoperability)."""
        return {k: _to_reference(k, self.grads[k].detach()).contiguous().clone() for k, _ in fashion_reference_shapes(self.L)}

    # ------------------------------------------------------------------ forward
    def _enqueue_forward(self, training: bool, use_noise_input: bool) -> None:
        B, L, P = self.B, self.L, self.prec
        p = self.params
        # ---- image encoder.  The 1-channel layers (conv1 here, the last transposed conv and their backward) run as direct
        # HBM-bound kernels (csrc/conv_small.cu): no im2col / cols buffers, the 64-channel activation is streamed once
        direct = self.direct_c1
        if direct:
            ops.conv_cin_fwd(self.x, p["image_encoder.features.0.weight"], self.c1_a, self.c1_h, B, 28, 28, 1, 64)
        else:
            ops.im2col_k4s2p1(self.x, self.cols1, B, 28, 28, 1)
            ops.gemm_batch([ops.gemm_desc(self.cols1, p["image_encoder.features.0.weight"], self.c1_a, B * 196, 64, 16,
                                          out2=self.c1_h, epilogue=ops.EPI_BIAS_SWISH)], P)
        imp = self.implicit_conv     # conv GEMM operands fetched by TMA im2col loads: no cols buffers in HBM
        v_c1 = ops.conv_view(B, 14, 14, 64) if imp else None
        if not imp:
            ops.im2col_k4s2p1(self.c1_h, self.cols2, B, 14, 14, 64)
        lt = self.label_table
        if lt:   # label encoder once per class (csrc/label_table.cu, side stream); the PoE kernels gather row text[b]
            emb, w2, b2, w3, b3 = self._label_encoder(0)
            with self._fork():
                ops.label_table_fwd(emb, w2, b2, w3, b3, self.tt_a2, self.tt_h2, self.enc_tab)
        else:
            ops.embedding_swish_fwd(p["text_encoder.net.0.weight"], self.text, None, self.te_h1)
        ops.gemm_batch([
            ops.gemm_desc(self.c1_h if imp else self.cols2, p["image_encoder.features.2.weight"], self.c2_a, B * 49, 128, 1024,
                          out2=self.c2_h, epilogue=ops.EPI_BIAS_SWISH, a_view=v_c1)] + ([] if lt else [
            ops.gemm_desc(self.te_h1, p["text_encoder.net.2.weight"], self.te_a2, B, 512, 512,
                          bias=p["text_encoder.net.2.bias"], out2=self.te_h2, epilogue=ops.EPI_BIAS_SWISH)]), P)
        # (classifier.0 has K = 6272 but only B/128 x 4 output tiles: fused split-K when that leaves SMs idle)
        self._gemm([
            self._D("cls0", self.c2_h.view(B, 6272), p["image_encoder.classifier.0.weight"], self.fc_a, B, 512, 6272,
                    bias=p["image_encoder.classifier.0.bias"], out2=self.fc_h, epilogue=ops.EPI_BIAS_SWISH)] + ([] if lt else [
            ops.gemm_desc(self.te_h2, p["text_encoder.net.4.weight"], self.enc_t, B, 2 * L, 512,
                          bias=p["text_encoder.net.4.bias"])]))
        self._gemm([self._D("cls2", self.fc_h, p["image_encoder.classifier.2.weight"], self.enc_i, B, 2 * L, 512,
                            bias=p["image_encoder.classifier.2.bias"])])
        # ---- PoE + reparametrise + KL (three passes)
        self._join()
        mu_e, lv_e, _, _, gather = self._label_experts()
        ops.poe_fwd(mu_e, lv_e, _PASS_MASKS, B, L, self.Z, variant=0, training=training,
                    noise=self.noise if (training and use_noise_input) else None,
                    noise_out=self.noise if (training and not use_noise_input) else None,
                    seed=self.seed * 1000003 + self.rank, offset=0, step_dev=self.step_count, kl_acc=self.acc[6:9],
                    gather=gather)
        # ---- decoders (image: rows [0,2B) of Z; text: rows [B,3B))
        zi, zt = self.Z[: 2 * B], self.Z[B:]
        ops.gemm_batch([
            ops.gemm_desc(zi, p["image_decoder.upsampler.0.weight"], self.u1_a, 2 * B, 512, L,
                          bias=p["image_decoder.upsampler.0.bias"], out2=self.u1_h, epilogue=ops.EPI_BIAS_SWISH),
            ops.gemm_desc(zt, p["text_decoder.net.0.weight"], self.td_a[0], 2 * B, 512, L,
                          bias=p["text_decoder.net.0.bias"], out2=self.td_h[0], epilogue=ops.EPI_BIAS_SWISH)], P)
        ops.gemm_batch([
            ops.gemm_desc(self.u1_h, p["image_decoder.upsampler.2.weight"], self.u2_a, 2 * B, 6272, 512,
                          bias=p["image_decoder.upsampler.2.bias"], out2=self.u2_h, epilogue=ops.EPI_BIAS_SWISH),
            ops.gemm_desc(self.td_h[0], p["text_decoder.net.2.weight"], self.td_a[1], 2 * B, 512, 512,
                          bias=p["text_decoder.net.2.bias"], out2=self.td_h[1], epilogue=ops.EPI_BIAS_SWISH)], P)
        txt3 = ops.gemm_desc(self.td_h[1], p["text_decoder.net.4.weight"], self.td_a[2], 2 * B, 512, 512,
                             bias=p["text_decoder.net.4.bias"], out2=self.td_h[2], epilogue=ops.EPI_BIAS_SWISH)
        if self.subpixel:
            descs = ops.subpixel_k4s2p1(self.u2_h, p["image_decoder.hallucinate.0.weight"], self.t1_a, 2 * B, 7, 7, 128, 64,
                                        out2=self.t1_h, epilogue=ops.EPI_BIAS_SWISH)
            ops.gemm_chain(descs + [txt3], [-1] * 5, self.chain_ws, P)
        else:
            ops.gemm_batch([
                ops.gemm_desc(self.u2_h.view(2 * B * 49, 128), p["image_decoder.hallucinate.0.weight"], self.colsT1,
                              2 * B * 49, 1024, 128), txt3], P)
            ops.col2im_k4s2p1(self.colsT1, self.t1_a, 2 * B, 7, 7, 64, out_act=self.t1_h)
        txt_last = ops.gemm_desc(self.td_h[2], p["text_decoder.net.6.weight"], self.logit_t, 2 * B, 10, 512,
                                 bias=p["text_decoder.net.6.bias"])
        if direct:
            ops.gemm_batch([txt_last], P)
            ops.convT_cout_fwd(self.t1_h, p["image_decoder.hallucinate.2.weight"], self.logit_i, 2 * B, 14, 14, 64, 1)
        else:
            ops.gemm_batch([
                ops.gemm_desc(self.t1_h, p["image_decoder.hallucinate.2.weight"], self.colsT2, 2 * B * 196, 16, 64), txt_last], P)
            ops.col2im_k4s2p1(self.colsT2, self.logit_i, 2 * B, 14, 14, 1)

    # ------------------------------------------------------------------ loss + backward
    def _enqueue_loss_and_backward(self, training: bool, b_global: int) -> None:
        B, L, P = self.B, self.L, self.prec
        p, g = self.params, self.grads

        def split_for(rows):   # ~16 k-blocks (512 rows) per wgrad tile, as many tiles as the reduction needs
            return max(1, min(rows // 512, 4096))

        with self._fork():
            ops.ce_fwd_bwd(self.logit_t, self.text, self.logit_t, 10, self.lam_t / b_global, self.acc[4:6], seg_rows=B)
        ops.bce_logits_fwd_bwd(self.logit_i, self.x, self.logit_i, self.lam_i / b_global, self.acc[0:3], seg_rows=B)
        self._join()
        # ---- last layers: convT2 (image) and net.6 (text)
        direct = self.direct_c1_bwd
        with self._fork():          # bias gradients are read by the optimizer only: beside the GEMMs, joined at the end
            ops.colsum_accumulate(self.logit_t, g["text_decoder.net.6.bias"])
        txt_last = [
            ops.gemm_desc(self.logit_t, self.td_h[2], g["text_decoder.net.6.weight"], 10, 512, 2 * B, a_mn=True, b_mn=True,
                          split_k=split_for(2 * B), accumulate=True),
            ops.gemm_desc(self.logit_t, p["text_decoder.net.6.weight"], self.td_dA[0], 2 * B, 512, 10, b_mn=True,
                          aux=self.td_a[2], epilogue=ops.EPI_MUL_DSWISH, colsum=g["text_decoder.net.4.bias"])]
        if direct:
            ops.convT_cout_bwd(self.logit_i, self.t1_h, self.t1_a, p["image_decoder.hallucinate.2.weight"], self.d_t1,
                               g["image_decoder.hallucinate.2.weight"], 2 * B, 14, 14, 64, 1)
            ops.gemm_batch(txt_last, P)
        else:
            ops.im2col_k4s2p1(self.logit_i, self.dcolsT2, 2 * B, 28, 28, 1)
            ops.gemm_batch([
                ops.gemm_desc(self.dcolsT2, self.t1_h, g["image_decoder.hallucinate.2.weight"], 16, 64, 2 * B * 196,
                              a_mn=True, b_mn=True, split_k=split_for(2 * B * 196), accumulate=True),
                ops.gemm_desc(self.dcolsT2, p["image_decoder.hallucinate.2.weight"], self.d_t1, 2 * B * 196, 64, 16, b_mn=True,
                              aux=self.t1_a, epilogue=ops.EPI_MUL_DSWISH)] + txt_last, P)
        # ---- convT1 (image) and net.4 (text): d cols = im2col(d_t1), implicit (TMA im2col) unless MVAE_IMPLICIT_CONV=0
        imp = self.implicit_conv
        v_t1 = ops.conv_view(2 * B, 14, 14, 64) if imp else None
        if not imp:
            ops.im2col_k4s2p1(self.d_t1, self.dcolsT1, 2 * B, 14, 14, 64)
        dcols = self.d_t1 if imp else self.dcolsT1
        ops.gemm_batch([
            ops.gemm_desc(dcols, self.u2_h.view(2 * B * 49, 128), g["image_decoder.hallucinate.0.weight"], 1024, 128,
                          2 * B * 49, a_mn=True, b_mn=True, split_k=split_for(2 * B * 49), accumulate=True, a_view=v_t1),
            ops.gemm_desc(dcols, p["image_decoder.hallucinate.0.weight"], self.d_u2.view(2 * B * 49, 128), 2 * B * 49,
                          128, 1024, b_mn=True, aux=self.u2_a.view(2 * B * 49, 128), epilogue=ops.EPI_MUL_DSWISH, a_view=v_t1),
            ops.gemm_desc(self.td_dA[0], self.td_h[1], g["text_decoder.net.4.weight"], 512, 512, 2 * B, a_mn=True, b_mn=True,
                          split_k=split_for(2 * B), accumulate=True),
            ops.gemm_desc(self.td_dA[0], p["text_decoder.net.4.weight"], self.td_dA[1], 2 * B, 512, 512, b_mn=True,
                          aux=self.td_a[1], epilogue=ops.EPI_MUL_DSWISH, colsum=g["text_decoder.net.2.bias"])], P)
        # ---- upsampler.2 (image) and net.2 (text)
        with self._fork():
            ops.colsum_accumulate(self.d_u2, g["image_decoder.upsampler.2.bias"])
        self._gemm([
            ops.gemm_desc(self.d_u2, self.u1_h, g["image_decoder.upsampler.2.weight"], 6272, 512, 2 * B, a_mn=True, b_mn=True,
                          split_k=split_for(2 * B), accumulate=True),
            self._D("dg_u2", self.d_u2, p["image_decoder.upsampler.2.weight"], self.d_u1, 2 * B, 512, 6272, share=3, b_mn=True,
                    aux=self.u1_a, epilogue=ops.EPI_MUL_DSWISH, colsum=g["image_decoder.upsampler.0.bias"]),
            ops.gemm_desc(self.td_dA[1], self.td_h[0], g["text_decoder.net.2.weight"], 512, 512, 2 * B, a_mn=True, b_mn=True,
                          split_k=split_for(2 * B), accumulate=True),
            ops.gemm_desc(self.td_dA[1], p["text_decoder.net.2.weight"], self.td_dA[0], 2 * B, 512, 512, b_mn=True,
                          aux=self.td_a[0], epilogue=ops.EPI_MUL_DSWISH, colsum=g["text_decoder.net.0.bias"])])
        # ---- first decoder layers -> dZ (zero-initialised, both decoders add)
        ops.gemm_batch([
            ops.gemm_desc(self.d_u1, self.Z[: 2 * B], g["image_decoder.upsampler.0.weight"], 512, L, 2 * B, a_mn=True,
                          b_mn=True, split_k=split_for(2 * B), accumulate=True),
            ops.gemm_desc(self.d_u1, p["image_decoder.upsampler.0.weight"], self.dZ[: 2 * B], 2 * B, L, 512, b_mn=True,
                          accumulate=True),
            ops.gemm_desc(self.td_dA[0], self.Z[B:], g["text_decoder.net.0.weight"], 512, L, 2 * B, a_mn=True, b_mn=True,
                          split_k=split_for(2 * B), accumulate=True),
            ops.gemm_desc(self.td_dA[0], p["text_decoder.net.0.weight"], self.dZ[B:], 2 * B, L, 512, b_mn=True,
                          accumulate=True)], P)
        # ---- PoE / reparam / KL backward
        mu_e, lv_e, dmu, dlv, gather = self._label_experts()
        ops.poe_bwd(mu_e, lv_e, _PASS_MASKS, B, L, self.dZ, dmu, dlv, kl_scale=1.0 / b_global, variant=0,
                    training=training, noise=self.noise if training else None, kl_scale_dev=self.beta_dev, gather=gather)
        lt = self.label_table
        # ---- encoders backward: heads
        with self._fork():
            ops.colsum_accumulate(self.d_enc_i, g["image_encoder.classifier.2.bias"])
            if not lt:
                ops.colsum_accumulate(self.d_enc_t, g["text_encoder.net.4.bias"])
        ops.gemm_batch([
            ops.gemm_desc(self.d_enc_i, self.fc_h, g["image_encoder.classifier.2.weight"], 2 * L, 512, B, a_mn=True, b_mn=True,
                          split_k=split_for(B), accumulate=True),
            ops.gemm_desc(self.d_enc_i, p["image_encoder.classifier.2.weight"], self.d_fc, B, 512, 2 * L, b_mn=True,
                          aux=self.fc_a, epilogue=ops.EPI_MUL_DSWISH, colsum=g["image_encoder.classifier.0.bias"])] + ([] if lt else [
            ops.gemm_desc(self.d_enc_t, self.te_h2, g["text_encoder.net.4.weight"], 2 * L, 512, B, a_mn=True, b_mn=True,
                          split_k=split_for(B), accumulate=True),
            ops.gemm_desc(self.d_enc_t, p["text_encoder.net.4.weight"], self.te_dA[0], B, 512, 2 * L, b_mn=True,
                          aux=self.te_a2, epilogue=ops.EPI_MUL_DSWISH, colsum=g["text_encoder.net.2.bias"])]), P)
        # ---- classifier.0 (image) and net.2 (text)
        ops.gemm_batch([
            ops.gemm_desc(self.d_fc, self.c2_h.view(B, 6272), g["image_encoder.classifier.0.weight"], 512, 6272, B,
                          a_mn=True, b_mn=True, split_k=split_for(B), accumulate=True),
            ops.gemm_desc(self.d_fc, p["image_encoder.classifier.0.weight"], self.d_c2.view(B, 6272), B, 6272, 512,
                          b_mn=True, aux=self.c2_a.view(B, 6272), epilogue=ops.EPI_MUL_DSWISH)] + ([] if lt else [
            ops.gemm_desc(self.te_dA[0], self.te_h1, g["text_encoder.net.2.weight"], 512, 512, B, a_mn=True, b_mn=True,
                          split_k=split_for(B), accumulate=True),
            ops.gemm_desc(self.te_dA[0], p["text_encoder.net.2.weight"], self.te_dA[1], B, 512, 512, b_mn=True)]), P)
        if lt:
            emb, w2, _, w3, _ = self._label_encoder(0)
            g_emb, g_w2, g_b2, g_w3, g_b3 = self._label_encoder(1)
            with self._fork():     # joined at the end of the backward pass
                ops.label_table_bwd(emb, w2, w3, self.tt_a2, self.tt_h2, self.d_tab, self.tt_dA2, g_emb, g_w2, g_b2, g_w3, g_b3)
        else:
            ops.embedding_swish_bwd(p["text_encoder.net.0.weight"], self.text, self.te_dA[1], g["text_encoder.net.0.weight"])
        # ---- conv2
        ops.gemm_batch([
            ops.gemm_desc(self.d_c2, self.c1_h if imp else self.cols2, g["image_encoder.features.2.weight"], 128, 1024, B * 49,
                          a_mn=True, b_mn=True, split_k=split_for(B * 49), accumulate=True,
                          b_view=ops.conv_view(B, 14, 14, 64) if imp else None),
            ] + ([] if self.subpixel else [
            ops.gemm_desc(self.d_c2, p["image_encoder.features.2.weight"], self.dcols2, B * 49, 1024, 128, b_mn=True)]), P)
        if self.subpixel:   # d c1 = ConvT(d c2; W2) * swish'(c1_a): four sub-pixel implicit GEMMs, rows stored in place
            descs = ops.subpixel_k4s2p1(self.d_c2, p["image_encoder.features.2.weight"], self.d_c1, B, 7, 7, 128, 64,
                                        w_is_conv=True, aux=self.c1_a, epilogue=ops.EPI_MUL_DSWISH)
            ops.gemm_chain(descs, [-1] * 4, self.chain_ws, P)
        else:
            ops.col2im_k4s2p1(self.dcols2, self.d_c1, B, 7, 7, 64, aux=self.c1_a)
        # ---- conv1 (no data gradient: the image is an input); direct whenever the forward was (no cols1 buffer then)
        if self.direct_c1:
            ops.conv_cin_wgrad(self.x, self.d_c1, g["image_encoder.features.0.weight"], B, 28, 28, 1, 64)
        else:
            ops.gemm_batch([ops.gemm_desc(self.d_c1, self.cols1, g["image_encoder.features.0.weight"], 64, 16, B * 196,
                                          a_mn=True, b_mn=True, split_k=split_for(B * 196), accumulate=True)], P)
        self._join()
