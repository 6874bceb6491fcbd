"""Fused training step of the CelebA-flavour MVAE: DCGAN conv image encoder/decoder with BatchNorm2d + Dropout(0.1),
attribute MLPs with BatchNorm1d, PoE variant B, per-pixel / per-attribute BCE (celeba/model.py:13-229,
celeba/train.py:22-65,167-212).

Structure follows ``trainer.MnistMVAETrainer`` (pass stacking, flat arenas, CUDA graph, one NCCL all-reduce).
Flavour specifics (all chosen so that the result equals the reference's three ``model()`` calls):
  * BatchNorm uses train-mode batch statistics PER model() CALL: every stacked tensor is split in row segments (one per
    pass) and statistics / running-stat updates are per segment, applied in the reference's call order
    (joint, image-only, attrs-only).  The encoders see the same input in two passes -> identical statistics: they are
    evaluated once and their running statistics are updated twice.
  * Dropout draws a fresh mask in the joint and in the image-only pass: the image encoder is shared up to the Dropout,
    the last Linear runs on the two masked copies ([2B,512]) and the two results enter the PoE as two experts.
  * Both decoders are evaluated for all three passes (3B rows): the reference also runs the decoder whose output the
    loss ignores, and that call updates BatchNorm running statistics.  Backward runs on the live row ranges only
    (image decoder rows [0,2B), attribute decoder rows [B,3B)).
  * Convs: NHWC + im2col/col2im around the tcgen05 GEMM (k4 s2 p1, and k4 s1 p0 for the 8x8 <-> 5x5 layers); parameters
    are stored in GEMM-operand order (see ``_to_internal``); the first attribute Linear is padded 18 -> 20 inputs (TMA).
Under data parallelism BatchNorm statistics are per rank shard (the north-star design has a single gradient all-reduce).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from .trainer import MnistMVAETrainer, _REF_TO_INTERNAL

N_ATTRS = 18
_PASS_MASKS3 = (0b001, 0b110, 0b100)   # experts: 0 = image enc (image-only mask), 1 = image enc (joint mask), 2 = attrs
_BN_ORDER3 = (1, 0, 2)                 # reference call order (joint, image-only, attrs-only) in internal segment ids


def _bn_names(prefix: str, c: int):
    return [(f"{prefix}.weight", (c,)), (f"{prefix}.bias", (c,))]


def celeba_param_shapes(L: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Reference PARAMETER names/shapes in registration order (celeba/model.py:76-92,113-126,145-153,172-183)."""
    e, d = "image_encoder.features", "image_decoder.hallucinate"
    out = [(f"{e}.0.weight", (32, 3, 4, 4)), (f"{e}.2.weight", (64, 32, 4, 4))] + _bn_names(f"{e}.3", 64)
    out += [(f"{e}.5.weight", (128, 64, 4, 4))] + _bn_names(f"{e}.6", 128)
    out += [(f"{e}.8.weight", (256, 128, 4, 4))] + _bn_names(f"{e}.9", 256)
    out += [("image_encoder.classifier.0.weight", (512, 6400)), ("image_encoder.classifier.0.bias", (512,)),
            ("image_encoder.classifier.3.weight", (2 * L, 512)), ("image_encoder.classifier.3.bias", (2 * L,)),
            ("image_decoder.upsample.0.weight", (6400, L)), ("image_decoder.upsample.0.bias", (6400,))]
    out += [(f"{d}.0.weight", (256, 128, 4, 4))] + _bn_names(f"{d}.1", 128)
    out += [(f"{d}.3.weight", (128, 64, 4, 4))] + _bn_names(f"{d}.4", 64)
    out += [(f"{d}.6.weight", (64, 32, 4, 4))] + _bn_names(f"{d}.7", 32) + [(f"{d}.9.weight", (32, 3, 4, 4))]
    a = "attrs_encoder.net"
    out += [(f"{a}.0.weight", (512, N_ATTRS)), (f"{a}.0.bias", (512,))] + _bn_names(f"{a}.1", 512)
    out += [(f"{a}.3.weight", (512, 512)), (f"{a}.3.bias", (512,))] + _bn_names(f"{a}.4", 512)
    out += [(f"{a}.6.weight", (2 * L, 512)), (f"{a}.6.bias", (2 * L,))]
    a = "attrs_decoder.net"
    out += [(f"{a}.0.weight", (512, L)), (f"{a}.0.bias", (512,))] + _bn_names(f"{a}.1", 512)
    out += [(f"{a}.3.weight", (512, 512)), (f"{a}.3.bias", (512,))] + _bn_names(f"{a}.4", 512)
    out += [(f"{a}.6.weight", (512, 512)), (f"{a}.6.bias", (512,))] + _bn_names(f"{a}.7", 512)
    out += [(f"{a}.9.weight", (N_ATTRS, 512)), (f"{a}.9.bias", (N_ATTRS,))]
    return out


_BN_LAYERS = {  # prefix -> channels
    "image_encoder.features.3": 64, "image_encoder.features.6": 128, "image_encoder.features.9": 256,
    "image_decoder.hallucinate.1": 128, "image_decoder.hallucinate.4": 64, "image_decoder.hallucinate.7": 32,
    "attrs_encoder.net.1": 512, "attrs_encoder.net.4": 512,
    "attrs_decoder.net.1": 512, "attrs_decoder.net.4": 512, "attrs_decoder.net.7": 512,
}
_CONV = {"image_encoder.features.0.weight", "image_encoder.features.2.weight", "image_encoder.features.5.weight",
         "image_encoder.features.8.weight"}
_CONVT = {"image_decoder.hallucinate.0.weight", "image_decoder.hallucinate.3.weight", "image_decoder.hallucinate.6.weight",
          "image_decoder.hallucinate.9.weight"}


def _internal_shape(name: str, shape, L: int):
    if name in _CONV:
        return (shape[0], 16 * shape[1])
    if name in _CONVT:
        return (16 * shape[1], shape[0])
    if name == "attrs_encoder.net.0.weight":
        return (512, 20)
    return tuple(shape)


def _to_internal(name: str, t: torch.Tensor) -> torch.Tensor:
    if name in _CONV:                                    # [Cout,Cin,kh,kw] -> [Cout,(kh,kw,ci)]
        return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)
    if name in _CONVT:                                   # [Cin,Cout,kh,kw] -> [(kh,kw,co),Cin]
        return t.permute(2, 3, 1, 0).reshape(-1, t.shape[0])
    if name == "image_encoder.classifier.0.weight":      # columns (c,h,w) -> (h,w,c)
        return t.reshape(512, 256, 5, 5).permute(0, 2, 3, 1).reshape(512, 6400)
    if name == "image_decoder.upsample.0.weight":        # rows (c,h,w) -> (h,w,c)
        return t.reshape(256, 5, 5, -1).permute(1, 2, 0, 3).reshape(6400, -1)
    if name == "image_decoder.upsample.0.bias":
        return t.reshape(256, 5, 5).permute(1, 2, 0).reshape(6400)
    if name == "attrs_encoder.net.0.weight":
        out = torch.zeros(512, 20, dtype=t.dtype, device=t.device)
        out[:, :N_ATTRS] = t
        return out
    return t


def _to_reference(name: str, t: torch.Tensor, ref_shape) -> torch.Tensor:
    if name in _CONV:
        co, ci = ref_shape[0], ref_shape[1]
        return t.reshape(co, 4, 4, ci).permute(0, 3, 1, 2)
    if name in _CONVT:
        ci, co = ref_shape[0], ref_shape[1]
        return t.reshape(4, 4, co, ci).permute(3, 2, 0, 1)
    if name == "image_encoder.classifier.0.weight":
        return t.reshape(512, 5, 5, 256).permute(0, 3, 1, 2).reshape(512, 6400)
    if name == "image_decoder.upsample.0.weight":
        return t.reshape(5, 5, 256, -1).permute(2, 0, 1, 3).reshape(6400, -1)
    if name == "image_decoder.upsample.0.bias":
        return t.reshape(5, 5, 256).permute(2, 0, 1).reshape(6400)
    if name == "attrs_encoder.net.0.weight":
        return t[:, :N_ATTRS]
    return t


class CelebAMVAETrainer(MnistMVAETrainer):
    """Whole-step trainer, CelebA flavour: ``step(image [B,3,64,64], attrs [B,18])``."""

    def __init__(self, n_latents: int = 100, batch_size: int = 128, lr: float = 1e-4, lambda_image: float = 1.0,
                 lambda_attrs: float = 10.0, **kw):
        kw.pop("lambda_text", None)
        kw["label_table"] = False        # the attribute encoder takes 2^18 distinct inputs: no class table here
        import os
        self.implicit_conv = os.environ.get("MVAE_IMPLICIT_CONV", "1") != "0"
        # transposed convolutions without cols / col2im (ops.subpixel_k4s2p1, ops.full_k4s1p0); see trainer_fashion
        self.subpixel = self.implicit_conv and os.environ.get("MVAE_SUBPIXEL", "1") != "0"
        super().__init__(n_latents=n_latents, batch_size=batch_size, lr=lr, lambda_image=lambda_image,
                         lambda_text=lambda_attrs, **kw)

    # ------------------------------------------------------------------ layout / buffers
    def _make_layout(self, L: int):
        self._ref_shapes = dict(celeba_param_shapes(L))
        return [(k, _internal_shape(k, s, L)) for k, s in celeba_param_shapes(L)]

    def _alloc_activations(self, f) -> None:
        B, L, dev = self.B, self.L, self.dev
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        # inputs: NCHW staging + NHWC image (also the BCE target), attributes padded 18 -> 20
        self.x_nchw = f(B, 3 * 4096)
        self.x = f(B, 12288)
        self.a_in = z(B, 20)
        self.drop_mask = torch.ones(2 * B, 512, dtype=torch.float32, device=dev)   # internal copy order (image-only, joint)
        self.logit_i = f(3 * B, 12288)
        self.logit_a_buf = z(3 * B, 20)
        self.logit_a = self.logit_a_buf[:, :N_ATTRS]
        self.dlogit_a_buf = z(3 * B, 20)
        self.dlogit_a = self.dlogit_a_buf[:, :N_ATTRS]
        self.enc_i2, self.d_enc_i2 = f(2 * B, 2 * L), f(2 * B, 2 * L)
        # BatchNorm state
        self.buffers: Dict[str, torch.Tensor] = {}
        self.bn_mean: Dict[str, torch.Tensor] = {}
        self.bn_invstd: Dict[str, torch.Tensor] = {}
        for prefix, c in _BN_LAYERS.items():
            self.buffers[prefix + ".running_mean"] = z(c)
            self.buffers[prefix + ".running_var"] = torch.ones(c, dtype=torch.float32, device=dev)
            self.bn_mean[prefix] = f(3, c); self.bn_invstd[prefix] = f(3, c)
        self.num_batches_tracked = {p: 0 for p in _BN_LAYERS}
        self.bn_acc = torch.zeros(3 * 512 * 2, dtype=torch.float64, device=dev)
        # ---- image encoder (B rows)
        self.cols1 = f(B * 1024, 48); self.c1_a, self.c1_h = f(B * 1024, 32), f(B * 1024, 32)
        self.cols2 = f(B * 256, 512); self.c2_x, self.c2_h = f(B * 256, 64), f(B * 256, 64)
        self.cols3 = f(B * 64, 1024); self.c3_x, self.c3_h = f(B * 64, 128), f(B * 64, 128)
        self.cols4 = f(B * 25, 2048); self.c4_x, self.c4_h = f(B * 25, 256), f(B * 25, 256)
        self.fc_a, self.fc_h = f(B, 512), f(B, 512)
        self.fcd = f(2 * B, 512)
        # ---- attrs encoder
        self.ae1_x, self.ae1_h, self.ae2_x, self.ae2_h = f(B, 512), f(B, 512), f(B, 512), f(B, 512)
        # ---- image decoder (3B rows)
        self.d0_a, self.d0_h = f(3 * B, 6400), f(3 * B, 6400)
        self.colsT1 = f(3 * B * 25, 2048); self.t1_x, self.t1_h = f(3 * B * 64, 128), f(3 * B * 64, 128)
        self.colsT2 = f(3 * B * 64, 1024); self.t2_x, self.t2_h = f(3 * B * 256, 64), f(3 * B * 256, 64)
        self.colsT3 = f(3 * B * 256, 512); self.t3_x, self.t3_h = f(3 * B * 1024, 32), f(3 * B * 1024, 32)
        self.colsT4 = f(3 * B * 1024, 48)
        # ---- attrs decoder (3B rows)
        self.ad_x = [f(3 * B, 512) for _ in range(3)]; self.ad_h = [f(3 * B, 512) for _ in range(3)]
        # ---- backward scratch (live rows only: 2B)
        self.dcolsT4 = f(2 * B * 1024, 48); self.d_t3h, self.d_t3x = f(2 * B * 1024, 32), f(2 * B * 1024, 32)
        self.dcolsT3 = f(2 * B * 256, 512); self.d_t2h, self.d_t2x = f(2 * B * 256, 64), f(2 * B * 256, 64)
        self.dcolsT2 = f(2 * B * 64, 1024); self.d_t1h, self.d_t1x = f(2 * B * 64, 128), f(2 * B * 64, 128)
        self.dcolsT1 = f(2 * B * 25, 2048); self.d_d0 = f(2 * B, 6400)
        self.d_adh = f(3 * B, 512); self.d_adx = f(3 * B, 512)
        self.d_fcd = f(2 * B, 512); self.d_fch = f(B, 512); self.d_fca = f(B, 512)
        self.d_c4h, self.d_c4x = f(B * 25, 256), f(B * 25, 256); self.dcols4 = f(B * 25, 2048)
        self.d_c3h, self.d_c3x = f(B * 64, 128), f(B * 64, 128); self.dcols3 = f(B * 64, 1024)
        self.d_c2h, self.d_c2x = f(B * 256, 64), f(B * 256, 64); self.dcols2 = f(B * 256, 512)
        self.d_c1a = f(B * 1024, 32)
        self.d_aeh, self.d_aex = f(B, 512), f(B, 512)

    # ------------------------------------------------------------------ parameters / state
    def init_parameters(self, seed: int = 0) -> None:
        g = torch.Generator(device="cpu").manual_seed(seed)
        sd = {}
        for name, shape in celeba_param_shapes(self.L):
            prefix = name.rsplit(".", 1)[0]
            if prefix in _BN_LAYERS:
                sd[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
                continue
            wshape = self._ref_shapes[prefix + ".weight"]
            fan_in = wshape[1] * 16 if name in _CONVT or (prefix + ".weight") in _CONVT else int(math.prod(wshape[1:]))
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        for prefix, c in _BN_LAYERS.items():
            sd[prefix + ".running_mean"] = torch.zeros(c); sd[prefix + ".running_var"] = torch.ones(c)
        self.load_state_dict(sd)
        self.adam_m.zero_(); self.adam_v.zero_(); self.step_count.zero_()

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k, _ in celeba_param_shapes(self.L):
            self.params[k].copy_(_to_internal(k, sd[k].to(torch.float32)).contiguous())
        for k in self.buffers:
            if k in sd:
                self.buffers[k].copy_(sd[k].to(torch.float32))
        for p in _BN_LAYERS:
            if p + ".num_batches_tracked" in sd:
                self.num_batches_tracked[p] = int(sd[p + ".num_batches_tracked"])

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k, shp in celeba_param_shapes(self.L):
            out[k] = _to_reference(k, self.params[k].detach(), shp).contiguous().clone()
            prefix = k.rsplit(".", 1)[0]
            if prefix in _BN_LAYERS and k.endswith(".bias"):
                out[prefix + ".running_mean"] = self.buffers[prefix + ".running_mean"].clone()
                out[prefix + ".running_var"] = self.buffers[prefix + ".running_var"].clone()
                out[prefix + ".num_batches_tracked"] = torch.tensor(self.num_batches_tracked[prefix], dtype=torch.int64)
        return out

    def export_grads(self) -> Dict[str, torch.Tensor]:
        return {k: _to_reference(k, self.grads[k].detach(), shp).contiguous().clone() for k, shp in celeba_param_shapes(self.L)}

    # ------------------------------------------------------------------ inputs
    def set_inputs(self, image, attrs, noise: Optional[torch.Tensor] = None, annealing_factor: float = 1.0,
                   drop_masks: Optional[torch.Tensor] = None) -> None:
        """image [B,3,64,64] (NCHW, like the reference), attrs [B,18] float {0,1}; optional noise [3,B,L] and dropout
        masks [2,B,512] in the REFERENCE's call order (joint, image-only[, attrs-only])."""
        B, L = self.B, self.L
        with torch.cuda.stream(self._stream):
            self.x_nchw.copy_(image.reshape(B, 3 * 4096), non_blocking=True)
            ops.nchw_to_nhwc(self.x_nchw, self.x, B, 3, 4096)
            self.a_in[:, :N_ATTRS].copy_(attrs.reshape(B, N_ATTRS).to(torch.float32), non_blocking=True)
            self._stage_beta(annealing_factor)
            if noise is not None:
                nz = self.noise.view(3, B, L)
                for ref_i, int_i in enumerate(_REF_TO_INTERNAL):
                    nz[int_i].copy_(noise[ref_i], non_blocking=True)
            self._masks_given = drop_masks is not None
            if drop_masks is not None:     # reference order (joint, image-only) -> internal copies (image-only, joint)
                self.drop_mask[:B].copy_(drop_masks[1], non_blocking=True)
                self.drop_mask[B:].copy_(drop_masks[0], non_blocking=True)

    def step(self, image, attrs, annealing_factor: float = 1.0, noise=None, training: bool = True, update: bool = True,
             sync: bool = True, drop_masks=None):
        self.set_inputs(image, attrs, noise, annealing_factor, drop_masks)
        self.run(training=training, noise_given=noise is not None, update=update)
        if training and update:
            for p in _BN_LAYERS:
                self.num_batches_tracked[p] += 3 if p.startswith(("image_decoder", "attrs_decoder")) else 2
        with torch.cuda.stream(self._stream):
            self.loss_host.copy_(self.loss_out, non_blocking=True)
        if sync:
            self._stream.synchronize()
            self.check_device_errors()
            return float(self.loss_host[0])
        return None

    # ---- device-resident dataset (celeba/datasets.py:38-135 + the DataLoader / H2D of celeba/train.py:163-186 replaced)
    def attach_dataset(self, images_u8: torch.Tensor, attrs: torch.Tensor) -> None:
        """Keep the whole uint8 dataset in HBM: ``images_u8`` [N,64,64,3] (HWC, as the image files decode) or [N,3,64,64]
        (CHW; converted once), ``attrs`` [N,18] in {0,1}.  A batch is then two gather launches: rows idx of the HWC bytes
        / 255 land directly in the NHWC activation layout the conv kernels read (ToTensor's /255 and the NCHW->NHWC staging
        of the host path both disappear), and the attribute rows become {0,1} floats."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
            raise _lib.MvaeError("attach_dataset: images must be a uint8 [N,64,64,3] or [N,3,64,64] tensor")
        if images_u8.shape[1] == 3 and images_u8.shape[-1] != 3:
            images_u8 = images_u8.permute(0, 2, 3, 1)
        n = images_u8.shape[0]
        if tuple(images_u8.shape[1:]) != (64, 64, 3) or tuple(attrs.shape) != (n, N_ATTRS):
            raise _lib.MvaeError("attach_dataset: expected images [N,64,64,3] and attrs [N,18]")
        self.ds_images = images_u8.contiguous().reshape(n, 64 * 64 * 3).to(self.dev)
        a = torch.zeros(n, 20, dtype=torch.uint8)                      # rows padded to 20 bytes (the a_in layout); 255 -> 1.0
        a[:, :N_ATTRS] = (attrs != 0).to(torch.uint8) * 255
        self.ds_attrs = a.to(self.dev)

    def epoch_permutation(self, seed: int) -> torch.Tensor:
        g = torch.Generator(device=self.dev).manual_seed(seed)
        return torch.randperm(self.ds_images.shape[0], generator=g, device=self.dev)

    def step_from_dataset(self, idx: torch.Tensor, annealing_factor: float = 1.0, training: bool = True,
                          update: bool = True, sync: bool = True):
        """One training iteration on rows ``idx`` (int64 [B], on the device) of the attached dataset."""
        if idx.numel() != self.B or idx.dtype != torch.int64 or not idx.is_cuda:
            raise _lib.MvaeError(f"step_from_dataset: idx must be a CUDA int64 tensor of {self.B} row indices")
        idx = idx.contiguous()
        with torch.cuda.stream(self._stream):
            ops.gather_batch_u8(self.ds_images, None, idx, self.x.view(self.B, 64 * 64 * 3), None)
            ops.gather_batch_u8(self.ds_attrs, None, idx, self.a_in, None)
            self._stage_beta(annealing_factor)
            self._masks_given = False
        self.run(training=training, noise_given=False, update=update)
        self._pipe_after_run(training, update)
        with torch.cuda.stream(self._stream):
            self.loss_host.copy_(self.loss_out, non_blocking=True)
        if sync:
            self._stream.synchronize()
            self.check_device_errors()
            return float(self.loss_host[0])
        return None

    # ---- pipelined host path (base class step_pipelined): one staged batch = NCHW image + [B,18] attrs
    def _pipe_slot_tensors(self):
        B, dev = self.B, self.dev
        return {"img": torch.empty(B, 3 * 4096, dtype=torch.float32, device=dev),
                "oth": torch.empty(B, N_ATTRS, dtype=torch.float32, device=dev)}

    def _pipe_consume(self, slot) -> None:
        ops.nchw_to_nhwc(slot["img"], self.x, self.B, 3, 4096)      # the layout change IS the copy out of the staging slot
        self.a_in[:, :N_ATTRS].copy_(slot["oth"], non_blocking=True)
        self._masks_given = False

    def _pipe_after_run(self, training: bool, update: bool) -> None:
        if training and update:
            for p in _BN_LAYERS:
                self.num_batches_tracked[p] += 3 if p.startswith(("image_decoder", "attrs_decoder")) else 2

    def run(self, training: bool = True, noise_given: bool = False, update: bool = True) -> None:
        # the graph key must also distinguish injected vs generated dropout masks
        self._graph_variant = bool(getattr(self, "_masks_given", False))
        key = (training, noise_given, update, self._graph_variant)
        if self.use_graph and key[:3] in self._graphs and self._graphs.get(("variant", key[:3])) != self._graph_variant:
            self._graphs.pop(key[:3])
        self._graphs[("variant", key[:3])] = self._graph_variant
        super().run(training=training, noise_given=noise_given, update=update)

    def _warmup_state(self):
        # BatchNorm running statistics live outside the arena and ARE updated by the eager warm-up step that precedes a
        # graph capture: without restoring them the first step of every graph key applied the momentum update twice
        return super()._warmup_state() + [self.buffers[k] for k in sorted(self.buffers)]

    # ------------------------------------------------------------------ helpers
    def _cols(self, x, buf, n, H, W, Cch, stride, pad):
        """Operand of a conv GEMM over the NHWC activation x [n,H,W,C]: (tensor, view).  Implicit (default, C % 32 == 0):
        the activation itself + its im2col view, fetched by TMA im2col loads inside the GEMM -- nothing is written to HBM.
        Otherwise the im2col matrix is materialised in `buf` (the 3-channel layers; MVAE_IMPLICIT_CONV=0)."""
        if self.implicit_conv and Cch % 32 == 0:
            return x, ops.conv_view(n, H, W, Cch, taps=4, stride=stride, pad=pad)
        ops.im2col_k4(x, buf, n, H, W, Cch, stride, pad)
        return buf, None

    def _bn_f(self, x, h, S, seg_rows, prefix, order, training, act=True):
        p = self.params
        ops.bn_forward(x, h, S, seg_rows, p[prefix + ".weight"], p[prefix + ".bias"], self.bn_mean[prefix],
                       self.bn_invstd[prefix], self.bn_acc, self.buffers[prefix + ".running_mean"],
                       self.buffers[prefix + ".running_var"], update_order=order, training=training, act=act)

    def _bn_b(self, x, dh, dx, S, seg_rows, seg0, nseg, prefix, act=True):
        p, g = self.params, self.grads
        ops.bn_backward(x, dh, dx, S, seg_rows, seg0, nseg, p[prefix + ".weight"], p[prefix + ".bias"],
                        self.bn_mean[prefix], self.bn_invstd[prefix], self.bn_acc, g[prefix + ".weight"],
                        g[prefix + ".bias"], act=act, training=self._bn_training)

    @staticmethod
    def _split(rows):
        """split-K of a wgrad whose reduction runs over `rows`: ~32 k-blocks (1024 rows) per tile.  (The first version
        capped the split at 64: the C=3 conv layers reduce over 2 M rows, which left 64 CTAs with 1024 k-blocks each --
        0.6 ms on the critical path of the step while 84 SMs idled.)"""
        return max(1, min(rows // 1024, 4096))

    # ------------------------------------------------------------------ forward
    def _enqueue_forward(self, training: bool, use_noise_input: bool) -> None:
        B, L, P = self.B, self.L, self.prec
        p = self.params
        e, d = "image_encoder.features", "image_decoder.hallucinate"
        G, D = ops.gemm_batch, ops.gemm_desc
        SW = ops.EPI_BIAS_SWISH
        # ---- image encoder (evaluated once; its BN running stats are updated twice: joint + image-only call)
        ops.im2col_k4(self.x, self.cols1, B, 64, 64, 3, 2, 1)
        G([D(self.cols1, p[f"{e}.0.weight"], self.c1_a, B * 1024, 32, 48, out2=self.c1_h, epilogue=SW)], P)
        c2, v2 = self._cols(self.c1_h, self.cols2, B, 32, 32, 32, 2, 1)
        G([D(c2, p[f"{e}.2.weight"], self.c2_x, B * 256, 64, 512, a_view=v2),
           D(self.a_in, p["attrs_encoder.net.0.weight"], self.ae1_x, B, 512, 20, bias=p["attrs_encoder.net.0.bias"])], P)
        self._bn_f(self.c2_x, self.c2_h, 1, B * 256, f"{e}.3", (0, 0), training)
        self._bn_f(self.ae1_x, self.ae1_h, 1, B, "attrs_encoder.net.1", (0, 0), training)
        c3, v3 = self._cols(self.c2_h, self.cols3, B, 16, 16, 64, 2, 1)
        G([D(c3, p[f"{e}.5.weight"], self.c3_x, B * 64, 128, 1024, a_view=v3),
           D(self.ae1_h, p["attrs_encoder.net.3.weight"], self.ae2_x, B, 512, 512, bias=p["attrs_encoder.net.3.bias"])], P)
        self._bn_f(self.c3_x, self.c3_h, 1, B * 64, f"{e}.6", (0, 0), training)
        self._bn_f(self.ae2_x, self.ae2_h, 1, B, "attrs_encoder.net.4", (0, 0), training)
        c4, v4 = self._cols(self.c3_h, self.cols4, B, 8, 8, 128, 1, 0)
        G([D(c4, p[f"{e}.8.weight"], self.c4_x, B * 25, 256, 2048, a_view=v4),
           D(self.ae2_h, p["attrs_encoder.net.6.weight"], self.enc_t, B, 2 * L, 512, bias=p["attrs_encoder.net.6.bias"])], P)
        self._bn_f(self.c4_x, self.c4_h, 1, B * 25, f"{e}.9", (0, 0), training)
        G([D(self.c4_h.view(B, 6400), p["image_encoder.classifier.0.weight"], self.fc_a, B, 512, 6400,
             bias=p["image_encoder.classifier.0.bias"], out2=self.fc_h, epilogue=SW)], P)
        if training:
            given = bool(getattr(self, "_masks_given", False))
            ops.dropout_fwd(self.fc_h, self.fcd, 2, 0.1, mask_out=None if given else self.drop_mask,
                            mask_in=self.drop_mask if given else None, seed=self.seed * 7919 + self.rank,
                            step_dev=self.step_count)
        else:
            self.fcd[:B].copy_(self.fc_h); self.fcd[B:].copy_(self.fc_h)
        G([D(self.fcd, p["image_encoder.classifier.3.weight"], self.enc_i2, 2 * B, 2 * L, 512,
             bias=p["image_encoder.classifier.3.bias"])], P)
        # ---- PoE (variant B) + reparametrise + KL: experts (image enc | image-only mask, image enc | joint mask, attrs)
        mu_e = [self.enc_i2[:B, :L], self.enc_i2[B:, :L], self.enc_t[:, :L]]
        lv_e = [self.enc_i2[:B, L:], self.enc_i2[B:, L:], self.enc_t[:, L:]]
        ops.poe_fwd(mu_e, lv_e, _PASS_MASKS3, B, L, self.Z, variant=1, training=training,
                    noise=self.noise if (training and use_noise_input) else None,
                    noise_out=self.noise if (training and not use_noise_input) else None,
                    seed=self.seed * 1000003 + self.rank, offset=0, step_dev=self.step_count, kl_acc=self.acc[6:9])
        # ---- decoders on all three passes (3B rows, BN statistics per pass)
        G([D(self.Z, p["image_decoder.upsample.0.weight"], self.d0_a, 3 * B, 6400, L,
             bias=p["image_decoder.upsample.0.bias"], out2=self.d0_h, epilogue=SW),
           D(self.Z, p["attrs_decoder.net.0.weight"], self.ad_x[0], 3 * B, 512, L, bias=p["attrs_decoder.net.0.bias"])], P)
        self._bn_f(self.ad_x[0], self.ad_h[0], 3, B, "attrs_decoder.net.1", _BN_ORDER3, training)
        sub = self.subpixel
        GC = lambda descs: ops.gemm_chain(descs, [-1] * len(descs), self.chain_ws, P)   # noqa: E731  (> 4 problems / row maps)
        ad1 = D(self.ad_h[0], p["attrs_decoder.net.3.weight"], self.ad_x[1], 3 * B, 512, 512, bias=p["attrs_decoder.net.3.bias"])
        # (the 5x5 -> 8x8 stride-1 layer stays GEMM -> cols -> col2im: as ONE implicit GEMM over the 8x8 outputs
        # (ops.full_k4s1p0) it multiplies mostly padding -- 2.6x the FLOPs; measured 953 vs 616 us at B = 1024)
        G([D(self.d0_h.view(3 * B * 25, 256), p[f"{d}.0.weight"], self.colsT1, 3 * B * 25, 2048, 256), ad1], P)
        ops.col2im_k4(self.colsT1, self.t1_x, 3 * B, 5, 5, 128, 1, 0)
        self._bn_f(self.t1_x, self.t1_h, 3, B * 64, f"{d}.1", _BN_ORDER3, training)
        self._bn_f(self.ad_x[1], self.ad_h[1], 3, B, "attrs_decoder.net.4", _BN_ORDER3, training)
        ad2 = D(self.ad_h[1], p["attrs_decoder.net.6.weight"], self.ad_x[2], 3 * B, 512, 512, bias=p["attrs_decoder.net.6.bias"])
        if sub:
            GC(ops.subpixel_k4s2p1(self.t1_h, p[f"{d}.3.weight"], self.t2_x, 3 * B, 8, 8, 128, 64) + [ad2])
        else:
            G([D(self.t1_h, p[f"{d}.3.weight"], self.colsT2, 3 * B * 64, 1024, 128), ad2], P)
            ops.col2im_k4(self.colsT2, self.t2_x, 3 * B, 8, 8, 64, 2, 1)
        self._bn_f(self.t2_x, self.t2_h, 3, B * 256, f"{d}.4", _BN_ORDER3, training)
        self._bn_f(self.ad_x[2], self.ad_h[2], 3, B, "attrs_decoder.net.7", _BN_ORDER3, training)
        ad3 = D(self.ad_h[2], p["attrs_decoder.net.9.weight"], self.logit_a, 3 * B, N_ATTRS, 512, bias=p["attrs_decoder.net.9.bias"])
        if sub:
            GC(ops.subpixel_k4s2p1(self.t2_h, p[f"{d}.6.weight"], self.t3_x, 3 * B, 16, 16, 64, 32) + [ad3])
        else:
            G([D(self.t2_h, p[f"{d}.6.weight"], self.colsT3, 3 * B * 256, 512, 64), ad3], P)
            ops.col2im_k4(self.colsT3, self.t3_x, 3 * B, 16, 16, 32, 2, 1)
        self._bn_f(self.t3_x, self.t3_h, 3, B * 1024, f"{d}.7", _BN_ORDER3, training)
        G([D(self.t3_h, p[f"{d}.9.weight"], self.colsT4, 3 * B * 1024, 48, 32)], P)
        ops.col2im_k4(self.colsT4, self.logit_i, 3 * B, 32, 32, 3, 2, 1)

    # ------------------------------------------------------------------ loss + backward (live rows only)
    def _enqueue_loss_and_backward(self, training: bool, b_global: int) -> None:
        B, L, P = self.B, self.L, self.prec
        p, g = self.params, self.grads
        e, d = "image_encoder.features", "image_decoder.hallucinate"
        G, D = ops.gemm_batch, ops.gemm_desc
        sp = self._split
        self._bn_training = training
        li = self.logit_i[: 2 * B]                       # image decoder live rows: passes (image-only, joint)
        ops.bce_logits_fwd_bwd(li, self.x, li, self.lam_i / b_global, self.acc[0:3], seg_rows=B)
        la, dla = self.logit_a[B:], self.dlogit_a[B:]    # attrs decoder live rows: passes (joint, attrs-only)
        ops.bce_logits_fwd_bwd(la, self.a_in[:, :N_ATTRS], dla, self.lam_t / b_global, self.acc[4:6], seg_rows=B)
        # ---- image decoder backward
        R = 2 * B
        ops.im2col_k4(li, self.dcolsT4, R, 64, 64, 3, 2, 1)
        ops.colsum_accumulate(dla, g["attrs_decoder.net.9.bias"])
        G([D(self.dcolsT4, self.t3_h[: R * 1024], g[f"{d}.9.weight"], 48, 32, R * 1024, a_mn=True, b_mn=True,
             split_k=sp(R * 1024), accumulate=True),
           D(self.dcolsT4, p[f"{d}.9.weight"], self.d_t3h, R * 1024, 32, 48, b_mn=True),
           D(dla, self.ad_h[2][B:], g["attrs_decoder.net.9.weight"], N_ATTRS, 512, R, a_mn=True, b_mn=True,
             split_k=sp(R), accumulate=True),
           D(dla, p["attrs_decoder.net.9.weight"], self.d_adh[B:], R, 512, N_ATTRS, b_mn=True)], P)
        self._bn_b(self.t3_x, self.d_t3h, self.d_t3x, 3, B * 1024, 0, 2, f"{d}.7")
        # d_t3h/d_t3x hold live rows [0, 2B*1024): bn_bwd indexes rows globally from segment 0 -> consistent
        self._bn_b(self.ad_x[2], self.d_adh, self.d_adx, 3, B, 1, 2, "attrs_decoder.net.7")
        dc3, w3 = self._cols(self.d_t3x, self.dcolsT3, R, 32, 32, 32, 2, 1)
        ops.colsum_accumulate(self.d_adx[B:], g["attrs_decoder.net.6.bias"])
        G([D(dc3, self.t2_h[: R * 256], g[f"{d}.6.weight"], 512, 64, R * 256, a_mn=True, b_mn=True,
             split_k=sp(R * 256), accumulate=True, a_view=w3),
           D(dc3, p[f"{d}.6.weight"], self.d_t2h, R * 256, 64, 512, b_mn=True, a_view=w3),
           D(self.d_adx[B:], self.ad_h[1][B:], g["attrs_decoder.net.6.weight"], 512, 512, R, a_mn=True, b_mn=True,
             split_k=sp(R), accumulate=True),
           D(self.d_adx[B:], p["attrs_decoder.net.6.weight"], self.d_adh[B:], R, 512, 512, b_mn=True)], P)
        self._bn_b(self.t2_x, self.d_t2h, self.d_t2x, 3, B * 256, 0, 2, f"{d}.4")
        self._bn_b(self.ad_x[1], self.d_adh, self.d_adx, 3, B, 1, 2, "attrs_decoder.net.4")
        dc2, w2 = self._cols(self.d_t2x, self.dcolsT2, R, 16, 16, 64, 2, 1)
        ops.colsum_accumulate(self.d_adx[B:], g["attrs_decoder.net.3.bias"])
        G([D(dc2, self.t1_h[: R * 64], g[f"{d}.3.weight"], 1024, 128, R * 64, a_mn=True, b_mn=True,
             split_k=sp(R * 64), accumulate=True, a_view=w2),
           D(dc2, p[f"{d}.3.weight"], self.d_t1h, R * 64, 128, 1024, b_mn=True, a_view=w2),
           D(self.d_adx[B:], self.ad_h[0][B:], g["attrs_decoder.net.3.weight"], 512, 512, R, a_mn=True, b_mn=True,
             split_k=sp(R), accumulate=True),
           D(self.d_adx[B:], p["attrs_decoder.net.3.weight"], self.d_adh[B:], R, 512, 512, b_mn=True)], P)
        self._bn_b(self.t1_x, self.d_t1h, self.d_t1x, 3, B * 64, 0, 2, f"{d}.1")
        self._bn_b(self.ad_x[0], self.d_adh, self.d_adx, 3, B, 1, 2, "attrs_decoder.net.1")
        dc1, w1 = self._cols(self.d_t1x, self.dcolsT1, R, 8, 8, 128, 1, 0)
        ops.colsum_accumulate(self.d_adx[B:], g["attrs_decoder.net.0.bias"])
        G([D(dc1, self.d0_h.view(3 * B * 25, 256)[: R * 25], g[f"{d}.0.weight"], 2048, 256, R * 25, a_mn=True,
             b_mn=True, split_k=sp(R * 25), accumulate=True, a_view=w1),
           D(dc1, p[f"{d}.0.weight"], self.d_d0.view(R * 25, 256), R * 25, 256, 2048, b_mn=True,
             aux=self.d0_a.view(3 * B * 25, 256)[: R * 25], epilogue=ops.EPI_MUL_DSWISH, a_view=w1),
           D(self.d_adx[B:], self.Z[B:], g["attrs_decoder.net.0.weight"], 512, L, R, a_mn=True, b_mn=True,
             split_k=sp(R), accumulate=True),
           D(self.d_adx[B:], p["attrs_decoder.net.0.weight"], self.dZ[B:], R, L, 512, b_mn=True, accumulate=True)], P)
        ops.colsum_accumulate(self.d_d0, g["image_decoder.upsample.0.bias"])
        G([D(self.d_d0, self.Z[:R], g["image_decoder.upsample.0.weight"], 6400, L, R, a_mn=True, b_mn=True,
             split_k=sp(R), accumulate=True),
           D(self.d_d0, p["image_decoder.upsample.0.weight"], self.dZ[:R], R, L, 6400, b_mn=True, accumulate=True)], P)
        # ---- PoE / reparam / KL backward
        mu_e = [self.enc_i2[:B, :L], self.enc_i2[B:, :L], self.enc_t[:, :L]]
        lv_e = [self.enc_i2[:B, L:], self.enc_i2[B:, L:], self.enc_t[:, L:]]
        dmu = [self.d_enc_i2[:B, :L], self.d_enc_i2[B:, :L], self.d_enc_t[:, :L]]
        dlv = [self.d_enc_i2[:B, L:], self.d_enc_i2[B:, L:], self.d_enc_t[:, L:]]
        ops.poe_bwd(mu_e, lv_e, _PASS_MASKS3, B, L, self.dZ, dmu, dlv, kl_scale=1.0 / b_global, variant=1,
                    training=training, noise=self.noise if training else None, kl_scale_dev=self.beta_dev)
        # ---- encoders: last Linear layers
        ops.colsum_accumulate(self.d_enc_i2, g["image_encoder.classifier.3.bias"])
        ops.colsum_accumulate(self.d_enc_t, g["attrs_encoder.net.6.bias"])
        G([D(self.d_enc_i2, self.fcd, g["image_encoder.classifier.3.weight"], 2 * L, 512, 2 * B, a_mn=True, b_mn=True,
             split_k=sp(2 * B), accumulate=True),
           D(self.d_enc_i2, p["image_encoder.classifier.3.weight"], self.d_fcd, 2 * B, 512, 2 * L, b_mn=True),
           D(self.d_enc_t, self.ae2_h, g["attrs_encoder.net.6.weight"], 2 * L, 512, B, a_mn=True, b_mn=True,
             split_k=sp(B), accumulate=True),
           D(self.d_enc_t, p["attrs_encoder.net.6.weight"], self.d_aeh, B, 512, 2 * L, b_mn=True)], P)
        if training:
            ops.dropout_bwd(self.d_fcd, self.drop_mask, self.d_fch, 2, 0.1)
        else:
            torch.add(self.d_fcd[:B], self.d_fcd[B:], out=self.d_fch)
        ops.swish_bwd(self.fc_a, self.d_fch, self.d_fca)
        self._bn_b(self.ae2_x, self.d_aeh, self.d_aex, 1, B, 0, 1, "attrs_encoder.net.4")
        ops.colsum_accumulate(self.d_fca, g["image_encoder.classifier.0.bias"])
        ops.colsum_accumulate(self.d_aex, g["attrs_encoder.net.3.bias"])
        G([D(self.d_fca, self.c4_h.view(B, 6400), g["image_encoder.classifier.0.weight"], 512, 6400, B, a_mn=True, b_mn=True,
             split_k=sp(B), accumulate=True),
           D(self.d_fca, p["image_encoder.classifier.0.weight"], self.d_c4h.view(B, 6400), B, 6400, 512, b_mn=True),
           D(self.d_aex, self.ae1_h, g["attrs_encoder.net.3.weight"], 512, 512, B, a_mn=True, b_mn=True, split_k=sp(B),
             accumulate=True),
           D(self.d_aex, p["attrs_encoder.net.3.weight"], self.d_aeh, B, 512, 512, b_mn=True)], P)
        self._bn_b(self.c4_x, self.d_c4h, self.d_c4x, 1, B * 25, 0, 1, f"{e}.9")
        self._bn_b(self.ae1_x, self.d_aeh, self.d_aex, 1, B, 0, 1, "attrs_encoder.net.1")
        ops.colsum_accumulate(self.d_aex, g["attrs_encoder.net.0.bias"])
        imp = self.implicit_conv
        V = ops.conv_view
        G([D(self.d_c4x, self.c3_h if imp else self.cols4, g[f"{e}.8.weight"], 256, 2048, B * 25, a_mn=True, b_mn=True,
             split_k=sp(B * 25), accumulate=True, b_view=V(B, 8, 8, 128, stride=1, pad=0) if imp else None),
           D(self.d_c4x, p[f"{e}.8.weight"], self.dcols4, B * 25, 2048, 256, b_mn=True),
           D(self.d_aex, self.a_in, g["attrs_encoder.net.0.weight"], 512, 20, B, a_mn=True, b_mn=True, split_k=sp(B),
             accumulate=True)], P)
        GC = lambda descs: ops.gemm_chain(descs, [-1] * len(descs), self.chain_ws, P)   # noqa: E731
        ops.col2im_k4(self.dcols4, self.d_c3h, B, 5, 5, 128, 1, 0)
        self._bn_b(self.c3_x, self.d_c3h, self.d_c3x, 1, B * 64, 0, 1, f"{e}.6")
        G([D(self.d_c3x, self.c2_h if imp else self.cols3, g[f"{e}.5.weight"], 128, 1024, B * 64, a_mn=True, b_mn=True,
             split_k=sp(B * 64), accumulate=True, b_view=V(B, 16, 16, 64) if imp else None),
           ] + ([] if self.subpixel else [D(self.d_c3x, p[f"{e}.5.weight"], self.dcols3, B * 64, 1024, 128, b_mn=True)]), P)
        if self.subpixel:
            GC(ops.subpixel_k4s2p1(self.d_c3x, p[f"{e}.5.weight"], self.d_c2h, B, 8, 8, 128, 64, w_is_conv=True))
        else:
            ops.col2im_k4(self.dcols3, self.d_c2h, B, 8, 8, 64, 2, 1)
        self._bn_b(self.c2_x, self.d_c2h, self.d_c2x, 1, B * 256, 0, 1, f"{e}.3")
        G([D(self.d_c2x, self.c1_h if imp else self.cols2, g[f"{e}.2.weight"], 64, 512, B * 256, a_mn=True, b_mn=True,
             split_k=sp(B * 256), accumulate=True, b_view=V(B, 32, 32, 32) if imp else None),
           ] + ([] if self.subpixel else [D(self.d_c2x, p[f"{e}.2.weight"], self.dcols2, B * 256, 512, 64, b_mn=True)]), P)
        if self.subpixel:
            GC(ops.subpixel_k4s2p1(self.d_c2x, p[f"{e}.2.weight"], self.d_c1a, B, 16, 16, 64, 32, w_is_conv=True, aux=self.c1_a,
                                   epilogue=ops.EPI_MUL_DSWISH))
        else:
            ops.col2im_k4(self.dcols2, self.d_c1a, B, 16, 16, 32, 2, 1, aux=self.c1_a)
        G([D(self.d_c1a, self.cols1, g[f"{e}.0.weight"], 32, 48, B * 1024, a_mn=True, b_mn=True, split_k=sp(B * 1024),
             accumulate=True)], P)

    def losses(self) -> Dict[str, float]:
        v = self.loss_host
        return {"total": float(v[0]), "joint": float(v[1 + _REF_TO_INTERNAL[0]]),
                "image": float(v[1 + _REF_TO_INTERNAL[1]]), "attrs": float(v[1 + _REF_TO_INTERNAL[2]])}
