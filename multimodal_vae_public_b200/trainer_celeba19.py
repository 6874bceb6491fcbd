"""Training step of the CelebA-19 MVAE: image + 18 single-attribute experts, 20 + approx_m ELBO terms per step
(joint, image-only, 18 single-attribute terms, approx_m sampled modality subsets) -- celeba19/model.py:14-226,
celeba19/train.py:26-142,240-309.

Built from the same kernels as the other flavours; what is specific here:
  * ONE PoE/reparametrise/KL launch fuses all P = 20 + approx_m passes over E = (#image passes) + 18 experts: every
    image-containing pass has its own Dropout mask, hence its own image expert; every attribute encoder is one expert.
  * Passes are ordered (image-only, joint, sampled-with-image, singles, sampled-without-image) so that the rows of Z that
    feed the image decoder's BACKWARD are contiguous.  The image decoder's FORWARD runs on all P passes because the
    reference calls it in every pass and each call updates the BatchNorm running statistics (P row segments, reference
    call order); the 18 attribute decoders have no BatchNorm, so each runs only on the passes whose loss uses it
    (joint, its single-attribute pass, sampled subsets containing it), gathered into a contiguous row block.
  * Loss weights follow the reference's quirk: the joint and image-only terms use the script's lambda_image /
    lambda_attrs, single-attribute and sampled terms call elbo_loss with its defaults (1.0).
  * The modality subsets change every step, so the step is enqueued eagerly (no CUDA graph capture); the host-side
    sampler (``sample_combinations``) is bit-compatible with the reference's numpy draws without materialising its
    524,267 x 19 pool.
"""
from __future__ import annotations

import math
from math import comb
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .trainer import FlatArena
from .trainer_celeba import (CelebAMVAETrainer, N_ATTRS, _BN_LAYERS, _CONVT, _internal_shape, _to_internal, _to_reference,
                             celeba_param_shapes)

_IMG_BN = {k: v for k, v in _BN_LAYERS.items() if k.startswith("image_")}


def celeba19_param_shapes(L: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Reference PARAMETER names/shapes (celeba19/model.py:21-31): image nets, 18 attr encoders, 18 attr decoders."""
    out = [(k, s) for k, s in celeba_param_shapes(L) if k.startswith("image_")]
    for i in range(N_ATTRS):
        p = f"attr_encoders.{i}.net"
        out += [(f"{p}.0.weight", (2, 512)), (f"{p}.2.weight", (512, 512)), (f"{p}.2.bias", (512,)),
                (f"{p}.4.weight", (2 * L, 512)), (f"{p}.4.bias", (2 * L,))]
    for i in range(N_ATTRS):
        p = f"attr_decoders.{i}.net"
        out += [(f"{p}.0.weight", (512, L)), (f"{p}.0.bias", (512,)), (f"{p}.2.weight", (512, 512)), (f"{p}.2.bias", (512,)),
                (f"{p}.4.weight", (512, 512)), (f"{p}.4.bias", (512,)), (f"{p}.6.weight", (1, 512)), (f"{p}.6.bias", (1,))]
    return out


def unrank_combination(n: int, k: int, idx: int) -> List[int]:
    """idx-th k-subset of range(n) in itertools.combinations order."""
    out, x = [], 0
    for j in range(k):
        while True:
            c = comb(n - x - 1, k - j - 1)
            if idx < c:
                out.append(x); x += 1
                break
            idx -= c; x += 1
    return out


def sample_combinations(n_modalities: int = 19, size: int = 1, rng=np.random) -> np.ndarray:
    """Host-side sampler of celeba19/train.py:111-142 (subset size ~ U{2..n-1}, then a uniform subset of that size),
    consuming the numpy RNG exactly like the reference's pool-based version (same draws, same result)."""
    sample_pool = rng.choice(np.arange(2, n_modalities), size, replace=True)
    dist = np.bincount(sample_pool, minlength=n_modalities)
    rows = []
    for k in range(n_modalities):
        if dist[k] > 0:
            for i in rng.choice(range(comb(n_modalities, k)), size=dist[k], replace=False):
                row = np.zeros(n_modalities, dtype=bool)
                row[unrank_combination(n_modalities, k, int(i))] = True
                rows.append(row)
    return np.stack(rows)


class CelebA19MVAETrainer(CelebAMVAETrainer):
    """``step(image [B,3,64,64], attrs [B,18], combos=None)``; combos = bool [approx_m, 19] (sampled when None)."""

    def __init__(self, n_latents: int = 100, batch_size: int = 64, approx_m: int = 1, **kw):
        kw.setdefault("dp_mode", "nccl")   # this flavour runs its own exchange (gradients + per-term losses) through NCCL
        self.approx_m = approx_m
        self.P = 20 + approx_m
        self.n_img_max = 2 + approx_m
        if self.P > 32 or self.n_img_max + N_ATTRS > 24:
            raise _lib.MvaeError("approx_m too large for the fused PoE kernel (P <= 32 passes, E <= 24 experts)")
        kw["use_graph"] = False           # the pass structure changes every step
        super().__init__(n_latents=n_latents, batch_size=batch_size, **kw)
        self.dZ = torch.zeros(self.P * batch_size, n_latents, dtype=torch.float32, device=self.dev)

    # ------------------------------------------------------------------ layout / buffers
    def _make_layout(self, L: int):
        self._ref_shapes = dict(celeba19_param_shapes(L))
        return [(k, _internal_shape(k, s, L)) for k, s in celeba19_param_shapes(L)]

    def _alloc_activations(self, f) -> None:
        B, L, dev, P, NI = self.B, self.L, self.dev, self.P, self.n_img_max
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        self.x_nchw, self.x = f(B, 3 * 4096), f(B, 12288)
        self.attrs_t = torch.zeros(N_ATTRS, B, dtype=torch.float32, device=dev)      # attribute columns, contiguous
        self.attrs_idx = torch.zeros(N_ATTRS, B, dtype=torch.int64, device=dev)
        self.drop_mask = torch.ones(NI * B, 512, dtype=torch.float32, device=dev)
        self.Z, self.noise = f(P * B, L), f(P * B, L)
        self.acc19 = torch.zeros(P + NI + N_ATTRS * (2 + self.approx_m), dtype=torch.float64, device=dev)
        self.buffers, self.bn_mean, self.bn_invstd = {}, {}, {}
        for prefix, c in _IMG_BN.items():
            self.buffers[prefix + ".running_mean"] = z(c)
            self.buffers[prefix + ".running_var"] = torch.ones(c, dtype=torch.float32, device=dev)
            self.bn_mean[prefix] = f(P, c); self.bn_invstd[prefix] = f(P, c)
        self.num_batches_tracked = {p: 0 for p in _IMG_BN}
        self.bn_acc = torch.zeros(P * 256 * 2, dtype=torch.float64, device=dev)
        # image encoder (B rows)
        self.cols1 = f(B * 1024, 48); self.c1_a, self.c1_h = f(B * 1024, 32), f(B * 1024, 32)
        self.cols2 = f(B * 256, 512); self.c2_x, self.c2_h = f(B * 256, 64), f(B * 256, 64)
        self.cols3 = f(B * 64, 1024); self.c3_x, self.c3_h = f(B * 64, 128), f(B * 64, 128)
        self.cols4 = f(B * 25, 2048); self.c4_x, self.c4_h = f(B * 25, 256), f(B * 25, 256)
        self.fc_a, self.fc_h = f(B, 512), f(B, 512)
        self.fcd = f(NI * B, 512)
        self.enc_img, self.d_enc_img = f(NI * B, 2 * L), f(NI * B, 2 * L)
        # attribute encoders (18 x B rows)
        self.ae_h1 = f(N_ATTRS, B, 512); self.ae_a2 = f(N_ATTRS, B, 512); self.ae_h2 = f(N_ATTRS, B, 512)
        self.enc_a, self.d_enc_a = f(N_ATTRS, B, 2 * L), f(N_ATTRS, B, 2 * L)
        self.ae_dA = f(N_ATTRS, B, 512); self.ae_dh1 = f(N_ATTRS, B, 512)
        # image decoder (forward on P passes, backward on the NI live ones)
        self.d0_a, self.d0_h = f(P * B, 6400), f(P * B, 6400)
        self.colsT1 = f(P * B * 25, 2048); self.t1_x, self.t1_h = f(P * B * 64, 128), f(P * B * 64, 128)
        self.colsT2 = f(P * B * 64, 1024); self.t2_x, self.t2_h = f(P * B * 256, 64), f(P * B * 256, 64)
        self.colsT3 = f(P * B * 256, 512); self.t3_x, self.t3_h = f(P * B * 1024, 32), f(P * B * 1024, 32)
        R = NI * B
        self.colsT4 = f(R * 1024, 48)          # the last ConvTranspose2d has no BatchNorm after it: live image passes only
        self.logit_i = f(R, 12288)
        self.dcolsT4 = f(R * 1024, 48); self.d_t3h, self.d_t3x = f(R * 1024, 32), f(R * 1024, 32)
        self.dcolsT3 = f(R * 256, 512); self.d_t2h, self.d_t2x = f(R * 256, 64), f(R * 256, 64)
        self.dcolsT2 = f(R * 64, 1024); self.d_t1h, self.d_t1x = f(R * 64, 128), f(R * 64, 128)
        self.dcolsT1 = f(R * 25, 2048); self.d_d0 = f(R, 6400)
        self.d_fcd, self.d_fch, self.d_fca = f(R, 512), f(B, 512), f(B, 512)
        self.d_c4h, self.d_c4x = f(B * 25, 256), f(B * 25, 256); self.dcols4 = f(B * 25, 2048)
        self.d_c3h, self.d_c3x = f(B * 64, 128), f(B * 64, 128); self.dcols3 = f(B * 64, 1024)
        self.d_c2h, self.d_c2x = f(B * 256, 64), f(B * 256, 64); self.dcols2 = f(B * 256, 512)
        self.d_c1a = f(B * 1024, 32)
        # attribute decoders: up to (2 + approx_m) live passes each, rows gathered contiguously
        nmax = (2 + self.approx_m) * B
        self.ad_z = f(N_ATTRS, nmax, L); self.ad_dz = f(N_ATTRS, nmax, L)
        self.ad_a = [f(N_ATTRS, nmax, 512) for _ in range(3)]; self.ad_h = [f(N_ATTRS, nmax, 512) for _ in range(3)]
        self.ad_logit = z(N_ATTRS, nmax, 4); self.ad_dlogit = z(N_ATTRS, nmax, 4)
        self.ad_dA = [f(N_ATTRS, nmax, 512) for _ in range(2)]
        self.ad_target = z(N_ATTRS, nmax, 1)               # attribute column i repeated for each live pass of decoder i
        self.loss19 = torch.zeros(1, dtype=torch.float32, device=dev)
        self.enc_t = self.d_enc_t = None   # unused base buffers

    # ------------------------------------------------------------------ parameters / state
    def init_parameters(self, seed: int = 0) -> None:
        g = torch.Generator(device="cpu").manual_seed(seed)
        sd = {}
        for name, shape in celeba19_param_shapes(self.L):
            prefix = name.rsplit(".", 1)[0]
            if prefix in _IMG_BN:
                sd[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            elif shape == (2, 512):
                sd[name] = torch.randn(shape, generator=g)
            else:
                wshape = self._ref_shapes[prefix + ".weight"]
                fan_in = wshape[1] * 16 if (prefix + ".weight") in _CONVT else int(math.prod(wshape[1:]))
                sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        self.load_state_dict(sd)
        self.adam_m.zero_(); self.adam_v.zero_(); self.step_count.zero_()

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        for k, _ in celeba19_param_shapes(self.L):
            self.params[k].copy_(_to_internal(k, sd[k].to(torch.float32)).contiguous())
        for k in self.buffers:
            if k in sd:
                self.buffers[k].copy_(sd[k].to(torch.float32))
        for p in _IMG_BN:
            if p + ".num_batches_tracked" in sd:
                self.num_batches_tracked[p] = int(sd[p + ".num_batches_tracked"])

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k, shp in celeba19_param_shapes(self.L):
            out[k] = _to_reference(k, self.params[k].detach(), shp).contiguous().clone()
            prefix = k.rsplit(".", 1)[0]
            if prefix in _IMG_BN and k.endswith(".bias"):
                out[prefix + ".running_mean"] = self.buffers[prefix + ".running_mean"].clone()
                out[prefix + ".running_var"] = self.buffers[prefix + ".running_var"].clone()
                out[prefix + ".num_batches_tracked"] = torch.tensor(self.num_batches_tracked[prefix], dtype=torch.int64)
        return out

    def export_grads(self) -> Dict[str, torch.Tensor]:
        return {k: _to_reference(k, self.grads[k].detach(), shp).contiguous().clone()
                for k, shp in celeba19_param_shapes(self.L)}

    # ------------------------------------------------------------------ pass structure of one step
    def _plan(self, combos: np.ndarray):
        """Internal pass order + bookkeeping.  Reference order: 0 joint, 1 image-only, 2..19 singles, 20.. sampled."""
        ref = [([True] * 19, True), ([True] + [False] * 18, True)]
        ref += [([False] + [k == i for k in range(N_ATTRS)], False) for i in range(N_ATTRS)]
        ref += [([bool(v) for v in c], False) for c in combos]
        samp_img = [20 + j for j, c in enumerate(combos) if c[0]]
        samp_no = [20 + j for j, c in enumerate(combos) if not c[0]]
        order = [1, 0] + samp_img + list(range(2, 20)) + samp_no          # internal index -> reference index
        n_img = 2 + len(samp_img)
        inv = {r: i for i, r in enumerate(order)}                         # reference index -> internal index
        masks = []
        for ip, r in enumerate(order):
            present, _ = ref[r]
            m = 0
            if present[0]:
                m |= 1 << ip                                              # image expert of this pass (ip < n_img)
            for i in range(N_ATTRS):
                if present[1 + i]:
                    m |= 1 << (n_img + i)
            masks.append(m)
        # reference call order of the image-containing passes (dropout masks / BN updates): joint, image-only, sampled
        img_call_order = [inv[0], inv[1]] + [inv[r] for r in samp_img]
        bn_order = [inv[r] for r in range(len(ref))]                      # image decoder BN updates: every pass, call order
        # attribute decoder i: live passes (internal ids), the joint pass first
        dec_passes = []
        for i in range(N_ATTRS):
            lp = [inv[0], inv[2 + i]] + [inv[20 + j] for j, c in enumerate(combos) if c[1 + i]]
            dec_passes.append(lp)
        # host-built index tables for the attribute decoders: Z rows to gather, ELBO term each BCE segment belongs to
        B, nseg = self.B, 2 + self.approx_m
        rowidx = np.zeros((N_ATTRS, nseg, B), dtype=np.int64)
        term_idx = np.zeros((N_ATTRS, nseg), dtype=np.int64)
        term_w = np.zeros((N_ATTRS, nseg), dtype=np.float64)
        for i, lp in enumerate(dec_passes):
            for k, ip in enumerate(lp):
                rowidx[i, k] = ip * B + np.arange(B)
                term_idx[i, k] = ip
                term_w[i, k] = self.lam_t if k == 0 else 1.0
        return {"ref": ref, "order": order, "inv": inv, "n_img": n_img, "masks": masks, "img_call_order": img_call_order,
                "bn_order": bn_order, "dec_passes": dec_passes, "P": len(order), "rowidx": rowidx.reshape(-1),
                "term_idx": term_idx.reshape(-1), "term_w": term_w.reshape(-1)}

    # ------------------------------------------------------------------ step
    def step(self, image, attrs, annealing_factor: float = 1.0, noise=None, training: bool = True, update: bool = True,
             sync: bool = True, drop_masks=None, combos: Optional[np.ndarray] = None):
        """noise: optional [P,B,L] in the REFERENCE's pass order; drop_masks: optional [n_img,B,512] in the reference's
        call order of the image-containing passes (joint, image-only, sampled...)."""
        B, L = self.B, self.L
        plan = self._sample_plan(combos)
        with torch.cuda.stream(self._stream):
            self.x_nchw.copy_(image.reshape(B, 3 * 4096), non_blocking=True)
            a = attrs.reshape(B, N_ATTRS).to(self.dev, non_blocking=True)
            self._stage_and_enqueue(plan, self.x_nchw, a, annealing_factor, noise, drop_masks, training, update)
        if sync:
            self._stream.synchronize()
            self.check_device_errors()
            return float(self.loss19.item())
        return None

    def _sample_plan(self, combos):
        """The step's pass structure: the sampled attribute subsets (celeba19/train.py:240-309; one draw per GLOBAL batch --
        rank 0's wins) and the plan of stacked passes built from them."""
        if combos is None:
            combos = sample_combinations(19, self.approx_m) if self.approx_m > 0 else np.zeros((0, 19), dtype=bool)
            if self.world > 1 and self.approx_m > 0:     # one global batch = one set of subsets: rank 0's draw wins
                import torch.distributed as dist
                t = torch.from_numpy(np.ascontiguousarray(combos).astype(np.uint8)).to(self.dev)
                dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
                combos = t.cpu().numpy().astype(bool)
        combos = np.asarray(combos, dtype=bool).reshape(-1, 19)
        if len(combos) != self.approx_m:
            raise _lib.MvaeError(f"expected {self.approx_m} sampled combinations, got {len(combos)}")
        plan = self._plan(combos)
        self._last_plan = plan
        return plan

    def _stage_and_enqueue(self, plan, img_nchw, a, annealing_factor, noise, drop_masks, training, update,
                           staged: Optional[torch.cuda.Event] = None) -> None:
        """(on the step stream) device-side staging of one batch -- NCHW -> NHWC, the attribute columns -- and the step's
        launches.  ``img_nchw`` [B, 3*4096] and ``a`` [B, 18] are device tensors; ``staged`` is recorded once they have
        been read (the pipelined path re-uses them for the upload after next)."""
        B, L = self.B, self.L
        ops.nchw_to_nhwc(img_nchw, self.x, B, 3, 4096)
        self.attrs_t.copy_(a.t().to(torch.float32)); self.attrs_idx.copy_(a.t().to(torch.int64))
        if staged is not None:
            staged.record(self._stream)
        if noise is not None:
            nz = self.noise.view(self.P, B, L)
            for ip, r in enumerate(plan["order"]):
                nz[ip].copy_(noise[r], non_blocking=True)
        if drop_masks is not None:
            for call_i, ip in enumerate(plan["img_call_order"]):
                self.drop_mask[ip * B:(ip + 1) * B].copy_(drop_masks[call_i], non_blocking=True)
        for k in ("rowidx", "term_idx", "term_w"):
            plan[k + "_dev"] = torch.as_tensor(plan[k], device=self.dev)
        n0 = _lib.launch_count()
        self._enqueue(plan, training, noise is not None, drop_masks is not None, float(annealing_factor), update)
        self.launches_per_step = _lib.launch_count() - n0
        if training and update:
            for p in _IMG_BN:
                self.num_batches_tracked[p] += plan["n_img"] if p.startswith("image_encoder") else plan["P"]

    def step_pipelined(self, image, attrs, annealing_factor: float = 1.0, training: bool = True, update: bool = True):
        """``step`` for HOST batches with the upload on a copy stream (it overlaps the previous step's kernels) and the loss
        read back one call late, like the other flavours' ``step_pipelined``; the pass structure is still re-planned on the
        host every call (the sampled subsets change the launches), which the asynchronous launches hide behind the
        previous step.  Returns the PREVIOUS call's loss (None on the first call); ``flush()`` returns the last one."""
        B = self.B
        if not hasattr(self, "_pipe"):
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._pipe = [{"img": torch.empty(B, 3 * 4096, dtype=torch.float32, device=self.dev),
                           "oth": torch.empty(B, N_ATTRS, dtype=torch.float32, device=self.dev),
                           "loss": torch.zeros(1, dtype=torch.float32).pin_memory(),
                           "up": torch.cuda.Event(), "free": torch.cuda.Event(), "done": torch.cuda.Event(), "busy": False}
                          for _ in range(2)]
            for slot in self._pipe:
                slot["free"].record(self._stream)
            self._pipe_idx = 0
        k = self._pipe_idx
        self._pipe_idx ^= 1
        cur, prev = self._pipe[k], self._pipe[k ^ 1]
        plan = self._sample_plan(None)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(cur["free"])            # the step stream finished reading this staging slot
            cur["img"].copy_(image.reshape(B, 3 * 4096), non_blocking=True)
            cur["oth"].copy_(attrs.reshape(B, N_ATTRS), non_blocking=True)
            cur["up"].record(self._copy_stream)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(cur["up"])
            self._stage_and_enqueue(plan, cur["img"], cur["oth"], annealing_factor, None, None, training, update,
                                    staged=cur["free"])
            cur["loss"].copy_(self.loss19.reshape(1), non_blocking=True)
            cur["done"].record(self._stream)
        cur["busy"] = True
        if prev["busy"]:
            prev["done"].synchronize()
            prev["busy"] = False
            return float(prev["loss"][0])
        return None

    def flush(self):
        """Wait for the in-flight pipelined step and return its loss."""
        out = None
        if hasattr(self, "_pipe"):
            for slot in (self._pipe[self._pipe_idx], self._pipe[self._pipe_idx ^ 1]):
                if slot["busy"]:
                    slot["done"].synchronize()
                    slot["busy"] = False
                    out = float(slot["loss"][0])
        self._stream.synchronize()
        self.check_device_errors()
        return out

    def losses(self):
        """Per-pass ELBO terms in the reference's order (after a synchronised step)."""
        plan = self._last_plan
        t = self._terms.cpu().tolist()
        return {"total": float(sum(t)), "terms": [t[plan["inv"][r]] for r in range(plan["P"])]}

    def _enqueue_step(self, training: bool, use_noise_input: bool, update: bool) -> None:
        """Re-enqueue the last step's pass structure (bench.py's per-kernel timing pass)."""
        self._enqueue(self._last_plan, training, use_noise_input, False, 1.0, update)

    def gemm_flops_last_step(self) -> float:
        """2*MAC of the GEMMs of the last step (depends on how many sampled subsets contained the image)."""
        B, L, plan = self.B, self.L, self._last_plan
        P, NI = plan["P"], plan["n_img"]
        enc_conv = 1024 * 32 * 48 + 256 * 64 * 512 + 64 * 128 * 1024 + 25 * 256 * 2048
        enc_i_f = enc_conv + 6400 * 512 + NI * 512 * 2 * L
        enc_i_b = 2 * enc_i_f - 1024 * 32 * 48
        enc_a = 3 * N_ATTRS * (512 * 512 + 512 * 2 * L)
        dec_i_bn = L * 6400 + 25 * 2048 * 256 + 64 * 1024 * 128 + 256 * 512 * 64
        dec_i = P * dec_i_bn + NI * 1024 * 48 * 32 + 2 * NI * (dec_i_bn + 1024 * 48 * 32)
        dec_a = 3 * sum(len(lp) for lp in plan["dec_passes"]) * (L * 512 + 2 * 512 * 512 + 512)
        return 2.0 * B * (enc_i_f + enc_i_b + enc_a + dec_i + dec_a)

    # ------------------------------------------------------------------ the launches of one step
    def _enqueue(self, plan, training: bool, noise_given: bool, masks_given: bool, beta: float, update: bool) -> None:
        B, L, Pp = self.B, self.L, self.prec
        P, NI = plan["P"], plan["n_img"]
        p, g = self.params, self.grads
        e, d = "image_encoder.features", "image_decoder.hallucinate"
        G, D = ops.gemm_batch, ops.gemm_desc
        SW, DS = ops.EPI_BIAS_SWISH, ops.EPI_MUL_DSWISH
        b_global = B * self.world
        sp = self._split
        self._bn_training = training
        self.grad_bucket.zero_(); self.dZ.zero_(); self.acc19.zero_(); self.ad_dz.zero_()
        if self.presplit:
            ops.split_lo(self.flat_params, self.params_lo)

        def batched(descs):
            for i in range(0, len(descs), 4):
                G(descs[i:i + 4], Pp)

        # ================================================================ forward
        # ---- image encoder (once; BN running stats updated once per image-containing pass)
        ops.im2col_k4(self.x, self.cols1, B, 64, 64, 3, 2, 1)
        G([D(self.cols1, p[f"{e}.0.weight"], self.c1_a, B * 1024, 32, 48, out2=self.c1_h, epilogue=SW)], Pp)
        c2, v2 = self._cols(self.c1_h, self.cols2, B, 32, 32, 32, 2, 1)
        G([D(c2, p[f"{e}.2.weight"], self.c2_x, B * 256, 64, 512, a_view=v2)], Pp)
        self._bn_f(self.c2_x, self.c2_h, 1, B * 256, f"{e}.3", (0,) * NI, training)
        c3, v3 = self._cols(self.c2_h, self.cols3, B, 16, 16, 64, 2, 1)
        G([D(c3, p[f"{e}.5.weight"], self.c3_x, B * 64, 128, 1024, a_view=v3)], Pp)
        self._bn_f(self.c3_x, self.c3_h, 1, B * 64, f"{e}.6", (0,) * NI, training)
        c4, v4 = self._cols(self.c3_h, self.cols4, B, 8, 8, 128, 1, 0)
        G([D(c4, p[f"{e}.8.weight"], self.c4_x, B * 25, 256, 2048, a_view=v4)], Pp)
        self._bn_f(self.c4_x, self.c4_h, 1, B * 25, f"{e}.9", (0,) * NI, training)
        G([D(self.c4_h.view(B, 6400), p["image_encoder.classifier.0.weight"], self.fc_a, B, 512, 6400,
             bias=p["image_encoder.classifier.0.bias"], out2=self.fc_h, epilogue=SW)], Pp)
        fcd = self.fcd[: NI * B]
        if training:
            ops.dropout_fwd(self.fc_h, fcd, NI, 0.1, mask_out=None if masks_given else self.drop_mask[: NI * B],
                            mask_in=self.drop_mask[: NI * B] if masks_given else None, seed=self.seed * 7919 + self.rank,
                            step_dev=self.step_count)
        else:
            for k in range(NI):
                fcd[k * B:(k + 1) * B].copy_(self.fc_h)
        enc_img = self.enc_img[: NI * B]
        G([D(fcd, p["image_encoder.classifier.3.weight"], enc_img, NI * B, 2 * L, 512,
             bias=p["image_encoder.classifier.3.bias"])], Pp)
        # ---- 18 attribute encoders
        for i in range(N_ATTRS):
            ops.embedding_swish_fwd(p[f"attr_encoders.{i}.net.0.weight"], self.attrs_idx[i], None, self.ae_h1[i])
        batched([D(self.ae_h1[i], p[f"attr_encoders.{i}.net.2.weight"], self.ae_a2[i], B, 512, 512,
                   bias=p[f"attr_encoders.{i}.net.2.bias"], out2=self.ae_h2[i], epilogue=SW) for i in range(N_ATTRS)])
        batched([D(self.ae_h2[i], p[f"attr_encoders.{i}.net.4.weight"], self.enc_a[i], B, 2 * L, 512,
                   bias=p[f"attr_encoders.{i}.net.4.bias"]) for i in range(N_ATTRS)])
        # ---- PoE (variant B) + reparametrise + KL, all passes in one launch
        mu_e = [enc_img[k * B:(k + 1) * B, :L] for k in range(NI)] + [self.enc_a[i][:, :L] for i in range(N_ATTRS)]
        lv_e = [enc_img[k * B:(k + 1) * B, L:] for k in range(NI)] + [self.enc_a[i][:, L:] for i in range(N_ATTRS)]
        kl_acc = self.acc19[:P]
        Z = self.Z[: P * B]
        ops.poe_fwd(mu_e, lv_e, plan["masks"], B, L, Z, variant=1, training=training,
                    noise=self.noise[: P * B] if (training and noise_given) else None,
                    noise_out=self.noise[: P * B] if (training and not noise_given) else None,
                    seed=self.seed * 1000003 + self.rank, offset=0, step_dev=self.step_count, kl_acc=kl_acc)
        # ---- image decoder forward on all P passes (BatchNorm statistics + running updates per pass, call order)
        bo = tuple(plan["bn_order"])
        G([D(Z, p["image_decoder.upsample.0.weight"], self.d0_a[: P * B], P * B, 6400, L,
             bias=p["image_decoder.upsample.0.bias"], out2=self.d0_h[: P * B], epilogue=SW)], Pp)
        G([D(self.d0_h[: P * B].view(P * B * 25, 256), p[f"{d}.0.weight"], self.colsT1[: P * B * 25], P * B * 25, 2048, 256)], Pp)
        ops.col2im_k4(self.colsT1, self.t1_x, P * B, 5, 5, 128, 1, 0)
        self._bn_f(self.t1_x[: P * B * 64], self.t1_h[: P * B * 64], P, B * 64, f"{d}.1", bo, training)
        GC = lambda descs: ops.gemm_chain(descs, [-1] * len(descs), self.chain_ws, Pp)   # noqa: E731
        if self.subpixel:   # stride-2 transposed convs as four sub-pixel implicit GEMMs (no cols, no col2im)
            GC(ops.subpixel_k4s2p1(self.t1_h, p[f"{d}.3.weight"], self.t2_x, P * B, 8, 8, 128, 64))
        else:
            G([D(self.t1_h[: P * B * 64], p[f"{d}.3.weight"], self.colsT2[: P * B * 64], P * B * 64, 1024, 128)], Pp)
            ops.col2im_k4(self.colsT2, self.t2_x, P * B, 8, 8, 64, 2, 1)
        self._bn_f(self.t2_x[: P * B * 256], self.t2_h[: P * B * 256], P, B * 256, f"{d}.4", bo, training)
        if self.subpixel:
            GC(ops.subpixel_k4s2p1(self.t2_h, p[f"{d}.6.weight"], self.t3_x, P * B, 16, 16, 64, 32))
        else:
            G([D(self.t2_h[: P * B * 256], p[f"{d}.6.weight"], self.colsT3[: P * B * 256], P * B * 256, 512, 64)], Pp)
            ops.col2im_k4(self.colsT3, self.t3_x, P * B, 16, 16, 32, 2, 1)
        self._bn_f(self.t3_x[: P * B * 1024], self.t3_h[: P * B * 1024], P, B * 1024, f"{d}.7", bo, training)
        R = NI * B                                     # image logits: only the passes whose loss has an image term
        G([D(self.t3_h[: R * 1024], p[f"{d}.9.weight"], self.colsT4[: R * 1024], R * 1024, 48, 32)], Pp)
        ops.col2im_k4(self.colsT4, self.logit_i, R, 32, 32, 3, 2, 1)
        # ---- attribute decoders on their live passes (rows gathered from Z)
        nrows = [len(lp) * B for lp in plan["dec_passes"]]
        torch.index_select(Z, 0, plan["rowidx_dev"], out=self.ad_z.view(-1, L))
        xin = [self.ad_z[i][: nrows[i]] for i in range(N_ATTRS)]
        for l, key in enumerate(("0", "2", "4")):
            K = L if l == 0 else 512
            batched([D(xin[i], p[f"attr_decoders.{i}.net.{key}.weight"], self.ad_a[l][i][: nrows[i]], nrows[i], 512, K,
                       bias=p[f"attr_decoders.{i}.net.{key}.bias"], out2=self.ad_h[l][i][: nrows[i]], epilogue=SW)
                     for i in range(N_ATTRS)])
            xin = [self.ad_h[l][i][: nrows[i]] for i in range(N_ATTRS)]
        batched([D(xin[i], p[f"attr_decoders.{i}.net.6.weight"], self.ad_logit[i][: nrows[i], :1], nrows[i], 1, 512,
                   bias=p[f"attr_decoders.{i}.net.6.bias"]) for i in range(N_ATTRS)])

        # ================================================================ losses (+ dlogits)
        li = self.logit_i[:R]
        acc_img = self.acc19[P:P + self.n_img_max]
        ops.bce_logits_fwd_bwd(li[: 2 * B], self.x, li[: 2 * B], self.lam_i / b_global, acc_img[:2], seg_rows=B)
        if NI > 2:
            ops.bce_logits_fwd_bwd(li[2 * B:], self.x, li[2 * B:], 1.0 / b_global, acc_img[2:], seg_rows=B)
        acc_attr = self.acc19[P + self.n_img_max:]                      # one BCE sum per (decoder, live pass) segment
        nseg = 2 + self.approx_m
        self.ad_target.copy_(self.attrs_t[:, None, :].expand(N_ATTRS, nseg, B).reshape(N_ATTRS, nseg * B, 1))
        ops.bce_logits_fwd_bwd(self.ad_logit.view(-1, 4)[:, :1], self.ad_target.view(-1, 1), self.ad_dlogit.view(-1, 4)[:, :1],
                               1.0 / b_global, acc_attr, seg_rows=B)
        self.ad_dlogit[:, :B, :1].mul_(self.lam_t)             # the joint term uses the script's lambda_attrs
        # per-pass ELBO terms (tiny device-side bookkeeping)
        w_img = torch.ones(self.n_img_max, dtype=torch.float64, device=self.dev); w_img[:2] = self.lam_i
        terms = beta * kl_acc.clone()
        terms[:NI] += (w_img * acc_img)[:NI]
        terms.index_add_(0, plan["term_idx_dev"], acc_attr * plan["term_w_dev"])
        self._terms = (terms / b_global).to(torch.float32)
        self.loss19.copy_(self._terms.sum().reshape(1))

        # ================================================================ backward
        # ---- attribute decoders
        dy = [self.ad_dlogit[i][: nrows[i], :1] for i in range(N_ATTRS)]
        for i in range(N_ATTRS):
            ops.colsum_accumulate(dy[i], g[f"attr_decoders.{i}.net.6.bias"])
        layer_keys = ("0", "2", "4", "6")
        for l in (3, 2, 1, 0):
            key = layer_keys[l]
            n_out = 1 if l == 3 else 512
            K = L if l == 0 else 512
            descs = []
            for i in range(N_ATTRS):
                n = nrows[i]
                x_in = self.ad_z[i][:n] if l == 0 else self.ad_h[l - 1][i][:n]
                descs.append(D(dy[i], x_in, g[f"attr_decoders.{i}.net.{key}.weight"], n_out, K, n, a_mn=True, b_mn=True,
                               split_k=sp(n), accumulate=True))
                if l > 0:
                    dx = self.ad_dA[l % 2][i][:n]
                    descs.append(D(dy[i], p[f"attr_decoders.{i}.net.{key}.weight"], dx, n, K, n_out, b_mn=True,
                                   aux=self.ad_a[l - 1][i][:n], epilogue=DS,
                                   colsum=g[f"attr_decoders.{i}.net.{layer_keys[l - 1]}.bias"]))
                else:
                    descs.append(D(dy[i], p[f"attr_decoders.{i}.net.0.weight"], self.ad_dz[i][:n], n, K, n_out, b_mn=True))
            batched(descs)
            if l > 0:
                dy = [self.ad_dA[l % 2][i][: nrows[i]] for i in range(N_ATTRS)]
        self.dZ.index_add_(0, plan["rowidx_dev"], self.ad_dz.view(-1, L))   # padding rows of ad_dz are zero
        # ---- image decoder (live rows [0, NI*B))
        ops.im2col_k4(li, self.dcolsT4, R, 64, 64, 3, 2, 1)
        G([D(self.dcolsT4[: R * 1024], self.t3_h[: R * 1024], g[f"{d}.9.weight"], 48, 32, R * 1024, a_mn=True, b_mn=True,
             split_k=sp(R * 1024), accumulate=True),
           D(self.dcolsT4[: R * 1024], p[f"{d}.9.weight"], self.d_t3h[: R * 1024], R * 1024, 32, 48, b_mn=True)], Pp)
        self._bn_b(self.t3_x, self.d_t3h, self.d_t3x, P, B * 1024, 0, NI, f"{d}.7")
        dc3, w3 = self._cols(self.d_t3x, self.dcolsT3, R, 32, 32, 32, 2, 1)
        G([D(dc3[: R * (256 if w3 is None else 1024)], self.t2_h[: R * 256], g[f"{d}.6.weight"], 512, 64, R * 256, a_mn=True, b_mn=True,
             split_k=sp(R * 256), accumulate=True, a_view=w3),
           D(dc3[: R * (256 if w3 is None else 1024)], p[f"{d}.6.weight"], self.d_t2h[: R * 256], R * 256, 64, 512, b_mn=True,
             a_view=w3)], Pp)
        self._bn_b(self.t2_x, self.d_t2h, self.d_t2x, P, B * 256, 0, NI, f"{d}.4")
        dc2, w2 = self._cols(self.d_t2x, self.dcolsT2, R, 16, 16, 64, 2, 1)
        G([D(dc2[: R * (64 if w2 is None else 256)], self.t1_h[: R * 64], g[f"{d}.3.weight"], 1024, 128, R * 64, a_mn=True, b_mn=True,
             split_k=sp(R * 64), accumulate=True, a_view=w2),
           D(dc2[: R * (64 if w2 is None else 256)], p[f"{d}.3.weight"], self.d_t1h[: R * 64], R * 64, 128, 1024, b_mn=True,
             a_view=w2)], Pp)
        self._bn_b(self.t1_x, self.d_t1h, self.d_t1x, P, B * 64, 0, NI, f"{d}.1")
        dc1, w1 = self._cols(self.d_t1x, self.dcolsT1, R, 8, 8, 128, 1, 0)
        d_d0 = self.d_d0[:R]
        G([D(dc1[: R * (25 if w1 is None else 64)], self.d0_h.view(-1, 256)[: R * 25], g[f"{d}.0.weight"], 2048, 256, R * 25, a_mn=True,
             b_mn=True, split_k=sp(R * 25), accumulate=True, a_view=w1),
           D(dc1[: R * (25 if w1 is None else 64)], p[f"{d}.0.weight"], d_d0.view(R * 25, 256), R * 25, 256, 2048, b_mn=True,
             aux=self.d0_a.view(-1, 256)[: R * 25], epilogue=DS, a_view=w1)], Pp)
        ops.colsum_accumulate(d_d0, g["image_decoder.upsample.0.bias"])
        G([D(d_d0, Z[:R], g["image_decoder.upsample.0.weight"], 6400, L, R, a_mn=True, b_mn=True, split_k=sp(R), accumulate=True),
           D(d_d0, p["image_decoder.upsample.0.weight"], self.dZ[:R], R, L, 6400, b_mn=True, accumulate=True)], Pp)
        # ---- PoE / reparam / KL backward -> every expert
        d_enc_img = self.d_enc_img[: NI * B]
        dmu = [d_enc_img[k * B:(k + 1) * B, :L] for k in range(NI)] + [self.d_enc_a[i][:, :L] for i in range(N_ATTRS)]
        dlv = [d_enc_img[k * B:(k + 1) * B, L:] for k in range(NI)] + [self.d_enc_a[i][:, L:] for i in range(N_ATTRS)]
        ops.poe_bwd(mu_e, lv_e, plan["masks"], B, L, self.dZ[: P * B], dmu, dlv, kl_scale=beta / b_global, variant=1,
                    training=training, noise=self.noise[: P * B] if training else None)
        # ---- attribute encoders
        for i in range(N_ATTRS):
            ops.colsum_accumulate(self.d_enc_a[i], g[f"attr_encoders.{i}.net.4.bias"])
        descs = []
        for i in range(N_ATTRS):
            descs.append(D(self.d_enc_a[i], self.ae_h2[i], g[f"attr_encoders.{i}.net.4.weight"], 2 * L, 512, B, a_mn=True,
                           b_mn=True, split_k=sp(B), accumulate=True))
            descs.append(D(self.d_enc_a[i], p[f"attr_encoders.{i}.net.4.weight"], self.ae_dA[i], B, 512, 2 * L, b_mn=True,
                           aux=self.ae_a2[i], epilogue=DS, colsum=g[f"attr_encoders.{i}.net.2.bias"]))
        batched(descs)
        descs = []
        for i in range(N_ATTRS):
            descs.append(D(self.ae_dA[i], self.ae_h1[i], g[f"attr_encoders.{i}.net.2.weight"], 512, 512, B, a_mn=True,
                           b_mn=True, split_k=sp(B), accumulate=True))
            descs.append(D(self.ae_dA[i], p[f"attr_encoders.{i}.net.2.weight"], self.ae_dh1[i], B, 512, 512, b_mn=True))
        batched(descs)
        for i in range(N_ATTRS):
            ops.embedding_swish_bwd(p[f"attr_encoders.{i}.net.0.weight"], self.attrs_idx[i], self.ae_dh1[i],
                                    g[f"attr_encoders.{i}.net.0.weight"])
        # ---- image encoder
        ops.colsum_accumulate(d_enc_img, g["image_encoder.classifier.3.bias"])
        d_fcd = self.d_fcd[: NI * B]
        G([D(d_enc_img, fcd, g["image_encoder.classifier.3.weight"], 2 * L, 512, NI * B, a_mn=True, b_mn=True,
             split_k=sp(NI * B), accumulate=True),
           D(d_enc_img, p["image_encoder.classifier.3.weight"], d_fcd, NI * B, 512, 2 * L, b_mn=True)], Pp)
        if training:
            ops.dropout_bwd(d_fcd, self.drop_mask[: NI * B], self.d_fch, NI, 0.1)
        else:
            self.d_fch.copy_(d_fcd.view(NI, B, 512).sum(0))
        ops.swish_bwd(self.fc_a, self.d_fch, self.d_fca)
        ops.colsum_accumulate(self.d_fca, g["image_encoder.classifier.0.bias"])
        G([D(self.d_fca, self.c4_h.view(B, 6400), g["image_encoder.classifier.0.weight"], 512, 6400, B, a_mn=True, b_mn=True,
             split_k=sp(B), accumulate=True),
           D(self.d_fca, p["image_encoder.classifier.0.weight"], self.d_c4h.view(B, 6400), B, 6400, 512, b_mn=True)], Pp)
        self._bn_b(self.c4_x, self.d_c4h, self.d_c4x, 1, B * 25, 0, 1, f"{e}.9")
        imp = self.implicit_conv
        V = ops.conv_view
        G([D(self.d_c4x, self.c3_h if imp else self.cols4, g[f"{e}.8.weight"], 256, 2048, B * 25, a_mn=True, b_mn=True,
             split_k=sp(B * 25), accumulate=True, b_view=V(B, 8, 8, 128, stride=1, pad=0) if imp else None),
           D(self.d_c4x, p[f"{e}.8.weight"], self.dcols4, B * 25, 2048, 256, b_mn=True)], Pp)
        ops.col2im_k4(self.dcols4, self.d_c3h, B, 5, 5, 128, 1, 0)
        self._bn_b(self.c3_x, self.d_c3h, self.d_c3x, 1, B * 64, 0, 1, f"{e}.6")
        G([D(self.d_c3x, self.c2_h if imp else self.cols3, g[f"{e}.5.weight"], 128, 1024, B * 64, a_mn=True, b_mn=True,
             split_k=sp(B * 64), accumulate=True, b_view=V(B, 16, 16, 64) if imp else None),
           ] + ([] if self.subpixel else [D(self.d_c3x, p[f"{e}.5.weight"], self.dcols3, B * 64, 1024, 128, b_mn=True)]), Pp)
        if self.subpixel:
            GC(ops.subpixel_k4s2p1(self.d_c3x, p[f"{e}.5.weight"], self.d_c2h, B, 8, 8, 128, 64, w_is_conv=True))
        else:
            ops.col2im_k4(self.dcols3, self.d_c2h, B, 8, 8, 64, 2, 1)
        self._bn_b(self.c2_x, self.d_c2h, self.d_c2x, 1, B * 256, 0, 1, f"{e}.3")
        G([D(self.d_c2x, self.c1_h if imp else self.cols2, g[f"{e}.2.weight"], 64, 512, B * 256, a_mn=True, b_mn=True,
             split_k=sp(B * 256), accumulate=True, b_view=V(B, 32, 32, 32) if imp else None),
           ] + ([] if self.subpixel else [D(self.d_c2x, p[f"{e}.2.weight"], self.dcols2, B * 256, 512, 64, b_mn=True)]), Pp)
        if self.subpixel:
            GC(ops.subpixel_k4s2p1(self.d_c2x, p[f"{e}.2.weight"], self.d_c1a, B, 16, 16, 64, 32, w_is_conv=True, aux=self.c1_a,
                                   epilogue=ops.EPI_MUL_DSWISH))
        else:
            ops.col2im_k4(self.dcols2, self.d_c1a, B, 16, 16, 32, 2, 1, aux=self.c1_a)
        G([D(self.d_c1a, self.cols1, g[f"{e}.0.weight"], 32, 48, B * 1024, a_mn=True, b_mn=True, split_k=sp(B * 1024),
             accumulate=True)], Pp)
        # ================================================================ exchange + update
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grad_bucket, group=self.pg)
            dist.all_reduce(self.loss19, group=self.pg)
        if update:
            self._enqueue_update()
