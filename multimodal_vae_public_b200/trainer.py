"""Fused MVAE training step for the MNIST-flavour (MLP) model on one B200 (one process per GPU).

This is the hot path the benchmark measures: the whole body of the reference's training loop
(mnist/train.py:196-219: zero_grad -> model(image,text), model(image), model(text=text) -> three
elbo_loss terms -> sum -> backward -> Adam.step) as ~35 kernel launches of libmvae_b200.so over
pre-allocated HBM buffers, captured once in a CUDA graph and replayed per step.

Work the reference does that cannot change the result is not executed:
  * the image/text encoders run once and feed both the joint and the uni-modal pass
    (no BatchNorm/Dropout in this flavour, so the duplicate evaluations are bit-identical);
  * the decoder whose output the loss ignores (text decoder in the image-only pass, image decoder in
    the text-only pass) is skipped: its output has zero gradient and is not part of the objective.
The three passes are stacked along the batch so each decoder layer is ONE GEMM over 2B rows.

Layout in HBM (fp32, row-major; B = per-rank batch, L = n_latents):
  flat arenas  params / grads / adam_m / adam_v   one contiguous bucket each (single NCCL all-reduce
               and single fused Adam over the bucket); the nn.Parameter views of the drop-in module
               alias `params`, their .grad alias `grads`.
  Z [3B, L]    rows [0,B) image-only pass, [B,2B) joint pass, [2B,3B) text-only pass, so the image
               decoder reads rows [0,2B) and the text decoder rows [B,3B) with no gather/concat.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._lib import PREC_3XTF32, PREC_TF32

# internal pass order (rows of Z) and the reference's order (joint, image, text) [mnist/train.py:200-202]
_PASS_MASKS = (0b01, 0b11, 0b10)        # expert 0 = image encoder, expert 1 = text encoder
_REF_TO_INTERNAL = (1, 0, 2)            # reference pass i lives at internal index _REF_TO_INTERNAL[i]


def mnist_layout(n_latents: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Arena order of the MNIST MVAE parameters (names = reference state_dict keys,
    mnist/model.py:75-78,95-98,116-119,136-139).  fc31/fc32 are adjacent so that the two encoder heads
    run as one N = 2L GEMM."""
    L = n_latents
    out = []
    for enc, first in (("image_encoder", ("fc1", (512, 784))), ("text_encoder", ("fc1", (10, 512)))):
        out.append((f"{enc}.{first[0]}.weight", first[1]))
        if enc == "image_encoder":
            out.append((f"{enc}.fc1.bias", (512,)))
        out += [(f"{enc}.fc2.weight", (512, 512)), (f"{enc}.fc2.bias", (512,)),
                (f"{enc}.fc31.weight", (L, 512)), (f"{enc}.fc32.weight", (L, 512)),
                (f"{enc}.fc31.bias", (L,)), (f"{enc}.fc32.bias", (L,))]
    for dec, n_out in (("image_decoder", 784), ("text_decoder", 10)):
        out += [(f"{dec}.fc1.weight", (512, L)), (f"{dec}.fc1.bias", (512,)),
                (f"{dec}.fc2.weight", (512, 512)), (f"{dec}.fc2.bias", (512,)),
                (f"{dec}.fc3.weight", (512, 512)), (f"{dec}.fc3.bias", (512,)),
                (f"{dec}.fc4.weight", (n_out, 512)), (f"{dec}.fc4.bias", (n_out,))]
    return out


def mnist_reference_order(n_latents: int) -> List[str]:
    """state_dict key order of the reference MVAE (registration order of mnist/model.py:20-27,75-78,...)."""
    keys = []
    for mod, layers in (("image_encoder", ("fc1", "fc2", "fc31", "fc32")), ("image_decoder", ("fc1", "fc2", "fc3", "fc4")),
                        ("text_encoder", ("fc1", "fc2", "fc31", "fc32")), ("text_decoder", ("fc1", "fc2", "fc3", "fc4"))):
        for l in layers:
            keys.append(f"{mod}.{l}.weight")
            if not (mod == "text_encoder" and l == "fc1"):
                keys.append(f"{mod}.{l}.bias")
    return keys


def adam_state_to_torch(names: Sequence[str], exp_avg: Dict[str, torch.Tensor], exp_avg_sq: Dict[str, torch.Tensor],
                        step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8) -> dict:
    """``torch.optim.Adam.state_dict()`` layout (what the reference stores under checkpoint['optimizer'],
    mnist/train.py:263-268) from named first / second moments; ``names`` in ``model.parameters()`` order."""
    state = {}
    if step > 0:
        for i, k in enumerate(names):
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": exp_avg[k].detach().clone().cpu(),
                        "exp_avg_sq": exp_avg_sq[k].detach().clone().cpu()}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
             "params": list(range(len(names)))}
    return {"state": state, "param_groups": [group]}


def adam_state_from_torch(sd: dict, names: Sequence[str]):
    """Inverse of ``adam_state_to_torch``: (exp_avg by name, exp_avg_sq by name, step, lr).  Accepts state dicts written
    by any torch version of the reference's era (``step`` as int or 0-dim tensor); parameters without state (never
    stepped) come back as None."""
    groups = sd["param_groups"]
    order = [i for g in groups for i in g["params"]]
    if len(order) != len(names):
        raise ValueError(f"optimizer state has {len(order)} parameters, the model has {len(names)}")
    m, v, step = {}, {}, 0
    for pos, k in zip(order, names):
        st = sd["state"].get(pos)
        if st is None:
            m[k] = v[k] = None
            continue
        m[k], v[k] = st["exp_avg"], st["exp_avg_sq"]
        step = max(step, int(st["step"].item() if torch.is_tensor(st["step"]) else st["step"]))
    return m, v, step, float(groups[0]["lr"])


class FlatArena:
    """One contiguous fp32 bucket with named, 16-byte aligned views."""

    def __init__(self, layout: Sequence[Tuple[str, Tuple[int, ...]]], device, n_buffers: int = 1, tail: int = 0,
                 alloc=None):
        self.offsets: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape in layout:
            n = int(math.prod(shape))
            self.offsets[name] = (off, tuple(shape))
            off += (n + 3) // 4 * 4
        self.numel = off
        # `tail` extra floats after the parameters (the gradient bucket carries the loss scalars there so that ONE
        # all-reduce moves gradients and loss)
        # `alloc(i, numel)` may supply buffer i (e.g. from NVLink-symmetric memory) -- it must come back zeroed
        self.buffers = []
        for i in range(n_buffers):
            t = alloc(i, off + tail) if alloc is not None else None
            self.buffers.append(t if t is not None else torch.zeros(off + tail, dtype=torch.float32, device=device))

    def view(self, buf: int, name: str) -> torch.Tensor:
        off, shape = self.offsets[name]
        return self.buffers[buf][off: off + int(math.prod(shape))].view(shape)

    def span(self, buf: int, first: str, last: str) -> torch.Tensor:
        """Contiguous 1-D slice covering parameters first..last (inclusive) -- they must be adjacent."""
        o0, _ = self.offsets[first]
        o1, s1 = self.offsets[last]
        return self.buffers[buf][o0: o1 + int(math.prod(s1))]


class MnistMVAETrainer:
    """Whole-step trainer (MNIST flavour).  ``step(image, text)`` == one iteration of the reference loop."""

    def __init__(self, n_latents: int = 64, batch_size: int = 4096, device="cuda", lr: float = 1e-3,
                 lambda_image: float = 1.0, lambda_text: float = 10.0, precision: int = PREC_3XTF32,
                 world_size: int = 1, seed: int = 0, rank: int = 0, use_graph: bool = True,
                 process_group=None, chain: Optional[bool] = None, dp_mode: Optional[str] = None,
                 label_table: Optional[bool] = None):
        _lib.load()  # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.MvaeError("MnistMVAETrainer needs a CUDA device (no CPU fallback)")
        self.L, self.B = n_latents, batch_size
        if n_latents % 4:
            raise _lib.MvaeError("n_latents must be a multiple of 4")
        self.dev = torch.device(device)
        self.lr, self.lam_i, self.lam_t = lr, lambda_image, lambda_text
        self.prec = precision
        self.world, self.rank, self.seed = world_size, rank, seed
        self.pg = process_group
        self.use_graph = use_graph
        # chain mode: each Linear stack (and its autograd chain) is ONE persistent launch whose tiles wait on row-block
        # completion counters instead of launch boundaries (mvae_gemm_chain); MVAE_CHAIN=0 restores one launch per layer
        self.chain = os.environ.get("MVAE_CHAIN", "1") != "0" if chain is None else bool(chain)
        self.chain_ws = ops.chain_workspace(torch.device(device))
        # label-table mode: the label encoder only ever sees 10 distinct inputs, so it is evaluated once per CLASS
        # (csrc/label_table.cu) and the PoE kernels gather row text[b]; MVAE_LABEL_TABLE=0 restores the per-sample GEMMs
        self.label_table = os.environ.get("MVAE_LABEL_TABLE", "1") != "0" if label_table is None else bool(label_table)
        # fused split-K (last-arriver epilogue) for problems with fewer output tiles than SMs; MVAE_FUSED_SPLIT=0 disables
        self.fused_split = os.environ.get("MVAE_FUSED_SPLIT", "1") != "0"
        self._split_ws: Dict[str, torch.Tensor] = {}
        self._sms = torch.cuda.get_device_properties(torch.device(device)).multi_processor_count
        self.layout = self._make_layout(n_latents)
        # Data-parallel exchange: "p2p" = ONE fused kernel per rank over NVLink peer memory (gradient reduce-scatter ->
        # Adam on the rank's slice -> parameter all-gather, csrc/dp_p2p.cu); "nccl" = ncclAllReduce + flat Adam.
        self.dp_mode = "none"
        if world_size > 1:
            # default "auto": the fused peer-memory kernel when every rank has NVLink peer access to every other rank (agreed
            # collectively), else NCCL.  Validated against NCCL and the single-process trajectory at 2, 4 and 8 GPUs
            # (tests/test_dp_p2p_gpu.py, profiles/r02_multi_gpu_*.txt); MVAE_DP=nccl / dp_mode="nccl" forces the collective.
            self.dp_mode = os.environ.get("MVAE_DP", "auto") if dp_mode is None else dp_mode
            if self.dp_mode not in ("nccl", "p2p", "auto"):
                raise _lib.MvaeError(f"dp_mode must be 'nccl', 'p2p' or 'auto', got {self.dp_mode!r}")
            if self.dp_mode in ("p2p", "auto"):
                # every rank must take the same path (a rank that fell back to NCCL on its own would deadlock both the
                # flag rendezvous and the collective): agree on the minimum over ranks
                self.dp_mode = "p2p" if self._p2p_agreed() else "nccl"
        self._symm = {}
        self.arena = FlatArena(self.layout, self.dev, n_buffers=4, tail=4,   # params, grads(+loss tail), adam m, adam v
                               alloc=self._symm_alloc if self.dp_mode == "p2p" else None)
        self.params = {k: self.arena.view(0, k) for k, _ in self.layout}
        self.grads = {k: self.arena.view(1, k) for k, _ in self.layout}
        n = self.arena.numel
        self.flat_params, self.adam_m, self.adam_v = (self.arena.buffers[i][:n] for i in (0, 2, 3))
        # pre-split weights (3xTF32): low halves of the parameters in a twin buffer, refreshed once per step and fetched by
        # TMA as the GEMMs' B_lo operand (ops.register_lo_arena) instead of being computed by the splitter warps per tile and
        # k-block.  Bit-identical results.  For 128-wide tiles it measured SLOWER on the same box (MNIST 0.727 vs 0.713
        # ms/step, FashionMNIST 6.70 vs 6.49: the extra 16 KiB of L2 -> shared-memory traffic per k-block costs more than
        # the ~300 splitter cycles it removes, profiles/r02_ab_mnist_fashion.txt); for narrow tiles (N <= 64: the sub-pixel
        # transposed convolutions of the conv flavours) the splitters pace the loop and B_lo is only 8 KiB.
        # MVAE_PRESPLIT = 0 (off) | 1 (every weight operand) | narrow (problems with N <= 64 only); flavour default below.
        mode = os.environ.get("MVAE_PRESPLIT", self._presplit_default)
        self.presplit = precision == PREC_3XTF32 and mode != "0"
        if self.presplit:
            self.params_lo = torch.zeros(n, dtype=torch.float32, device=self.dev)
            ops.register_lo_arena(self.flat_params, self.params_lo, max_n=64 if mode == "narrow" else 0)
        self.grad_bucket = self.arena.buffers[1]          # gradients + 4 loss floats: the all-reduce payload
        self.flat_grads = self.grad_bucket[:n]
        B, L, dev = batch_size, n_latents, self.dev
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
        # inputs (device-resident copies; step() fills them from host or device tensors)
        self.x = f(B, 784)
        self.text = torch.zeros(B, dtype=torch.int64, device=dev)
        self.noise = f(3 * B, L)           # internal pass order
        self.enc_i, self.enc_t = f(B, 2 * L), f(B, 2 * L)
        self.d_enc_i, self.d_enc_t = f(B, 2 * L), f(B, 2 * L)
        self.Z = f(3 * B, L)
        self.logit_i = f(2 * B, 784)
        self.logit_t_buf = f(2 * B, 16)    # N = 10 padded to ld 16 for TMA
        self.logit_t = self.logit_t_buf[:, :10]
        self._alloc_activations(f)
        # ONE zero-initialised region per step (a single memset): dZ, the class-wise label-encoder gradient table and the
        # loss accumulators
        V = self._n_classes()
        n_dz, n_tab = 3 * B * L, V * 2 * L
        n_da2 = V * 512
        self.zero_region = torch.zeros(n_dz + n_tab + n_da2 + 32, dtype=torch.float32, device=dev)
        self.dZ = self.zero_region[:n_dz].view(3 * B, L)
        self.d_tab = self.zero_region[n_dz:n_dz + n_tab].view(V, 2 * L)
        self.tt_dA2 = self.zero_region[n_dz + n_tab:n_dz + n_tab + n_da2].view(V, 512)   # accumulation target of the table bwd
        o = n_dz + n_tab + n_da2
        self.acc = self.zero_region[o:o + 18].view(torch.float64)  # recon_img[3], recon_txt[3], kl[3]
        # label table: pre-activation / activation of the hidden layer and the (mu | logvar) rows, one per class
        self.tt_a2, self.tt_h2 = f(V, 512), f(V, 512)
        self.enc_tab = f(V, 2 * L)
        self.loss_tail = self.grad_bucket[n:n + 4]          # total, internal passes 0..2 (tail of the gradient bucket)
        self.loss_out = self.loss_tail                      # what is copied to the host (the sums over ranks)
        if self.dp_mode == "p2p":
            self._p2p_finish_setup(n, 4)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.beta_dev = torch.ones(1, dtype=torch.float32, device=dev)   # KL annealing factor
        # pinned staging ring for the annealing factor: an asynchronous H2D copy reads the pinned word when the COPY runs,
        # so one buffer reused by back-to-back sync=False steps could hand a step the next step's KL weight
        self._beta_ring = [(torch.ones(1, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
        self._beta_idx = 0
        self.loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self._graphs: Dict[Tuple[bool, bool], object] = {}
        self._stream = torch.cuda.Stream(device=dev)
        # independent small kernels (the label encoder's class table, the label-side loss) run on a side stream next to
        # the GEMM chains -- inside the captured graph they become parallel branches; MVAE_OVERLAP=0 serialises them
        self._side_stream = torch.cuda.Stream(device=dev)
        self.overlap = os.environ.get("MVAE_OVERLAP", "1") != "0"
        self.launches_per_step = 0
        self.init_parameters(seed)

    _presplit_default = "0"

    def __del__(self):
        try:
            if getattr(self, "presplit", False):
                ops.unregister_lo_arena(self.flat_params)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    # ------------------------------------------------------------------ flavour hooks
    def _make_layout(self, n_latents: int):
        return mnist_layout(n_latents)

    def _n_classes(self) -> int:
        return 10

    def _label_encoder(self, buf: int):
        """(embedding, w2, b2, heads w [2L,512], heads b [2L]) of the label encoder in arena buffer `buf` (0 = parameters,
        1 = gradients)."""
        L, a = self.L, self.arena
        return (a.view(buf, "text_encoder.fc1.weight"), a.view(buf, "text_encoder.fc2.weight"),
                a.view(buf, "text_encoder.fc2.bias"),
                a.span(buf, "text_encoder.fc31.weight", "text_encoder.fc32.weight").view(2 * L, 512),
                a.span(buf, "text_encoder.fc31.bias", "text_encoder.fc32.bias"))

    def _label_experts(self):
        """Expert tensors of the PoE kernels: (mu_e, lv_e, dmu_e, dlv_e, gather) with the label encoder either as a
        per-sample [B, 2L] matrix or as a [V, 2L] table indexed by the labels."""
        L = self.L
        if self.label_table:
            et, dt, gather = self.enc_tab, self.d_tab, [None, self.text]
        else:
            et, dt, gather = self.enc_t, self.d_enc_t, None
        return ([self.enc_i[:, :L], et[:, :L]], [self.enc_i[:, L:], et[:, L:]],
                [self.d_enc_i[:, :L], dt[:, :L]], [self.d_enc_i[:, L:], dt[:, L:]], gather)

    def _alloc_activations(self, f) -> None:
        B = self.B
        # encoders
        self.ie_a1, self.ie_h1, self.ie_a2, self.ie_h2 = f(B, 512), f(B, 512), f(B, 512), f(B, 512)
        self.te_h1, self.te_a2, self.te_h2 = f(B, 512), f(B, 512), f(B, 512)
        # decoders (2B rows each)
        self.id_a = [f(2 * B, 512) for _ in range(3)]; self.id_h = [f(2 * B, 512) for _ in range(3)]
        self.td_a = [f(2 * B, 512) for _ in range(3)]; self.td_h = [f(2 * B, 512) for _ in range(3)]
        # backward scratch
        # one dA buffer per hidden layer: inside a chained launch no buffer may be rewritten while another problem of the
        # same launch (the wgrad of the layer above) still reads it
        self.id_dA = [f(2 * B, 512) for _ in range(3)]; self.td_dA = [f(2 * B, 512) for _ in range(3)]
        self.ie_dA = [f(B, 512) for _ in range(2)]; self.te_dA = [f(B, 512) for _ in range(2)]

    # ------------------------------------------------------------------ parallel branches (side stream)
    def _fork(self):
        """Context manager: the enclosed launches go to the side stream, ordered after everything enqueued so far on the
        step stream; ``_join()`` orders the step stream after them.  With MVAE_OVERLAP=0 a no-op (same stream)."""
        import contextlib
        if not self.overlap:
            return contextlib.nullcontext()
        self._side_stream.wait_stream(torch.cuda.current_stream())
        self._forked = True
        return torch.cuda.stream(self._side_stream)

    def _join(self) -> None:
        # (only after a fork: waiting on a side stream that holds no work of THIS capture is a capture-isolation error)
        if self.overlap and getattr(self, "_forked", False):
            torch.cuda.current_stream().wait_stream(self._side_stream)
            self._forked = False

    def _side_mark(self):
        """Inside a ``_fork()`` block: an event after the side-stream work enqueued so far, for ``_wait_mark`` -- the step
        stream can then wait for THAT work only while later side-stream launches (bias-gradient column sums, which nothing
        but the optimizer reads) keep running beside the next GEMM chain until the final ``_join()``."""
        if not self.overlap:
            return None
        ev = torch.cuda.Event()
        ev.record(self._side_stream)
        return ev

    def _wait_mark(self, ev) -> None:
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    # ------------------------------------------------------------------ GEMM problem helpers
    def _D(self, key: str, A, Bm, Cm, M: int, N: int, K: int, share: int = 1, **kw):
        """``ops.gemm_desc`` for a forward / dgrad problem, with FUSED SPLIT-K chosen automatically when the problem has
        fewer output tiles than its share of the SMs (small per-GPU batches, the K = 6272 classifier layers): the k range
        of every tile is split over several CTAs, the last arriver applies the epilogue (include/mvae_b200.h, split_ws).
        ``share`` = number of sibling problems of the same launch that want SMs at the same time; ``key`` names the
        problem's scratch matrix."""
        if self.fused_split and N % 4 == 0 and not kw.get("accumulate", False) and kw.get("split_k", 1) == 1:
            bn = 128 if N >= 128 else (N + 31) // 32 * 32 if kw.get("b_mn") else (N + 15) // 16 * 16
            tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
            kb = (K + 31) // 32
            budget = self._sms       # (siblings with short reductions drain quickly; a long one owns the machine's tail)
            # measured (profiles/r02_fused_split_ab.txt): worth it for LONG reductions only -- the K = 6272 classifier layers
            # at 512 rows/GPU gain 5 % of the FashionMNIST step, while splitting the K = 512 / 784 MNIST layers costs 6 %
            # (pipeline fill + the scratch round trip outweigh the shorter k loop): keep >= 16 k-blocks per split
            if 2 * tiles <= budget and kb >= 48:
                split = min(budget // tiles, kb // 16, 16)
                if split > 1:
                    ws = self._split_ws.get(key)
                    if ws is None or ws.numel() < M * ((N + 3) // 4 * 4):
                        ws = self._split_ws[key] = torch.zeros(M * ((N + 3) // 4 * 4), dtype=torch.float32, device=self.dev)
                    kw.update(split_k=split, split_ws=ws)
        return ops.gemm_desc(A, Bm, Cm, M, N, K, **kw)

    def _gemm(self, descs) -> None:
        """Independent problems in one launch (a chained launch without dependencies when a problem uses fused split-K,
        which needs the counter workspace)."""
        if any(d.split_ws for d in descs):
            ops.gemm_chain(descs, [-1] * len(descs), self.chain_ws, self.prec)
        else:
            ops.gemm_batch(descs, self.prec)

    # ------------------------------------------------------------------ parameters
    def init_parameters(self, seed: int = 0) -> None:
        """PyTorch-default initialisation scales (nn.Linear: U(+-1/sqrt(fan_in)); nn.Embedding: N(0,1))."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name, shape in self.layout:
            if name == "text_encoder.fc1.weight":
                v = torch.randn(shape, generator=g)
            else:
                wname = name.replace(".bias", ".weight")
                fan_in = dict(self.layout)[wname][1]
                bound = 1.0 / math.sqrt(fan_in)
                v = (torch.rand(shape, generator=g) * 2 - 1) * bound
            self.params[name].copy_(v)
        self.adam_m.zero_(); self.adam_v.zero_(); self.step_count.zero_()

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        missing = [k for k, _ in self.layout if k not in sd]
        if missing:
            raise KeyError(f"missing keys: {missing}")
        for k, _ in self.layout:
            self.params[k].copy_(sd[k].to(torch.float32))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Copies of the parameters under the reference's state_dict keys, in the reference's order
        (mnist/model.py:20-27), so ``ref_model.load_state_dict(trainer.state_dict())`` works."""
        return {k: self.params[k].detach().clone() for k in mnist_reference_order(self.L)}

    # ------------------------------------------------------------------ optimizer state (checkpoint['optimizer'])
    def _with_arena_buffer(self, buf: int, fn):
        """Run a state_dict-style method with ``self.params`` pointing at another arena buffer (Adam moments), so the
        flavour's layout permutations (conv weights, FC column orders) are applied to the moments exactly as to the
        parameters."""
        saved = self.params
        self.params = {k: self.arena.view(buf, k) for k, _ in self.layout}
        try:
            return fn()
        finally:
            self.params = saved

    def optimizer_state_dict(self) -> dict:
        """Adam state in ``torch.optim.Adam.state_dict()`` format, tensors in the reference's parameter layouts and
        ``model.parameters()`` order: drop it into the reference's checkpoint dict (mnist/train.py:263-268)."""
        names = [k for k in self.state_dict() if k in self.arena.offsets]
        m = self._with_arena_buffer(2, self.state_dict)
        v = self._with_arena_buffer(3, self.state_dict)
        self._stream.synchronize()
        return adam_state_to_torch(names, m, v, int(self.step_count.item()), self.lr)

    def load_optimizer_state_dict(self, sd: dict) -> None:
        """Resume from a reference checkpoint's optimizer state (moments, step count, learning rate)."""
        names = [k for k in self.state_dict() if k in self.arena.offsets]
        m, v, step, lr = adam_state_from_torch(sd, names)
        shapes = {k: t.shape for k, t in self.state_dict().items()}
        zeros = lambda k: torch.zeros(shapes[k], dtype=torch.float32)  # noqa: E731
        m = {k: (t if t is not None else zeros(k)).to(self.dev) for k, t in m.items()}
        v = {k: (t if t is not None else zeros(k)).to(self.dev) for k, t in v.items()}
        full = self.state_dict()          # buffers (BatchNorm running statistics) pass through unchanged
        self._with_arena_buffer(2, lambda: self.load_state_dict({**full, **m}))
        self._with_arena_buffer(3, lambda: self.load_state_dict({**full, **v}))
        self.step_count.fill_(step)
        self.lr = lr
        self._graphs.clear()              # the learning rate is baked into captured launches

    # ------------------------------------------------------------------ one step worth of launches
    def _p(self, name):
        return self.params[name]

    def _g(self, name):
        return self.grads[name]

    def _enqueue_forward(self, training: bool, use_noise_input: bool) -> None:
        B, L, P = self.B, self.L, self.prec
        p = self.params
        wi = self.arena.span(0, "image_encoder.fc31.weight", "image_encoder.fc32.weight").view(2 * L, 512)
        bi = self.arena.span(0, "image_encoder.fc31.bias", "image_encoder.fc32.bias")
        wt = self.arena.span(0, "text_encoder.fc31.weight", "text_encoder.fc32.weight").view(2 * L, 512)
        bt = self.arena.span(0, "text_encoder.fc31.bias", "text_encoder.fc32.bias")
        D = ops.gemm_desc
        ch = self.chain      # (fused split-K lives in chained launches)
        sh = 1 if self.label_table else 2
        S = (lambda key, *a, **k: self._D(key, *a, **k)) if ch else (lambda key, *a, share=1, **k: D(*a, **k))
        fc1_i = S("fc1_i", self.x, p["image_encoder.fc1.weight"], self.ie_a1, B, 512, 784, bias=p["image_encoder.fc1.bias"],
                  out2=self.ie_h1, epilogue=ops.EPI_BIAS_SWISH)
        fc2_i = S("fc2_i", self.ie_h1, p["image_encoder.fc2.weight"], self.ie_a2, B, 512, 512, share=sh,
                  bias=p["image_encoder.fc2.bias"], out2=self.ie_h2, epilogue=ops.EPI_BIAS_SWISH)
        fc2_t = S("fc2_t", self.te_h1, p["text_encoder.fc2.weight"], self.te_a2, B, 512, 512, share=2,
                  bias=p["text_encoder.fc2.bias"], out2=self.te_h2, epilogue=ops.EPI_BIAS_SWISH)
        heads_i = S("heads_i", self.ie_h2, wi, self.enc_i, B, 2 * L, 512, share=sh, bias=bi)
        heads_t = S("heads_t", self.te_h2, wt, self.enc_t, B, 2 * L, 512, share=2, bias=bt)
        if self.label_table:
            emb, w2, b2, w3, b3 = self._label_encoder(0)
            with self._fork():      # (the class table needs 10 rows of work: next to the image encoder chain, not before it)
                ops.label_table_fwd(emb, w2, b2, w3, b3, self.tt_a2, self.tt_h2, self.enc_tab)
            if self.chain:
                ops.gemm_chain([fc1_i, fc2_i, heads_i], [-1, 0, 1], self.chain_ws, P)
            else:
                ops.gemm_batch([fc1_i], P); ops.gemm_batch([fc2_i], P); ops.gemm_batch([heads_i], P)
            self._join()
        else:
            ops.embedding_swish_fwd(p["text_encoder.fc1.weight"], self.text, None, self.te_h1)
        if self.label_table:
            pass
        elif self.chain:
            # both encoders in ONE launch: the two independent first layers fill the machine together, then the
            # dependent layers follow tile by tile
            ops.gemm_chain([fc1_i, fc2_t, fc2_i, heads_t, heads_i], [-1, -1, 0, 1, 2], self.chain_ws, P)
        else:
            ops.gemm_batch([fc1_i], P)
            ops.gemm_batch([fc2_i, fc2_t], P)
            ops.gemm_batch([heads_i, heads_t], P)
        # PoE + reparametrise + KL for the three passes
        mu_e, lv_e, _, _, gather = self._label_experts()
        ops.poe_fwd(mu_e, lv_e, _PASS_MASKS, B, L, self.Z, variant=0, training=training,
                    noise=self.noise if (training and use_noise_input) else None,
                    noise_out=self.noise if (training and not use_noise_input) else None,
                    seed=self.seed * 1000003 + self.rank, offset=0, step_dev=self.step_count,
                    kl_acc=self.acc[6:9], gather=gather)
        # decoders: image decoder on rows [0,2B) (image-only, joint), text decoder on rows [B,3B) (joint, text-only)
        zi, zt = self.Z[: 2 * B], self.Z[B:]
        xin_i, xin_t = zi, zt
        layers = []
        for l in range(3):
            K = L if l == 0 else 512
            layers.append([
                S(f"id{l}", xin_i, p[f"image_decoder.fc{l + 1}.weight"], self.id_a[l], 2 * B, 512, K, share=2,
                  bias=p[f"image_decoder.fc{l + 1}.bias"], out2=self.id_h[l], epilogue=ops.EPI_BIAS_SWISH),
                S(f"td{l}", xin_t, p[f"text_decoder.fc{l + 1}.weight"], self.td_a[l], 2 * B, 512, K, share=2,
                  bias=p[f"text_decoder.fc{l + 1}.bias"], out2=self.td_h[l], epilogue=ops.EPI_BIAS_SWISH)])
            xin_i, xin_t = self.id_h[l], self.td_h[l]
        layers.append([
            S("id3", xin_i, p["image_decoder.fc4.weight"], self.logit_i, 2 * B, 784, 512, bias=p["image_decoder.fc4.bias"]),
            D(xin_t, p["text_decoder.fc4.weight"], self.logit_t, 2 * B, 10, 512, bias=p["text_decoder.fc4.bias"])])
        if self.chain:   # both decoders, all four layers: one launch (problem 2l+d reads the output of problem 2(l-1)+d)
            ops.gemm_chain([d for pair in layers for d in pair], [-1, -1, 0, 1, 2, 3, 4, 5], self.chain_ws, P)
        else:
            for pair in layers:
                ops.gemm_batch(pair, P)

    def _enqueue_loss_and_backward(self, training: bool, b_global: int) -> None:
        B, L, P = self.B, self.L, self.prec
        p, g = self.params, self.grads
        # reconstruction losses + dlogits (in place)
        dyi, dyt = self.logit_i, self.logit_t
        with self._fork():          # label term next to the image term
            ops.ce_fwd_bwd(self.logit_t, self.text, self.logit_t, 10, self.lam_t / b_global, self.acc[4:6], seg_rows=B)
            ce_done = self._side_mark()
            ops.colsum_accumulate(dyt, g["text_decoder.fc4.bias"])
        ops.bce_logits_fwd_bwd(self.logit_i, self.x, self.logit_i, self.lam_i / b_global, self.acc[0:3], seg_rows=B)
        with self._fork():          # the last layers' bias gradients: beside the decoder backward chain, joined at the end
            ops.colsum_accumulate(dyi, g["image_decoder.fc4.bias"])
        self._wait_mark(ce_done)
        # ---- decoders backward, the two decoders batched per layer
        nk = max(1, (2 * B) // 32)  # k-blocks of a decoder wgrad
        D = ops.gemm_desc
        S = (lambda key, *a, **k: self._D(key, *a, **k)) if self.chain else (lambda key, *a, share=1, **k: D(*a, **k))
        split = max(1, min(nk // 16, 32))   # ~16 k-blocks per wgrad tile, like the dgrad tiles of the same launch
        chain_descs, chain_deps = [], []
        for l in (4, 3, 2, 1):
            n_i = 784 if l == 4 else 512
            n_t = 10 if l == 4 else 512
            K = L if l == 1 else 512
            x_i = self.Z[: 2 * B] if l == 1 else self.id_h[l - 2]
            x_t = self.Z[B:] if l == 1 else self.td_h[l - 2]
            wgrads = [
                D(dyi, x_i, g[f"image_decoder.fc{l}.weight"], n_i, K, 2 * B, a_mn=True, b_mn=True,
                  split_k=split, accumulate=True),
                D(dyt, x_t, g[f"text_decoder.fc{l}.weight"], n_t, K, 2 * B, a_mn=True, b_mn=True,
                  split_k=split, accumulate=True)]
            if l > 1:  # dA_{l-1} = (dy W_l) * swish'(a_{l-1}); its column sums are the bias gradient of layer l-1
                dxi, dxt = self.id_dA[l - 2], self.td_dA[l - 2]
                dgrads = [   # (share 4: the two dgrads compete with the two split-K wgrads of the layer for the SMs)
                    S(f"dgi{l}", dyi, p[f"image_decoder.fc{l}.weight"], dxi, 2 * B, K, n_i, share=4, b_mn=True,
                      aux=self.id_a[l - 2], epilogue=ops.EPI_MUL_DSWISH, colsum=g[f"image_decoder.fc{l - 1}.bias"]),
                    S(f"dgt{l}", dyt, p[f"text_decoder.fc{l}.weight"], dxt, 2 * B, K, n_t, share=4, b_mn=True,
                      aux=self.td_a[l - 2], epilogue=ops.EPI_MUL_DSWISH, colsum=g[f"text_decoder.fc{l - 1}.bias"])]
            else:  # dZ is zero-initialised; both decoders add into it (the joint rows get both)
                dxi, dxt = self.dZ[: 2 * B], self.dZ[B:]
                dgrads = [
                    D(dyi, p["image_decoder.fc1.weight"], dxi, 2 * B, K, n_i, b_mn=True, accumulate=True),
                    D(dyt, p["text_decoder.fc1.weight"], dxt, 2 * B, K, n_t, b_mn=True, accumulate=True)]
            if self.chain:
                # per layer: [dgrad_i, dgrad_t, wgrad_i, wgrad_t]; dy of layer l is the dgrad output of layer l+1, four
                # problems back (the dgrads come first so the dependent chain advances ahead of the filler wgrads)
                base = len(chain_descs)
                dep_i, dep_t = (-1, -1) if l == 4 else (base - 4, base - 3)
                chain_descs += dgrads + wgrads
                chain_deps += [dep_i, dep_t, dep_i, dep_t]
            else:
                ops.gemm_batch(wgrads + dgrads, P)
            dyi, dyt = dxi, dxt
        if self.chain:   # the whole decoder backward (16 problems) is one launch
            ops.gemm_chain(chain_descs, chain_deps, self.chain_ws, P)
        # ---- PoE / reparam / KL backward -> gradients of both encoders' outputs (summed over passes)
        mu_e, lv_e, dmu, dlv, gather = self._label_experts()
        ops.poe_bwd(mu_e, lv_e, _PASS_MASKS, B, L, self.dZ, dmu, dlv, kl_scale=1.0 / b_global, variant=0,
                    training=training, noise=self.noise if training else None, kl_scale_dev=self.beta_dev, gather=gather)
        # ---- encoders backward
        nk = max(1, B // 32)
        split = max(1, min(nk // 16, 32))
        arena = self.arena
        gwi = arena.span(1, "image_encoder.fc31.weight", "image_encoder.fc32.weight").view(2 * L, 512)
        gbi = arena.span(1, "image_encoder.fc31.bias", "image_encoder.fc32.bias")
        gwt = arena.span(1, "text_encoder.fc31.weight", "text_encoder.fc32.weight").view(2 * L, 512)
        gbt = arena.span(1, "text_encoder.fc31.bias", "text_encoder.fc32.bias")
        wi = arena.span(0, "image_encoder.fc31.weight", "image_encoder.fc32.weight").view(2 * L, 512)
        wt = arena.span(0, "text_encoder.fc31.weight", "text_encoder.fc32.weight").view(2 * L, 512)
        with self._fork():          # bias gradients of the encoder heads: beside the encoder backward chain
            ops.colsum_accumulate(self.d_enc_i, gbi)
            if not self.label_table:
                ops.colsum_accumulate(self.d_enc_t, gbt)
        wg_hi = D(self.d_enc_i, self.ie_h2, gwi, 2 * L, 512, B, a_mn=True, b_mn=True, split_k=split, accumulate=True)
        wg_ht = D(self.d_enc_t, self.te_h2, gwt, 2 * L, 512, B, a_mn=True, b_mn=True, split_k=split, accumulate=True)
        dg_hi = D(self.d_enc_i, wi, self.ie_dA[0], B, 512, 2 * L, b_mn=True, aux=self.ie_a2,
                  epilogue=ops.EPI_MUL_DSWISH, colsum=g["image_encoder.fc2.bias"])   # (K = 128: 4 k-blocks, nothing to split)
        dg_ht = D(self.d_enc_t, wt, self.te_dA[0], B, 512, 2 * L, b_mn=True, aux=self.te_a2,
                  epilogue=ops.EPI_MUL_DSWISH, colsum=g["text_encoder.fc2.bias"])
        wg_2i = D(self.ie_dA[0], self.ie_h1, g["image_encoder.fc2.weight"], 512, 512, B, a_mn=True, b_mn=True,
                  split_k=split, accumulate=True)
        wg_2t = D(self.te_dA[0], self.te_h1, g["text_encoder.fc2.weight"], 512, 512, B, a_mn=True, b_mn=True,
                  split_k=split, accumulate=True)
        dg_2i = S("dg_2i", self.ie_dA[0], p["image_encoder.fc2.weight"], self.ie_dA[1], B, 512, 512, share=2, b_mn=True,
                  aux=self.ie_a1, epilogue=ops.EPI_MUL_DSWISH, colsum=g["image_encoder.fc1.bias"])
        dg_2t = D(self.te_dA[0], p["text_encoder.fc2.weight"], self.te_dA[1], B, 512, 512, b_mn=True)
        wg_1i = D(self.ie_dA[1], self.x, g["image_encoder.fc1.weight"], 512, 784, B, a_mn=True, b_mn=True,
                  split_k=split, accumulate=True)
        if self.label_table:
            # the label encoder's backward on its class table (side stream) next to the image encoder's chained backward
            emb, w2, _, w3, _ = self._label_encoder(0)
            g_emb, g_w2, g_b2, g_w3, g_b3 = self._label_encoder(1)
            with self._fork():
                ops.label_table_bwd(emb, w2, w3, self.tt_a2, self.tt_h2, self.d_tab, self.tt_dA2, g_emb, g_w2, g_b2, g_w3, g_b3)
            if self.chain:
                ops.gemm_chain([dg_hi, wg_hi, dg_2i, wg_2i, wg_1i], [-1, -1, 0, 0, 2], self.chain_ws, P)
            else:
                ops.gemm_batch([wg_hi, dg_hi], P); ops.gemm_batch([wg_2i, dg_2i], P); ops.gemm_batch([wg_1i], P)
            self._join()
            return
        if self.chain:   # both encoders' backward: one launch
            ops.gemm_chain([dg_hi, dg_ht, wg_hi, wg_ht, dg_2i, dg_2t, wg_2i, wg_2t, wg_1i],
                           [-1, -1, -1, -1, 0, 1, 0, 1, 4], self.chain_ws, P)
        else:
            ops.gemm_batch([wg_hi, wg_ht, dg_hi, dg_ht], P)
            ops.gemm_batch([wg_2i, wg_2t, dg_2i, dg_2t], P)
            ops.gemm_batch([wg_1i], P)
        ops.embedding_swish_bwd(p["text_encoder.fc1.weight"], self.text, self.te_dA[1], g["text_encoder.fc1.weight"])

    def _enqueue_fwd_bwd(self, training: bool, use_noise_input: bool) -> None:
        b_global = self.B * self.world
        self.grad_bucket.zero_()
        self.zero_region.zero_()
        if self.presplit:
            ops.split_lo(self.flat_params, self.params_lo)
        self._enqueue_forward(training, use_noise_input)
        self._enqueue_loss_and_backward(training, b_global)
        self._join()                # side-stream bias-gradient sums / label-table backward
        ops.elbo_finalize(self.acc[0:3], self.acc[3:6], self.acc[6:9], 3, self.lam_i, self.lam_t, 1.0, 1.0 / b_global,
                          self.loss_tail, beta_dev=self.beta_dev)

    def _enqueue_allreduce(self, update: bool = True) -> None:
        """The one exchange step of the data-parallel path: SUM all-reduce of the flat gradient bucket (+ loss tail)
        over NCCL.  Kept OUT of the CUDA graphs (capturing NCCL needs every rank's watchdog to stay quiet)."""
        if self.world > 1 and (self.dp_mode != "p2p" or not update):
            import torch.distributed as dist
            dist.all_reduce(self.grad_bucket, group=self.pg)
            if self.dp_mode == "p2p":          # gradient-only step in p2p mode: the fused kernel is not run
                self.loss_sum.copy_(self.loss_tail)

    def _enqueue_update(self) -> None:
        if self.dp_mode == "p2p":
            ops.allreduce_adam_p2p(self._p2p_ptrs["grads"], self._p2p_ptrs["params"], self._p2p_ptrs["flags"], self.adam_m,
                                   self.adam_v, self.arena.numel, 4, self.loss_sum, self.rank, self.world, self.step_count,
                                   lr=self.lr)
        else:
            ops.adam_flat(self.flat_params, self.flat_grads, self.adam_m, self.adam_v, self.step_count, lr=self.lr)

    # ------------------------------------------------------------------ NVLink peer-memory plumbing (torch symmetric memory)
    def _group(self):
        import torch.distributed as dist
        return self.pg if self.pg is not None else dist.group.WORLD

    def _p2p_agreed(self) -> bool:
        """Peer memory needs one GPU per rank on one node with P2P access (NVLink / NVSwitch) -- decided collectively: the
        ranks exchange their device ordinals, every rank checks access to every other rank's device, and the answer is the
        MIN over ranks."""
        import torch.distributed as dist
        if not dist.is_initialized():
            return False
        ok = 1
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
            if dist.get_backend(self._group()) != "nccl":
                ok = 0
        except Exception:  # noqa: BLE001
            ok = 0
        if ok == 0 and dist.get_backend(self._group()) != "nccl":
            return False           # (a gloo group on CPU tensors: every rank sees the same backend, no vote needed)
        me = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
        mine = torch.tensor([me], dtype=torch.int64, device=self.dev)
        devs = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(devs, mine, group=self._group())
        devs = [int(d.item()) for d in devs]
        if len(set(devs)) != self.world:          # two ranks on one GPU: no peer mapping to speak of
            ok = 0
        else:
            try:
                ok = min(ok, int(all(torch.cuda.can_device_access_peer(me, d) for d in devs if d != me)))
            except Exception:  # noqa: BLE001
                ok = 0
        vote = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=self._group())
        return bool(vote.item())

    def _symm_alloc(self, i: int, numel: int):
        """Parameters (buffer 0) and gradients (buffer 1) live in symmetric memory so that every rank can address every
        other rank's copy; the Adam moments stay private."""
        if i > 1:
            return None
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(numel, dtype=torch.float32, device=self.dev)
        t.zero_()
        self._symm["params" if i == 0 else "grads"] = t
        return t

    def _p2p_finish_setup(self, n: int, tail: int) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        flags = symm.empty(2 * self.world + 4, dtype=torch.int32, device=self.dev)
        flags.zero_()
        self._symm["flags"] = flags
        torch.cuda.synchronize(self.dev)
        self._p2p_ptrs = {}
        for k in ("params", "grads", "flags"):
            h = symm.rendezvous(self._symm[k], self._group())
            self._p2p_ptrs[k] = [int(x) for x in h.buffer_ptrs]
            self._symm[k + "_handle"] = h          # keeps the mappings alive
        self.loss_sum = torch.zeros(tail, dtype=torch.float32, device=self.dev)
        self.loss_out = self.loss_sum              # what step() copies to the host: the sums over ranks
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self._group())          # every rank's flags are zero before anyone can signal

    def _enqueue_step(self, training: bool, use_noise_input: bool, update: bool) -> None:
        self._enqueue_fwd_bwd(training, use_noise_input)
        self._enqueue_allreduce(update)
        if update:
            self._enqueue_update()

    # ------------------------------------------------------------------ public API
    def _stage_beta(self, annealing_factor: float) -> None:
        """KL annealing factor -> device scalar (call with the step stream current).  Ring of pinned words guarded by events:
        a slot is rewritten only after the copy that read it has executed."""
        buf, ev = self._beta_ring[self._beta_idx]
        self._beta_idx = (self._beta_idx + 1) % len(self._beta_ring)
        ev.synchronize()                 # (never recorded -> returns at once)
        buf[0] = float(annealing_factor)
        self.beta_dev.copy_(buf, non_blocking=True)
        ev.record(self._stream)

    def check_device_errors(self) -> None:
        """Raise if a bounded device-side wait gave up since the last check: the chained GEMM's producer wait
        (chain_ws[1]) or the peer-memory exchange's rendezvous (flag word 2*world).  Called at every host
        synchronisation point (step(sync=True), flush(), synchronize()); costs one tiny D2H copy."""
        bad = int(self.chain_ws[1].item())
        if bad:
            self.chain_ws[1] = 0
            raise _lib.MvaeError("mvae_gemm_chain: a dependency wait timed out on the device (results of the last steps are "
                                 "invalid); was the GPU time-sliced or a kernel of the chain preempted?")
        if self.dp_mode == "p2p":
            err = int(self._symm["flags"][2 * self.world].item())
            if err:
                raise _lib.MvaeError(f"mvae_allreduce_adam_p2p: rank {self.rank} gave up waiting for its peers (code {err}: "
                                     "1 = timeout, 2 = a peer is at a different step); that step's update was skipped")

    def set_inputs(self, image: torch.Tensor, text: torch.Tensor, noise: Optional[torch.Tensor] = None,
                   annealing_factor: float = 1.0) -> None:
        """Stage one batch (host or device tensors) into the device-resident input buffers on the step stream.
        ``noise``: optional [3,B,L] N(0,1) draws in the REFERENCE's pass order (joint, image, text)."""
        B, L = self.B, self.L
        with torch.cuda.stream(self._stream):
            self.x.copy_(image.reshape(B, 784), non_blocking=True)
            self.text.copy_(text.reshape(B), non_blocking=True)
            self._stage_beta(annealing_factor)
            if noise is not None:
                nz = self.noise.view(3, B, L)
                for ref_i, int_i in enumerate(_REF_TO_INTERNAL):
                    nz[int_i].copy_(noise[ref_i], non_blocking=True)

    def _warmup_state(self) -> List[torch.Tensor]:
        """Every tensor an (eager, real) warm-up step may change and a later step reads: parameters, Adam moments and
        step counter here; flavours add their non-arena state (BatchNorm running statistics ...)."""
        return [self.flat_params, self.adam_m, self.adam_v, self.step_count]

    def _warmup_snapshot(self):
        return [t.clone() for t in self._warmup_state()]

    def _warmup_restore(self, saved) -> None:
        for t, s in zip(self._warmup_state(), saved):
            t.copy_(s)

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(g, stream=self._stream, capture_error_mode="thread_local"):
            fn()
        return g, _lib.launch_count() - n0

    def run(self, training: bool = True, noise_given: bool = False, update: bool = True) -> None:
        """Enqueue one step (graph replay when enabled).  Does not synchronise.
        world == 1: one graph for the whole step.  world > 1: graph(fwd+bwd) -> NCCL all-reduce -> graph(Adam)."""
        key = (training, noise_given, update)
        with torch.cuda.stream(self._stream):
            if not self.use_graph:
                n0 = _lib.launch_count()
                self._enqueue_step(training, noise_given, update)
                self.launches_per_step = _lib.launch_count() - n0
                return
            gr = self._graphs.get(key)
            if gr is None:
                # warm-up once eagerly (sets func attributes, loads modules, inits NCCL), then capture
                saved = self._warmup_snapshot()
                self._enqueue_step(training, noise_given, update)
                self._stream.synchronize()
                self._warmup_restore(saved)
                self._stream.synchronize()
                if self.world == 1 or (self.dp_mode == "p2p" and update):
                    # (p2p: the exchange is one of OUR kernels, so the whole step -- including it -- is one graph)
                    g, n = self._capture(lambda: self._enqueue_step(training, noise_given, update))
                    gr = (g, None)
                else:
                    g1, n1 = self._capture(lambda: self._enqueue_fwd_bwd(training, noise_given))
                    g2, n2 = self._capture(self._enqueue_update) if update else (None, 0)
                    gr, n = (g1, g2), n1 + n2
                self.launches_per_step = n
                self._graphs[key] = gr
            gr[0].replay()
            if self.world > 1 and not (self.dp_mode == "p2p" and update):
                self._enqueue_allreduce(update)
                if gr[1] is not None:
                    gr[1].replay()

    def step(self, image: torch.Tensor, text: torch.Tensor, annealing_factor: float = 1.0,
             noise: Optional[torch.Tensor] = None, training: bool = True, update: bool = True,
             sync: bool = True) -> Optional[float]:
        """One training iteration of mnist/train.py:196-219 on this rank's shard.  Returns the loss
        (joint + image + text ELBO, global-batch mean) if ``sync`` else None (read ``loss_host`` later)."""
        self.set_inputs(image, text, noise, annealing_factor)
        self.run(training=training, noise_given=noise is not None, update=update)
        with torch.cuda.stream(self._stream):
            self.loss_host.copy_(self.loss_out, non_blocking=True)
        if sync:
            self._stream.synchronize()
            self.check_device_errors()
            return float(self.loss_host[0])
        return None

    # ------------------------------------------------------------------ device-resident dataset (no per-step H2D at all)
    def attach_dataset(self, images_u8: torch.Tensor, labels: torch.Tensor) -> None:
        """Keep the whole uint8 dataset in HBM (MNIST: 47 MB): ``images_u8`` [N, 784] / [N,1,28,28] uint8 as stored on
        disk, ``labels`` [N].  Batches are then built on the device by one gather launch (ToTensor's /255 and the label
        lookup fused), replacing DataLoader + H2D of mnist/train.py:159-165,188-193."""
        if images_u8.dtype != torch.uint8:
            raise _lib.MvaeError("attach_dataset: images must be uint8 (the /255 of ToTensor happens on the device)")
        n = images_u8.shape[0]
        self.ds_images = images_u8.reshape(n, -1).contiguous().to(self.dev)
        if self.ds_images.shape[1] != self.x.shape[1]:
            raise _lib.MvaeError(f"attach_dataset: rows of {self.ds_images.shape[1]} bytes, the model takes {self.x.shape[1]}")
        self.ds_labels = labels.reshape(n).to(torch.int64).to(self.dev)

    def epoch_permutation(self, seed: int) -> torch.Tensor:
        """A device-side shuffle of the attached dataset (DataLoader(shuffle=True)); slice it into batches of B."""
        g = torch.Generator(device=self.dev).manual_seed(seed)
        return torch.randperm(self.ds_images.shape[0], generator=g, device=self.dev)

    def step_from_dataset(self, idx: torch.Tensor, annealing_factor: float = 1.0, training: bool = True,
                          update: bool = True, sync: bool = True) -> Optional[float]:
        """One training iteration on rows ``idx`` (int64 [B], on the device) of the attached dataset."""
        if idx.numel() != self.B or idx.dtype != torch.int64 or not idx.is_cuda:
            raise _lib.MvaeError(f"step_from_dataset: idx must be a CUDA int64 tensor of {self.B} row indices")
        with torch.cuda.stream(self._stream):
            ops.gather_batch_u8(self.ds_images, self.ds_labels, idx.contiguous(), self.x, self.text)
            self._stage_beta(annealing_factor)
        self.run(training=training, noise_given=False, update=update)
        with torch.cuda.stream(self._stream):
            self.loss_host.copy_(self.loss_out, non_blocking=True)
        if sync:
            self._stream.synchronize()
            self.check_device_errors()
            return float(self.loss_host[0])
        return None

    # ------------------------------------------------------------------ host-fed, double-buffered stepping
    # flavour hooks of the pipelined path: shapes of one staged batch, and how a staged batch becomes the step's inputs
    def _pipe_slot_tensors(self) -> Dict[str, torch.Tensor]:
        B, dev = self.B, self.dev
        return {"img": torch.empty(B, 784, dtype=torch.float32, device=dev),
                "oth": torch.empty(B, dtype=torch.int64, device=dev)}

    def _pipe_consume(self, slot) -> None:
        """(step stream current) staged batch -> the device-resident input buffers the step reads."""
        self.x.copy_(slot["img"], non_blocking=True)
        self.text.copy_(slot["oth"], non_blocking=True)

    def _pipe_after_run(self, training: bool, update: bool) -> None:
        pass

    def _pipe_init(self) -> None:
        self._copy_stream = torch.cuda.Stream(device=self.dev)
        self._pipe = []
        for _ in range(2):
            slot = dict(self._pipe_slot_tensors())
            slot.update({"beta": torch.ones(1, dtype=torch.float32).pin_memory(),
                         "loss": torch.zeros(4, dtype=torch.float32).pin_memory(),
                         "up": torch.cuda.Event(), "free": torch.cuda.Event(), "done": torch.cuda.Event(), "busy": False})
            slot["free"].record(self._stream)
            self._pipe.append(slot)
        self._pipe_idx = 0

    def step_pipelined(self, image: torch.Tensor, text: torch.Tensor, annealing_factor: float = 1.0,
                       training: bool = True, update: bool = True) -> Optional[float]:
        """Like ``step`` for HOST batches (pinned memory recommended), but the upload of this batch runs on a copy
        stream while the previous step is still computing, and the loss that is read back (and returned) is the
        PREVIOUS call's -- the usual one-step-lagged logging of an asynchronous training loop.  Returns None on the
        first call; ``flush()`` returns the last loss.  Every call still moves one batch host->device and one loss
        device->host.  (``text`` = labels [B] for the MNIST-shape flavours, attrs [B,18] for CelebA.)"""
        if not hasattr(self, "_pipe"):
            self._pipe_init()
        k = self._pipe_idx
        self._pipe_idx ^= 1
        cur, prev = self._pipe[k], self._pipe[k ^ 1]
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(cur["free"])            # the compute stream finished reading this staging slot
            cur["img"].copy_(image.reshape(cur["img"].shape), non_blocking=True)
            cur["oth"].copy_(text.reshape(cur["oth"].shape), non_blocking=True)
            cur["up"].record(self._copy_stream)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(cur["up"])
            self._pipe_consume(cur)
            cur["free"].record(self._stream)
            cur["beta"][0] = float(annealing_factor)
            self.beta_dev.copy_(cur["beta"], non_blocking=True)
        self.run(training=training, noise_given=False, update=update)
        self._pipe_after_run(training, update)
        with torch.cuda.stream(self._stream):
            cur["loss"].copy_(self.loss_out, non_blocking=True)
            cur["done"].record(self._stream)
        cur["busy"] = True
        if prev["busy"]:
            prev["done"].synchronize()
            prev["busy"] = False
            return float(prev["loss"][0])
        return None

    def flush(self) -> Optional[float]:
        """Wait for the in-flight pipelined step and return its loss."""
        out = None
        if hasattr(self, "_pipe"):
            for slot in (self._pipe[self._pipe_idx], self._pipe[self._pipe_idx ^ 1]):
                if slot["busy"]:
                    slot["done"].synchronize()
                    slot["busy"] = False
                    out = float(slot["loss"][0])
                    self.loss_host.copy_(slot["loss"])
        self.check_device_errors()
        return out

    def losses(self) -> Dict[str, float]:
        """(after a synchronised step) the reference's three ELBO terms and their sum."""
        v = self.loss_host
        return {"total": float(v[0]), "joint": float(v[1 + _REF_TO_INTERNAL[0]]),
                "image": float(v[1 + _REF_TO_INTERNAL[1]]), "text": float(v[1 + _REF_TO_INTERNAL[2]])}

    def synchronize(self) -> None:
        self._stream.synchronize()
        self.check_device_errors()
