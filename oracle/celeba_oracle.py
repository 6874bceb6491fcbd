"""CPU restatement of the CelebA-flavour MVAE (celeba/model.py, celeba/train.py) -- test infrastructure only.

Functional torch-CPU code (fp32 or fp64): conv stacks with train-mode BatchNorm (per call batch statistics + running
statistic updates in call order), Dropout with INJECTED masks, PoE variant B, three-pass objective.  Pinned to the
unmodified reference by tests/golden/celeba_golden.npz (tests/golden/make_golden.py replays the reference's RNG draws
to obtain the dropout masks and the reparametrisation noise)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .mvae_oracle import bce_with_logits, kl_rows, product_of_experts, prior_expert, reparametrize, swish

F = torch.nn.functional
Tensor = torch.Tensor
N_ATTRS = 18


def celeba_state_shapes(L: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Every state_dict entry (parameters AND BatchNorm buffers) in the reference's order (celeba/model.py:21-26,
    76-92,113-126,145-153,172-183)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(prefix, c):
        return [(f"{prefix}.weight", (c,)), (f"{prefix}.bias", (c,)), (f"{prefix}.running_mean", (c,)),
                (f"{prefix}.running_var", (c,)), (f"{prefix}.num_batches_tracked", ())]
    e = "image_encoder.features"
    out += [(f"{e}.0.weight", (32, 3, 4, 4)), (f"{e}.2.weight", (64, 32, 4, 4))] + bn(f"{e}.3", 64)
    out += [(f"{e}.5.weight", (128, 64, 4, 4))] + bn(f"{e}.6", 128) + [(f"{e}.8.weight", (256, 128, 4, 4))] + bn(f"{e}.9", 256)
    out += [("image_encoder.classifier.0.weight", (512, 6400)), ("image_encoder.classifier.0.bias", (512,)),
            ("image_encoder.classifier.3.weight", (2 * L, 512)), ("image_encoder.classifier.3.bias", (2 * L,))]
    d = "image_decoder.hallucinate"
    out += [("image_decoder.upsample.0.weight", (6400, L)), ("image_decoder.upsample.0.bias", (6400,))]
    out += [(f"{d}.0.weight", (256, 128, 4, 4))] + bn(f"{d}.1", 128) + [(f"{d}.3.weight", (128, 64, 4, 4))] + bn(f"{d}.4", 64)
    out += [(f"{d}.6.weight", (64, 32, 4, 4))] + bn(f"{d}.7", 32) + [(f"{d}.9.weight", (32, 3, 4, 4))]
    a = "attrs_encoder.net"
    out += [(f"{a}.0.weight", (512, N_ATTRS)), (f"{a}.0.bias", (512,))] + bn(f"{a}.1", 512)
    out += [(f"{a}.3.weight", (512, 512)), (f"{a}.3.bias", (512,))] + bn(f"{a}.4", 512)
    out += [(f"{a}.6.weight", (2 * L, 512)), (f"{a}.6.bias", (2 * L,))]
    a = "attrs_decoder.net"
    out += [(f"{a}.0.weight", (512, L)), (f"{a}.0.bias", (512,))] + bn(f"{a}.1", 512)
    out += [(f"{a}.3.weight", (512, 512)), (f"{a}.3.bias", (512,))] + bn(f"{a}.4", 512)
    out += [(f"{a}.6.weight", (512, 512)), (f"{a}.6.bias", (512,))] + bn(f"{a}.7", 512)
    out += [(f"{a}.9.weight", (N_ATTRS, 512)), (f"{a}.9.bias", (N_ATTRS,))]
    return out


def is_buffer(name: str) -> bool:
    return name.endswith("running_mean") or name.endswith("running_var") or name.endswith("num_batches_tracked")


def celeba_param_shapes(L: int):
    return [(k, s) for k, s in celeba_state_shapes(L) if not is_buffer(k)]


def make_celeba_state(L: int, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic state (numpy RandomState stream): conv/linear U(+-1/sqrt(fan_in)), BN weight U(0.5,1.5), BN bias
    U(-0.5,0.5), running_mean N(0,0.1), running_var U(0.5,1.5) -- non-trivial values so that layout bugs show."""
    import math
    import numpy as np
    rs = np.random.RandomState(seed)
    shapes = celeba_state_shapes(L)
    is_bn = {k[: -len(".running_mean")] for k, _ in shapes if k.endswith(".running_mean")}
    out: Dict[str, Tensor] = {}
    fan = {}
    for name, shape in shapes:
        prefix = name.rsplit(".", 1)[0]
        if name.endswith("num_batches_tracked"):
            out[name] = torch.tensor(0, dtype=torch.int64); continue
        if prefix in is_bn:
            kind = name.rsplit(".", 1)[1]
            v = {"weight": rs.uniform(0.5, 1.5, shape), "bias": rs.uniform(-0.5, 0.5, shape),
                 "running_mean": 0.1 * rs.standard_normal(shape), "running_var": rs.uniform(0.5, 1.5, shape)}[kind]
        elif name.endswith(".weight"):
            if "hallucinate" in name and len(shape) == 4:
                fan_in = shape[1] * 16
            else:
                fan_in = int(np.prod(shape[1:]))
            fan[prefix] = fan_in
            v = rs.uniform(-1, 1, shape) / math.sqrt(fan_in)
        else:
            v = rs.uniform(-1, 1, shape) / math.sqrt(fan.get(prefix, shape[0]))
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return out


def _bn(st, prefix, x, training):
    """nn.BatchNorm{1,2}d forward; in training mode updates st[prefix.running_*] in place (momentum 0.1)."""
    return F.batch_norm(x, st[prefix + ".running_mean"], st[prefix + ".running_var"], st[prefix + ".weight"],
                        st[prefix + ".bias"], training, 0.1, 1e-5)


def _lin(st, name, x):
    return torch.addmm(st[name + ".bias"], x, st[name + ".weight"].t())


def image_encoder(st, x, L, training, drop_mask: Optional[Tensor]):
    """celeba/model.py:94-100.  drop_mask: [B,512] of {0,1} (train) or None (eval / p = 0)."""
    e = "image_encoder.features"
    h = swish(F.conv2d(x, st[e + ".0.weight"], None, 2, 1))
    h = swish(_bn(st, e + ".3", F.conv2d(h, st[e + ".2.weight"], None, 2, 1), training))
    h = swish(_bn(st, e + ".6", F.conv2d(h, st[e + ".5.weight"], None, 2, 1), training))
    h = swish(_bn(st, e + ".9", F.conv2d(h, st[e + ".8.weight"], None, 1, 0), training))
    h = swish(_lin(st, "image_encoder.classifier.0", h.reshape(-1, 256 * 5 * 5)))
    if drop_mask is not None:
        h = h * drop_mask / 0.9
    o = _lin(st, "image_encoder.classifier.3", h)
    return o[:, :L], o[:, L:]


def image_decoder(st, z, training):
    """celeba/model.py:128-133."""
    d = "image_decoder.hallucinate"
    h = swish(_lin(st, "image_decoder.upsample.0", z)).reshape(-1, 256, 5, 5)
    h = swish(_bn(st, d + ".1", F.conv_transpose2d(h, st[d + ".0.weight"], None, 1, 0), training))
    h = swish(_bn(st, d + ".4", F.conv_transpose2d(h, st[d + ".3.weight"], None, 2, 1), training))
    h = swish(_bn(st, d + ".7", F.conv_transpose2d(h, st[d + ".6.weight"], None, 2, 1), training))
    return F.conv_transpose2d(h, st[d + ".9.weight"], None, 2, 1)


def attrs_encoder(st, a, L, training):
    """celeba/model.py:157-160."""
    n = "attrs_encoder.net"
    h = swish(_bn(st, n + ".1", _lin(st, n + ".0", a), training))
    h = swish(_bn(st, n + ".4", _lin(st, n + ".3", h), training))
    o = _lin(st, n + ".6", h)
    return o[:, :L], o[:, L:]


def attrs_decoder(st, z, training):
    """celeba/model.py:185-190."""
    n = "attrs_decoder.net"
    h = swish(_bn(st, n + ".1", _lin(st, n + ".0", z), training))
    h = swish(_bn(st, n + ".4", _lin(st, n + ".3", h), training))
    h = swish(_bn(st, n + ".7", _lin(st, n + ".6", h), training))
    return _lin(st, n + ".9", h)


def forward(st, image, attrs, L, noise, training, drop_mask):
    """celeba/model.py:35-63 (PoE variant B, :200-207)."""
    B = image.size(0) if image is not None else attrs.size(0)
    dtype = st["image_encoder.features.0.weight"].dtype
    mu, logvar = prior_expert((1, B, L), dtype)
    if image is not None:
        m, lv = image_encoder(st, image, L, training, drop_mask)
        mu = torch.cat((mu, m.unsqueeze(0)), 0); logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    if attrs is not None:
        m, lv = attrs_encoder(st, attrs, L, training)
        mu = torch.cat((mu, m.unsqueeze(0)), 0); logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    mu, logvar = product_of_experts(mu, logvar, variant="B")
    z = reparametrize(mu, logvar, noise if training else None)
    return image_decoder(st, z, training), attrs_decoder(st, z, training), mu, logvar


def elbo(recon_image, image, recon_attrs, attrs, mu, logvar, lam_i, lam_a, beta):
    """celeba/train.py:22-65."""
    B = mu.size(0)
    img = 0
    att = 0
    if recon_image is not None and image is not None:
        img = torch.sum(bce_with_logits(recon_image.reshape(B, -1), image.reshape(B, -1)), dim=1)
    if recon_attrs is not None and attrs is not None:
        for i in range(N_ATTRS):
            att = att + bce_with_logits(recon_attrs[:, i], attrs[:, i])
    return torch.mean(lam_i * img + lam_a * att + beta * kl_rows(mu, logvar))


def step_grads(state: Dict[str, Tensor], image, attrs, L, noises: Sequence[Optional[Tensor]],
               drop_masks: Sequence[Optional[Tensor]], lam_i=1.0, lam_a=10.0, beta=1.0, training=True):
    """Three-pass objective + gradients (celeba/train.py:189-211).  ``state`` is NOT modified; returns
    (loss, (joint, image, attrs), grads, new_buffers, aux).  noises: 3 draws (joint, image, attrs);
    drop_masks: 2 masks (joint pass, image-only pass)."""
    st = {}
    for k, v in state.items():
        if is_buffer(k):
            st[k] = v.detach().clone()
        else:
            st[k] = v.detach().clone().requires_grad_(True)
    r1 = forward(st, image, attrs, L, noises[0], training, drop_masks[0] if training else None)
    r2 = forward(st, image, None, L, noises[1], training, drop_masks[1] if training else None)
    r3 = forward(st, None, attrs, L, noises[2], training, None)
    j = elbo(r1[0], image, r1[1], attrs, r1[2], r1[3], lam_i, lam_a, beta)
    i = elbo(r2[0], image, None, None, r2[2], r2[3], lam_i, lam_a, beta)
    a = elbo(None, None, r3[1], attrs, r3[2], r3[3], lam_i, lam_a, beta)
    loss = j + i + a
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in st.items() if not is_buffer(k)}
    bufs = {k: v for k, v in st.items() if is_buffer(k) and not k.endswith("num_batches_tracked")}
    aux = {"mu": (r1[2], r2[2], r3[2]), "logvar": (r1[3], r2[3], r3[3]), "recon_image": (r1[0], r2[0], r3[0]),
           "recon_attrs": (r1[1], r2[1], r3[1])}
    return loss.detach(), (j.detach(), i.detach(), a.detach()), grads, bufs, aux
