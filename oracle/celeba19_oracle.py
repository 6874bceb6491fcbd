"""CPU restatement of the CelebA-19 MVAE (19 modalities: image + 18 single-attribute experts; celeba19/model.py,
celeba19/train.py) -- test infrastructure only.  Image encoder/decoder are the CelebA DCGAN nets (shared code with
``celeba_oracle``); each attribute has its own Embedding(2,512)->512->2L encoder and L->512->512->512->1 decoder."""
from __future__ import annotations

from math import comb
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import celeba_oracle as C
from .mvae_oracle import bce_with_logits, kl_rows, product_of_experts, prior_expert, reparametrize, swish

Tensor = torch.Tensor
N_ATTRS = 18


def celeba19_state_shapes(L: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict order of celeba19/model.py:21-31: image nets (as celeba), 18 attr encoders, 18 attr decoders."""
    out = [(k, s) for k, s in C.celeba_state_shapes(L) if k.startswith("image_")]
    for i in range(N_ATTRS):
        p = f"attr_encoders.{i}.net"
        out += [(f"{p}.0.weight", (2, 512)), (f"{p}.2.weight", (512, 512)), (f"{p}.2.bias", (512,)),
                (f"{p}.4.weight", (2 * L, 512)), (f"{p}.4.bias", (2 * L,))]
    for i in range(N_ATTRS):
        p = f"attr_decoders.{i}.net"
        out += [(f"{p}.0.weight", (512, L)), (f"{p}.0.bias", (512,)), (f"{p}.2.weight", (512, 512)), (f"{p}.2.bias", (512,)),
                (f"{p}.4.weight", (512, 512)), (f"{p}.4.bias", (512,)), (f"{p}.6.weight", (1, 512)), (f"{p}.6.bias", (1,))]
    return out


def celeba19_param_shapes(L: int):
    return [(k, s) for k, s in celeba19_state_shapes(L) if not C.is_buffer(k)]


def make_celeba19_state(L: int, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    import math
    st = {k: v for k, v in C.make_celeba_state(L, seed, dtype).items() if k.startswith("image_")}
    rs = np.random.RandomState(seed + 1)
    fan = {}
    for name, shape in celeba19_state_shapes(L):
        if name in st:
            continue
        prefix = name.rsplit(".", 1)[0]
        if name.endswith("net.0.weight") and shape == (2, 512):
            v = rs.standard_normal(shape)
        elif name.endswith(".weight"):
            fan[prefix] = shape[1]
            v = rs.uniform(-1, 1, shape) / math.sqrt(shape[1])
        else:
            v = rs.uniform(-1, 1, shape) / math.sqrt(fan[prefix])
        st[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return {k: st[k] for k, _ in celeba19_state_shapes(L)}


def attr_encoder(st, i: int, a: Tensor, L: int):
    """celeba19/model.py:180-184 (input cast with .long())."""
    p = f"attr_encoders.{i}.net"
    h = swish(st[p + ".0.weight"][a.long()])
    h = swish(torch.addmm(st[p + ".2.bias"], h, st[p + ".2.weight"].t()))
    o = torch.addmm(st[p + ".4.bias"], h, st[p + ".4.weight"].t())
    return o[:, :L], o[:, L:]


def attr_decoder(st, i: int, z: Tensor):
    """celeba19/model.py:207-209 -> [B,1]."""
    p = f"attr_decoders.{i}.net"
    h = z
    for l in (0, 2, 4):
        h = swish(torch.addmm(st[f"{p}.{l}.bias"], h, st[f"{p}.{l}.weight"].t()))
    return torch.addmm(st[p + ".6.bias"], h, st[p + ".6.weight"].t())


def forward(st, image, attrs: Sequence[Optional[Tensor]], L, noise, training, drop_mask):
    """celeba19/model.py:41-89: experts in the order image, attr 0..17 (present ones); ALL 19 decoders run."""
    B = image.size(0) if image is not None else next(a for a in attrs if a is not None).size(0)
    dtype = st["image_encoder.features.0.weight"].dtype
    mu, logvar = prior_expert((1, B, L), dtype)
    if image is not None:
        m, lv = C.image_encoder(st, image, L, training, drop_mask)
        mu = torch.cat((mu, m.unsqueeze(0)), 0); logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    for i in range(N_ATTRS):
        if attrs[i] is not None:
            m, lv = attr_encoder(st, i, attrs[i], L)
            mu = torch.cat((mu, m.unsqueeze(0)), 0); logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    mu, logvar = product_of_experts(mu, logvar, variant="B")
    z = reparametrize(mu, logvar, noise if training else None)
    return C.image_decoder(st, z, training), [attr_decoder(st, i, z).squeeze(1) for i in range(N_ATTRS)], mu, logvar


def elbo19(recon: Sequence[Tensor], data: Sequence[Tensor], mu, logvar, lam_i=1.0, lam_a=1.0, beta=1.0):
    """celeba19/train.py:26-60: >1-D modality = image (sum over pixels), else attribute (element-wise [B])."""
    assert len(recon) == len(data), "must supply ground truth for every modality."
    B = mu.size(0)
    bce = 0
    for r, d in zip(recon, data):
        if r.dim() > 1:
            bce = bce + lam_i * torch.sum(bce_with_logits(r.reshape(B, -1), d.reshape(B, -1)), dim=1)
        else:
            bce = bce + lam_a * bce_with_logits(r, d)
    return torch.mean(bce + beta * kl_rows(mu, logvar))


def pass_list(combos: Sequence[Sequence[bool]]):
    """Modality sets of one step in the reference's call order (celeba19/train.py:262-302): joint, image-only,
    18 single attributes, then the sampled combinations.  Entry = (bool[19] present, uses_script_lambdas)."""
    P = [([True] * 19, True), ([True] + [False] * 18, True)]
    for i in range(N_ATTRS):
        P.append(([False] + [k == i for k in range(N_ATTRS)], False))
    for c in combos:
        P.append(([bool(v) for v in c], False))
    return P


def step_grads(state, image, attrs, L, noises, drop_masks, combos, lam_i=1.0, lam_a=10.0, beta=1.0, training=True):
    """One training step's objective + gradients (celeba19/train.py:240-309).  attrs [B,18] float {0,1};
    noises: one [B,L] draw per pass; drop_masks: one [B,512] mask per pass that contains the image (call order).
    Quirk preserved: single-attribute and sampled terms use lambda = 1 (elbo_loss defaults, :281-300)."""
    st = {}
    for k, v in state.items():
        st[k] = v.detach().clone() if C.is_buffer(k) else v.detach().clone().requires_grad_(True)
    cols = [attrs[:, i] for i in range(N_ATTRS)]
    total = 0
    terms = []
    mi = 0
    for pi, (present, script_lam) in enumerate(pass_list(combos)):
        img = image if present[0] else None
        al = [cols[i] if present[1 + i] else None for i in range(N_ATTRS)]
        dm = None
        if present[0] and training:
            dm = drop_masks[mi]; mi += 1
        ri, ra, mu, lv = forward(st, img, al, L, noises[pi] if training else None, training, dm)
        recon, data = [], []
        if present[0]:
            recon.append(ri); data.append(image)
        for i in range(N_ATTRS):
            if present[1 + i]:
                recon.append(ra[i]); data.append(cols[i])
        li, la = (lam_i, lam_a) if script_lam else (1.0, 1.0)
        t = elbo19(recon, data, mu, lv, li, la, beta)
        terms.append(t.detach())
        total = total + t
    total.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in st.items() if not C.is_buffer(k)}
    bufs = {k: v for k, v in st.items() if C.is_buffer(k) and not k.endswith("num_batches_tracked")}
    return total.detach(), terms, grads, bufs


# ---------------------------------------------------------------------------- host-side sampler (celeba19/train.py:87-142)
def unrank_combination(n: int, k: int, idx: int) -> List[int]:
    """idx-th k-subset of range(n) in itertools.combinations (lexicographic) order."""
    out = []
    x = 0
    for j in range(k):
        while True:
            c = comb(n - x - 1, k - j - 1)
            if idx < c:
                out.append(x); x += 1
                break
            idx -= c; x += 1
    return out


def sample_combinations_fast(n_modalities: int = 19, size: int = 1, rng=np.random) -> np.ndarray:
    """Same draws and same result as the reference's ``sample_combinations(enumerate_combinations(19), size)`` but
    without materialising the 524,267 x 19 pool: subset size ~ U{2..n-1}, then a uniform subset of that size; the
    second draw replays ``np.random.choice(range(C(n,k)), cnt, replace=False)`` and un-ranks the indices."""
    pool_space = np.arange(2, n_modalities)                       # sizes that occur in the pool
    sample_pool = rng.choice(pool_space, size, replace=True)
    dist = np.bincount(sample_pool, minlength=n_modalities)
    out = []
    for k in range(n_modalities):
        if dist[k] > 0:
            idx = rng.choice(range(comb(n_modalities, k)), size=dist[k], replace=False)
            for i in idx:
                row = np.zeros(n_modalities, dtype=bool)
                row[unrank_combination(n_modalities, k, int(i))] = True
                out.append(row)
    return np.stack(out)
