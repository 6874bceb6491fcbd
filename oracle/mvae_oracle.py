"""CPU restatement of the reference MVAE training-step math (test infrastructure).

Every function cites the reference file:line it follows (paths relative to the
upstream repo mhw32/multimodal-vae-public).  The arithmetic is plain torch on the
CPU in the dtype of the inputs (fp32 to mirror the reference, fp64 to budget
tolerances); gradients come from torch autograd, exactly as in the reference.
A second, independent numpy restatement of the element-wise pieces with
*analytic* gradients lives in ``oracle/elementwise_np.py`` and is used to check
the formulas the CUDA kernels implement.

This file is NOT on the product path (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# ----------------------------------------------------------------------------
# parameters: deterministic, torch-RNG independent (numpy legacy MT19937 stream)
# ----------------------------------------------------------------------------

def mnist_param_shapes(n_latents: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict names/shapes of the MNIST MVAE, reference order
    (mnist/model.py:20-27, 75-78, 95-98, 116-119, 136-139)."""
    L = n_latents
    return [
        ("image_encoder.fc1.weight", (512, 784)), ("image_encoder.fc1.bias", (512,)),
        ("image_encoder.fc2.weight", (512, 512)), ("image_encoder.fc2.bias", (512,)),
        ("image_encoder.fc31.weight", (L, 512)), ("image_encoder.fc31.bias", (L,)),
        ("image_encoder.fc32.weight", (L, 512)), ("image_encoder.fc32.bias", (L,)),
        ("image_decoder.fc1.weight", (512, L)), ("image_decoder.fc1.bias", (512,)),
        ("image_decoder.fc2.weight", (512, 512)), ("image_decoder.fc2.bias", (512,)),
        ("image_decoder.fc3.weight", (512, 512)), ("image_decoder.fc3.bias", (512,)),
        ("image_decoder.fc4.weight", (784, 512)), ("image_decoder.fc4.bias", (784,)),
        ("text_encoder.fc1.weight", (10, 512)),
        ("text_encoder.fc2.weight", (512, 512)), ("text_encoder.fc2.bias", (512,)),
        ("text_encoder.fc31.weight", (L, 512)), ("text_encoder.fc31.bias", (L,)),
        ("text_encoder.fc32.weight", (L, 512)), ("text_encoder.fc32.bias", (L,)),
        ("text_decoder.fc1.weight", (512, L)), ("text_decoder.fc1.bias", (512,)),
        ("text_decoder.fc2.weight", (512, 512)), ("text_decoder.fc2.bias", (512,)),
        ("text_decoder.fc3.weight", (512, 512)), ("text_decoder.fc3.bias", (512,)),
        ("text_decoder.fc4.weight", (10, 512)), ("text_decoder.fc4.bias", (10,)),
    ]


def make_params(shapes: Sequence[Tuple[str, Tuple[int, ...]]], seed: int = 0,
                dtype=torch.float32,
                embedding_names: Sequence[str] = ("text_encoder.fc1.weight", "text_encoder.net.0.weight")) -> Params:
    """Deterministic parameters with PyTorch-default *scales* (U(+-1/sqrt(fan_in)) for
    Linear, N(0,1) for Embedding) drawn from numpy's frozen RandomState stream so
    the same values can be rebuilt on any box without the reference or torch's RNG."""
    rs = np.random.RandomState(seed)
    out: Params = {}
    fan_in_of_weight = {}
    for name, shape in shapes:
        if name in embedding_names:
            v = rs.standard_normal(size=shape)
        elif name.endswith(".weight"):
            fan_in = int(np.prod(shape[1:]))
            fan_in_of_weight[name[: -len(".weight")]] = fan_in
            b = 1.0 / math.sqrt(fan_in)
            v = rs.uniform(-b, b, size=shape)
        else:  # bias
            fan_in = fan_in_of_weight.get(name[: -len(".bias")], int(shape[0]))
            b = 1.0 / math.sqrt(fan_in)
            v = rs.uniform(-b, b, size=shape)
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return out


# ----------------------------------------------------------------------------
# element-wise pieces
# ----------------------------------------------------------------------------

def swish(x: Tensor) -> Tensor:
    """mnist/model.py:166-169  x * sigmoid(x)."""
    return x * torch.sigmoid(x)


def product_of_experts(mu: Tensor, logvar: Tensor, eps: float = 1e-8, variant: str = "A") -> Tuple[Tensor, Tensor]:
    """[M,B,L] stacks -> fused ([B,L],[B,L]).

    variant "A": mnist/model.py:156-163, fashionmnist/model.py:175-182
        var = exp(lv)+eps ; T = 1/(var+eps) ; pd_logvar = log(pd_var+eps)
    variant "B": celeba/model.py:200-207, celeba19/model.py:219-226
        var = exp(lv)+eps ; T = 1/var       ; pd_logvar = log(pd_var)
    """
    var = torch.exp(logvar) + eps
    T = 1.0 / (var + eps) if variant == "A" else 1.0 / var
    sT = torch.sum(T, dim=0)
    pd_mu = torch.sum(mu * T, dim=0) / sT
    pd_var = 1.0 / sT
    pd_logvar = torch.log(pd_var + eps) if variant == "A" else torch.log(pd_var)
    return pd_mu, pd_logvar


def prior_expert(size, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """mnist/model.py:172-185: N(0,1) prior, mu = 0, logvar = 0 (celeba variant writes
    log(ones) which is the same value, celeba/model.py:225-226)."""
    return torch.zeros(size, dtype=dtype), torch.zeros(size, dtype=dtype)


def bce_with_logits(x: Tensor, t: Tensor) -> Tensor:
    """mnist/train.py:62-74: max(x,0) - x*t + log(1+exp(-|x|)), element-wise."""
    if t.size() != x.size():
        raise ValueError("Target size ({}) must be the same as input size ({})".format(t.size(), x.size()))
    return torch.clamp(x, 0) - x * t + torch.log(1 + torch.exp(-torch.abs(x)))


def cross_entropy_rows(x: Tensor, target: Tensor, eps: float = 1e-6) -> Tensor:
    """mnist/train.py:77-94: -onehot(target) * log_softmax(x+eps) -> [N,K]."""
    if target.size(0) != x.size(0):
        raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(0), x.size(0)))
    logp = torch.log_softmax(x + eps, dim=1)
    onehot = torch.zeros_like(logp).scatter(1, target.unsqueeze(1), 1)
    return -(onehot * logp)


def kl_rows(mu: Tensor, logvar: Tensor) -> Tensor:
    """mnist/train.py:56: -0.5 * sum(1 + lv - mu^2 - exp(lv), dim=1)."""
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1)


def elbo_loss_bimodal(recon_image, image, recon_text, text, mu, logvar,
                      lambda_image=1.0, lambda_text=1.0, annealing_factor=1.0) -> Tensor:
    """mnist/train.py:20-59 (fashionmnist identical).  ``None`` pairs skip a term."""
    B = mu.size(0)
    image_bce = 0
    text_bce = 0
    if recon_image is not None and image is not None:
        image_bce = torch.sum(bce_with_logits(recon_image.reshape(B, -1), image.reshape(B, -1)), dim=1)
    if recon_text is not None and text is not None:
        text_bce = torch.sum(cross_entropy_rows(recon_text, text), dim=1)
    return torch.mean(lambda_image * image_bce + lambda_text * text_bce + annealing_factor * kl_rows(mu, logvar))


def elbo_loss_celeba(recon_image, image, recon_attrs, attrs, mu, logvar,
                     lambda_image=1.0, lambda_attrs=1.0, annealing_factor=1.0) -> Tensor:
    """celeba/train.py:22-65: image BCE summed over pixels; attrs = column loop of
    element-wise BCE accumulated per row (== row sum over the 18 columns)."""
    B = mu.size(0)
    image_bce = 0
    attrs_bce = 0
    if recon_image is not None and image is not None:
        image_bce = torch.sum(bce_with_logits(recon_image.reshape(B, -1), image.reshape(B, -1)), dim=1)
    if recon_attrs is not None and attrs is not None:
        for i in range(attrs.size(1)):
            attrs_bce = attrs_bce + bce_with_logits(recon_attrs[:, i], attrs[:, i])
    return torch.mean(lambda_image * image_bce + lambda_attrs * attrs_bce + annealing_factor * kl_rows(mu, logvar))


def annealing_factor(epoch: int, batch_idx: int, n_mini_batches: int, annealing_epochs: int,
                     flavor: str = "mnist") -> float:
    """mnist/train.py:180-186 ; fashionmnist/train.py:182 uses ``epoch * N`` (no -1)."""
    if epoch < annealing_epochs:
        e = epoch if flavor == "fashionmnist" else (epoch - 1)
        return float(batch_idx + e * n_mini_batches + 1) / float(annealing_epochs * n_mini_batches)
    return 1.0


# ----------------------------------------------------------------------------
# MNIST MVAE (MLP) forward -- mnist/model.py
# ----------------------------------------------------------------------------

def _lin(p: Params, name: str, x: Tensor) -> Tensor:
    return torch.addmm(p[name + ".bias"], x, p[name + ".weight"].t())


def mnist_image_encoder(p: Params, x: Tensor):
    """mnist/model.py:81-84."""
    h = swish(_lin(p, "image_encoder.fc1", x.reshape(-1, 784)))
    h = swish(_lin(p, "image_encoder.fc2", h))
    return _lin(p, "image_encoder.fc31", h), _lin(p, "image_encoder.fc32", h)


def mnist_text_encoder(p: Params, y: Tensor):
    """mnist/model.py:122-125 (fc1 is an Embedding(10,512))."""
    h = swish(p["text_encoder.fc1.weight"][y])
    h = swish(_lin(p, "text_encoder.fc2", h))
    return _lin(p, "text_encoder.fc31", h), _lin(p, "text_encoder.fc32", h)


def mnist_decoder(p: Params, which: str, z: Tensor) -> Tensor:
    """mnist/model.py:101-105 / 142-146: three Swish layers then logits."""
    h = swish(_lin(p, which + ".fc1", z))
    h = swish(_lin(p, which + ".fc2", h))
    h = swish(_lin(p, which + ".fc3", h))
    return _lin(p, which + ".fc4", h)


def mnist_infer(p: Params, image: Optional[Tensor], text: Optional[Tensor], n_latents: int, variant: str = "A"):
    """mnist/model.py:46-64: prior + available experts -> PoE."""
    B = image.size(0) if image is not None else text.size(0)
    dtype = p["image_encoder.fc1.weight"].dtype
    mu, logvar = prior_expert((1, B, n_latents), dtype)
    if image is not None:
        m, lv = mnist_image_encoder(p, image)
        mu = torch.cat((mu, m.unsqueeze(0)), 0)
        logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    if text is not None:
        m, lv = mnist_text_encoder(p, text)
        mu = torch.cat((mu, m.unsqueeze(0)), 0)
        logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    return product_of_experts(mu, logvar, variant=variant)


def reparametrize(mu: Tensor, logvar: Tensor, noise: Optional[Tensor]) -> Tensor:
    """mnist/model.py:29-35: train -> eps*exp(0.5*lv)+mu ; eval (noise None) -> mu."""
    if noise is None:
        return mu
    return noise * torch.exp(0.5 * logvar) + mu


def mnist_forward(p: Params, image, text, n_latents: int, noise: Optional[Tensor]):
    """mnist/model.py:37-44: infer -> reparametrize -> BOTH decoders."""
    mu, logvar = mnist_infer(p, image, text, n_latents)
    z = reparametrize(mu, logvar, noise)
    return mnist_decoder(p, "image_decoder", z), mnist_decoder(p, "text_decoder", z), mu, logvar


def mnist_step_losses(p: Params, image: Tensor, text: Tensor, n_latents: int,
                      noises: Sequence[Optional[Tensor]], lambda_image=1.0, lambda_text=10.0,
                      annealing=1.0):
    """The three-pass objective of mnist/train.py:196-214 (joint, image-only, text-only).
    ``noises`` = the three reparametrisation draws in call order (None -> eval mode)."""
    ri1, rt1, mu1, lv1 = mnist_forward(p, image, text, n_latents, noises[0])
    ri2, rt2, mu2, lv2 = mnist_forward(p, image, None, n_latents, noises[1])
    ri3, rt3, mu3, lv3 = mnist_forward(p, None, text, n_latents, noises[2])
    joint = elbo_loss_bimodal(ri1, image, rt1, text, mu1, lv1, lambda_image, lambda_text, annealing)
    img = elbo_loss_bimodal(ri2, image, None, None, mu2, lv2, lambda_image, lambda_text, annealing)
    txt = elbo_loss_bimodal(None, None, rt3, text, mu3, lv3, lambda_image, lambda_text, annealing)
    aux = {"mu": (mu1, mu2, mu3), "logvar": (lv1, lv2, lv3), "recon_image": (ri1, ri2, ri3),
           "recon_text": (rt1, rt2, rt3)}
    return joint + img + txt, (joint, img, txt), aux


def mnist_step_grads(p: Params, image, text, n_latents, noises, lambda_image=1.0, lambda_text=10.0,
                     annealing=1.0):
    """loss, (joint,image,text) ELBOs and d loss / d param for every parameter
    (mnist/train.py:196-218: zero_grad -> 3 forwards -> 3 ELBOs -> sum -> backward)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss, terms, aux = mnist_step_losses(q, image, text, n_latents, noises, lambda_image, lambda_text, annealing)
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    return loss.detach(), tuple(t.detach() for t in terms), grads, aux


# ----------------------------------------------------------------------------
# FashionMNIST MVAE (conv) forward -- fashionmnist/model.py
# ----------------------------------------------------------------------------

def fashion_param_shapes(n_latents: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict names/shapes of the FashionMNIST MVAE in the reference's registration order
    (fashionmnist/model.py:24-31, 76-87, 104-116, 131-137, 153-161).  Convs have bias=False."""
    L = n_latents
    return [
        ("image_encoder.features.0.weight", (64, 1, 4, 4)),
        ("image_encoder.features.2.weight", (128, 64, 4, 4)),
        ("image_encoder.classifier.0.weight", (512, 6272)), ("image_encoder.classifier.0.bias", (512,)),
        ("image_encoder.classifier.2.weight", (2 * L, 512)), ("image_encoder.classifier.2.bias", (2 * L,)),
        ("image_decoder.upsampler.0.weight", (512, L)), ("image_decoder.upsampler.0.bias", (512,)),
        ("image_decoder.upsampler.2.weight", (6272, 512)), ("image_decoder.upsampler.2.bias", (6272,)),
        ("image_decoder.hallucinate.0.weight", (128, 64, 4, 4)),   # ConvTranspose2d: [Cin, Cout, kh, kw]
        ("image_decoder.hallucinate.2.weight", (64, 1, 4, 4)),
        ("text_encoder.net.0.weight", (10, 512)),
        ("text_encoder.net.2.weight", (512, 512)), ("text_encoder.net.2.bias", (512,)),
        ("text_encoder.net.4.weight", (2 * L, 512)), ("text_encoder.net.4.bias", (2 * L,)),
        ("text_decoder.net.0.weight", (512, L)), ("text_decoder.net.0.bias", (512,)),
        ("text_decoder.net.2.weight", (512, 512)), ("text_decoder.net.2.bias", (512,)),
        ("text_decoder.net.4.weight", (512, 512)), ("text_decoder.net.4.bias", (512,)),
        ("text_decoder.net.6.weight", (10, 512)), ("text_decoder.net.6.bias", (10,)),
    ]


FASHION_EMBEDDINGS = ("text_encoder.net.0.weight",)


def fashion_image_encoder(p: Params, x: Tensor, L: int):
    """fashionmnist/model.py:89-94: conv(1->64) Swish conv(64->128) Swish, flatten (C,H,W), FC 6272->512 Swish, FC -> 2L."""
    F = torch.nn.functional
    h = swish(F.conv2d(x.reshape(-1, 1, 28, 28), p["image_encoder.features.0.weight"], None, 2, 1))
    h = swish(F.conv2d(h, p["image_encoder.features.2.weight"], None, 2, 1))
    h = swish(_lin(p, "image_encoder.classifier.0", h.reshape(h.size(0), -1)))
    o = _lin(p, "image_encoder.classifier.2", h)
    return o[:, :L], o[:, L:]


def fashion_image_decoder(p: Params, z: Tensor):
    """fashionmnist/model.py:117-121: FC L->512 Swish, FC 512->6272 Swish, view(128,7,7), convT(128->64) Swish,
    convT(64->1) -> logits [B,1,28,28]."""
    F = torch.nn.functional
    h = swish(_lin(p, "image_decoder.upsampler.0", z))
    h = swish(_lin(p, "image_decoder.upsampler.2", h)).reshape(-1, 128, 7, 7)
    h = swish(F.conv_transpose2d(h, p["image_decoder.hallucinate.0.weight"], None, 2, 1))
    return F.conv_transpose2d(h, p["image_decoder.hallucinate.2.weight"], None, 2, 1)


def fashion_text_encoder(p: Params, y: Tensor, L: int):
    """fashionmnist/model.py:139-143."""
    h = swish(p["text_encoder.net.0.weight"][y])
    h = swish(_lin(p, "text_encoder.net.2", h))
    o = _lin(p, "text_encoder.net.4", h)
    return o[:, :L], o[:, L:]


def fashion_text_decoder(p: Params, z: Tensor):
    """fashionmnist/model.py:163-165."""
    h = swish(_lin(p, "text_decoder.net.0", z))
    h = swish(_lin(p, "text_decoder.net.2", h))
    h = swish(_lin(p, "text_decoder.net.4", h))
    return _lin(p, "text_decoder.net.6", h)


def fashion_forward(p: Params, image, text, L: int, noise: Optional[Tensor]):
    """fashionmnist/model.py:41-67 (same control flow as mnist)."""
    B = image.size(0) if image is not None else text.size(0)
    dtype = p["image_encoder.features.0.weight"].dtype
    mu, logvar = prior_expert((1, B, L), dtype)
    if image is not None:
        m, lv = fashion_image_encoder(p, image, L)
        mu = torch.cat((mu, m.unsqueeze(0)), 0); logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    if text is not None:
        m, lv = fashion_text_encoder(p, text, L)
        mu = torch.cat((mu, m.unsqueeze(0)), 0); logvar = torch.cat((logvar, lv.unsqueeze(0)), 0)
    mu, logvar = product_of_experts(mu, logvar, variant="A")
    z = reparametrize(mu, logvar, noise)
    return fashion_image_decoder(p, z), fashion_text_decoder(p, z), mu, logvar


def fashion_step_grads(p: Params, image, text, L, noises, lambda_image=1.0, lambda_text=10.0, annealing=1.0):
    """Three-pass objective + gradients of fashionmnist/train.py:196-218 (identical to mnist's)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    r1 = fashion_forward(q, image, text, L, noises[0])
    r2 = fashion_forward(q, image, None, L, noises[1])
    r3 = fashion_forward(q, None, text, L, noises[2])
    joint = elbo_loss_bimodal(r1[0], image, r1[1], text, r1[2], r1[3], lambda_image, lambda_text, annealing)
    img = elbo_loss_bimodal(r2[0], image, None, None, r2[2], r2[3], lambda_image, lambda_text, annealing)
    txt = elbo_loss_bimodal(None, None, r3[1], text, r3[2], r3[3], lambda_image, lambda_text, annealing)
    loss = joint + img + txt
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    aux = {"mu": (r1[2], r2[2], r3[2]), "logvar": (r1[3], r2[3], r3[3]), "recon_image": (r1[0], r2[0], r3[0]),
           "recon_text": (r1[1], r2[1], r3[1])}
    return loss.detach(), (joint.detach(), img.detach(), txt.detach()), grads, aux


# ----------------------------------------------------------------------------
# Adam -- torch.optim.Adam defaults as used at mnist/train.py:168,219
# ----------------------------------------------------------------------------

def adam_update(params: Params, grads: Params, state: Dict[str, Dict[str, Tensor]], step: int,
                lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8) -> None:
    """In-place Adam (no weight decay, no amsgrad); ``step`` is the 1-based step count."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for k, w in params.items():
        g = grads[k]
        st = state.setdefault(k, {"m": torch.zeros_like(w), "v": torch.zeros_like(w)})
        st["m"].mul_(beta1).add_(g, alpha=1 - beta1)
        st["v"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        w.addcdiv_(st["m"], denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------
# CPU baseline step (timed by bench.py as cpu_baseline / --impl reference, kind "port")
# ----------------------------------------------------------------------------

class MnistCpuBaseline:
    """The reference step body (mnist/train.py:196-219) on CPU torch with the same ATen ops the
    reference issues: 3 forwards, 3 ELBOs, backward, ``optim.Adam.step``."""

    def __init__(self, n_latents=64, seed=0, lr=1e-3):
        self.L = n_latents
        self.p = {k: v.requires_grad_(True) for k, v in make_params(mnist_param_shapes(n_latents), seed).items()}
        self.opt = torch.optim.Adam(list(self.p.values()), lr=lr)

    def step(self, image: Tensor, text: Tensor, lambda_image=1.0, lambda_text=10.0, annealing=0.5) -> float:
        B = image.size(0)
        self.opt.zero_grad()
        noises = [torch.empty(B, self.L).normal_() for _ in range(3)]
        loss, _, _ = mnist_step_losses(self.p, image, text, self.L, noises, lambda_image, lambda_text, annealing)
        v = float(loss.detach())
        loss.backward()
        self.opt.step()
        return v
