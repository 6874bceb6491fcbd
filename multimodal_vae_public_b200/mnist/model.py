"""B200-native MVAE for MNIST image + label -- same classes, constructor arguments, attribute names,
``state_dict`` keys, call signatures and return tuples as the reference's ``mnist/model.py``
(MVAE :14-64, ImageEncoder :67-84, ImageDecoder :87-105, TextEncoder :108-125, TextDecoder :128-146,
ProductOfExperts :149-163, Swish :166-169, prior_expert :172-185), but every forward/backward op runs
in hand-written sm_100a kernels (libmvae_b200.so) through ``multimodal_vae_public_b200.functional``.

For throughput use ``multimodal_vae_public_b200.trainer.MnistMVAETrainer`` (whole step fused and graph
captured); these modules exist so that code written against the reference (``model(image, text)``,
``model.infer``, ``loss.backward()``, ``load_state_dict``) keeps working unchanged on the GPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F


class Swish(nn.Module):
    """x * sigmoid(x)."""

    def forward(self, x):
        return F.swish(x)


class ProductOfExperts(nn.Module):
    """Parameters of the product of M independent Gaussian experts (stack along dim 0).

    @param mu: M x B x D ; @param logvar: M x B x D  ->  (B x D, B x D).  eps must be the reference's 1e-8."""

    variant = 0  # 0 = mnist/fashionmnist formula, 1 = celeba formula

    def forward(self, mu, logvar, eps=1e-8):
        if abs(eps - 1e-8) > 1e-20:
            raise ValueError("the fused kernel implements the reference's eps=1e-8 only")
        M = mu.size(0)
        return F.product_of_experts([mu[i] for i in range(M)], [logvar[i] for i in range(M)], variant=self.variant,
                                    with_prior=False)


def prior_expert(size, use_cuda=False):
    """Universal N(0,1) prior expert: mu = 0, logvar = 0 of the requested size."""
    dev = "cuda" if use_cuda else "cpu"
    return torch.zeros(size, device=dev), torch.zeros(size, device=dev)


class ImageEncoder(nn.Module):
    """q(z|x): 784 -> 512 -> 512 -> 2 x n_latents."""

    def __init__(self, n_latents):
        super().__init__()
        self.fc1 = nn.Linear(784, 512)
        self.fc2 = nn.Linear(512, 512)
        self.fc31 = nn.Linear(512, n_latents)
        self.fc32 = nn.Linear(512, n_latents)
        self.swish = Swish()

    def forward(self, x):
        h = F.linear_swish(x.reshape(-1, 784), self.fc1.weight, self.fc1.bias)
        h = F.linear_swish(h, self.fc2.weight, self.fc2.bias)
        return F.linear(h, self.fc31.weight, self.fc31.bias), F.linear(h, self.fc32.weight, self.fc32.bias)


class _Decoder(nn.Module):
    def __init__(self, n_latents, n_out):
        super().__init__()
        self.fc1 = nn.Linear(n_latents, 512)
        self.fc2 = nn.Linear(512, 512)
        self.fc3 = nn.Linear(512, 512)
        self.fc4 = nn.Linear(512, n_out)
        self.swish = Swish()

    def forward(self, z):
        h = F.linear_swish(z, self.fc1.weight, self.fc1.bias)
        h = F.linear_swish(h, self.fc2.weight, self.fc2.bias)
        h = F.linear_swish(h, self.fc3.weight, self.fc3.bias)
        return F.linear(h, self.fc4.weight, self.fc4.bias)  # logits: no sigmoid / softmax here


class ImageDecoder(_Decoder):
    """p(x|z): n_latents -> 512 -> 512 -> 512 -> 784 logits."""

    def __init__(self, n_latents):
        super().__init__(n_latents, 784)


class TextEncoder(nn.Module):
    """q(z|y): Embedding(10,512) -> 512 -> 2 x n_latents."""

    def __init__(self, n_latents):
        super().__init__()
        self.fc1 = nn.Embedding(10, 512)
        self.fc2 = nn.Linear(512, 512)
        self.fc31 = nn.Linear(512, n_latents)
        self.fc32 = nn.Linear(512, n_latents)
        self.swish = Swish()

    def forward(self, x):
        h = F.embedding_swish(x, self.fc1.weight)
        h = F.linear_swish(h, self.fc2.weight, self.fc2.bias)
        return F.linear(h, self.fc31.weight, self.fc31.bias), F.linear(h, self.fc32.weight, self.fc32.bias)


class TextDecoder(_Decoder):
    """p(y|z): n_latents -> 512 -> 512 -> 512 -> 10 logits."""

    def __init__(self, n_latents):
        super().__init__(n_latents, 10)


class MVAE(nn.Module):
    """Multimodal VAE.  ``forward(image=None, text=None) -> (img_recon, txt_recon, mu, logvar)``."""

    def __init__(self, n_latents):
        super().__init__()
        self.image_encoder = ImageEncoder(n_latents)
        self.image_decoder = ImageDecoder(n_latents)
        self.text_encoder = TextEncoder(n_latents)
        self.text_decoder = TextDecoder(n_latents)
        self.experts = ProductOfExperts()
        self.n_latents = n_latents

    def reparametrize(self, mu, logvar):
        if self.training:
            return F.reparametrize(mu, logvar)
        return mu

    def forward(self, image=None, text=None):
        mu, logvar = self.infer(image, text)
        z = self.reparametrize(mu, logvar)
        return self.image_decoder(z), self.text_decoder(z), mu, logvar

    def infer(self, image=None, text=None):
        if image is None and text is None:
            raise ValueError("at least one modality is required")
        mus, lvs = [], []
        if image is not None:
            m, lv = self.image_encoder(image)
            mus.append(m); lvs.append(lv)
        if text is not None:
            m, lv = self.text_encoder(text)
            mus.append(m); lvs.append(lv)
        # prior expert N(0,1) is folded into the kernel (never materialised / concatenated)
        return F.product_of_experts(mus, lvs, variant=self.experts.variant, with_prior=True)
