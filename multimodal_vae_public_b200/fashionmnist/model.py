"""B200-native MVAE for FashionMNIST image + label: same class names, ``nn.Sequential`` sub-module names and
``state_dict`` keys, signatures and return tuples as the reference's ``fashionmnist/model.py`` (MVAE :19-67,
ImageEncoder :70-94, ImageDecoder :97-121, TextEncoder :124-143, TextDecoder :146-165, ProductOfExperts :168-182,
Swish, prior_expert), computed by the sm_100a kernels of libmvae_b200.so (conv = im2col/col2im + tcgen05 GEMM).
For throughput use ``trainer_fashion.FashionMVAETrainer``.
"""
from __future__ import annotations

import torch.nn as nn

from .. import functional as F
from ..mnist.model import ProductOfExperts, Swish, prior_expert  # same formulas (variant "A")  # noqa: F401

LABEL_IX_TO_STRING = {0: 'T-shirt/top', 1: 'Trouser', 2: 'Pullover', 3: 'Dress', 4: 'Coat', 5: 'Sandal', 6: 'Shirt',
                      7: 'Sneaker', 8: 'Bag', 9: 'Ankle boot'}


class ImageEncoder(nn.Module):
    """q(z|x): conv 1->64->128 (k4 s2 p1, no bias) + Swish, FC 6272->512 + Swish, FC 512->2*n_latents."""

    def __init__(self, n_latents):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(1, 64, 4, 2, 1, bias=False), Swish(),
                                      nn.Conv2d(64, 128, 4, 2, 1, bias=False), Swish())
        self.classifier = nn.Sequential(nn.Linear(128 * 7 * 7, 512), Swish(), nn.Linear(512, n_latents * 2))
        self.n_latents = n_latents

    def forward(self, x):
        n = self.n_latents
        h = F.conv4x4s2(x.reshape(-1, 1, 28, 28), self.features[0].weight, swish_act=True)
        h = F.conv4x4s2(h, self.features[2].weight, swish_act=True)
        h = F.linear_swish(h.reshape(h.size(0), -1), self.classifier[0].weight, self.classifier[0].bias)
        o = F.linear(h, self.classifier[2].weight, self.classifier[2].bias)
        return o[:, :n], o[:, n:]


class ImageDecoder(nn.Module):
    """p(x|z): FC n_latents->512->6272 + Swish, convT 128->64->1 (k4 s2 p1, no bias); logits [B,1,28,28]."""

    def __init__(self, n_latents):
        super().__init__()
        self.n_latents = n_latents
        self.upsampler = nn.Sequential(nn.Linear(n_latents, 512), Swish(), nn.Linear(512, 128 * 7 * 7), Swish())
        self.hallucinate = nn.Sequential(nn.ConvTranspose2d(128, 64, 4, 2, 1, bias=False), Swish(),
                                         nn.ConvTranspose2d(64, 1, 4, 2, 1, bias=False))

    def forward(self, z):
        h = F.linear_swish(z, self.upsampler[0].weight, self.upsampler[0].bias)
        h = F.linear_swish(h, self.upsampler[2].weight, self.upsampler[2].bias).reshape(-1, 128, 7, 7)
        h = F.conv_transpose4x4s2(h, self.hallucinate[0].weight, swish_act=True)
        return F.conv_transpose4x4s2(h, self.hallucinate[2].weight)


class TextEncoder(nn.Module):
    """q(z|y): Embedding(10,512) + Swish, 512->512 + Swish, 512->2*n_latents."""

    def __init__(self, n_latents):
        super().__init__()
        self.net = nn.Sequential(nn.Embedding(10, 512), Swish(), nn.Linear(512, 512), Swish(),
                                 nn.Linear(512, n_latents * 2))
        self.n_latents = n_latents

    def forward(self, x):
        n = self.n_latents
        h = F.embedding_swish(x, self.net[0].weight)
        h = F.linear_swish(h, self.net[2].weight, self.net[2].bias)
        o = F.linear(h, self.net[4].weight, self.net[4].bias)
        return o[:, :n], o[:, n:]


class TextDecoder(nn.Module):
    """p(y|z): n_latents->512->512->512->10 logits."""

    def __init__(self, n_latents):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(n_latents, 512), Swish(), nn.Linear(512, 512), Swish(),
                                 nn.Linear(512, 512), Swish(), nn.Linear(512, 10))

    def forward(self, z):
        h = F.linear_swish(z, self.net[0].weight, self.net[0].bias)
        h = F.linear_swish(h, self.net[2].weight, self.net[2].bias)
        h = F.linear_swish(h, self.net[4].weight, self.net[4].bias)
        return F.linear(h, self.net[6].weight, self.net[6].bias)


class MVAE(nn.Module):
    """``forward(image=None, text=None) -> (img_recon [B,1,28,28], txt_recon [B,10], mu, logvar)``."""

    def __init__(self, n_latents):
        super().__init__()
        self.image_encoder = ImageEncoder(n_latents)
        self.image_decoder = ImageDecoder(n_latents)
        self.text_encoder = TextEncoder(n_latents)
        self.text_decoder = TextDecoder(n_latents)
        self.experts = ProductOfExperts()
        self.n_latents = n_latents

    def reparametrize(self, mu, logvar):
        return F.reparametrize(mu, logvar) if self.training else mu

    def forward(self, image=None, text=None):
        mu, logvar = self.infer(image, text)
        z = self.reparametrize(mu, logvar)
        return self.image_decoder(z), self.text_decoder(z), mu, logvar

    def infer(self, image=None, text=None):
        if image is None and text is None:
            raise ValueError("at least one modality is required")
        mus, lvs = [], []
        if image is not None:
            m, lv = self.image_encoder(image); mus.append(m); lvs.append(lv)
        if text is not None:
            m, lv = self.text_encoder(text); mus.append(m); lvs.append(lv)
        return F.product_of_experts(mus, lvs, variant=self.experts.variant, with_prior=True)
