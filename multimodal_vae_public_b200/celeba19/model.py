"""B200-native CelebA-19 MVAE: image + one inference/generative network PER attribute (18), same class names,
``state_dict`` keys, signatures and return values as the reference's ``celeba19/model.py`` (MVAE :14-89 with
``attr_encoders`` / ``attr_decoders`` ModuleLists, AttributeEncoder :160-183 [Embedding(2,512) -> 512 -> 2L],
AttributeDecoder :186-209 [L -> 512 -> 512 -> 512 -> 1], ProductOfExperts :212-225), computed by libmvae_b200.so.
The image networks are the CelebA DCGAN ones.  For throughput use ``trainer_celeba19.CelebA19MVAETrainer``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F
from ..celeba.model import ImageDecoder, ImageEncoder, N_ATTRS, ProductOfExperts, Swish, prior_expert  # noqa: F401


class AttributeEncoder(nn.Module):
    """q(z|y_i) for ONE binary attribute: Embedding(2,512), Swish, Linear(512,512), Swish, Linear(512, 2L)."""

    def __init__(self, n_latents):
        super().__init__()
        self.net = nn.Sequential(nn.Embedding(2, 512), Swish(), nn.Linear(512, 512), Swish(),
                                 nn.Linear(512, n_latents * 2))
        self.n_latents = n_latents

    def forward(self, x):
        n, s = self.n_latents, self.net
        h = F.embedding_swish(x.long(), s[0].weight)
        h = F.linear_swish(h, s[2].weight, s[2].bias)
        o = F.linear(h, s[4].weight, s[4].bias)
        return o[:, :n], o[:, n:]


class AttributeDecoder(nn.Module):
    """p(y_i|z): L -> 512 -> 512 -> 512 -> 1 logit (no sigmoid)."""

    def __init__(self, n_latents):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(n_latents, 512), Swish(), nn.Linear(512, 512), Swish(),
                                 nn.Linear(512, 512), Swish(), nn.Linear(512, 1))

    def forward(self, z):
        s = self.net
        h = F.linear_swish(z, s[0].weight, s[0].bias)
        h = F.linear_swish(h, s[2].weight, s[2].bias)
        h = F.linear_swish(h, s[4].weight, s[4].bias)
        return F.linear(h, s[6].weight, s[6].bias)


class MVAE(nn.Module):
    """``forward(image=None, attrs=[None]*18) -> (image_recon, [18 x attr_recon [B]], mu, logvar)``."""

    def __init__(self, n_latents):
        super().__init__()
        self.image_encoder = ImageEncoder(n_latents)
        self.image_decoder = ImageDecoder(n_latents)
        self.attr_encoders = nn.ModuleList([AttributeEncoder(n_latents) for _ in range(N_ATTRS)])
        self.attr_decoders = nn.ModuleList([AttributeDecoder(n_latents) for _ in range(N_ATTRS)])
        self.experts = ProductOfExperts()
        self.n_latents = n_latents

    def reparametrize(self, mu, logvar):
        return F.reparametrize(mu, logvar) if self.training else mu

    def forward(self, image=None, attrs=None):
        attrs = [None] * N_ATTRS if attrs is None else attrs
        mu, logvar = self.infer(image, attrs)
        z = self.reparametrize(mu, logvar)
        image_recon = self.image_decoder(z)
        attr_recons = [self.attr_decoders[i](z).squeeze(1) for i in range(N_ATTRS)]
        return image_recon, attr_recons, mu, logvar

    def infer(self, image=None, attrs=None):
        attrs = [None] * N_ATTRS if attrs is None else attrs
        if image is None and all(a is None for a in attrs):
            raise ValueError("at least one modality is required")
        mus, lvs = [], []
        if image is not None:
            m, lv = self.image_encoder(image); mus.append(m); lvs.append(lv)
        for i in range(N_ATTRS):
            if attrs[i] is not None:
                m, lv = self.attr_encoders[i](attrs[i]); mus.append(m); lvs.append(lv)
        return F.product_of_experts(mus, lvs, variant=self.experts.variant, with_prior=True)
