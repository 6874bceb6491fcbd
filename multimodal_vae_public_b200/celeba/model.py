"""B200-native MVAE for CelebA image (3x64x64) + 18 binary attributes: same class names, ``nn.Sequential`` layouts /
``state_dict`` keys (including BatchNorm buffers), signatures and return tuples as the reference's ``celeba/model.py``
(MVAE :13-63, ImageEncoder :66-100, ImageDecoder :103-133, AttributeEncoder :136-160, AttributeDecoder :163-190,
ProductOfExperts :193-207 [no extra eps], Swish, prior_expert), computed by libmvae_b200.so: conv = im2col/col2im +
tcgen05 GEMM on channels-last activations, BatchNorm/Dropout/PoE/reparametrise in fused element-wise kernels.
For throughput use ``trainer_celeba.CelebAMVAETrainer``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F
from ..mnist.model import Swish, prior_expert  # noqa: F401

N_ATTRS = 18


class ProductOfExperts(nn.Module):
    """celeba variant: var = exp(logvar) + eps; T = 1/var; logvar_out = log(1/sum T)."""

    variant = 1

    def forward(self, mu, logvar, eps=1e-8):
        if abs(eps - 1e-8) > 1e-20:
            raise ValueError("the fused kernel implements the reference's eps=1e-8 only")
        M = mu.size(0)
        return F.product_of_experts([mu[i] for i in range(M)], [logvar[i] for i in range(M)], variant=self.variant,
                                    with_prior=False)


def _bn(x, bn: nn.modules.batchnorm._BatchNorm, swish_act=True):
    if bn.training:
        bn.num_batches_tracked += 1
    return F.batch_norm_act(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.training, swish_act)


class ImageEncoder(nn.Module):
    """q(z|x), DCGAN encoder: conv 3->32->64->128 (k4 s2 p1) -> 256 (k4 s1 p0), BatchNorm + Swish, FC, Dropout, FC."""

    def __init__(self, n_latents):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(3, 32, 4, 2, 1, bias=False), Swish(),
            nn.Conv2d(32, 64, 4, 2, 1, bias=False), nn.BatchNorm2d(64), Swish(),
            nn.Conv2d(64, 128, 4, 2, 1, bias=False), nn.BatchNorm2d(128), Swish(),
            nn.Conv2d(128, 256, 4, 1, 0, bias=False), nn.BatchNorm2d(256), Swish())
        self.classifier = nn.Sequential(nn.Linear(256 * 5 * 5, 512), Swish(), nn.Dropout(p=0.1),
                                        nn.Linear(512, n_latents * 2))
        self.n_latents = n_latents

    def forward(self, x):
        n, f, c = self.n_latents, self.features, self.classifier
        h = x.reshape(-1, 3, 64, 64).permute(0, 2, 3, 1)                       # channels-last from here on
        h = F.conv4x4(h, f[0].weight, 2, 1, swish_act=True, nhwc=True)
        h = _bn(F.conv4x4(h, f[2].weight, 2, 1, nhwc=True), f[3])
        h = _bn(F.conv4x4(h, f[5].weight, 2, 1, nhwc=True), f[6])
        h = _bn(F.conv4x4(h, f[8].weight, 1, 0, nhwc=True), f[9])              # [B,5,5,256]
        h = h.permute(0, 3, 1, 2).reshape(-1, 256 * 5 * 5)                      # reference flatten order (c,h,w)
        h = F.linear_swish(h, c[0].weight, c[0].bias)
        h = F.dropout(h, c[2].p, self.training)
        o = F.linear(h, c[3].weight, c[3].bias)
        return o[:, :n], o[:, n:]


class ImageDecoder(nn.Module):
    """p(x|z), DCGAN decoder: FC, convT 256->128 (k4 s1 p0) ->64->32->3 (k4 s2 p1), BatchNorm + Swish; logits."""

    def __init__(self, n_latents):
        super().__init__()
        self.upsample = nn.Sequential(nn.Linear(n_latents, 256 * 5 * 5), Swish())
        self.hallucinate = nn.Sequential(
            nn.ConvTranspose2d(256, 128, 4, 1, 0, bias=False), nn.BatchNorm2d(128), Swish(),
            nn.ConvTranspose2d(128, 64, 4, 2, 1, bias=False), nn.BatchNorm2d(64), Swish(),
            nn.ConvTranspose2d(64, 32, 4, 2, 1, bias=False), nn.BatchNorm2d(32), Swish(),
            nn.ConvTranspose2d(32, 3, 4, 2, 1, bias=False))

    def forward(self, z):
        u, d = self.upsample, self.hallucinate
        h = F.linear_swish(z, u[0].weight, u[0].bias).reshape(-1, 256, 5, 5).permute(0, 2, 3, 1)
        h = _bn(F.conv_transpose4x4(h, d[0].weight, 1, 0, nhwc=True), d[1])
        h = _bn(F.conv_transpose4x4(h, d[3].weight, 2, 1, nhwc=True), d[4])
        h = _bn(F.conv_transpose4x4(h, d[6].weight, 2, 1, nhwc=True), d[7])
        return F.conv_transpose4x4(h, d[9].weight, 2, 1, nhwc=True).permute(0, 3, 1, 2)   # NCHW logits


class AttributeEncoder(nn.Module):
    """q(z|y): 18 -> 512 -> 512 -> 2*n_latents with BatchNorm1d + Swish."""

    def __init__(self, n_latents):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(N_ATTRS, 512), nn.BatchNorm1d(512), Swish(),
                                 nn.Linear(512, 512), nn.BatchNorm1d(512), Swish(), nn.Linear(512, n_latents * 2))
        self.n_latents = n_latents

    def forward(self, x):
        n, s = self.n_latents, self.net
        h = _bn(F.linear(x.to(torch.float32), s[0].weight, s[0].bias), s[1])
        h = _bn(F.linear(h, s[3].weight, s[3].bias), s[4])
        o = F.linear(h, s[6].weight, s[6].bias)
        return o[:, :n], o[:, n:]


class AttributeDecoder(nn.Module):
    """p(y|z): n_latents -> 512 -> 512 -> 512 -> 18 logits with BatchNorm1d + Swish."""

    def __init__(self, n_latents):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(n_latents, 512), nn.BatchNorm1d(512), Swish(),
                                 nn.Linear(512, 512), nn.BatchNorm1d(512), Swish(),
                                 nn.Linear(512, 512), nn.BatchNorm1d(512), Swish(), nn.Linear(512, N_ATTRS))

    def forward(self, z):
        s = self.net
        h = _bn(F.linear(z, s[0].weight, s[0].bias), s[1])
        h = _bn(F.linear(h, s[3].weight, s[3].bias), s[4])
        h = _bn(F.linear(h, s[6].weight, s[6].bias), s[7])
        return F.linear(h, s[9].weight, s[9].bias)


class MVAE(nn.Module):
    """``forward(image=None, attrs=None) -> (image_recon [B,3,64,64], attrs_recon [B,18], mu, logvar)``."""

    def __init__(self, n_latents):
        super().__init__()
        self.image_encoder = ImageEncoder(n_latents)
        self.image_decoder = ImageDecoder(n_latents)
        self.attrs_encoder = AttributeEncoder(n_latents)
        self.attrs_decoder = AttributeDecoder(n_latents)
        self.experts = ProductOfExperts()
        self.n_latents = n_latents

    def reparametrize(self, mu, logvar):
        return F.reparametrize(mu, logvar) if self.training else mu

    def forward(self, image=None, attrs=None):
        mu, logvar = self.infer(image, attrs)
        z = self.reparametrize(mu, logvar)
        return self.image_decoder(z), self.attrs_decoder(z), mu, logvar

    def infer(self, image=None, attrs=None):
        if image is None and attrs is None:
            raise ValueError("at least one modality is required")
        mus, lvs = [], []
        if image is not None:
            m, lv = self.image_encoder(image); mus.append(m); lvs.append(lv)
        if attrs is not None:
            m, lv = self.attrs_encoder(attrs); mus.append(m); lvs.append(lv)
        return F.product_of_experts(mus, lvs, variant=self.experts.variant, with_prior=True)
