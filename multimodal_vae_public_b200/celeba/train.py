"""``celeba/train.py`` surface: elbo_loss with the (image, attrs) signature (celeba/train.py:22-65),
binary_cross_entropy_with_logits (:68-80), AverageMeter, save/load_checkpoint -- fused CUDA kernels underneath."""
from __future__ import annotations

import torch

from .. import functional as F
from ..mnist.train import AverageMeter, binary_cross_entropy_with_logits, save_checkpoint  # noqa: F401
from .model import MVAE, N_ATTRS  # noqa: F401


def elbo_loss(recon_image, image, recon_attrs, attrs, mu, logvar, lambda_image=1.0, lambda_attrs=1.0, annealing_factor=1):
    """mean_b [ lambda_image * sum_pix BCE + lambda_attrs * sum_18 BCE + annealing_factor * KL ]; None pairs drop a term."""
    B = mu.size(0)
    total = annealing_factor * F.kl_sum(mu, logvar)
    if recon_image is not None and image is not None:
        total = total + lambda_image * F.bce_with_logits_sum(recon_image.reshape(B, -1), image.reshape(B, -1))
    if recon_attrs is not None and attrs is not None:
        total = total + lambda_attrs * F.bce_with_logits_sum(recon_attrs, attrs.to(torch.float32))
    return total / B


def load_checkpoint(file_path, use_cuda=False):
    ckpt = torch.load(file_path, map_location=None if use_cuda else "cpu")
    model = MVAE(ckpt["n_latents"])
    model.load_state_dict(ckpt["state_dict"])
    return model.cuda() if use_cuda else model
