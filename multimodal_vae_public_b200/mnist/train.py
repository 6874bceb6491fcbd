"""Loss functions and helpers with the reference's ``mnist/train.py`` names and signatures
(elbo_loss :20-59, binary_cross_entropy_with_logits :62-74, cross_entropy :77-94, AverageMeter :97-112,
save_checkpoint/load_checkpoint :115-129), evaluated by fused CUDA kernels (forward + analytic gradient).
"""
from __future__ import annotations

import os
import shutil

import torch

from .. import functional as F
from .model import MVAE


def elbo_loss(recon_image, image, recon_text, text, mu, logvar, lambda_image=1.0, lambda_text=1.0, annealing_factor=1):
    """Bimodal ELBO: mean_b [ lambda_image * sum_pix BCE + lambda_text * CE + annealing_factor * KL ].
    A ``None`` reconstruction/target pair drops that term."""
    B = mu.size(0)
    total = annealing_factor * F.kl_sum(mu, logvar)
    if recon_image is not None and image is not None:
        total = total + lambda_image * F.bce_with_logits_sum(recon_image.reshape(-1, 784), image.reshape(-1, 784))
    if recon_text is not None and text is not None:
        total = total + lambda_text * F.cross_entropy_sum(recon_text, text)
    return total / B


def binary_cross_entropy_with_logits(input, target):
    """Element-wise sigmoid + BCE (same-shape output).  Raises ValueError on a shape mismatch."""
    if not (target.size() == input.size()):
        raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(), input.size()))
    return F.bce_with_logits(input, target)


def cross_entropy(input, target, eps=1e-6):
    """-onehot(target) * log_softmax(input + eps) -> [N, K] (caller sums dim 1)."""
    if not (target.size(0) == input.size(0)):
        raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(0), input.size(0)))
    if abs(eps - 1e-6) > 1e-12:
        raise ValueError("the fused kernel implements the reference's eps=1e-6 only")
    return F.cross_entropy_rows(input, target)


class AverageMeter(object):
    """Running average of a scalar."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def save_checkpoint(state, is_best, folder="./", filename="checkpoint.pth.tar"):
    os.makedirs(folder, exist_ok=True)
    torch.save(state, os.path.join(folder, filename))
    if is_best:
        shutil.copyfile(os.path.join(folder, filename), os.path.join(folder, "model_best.pth.tar"))


def load_checkpoint(file_path, use_cuda=False):
    ckpt = torch.load(file_path, map_location=None if use_cuda else "cpu")
    model = MVAE(ckpt["n_latents"])
    model.load_state_dict(ckpt["state_dict"])
    return model.cuda() if use_cuda else model
