"""``celeba19/train.py`` surface: list-based elbo_loss (celeba19/train.py:26-60), binary_cross_entropy_with_logits
(:63-76), tensor_2d_to_list (:79-85), enumerate_combinations / sample_combinations (:88-142), AverageMeter,
save/load_checkpoint -- losses by fused CUDA kernels; the subset sampler is host code that consumes numpy's global RNG
exactly like the reference (same draws, same subsets) without materialising the 524,267 x 19 pool."""
from __future__ import annotations

from math import comb

import numpy as np
import torch

from .. import functional as F
from ..mnist.train import AverageMeter, binary_cross_entropy_with_logits, save_checkpoint  # noqa: F401
from ..trainer_celeba19 import sample_combinations as _sample_by_size, unrank_combination
from .model import MVAE, N_ATTRS  # noqa: F401


def elbo_loss(recon, data, mu, logvar, lambda_image=1.0, lambda_attrs=1.0, annealing_factor=1.0):
    """ELBO over an arbitrary list of modalities: a >1-D entry is an image (BCE summed over pixels, lambda_image), a
    1-D entry is one attribute (element-wise BCE, lambda_attrs); mean over the batch of [BCE + annealing * KL]."""
    assert len(recon) == len(data), "must supply ground truth for every modality."
    B = mu.size(0)
    total = annealing_factor * F.kl_sum(mu, logvar)
    for r, d in zip(recon, data):
        if r.dim() > 1:
            total = total + lambda_image * F.bce_with_logits_sum(r.reshape(B, -1), d.reshape(B, -1))
        else:
            total = total + lambda_attrs * F.bce_with_logits_sum(r.reshape(B, 1), d.reshape(B, 1).to(torch.float32))
    return total / B


def tensor_2d_to_list(x):
    return [x[:, i] for i in range(x.size(1))]


class CombinationPool:
    """Lazy stand-in for the reference's boolean [524267, 19] pool: all subsets of n modalities with 2..n-1 members,
    ordered by size and then lexicographically; rows are materialised on demand."""

    def __init__(self, n: int):
        self.n = n
        self.sizes = list(range(2, n))
        self.counts = [comb(n, k) for k in self.sizes]
        self.shape = (sum(self.counts), n)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx: int) -> np.ndarray:
        idx = int(idx)
        if idx < 0:
            idx += len(self)
        for k, c in zip(self.sizes, self.counts):
            if idx < c:
                row = np.zeros(self.n, dtype=bool)
                row[unrank_combination(self.n, k, idx)] = True
                return row
            idx -= c
        raise IndexError("combination index out of range")


def enumerate_combinations(n):
    return CombinationPool(n)


def sample_combinations(pool, size=1):
    """``size`` subsets: the subset SIZE is uniform over the sizes present in the pool, then a subset of that size is
    drawn without replacement.  ``pool`` is a ``CombinationPool`` (O(1) memory) or an explicit boolean array."""
    if isinstance(pool, CombinationPool):
        return _sample_by_size(pool.n, size, np.random)
    pool = np.asarray(pool).astype(bool)
    n = pool.shape[1]
    sums = pool.sum(axis=1)
    present = np.flatnonzero(np.bincount(sums))
    drawn = np.bincount(np.random.choice(present, size, replace=True), minlength=n)
    picked = []
    for k in range(n):
        if drawn[k] > 0:
            rows = pool[sums == k]
            picked.append(rows[np.random.choice(range(rows.shape[0]), size=drawn[k], replace=False)])
    return np.concatenate(picked)


def load_checkpoint(file_path, use_cuda=False):
    ckpt = torch.load(file_path, map_location=None if use_cuda else "cpu")
    model = MVAE(ckpt["n_latents"])
    model.load_state_dict(ckpt["state_dict"])
    return model.cuda() if use_cuda else model
