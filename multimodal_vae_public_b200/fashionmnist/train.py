"""``fashionmnist/train.py`` surface: identical loss functions to mnist (fashionmnist/train.py:20-94); the only
difference in the script is the annealing schedule (``epoch * N`` instead of ``(epoch - 1) * N``, :182)."""
from ..mnist.train import (AverageMeter, binary_cross_entropy_with_logits, cross_entropy, elbo_loss,  # noqa: F401
                           save_checkpoint)
from .model import MVAE


def annealing_factor(epoch, batch_idx, n_mini_batches, annealing_epochs):
    """KL annealing of fashionmnist/train.py:180-186."""
    if epoch < annealing_epochs:
        return float(batch_idx + epoch * n_mini_batches + 1) / float(annealing_epochs * n_mini_batches)
    return 1.0


def load_checkpoint(file_path, use_cuda=False):
    import torch
    ckpt = torch.load(file_path, map_location=None if use_cuda else "cpu")
    model = MVAE(ckpt["n_latents"])
    model.load_state_dict(ckpt["state_dict"])
    return model.cuda() if use_cuda else model
