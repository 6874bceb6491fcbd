"""Conditional / unconditional generation with the reference's ``mnist/sample.py`` semantics (:66-112), as a function
instead of a script: the latent Gaussian is N(0, 1) (mode 1) or the posterior inferred from an image and / or a label
(modes 2-4, ``model.infer`` in eval mode), ``n_samples`` draws ``z = eps * std + mu`` are decoded by both decoders, and the
outputs are ``sigmoid(image logits)`` and ``log_softmax(label logits)`` exactly as the script saves them.  Everything runs
through the CUDA kernels of the drop-in modules (encoders, fused PoE, decoders); nothing falls back to the CPU.

Reference quirk kept out: the script tests ``if not args.condition_on_text`` on an int, so "condition on label 0" silently
becomes unconditional there; here ``None`` means "not given" and 0 is a label.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


@torch.no_grad()
def generate(model, n_samples: int = 64, image: Optional[torch.Tensor] = None, text=None,
             noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns ``(img_recon [n,1,28,28] in (0,1), txt_logprob [n,10], z [n,L])``.

    image: one image ([1,28,28], [1,1,28,28] or [784]) to condition on; text: one label (int or 1-element tensor);
    noise: optional [n_samples, n_latents] N(0,1) draws (the script's ``torch.randn``), for reproducible tests."""
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("generate() runs on the CUDA kernels of the drop-in model: move the model to a GPU first")
    was_training = model.training
    model.eval()
    try:
        L = model.n_latents
        if image is None and text is None:                      # mode 1: the prior
            mu = torch.zeros(1, L, device=dev)
            std = torch.ones(1, L, device=dev)
        else:                                                   # modes 2-4: the (product-of-experts) posterior
            img = None if image is None else image.to(dev, torch.float32).reshape(1, 1, 28, 28)
            txt = None if text is None else torch.as_tensor(text, device=dev).reshape(1).long()
            mu, logvar = model.infer(image=img, text=txt)
            std = logvar.mul(0.5).exp()
        eps = torch.randn(n_samples, L, device=dev) if noise is None else noise.to(dev, torch.float32)
        if eps.shape != (n_samples, L):
            raise ValueError(f"noise must be [{n_samples}, {L}]")
        z = eps * std.expand_as(eps) + mu.expand_as(eps)
        img_recon = torch.sigmoid(model.image_decoder(z)).reshape(n_samples, 1, 28, 28)
        txt_recon = torch.log_softmax(model.text_decoder(z), dim=1)
        return img_recon, txt_recon, z
    finally:
        model.train(was_training)
