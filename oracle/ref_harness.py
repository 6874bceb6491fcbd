"""Run the UNMODIFIED reference (``oracle/_ref/``, see make_ref.py) -- baseline infrastructure, never the product path.

The reference's training loops live under ``if __name__ == "__main__"`` (argparse, dataset downloads, ``.data[0]``) and
cannot be imported, so ``RefStep.step`` replays the loop BODY verbatim around the reference's own ``MVAE`` and
``elbo_loss`` (mnist/train.py:196-219, fashionmnist/train.py:196-219, celeba/train.py:189-212,
celeba19/train.py:254-309) with ``torch.optim.Adam``, on in-memory tensors.  Harness-side shims only (the files are
byte-identical copies): ``builtins.xrange = range``, ``np.int = int``, a stub ``datasets`` module (``N_ATTRS = 18``).

``device="cpu"`` is the reference's CPU path (what `python train.py` runs without --cuda); ``device="cuda"`` is its only
GPU path: the same modules under stock PyTorch eager (cuBLAS / cuDNN / ATen), as `python train.py --cuda` would.
"""
from __future__ import annotations

import builtins
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FLAVOURS = {"mnist": "mnist", "fashion": "fashionmnist", "celeba": "celeba", "celeba19": "celeba19"}


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, d, f)) for d in FLAVOURS.values() for f in ("model.py", "train.py"))


def _load(subdir: str, fname: str, modname: str, extra: dict):
    builtins.xrange = range
    if not hasattr(np, "int"):
        np.int = int
    keys = ("model", "datasets")
    saved = {k: sys.modules.get(k) for k in keys}
    try:
        sys.modules.update(extra)
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_DIR, subdir, fname + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_flavour(flavour: str):
    """(model module, train module) of one reference experiment directory."""
    d = FLAVOURS[flavour]
    ds = types.ModuleType("datasets")
    ds.N_ATTRS = 18
    ds.CelebAttributes = object
    ds.FashionMNIST = object
    model = _load(d, "model", f"mvae_ref_{d}_model", {"datasets": ds})
    train = _load(d, "train", f"mvae_ref_{d}_train", {"datasets": ds, "model": model})
    return model, train


class RefStep:
    """One reference training iteration per ``step`` call."""

    def __init__(self, flavour: str, n_latents: int, lr: float, lambda_image: float = 1.0, lambda_other: float = 10.0,
                 device: str = "cpu", approx_m: int = 1, seed: int = 0):
        if not available():
            raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
        self.flavour, self.dev = flavour, torch.device(device)
        self.model_mod, self.train_mod = load_flavour(flavour)
        torch.manual_seed(seed)
        self.model = self.model_mod.MVAE(n_latents)
        if self.dev.type == "cuda":
            self.model.cuda()
        self.model.train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr)
        self.lam_i, self.lam_o = lambda_image, lambda_other
        self.approx_m = approx_m
        if flavour == "celeba19":
            self.pool = self.train_mod.enumerate_combinations(19)

    def step(self, image: torch.Tensor, other: torch.Tensor, annealing_factor: float = 1.0) -> torch.Tensor:
        """image [B,C,H,W]; other = labels [B] int64 (mnist / fashion) or attrs [B,18] float (celeba*).  Returns the
        0-dim loss tensor (not synchronised)."""
        T, model, f = self.train_mod, self.model, self.flavour
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.opt.zero_grad()
            if f in ("mnist", "fashion"):
                kw = dict(lambda_image=self.lam_i, lambda_text=self.lam_o, annealing_factor=annealing_factor)
                r1 = model(image, other); r2 = model(image); r3 = model(text=other)
                loss = (T.elbo_loss(r1[0], image, r1[1], other, r1[2], r1[3], **kw) +
                        T.elbo_loss(r2[0], image, None, None, r2[2], r2[3], **kw) +
                        T.elbo_loss(None, None, r3[1], other, r3[2], r3[3], **kw))
            elif f == "celeba":
                kw = dict(lambda_image=self.lam_i, lambda_attrs=self.lam_o, annealing_factor=annealing_factor)
                r1 = model(image, other); r2 = model(image); r3 = model(attrs=other)
                loss = (T.elbo_loss(r1[0], image, r1[1], other, r1[2], r1[3], **kw) +
                        T.elbo_loss(r2[0], image, None, None, r2[2], r2[3], **kw) +
                        T.elbo_loss(None, None, r3[1], other, r3[2], r3[3], **kw))
            else:
                attrs = T.tensor_2d_to_list(other)
                n = len(attrs)
                ri, ra, mu, lv = model(image, attrs)
                loss = T.elbo_loss([ri] + ra, [image] + attrs, mu, lv, lambda_image=self.lam_i, lambda_attrs=self.lam_o,
                                   annealing_factor=annealing_factor)
                ri, _, mu, lv = model(image=image)
                loss = loss + T.elbo_loss([ri], [image], mu, lv, lambda_image=self.lam_i, lambda_attrs=self.lam_o,
                                          annealing_factor=annealing_factor)
                for ix in range(n):
                    _, ra, mu, lv = model(attrs=[attrs[k] if k == ix else None for k in range(n)])
                    loss = loss + T.elbo_loss([ra[ix]], [attrs[ix]], mu, lv, annealing_factor=annealing_factor)
                if self.approx_m > 0:
                    for combo in T.sample_combinations(self.pool, size=self.approx_m):
                        ac = combo[1:]
                        ri, ra, mu, lv = model(image=image if combo[0] else None,
                                               attrs=[attrs[ix] if ac[ix] else None for ix in range(ac.size)])
                        rec = [ra[ix] for ix in range(ac.size) if ac[ix]]
                        dat = [attrs[ix] for ix in range(ac.size) if ac[ix]]
                        if combo[0]:
                            rec, dat = [ri] + rec, [image] + dat
                        loss = loss + T.elbo_loss(rec, dat, mu, lv, annealing_factor=annealing_factor)
            loss.backward()
            self.opt.step()
        return loss.detach()
