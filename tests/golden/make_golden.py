"""Generate golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py        # needs /root/reference (build container only)

The reference modules are imported with importlib under unique names; harness-side
shims only (never edits the reference): ``builtins.xrange = range``, ``np.int = int``,
a stub ``datasets`` module providing ``N_ATTRS = 18`` for celeba.  Parameters come
from ``oracle.mvae_oracle.make_params`` (numpy RandomState stream), loaded into the
reference ``MVAE`` with ``load_state_dict`` so the fixture does not need to carry 10 MB
of weights.  Output: ``tests/golden/mnist_golden.npz`` and ``elementwise_golden.npz``.
"""
import builtins
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import mvae_oracle as O  # noqa: E402

REF = os.environ.get("MVAE_REFERENCE", "/root/reference")
warnings.filterwarnings("ignore")


def load_ref(subdir, fname, modname, extra_modules=None):
    """import /root/reference/<subdir>/<fname>.py as <modname>, with `model`/`datasets`
    resolved to that directory's own files."""
    builtins.xrange = range
    if not hasattr(np, "int"):
        np.int = int
    saved = {k: sys.modules.get(k) for k in ("model", "datasets")}
    try:
        for k, v in (extra_modules or {}).items():
            sys.modules[k] = v
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, subdir, fname + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def tensor_digest(t):
    a = t.detach().double().reshape(-1)
    return np.array([a.sum().item(), a.abs().sum().item(), a.pow(2).sum().sqrt().item()], np.float64)


def main():
    out = {}
    # ---------------------------------------------------------------- mnist
    ref_model = load_ref("mnist", "model", "ref_mnist_model")
    ref_train = load_ref("mnist", "train", "ref_mnist_train", {"model": ref_model})
    L, B = 64, 8
    params = O.make_params(O.mnist_param_shapes(L), seed=0)
    model = ref_model.MVAE(L)
    assert [k for k, _ in model.state_dict().items()] == [k for k, _ in O.mnist_param_shapes(L)]
    model.load_state_dict(params)
    rs = np.random.RandomState(1234)
    image = torch.from_numpy(rs.uniform(0, 1, size=(B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, size=(B,)).astype(np.int64))
    lam_i, lam_t, beta = 1.0, 10.0, 0.5

    def step(train_mode, seed=77):
        model.train(train_mode)
        model.zero_grad()
        torch.manual_seed(seed)
        r1 = model(image, text); r2 = model(image); r3 = model(text=text)
        j = ref_train.elbo_loss(r1[0], image, r1[1], text, r1[2], r1[3], lambda_image=lam_i, lambda_text=lam_t, annealing_factor=beta)
        i = ref_train.elbo_loss(r2[0], image, None, None, r2[2], r2[3], lambda_image=lam_i, lambda_text=lam_t, annealing_factor=beta)
        t = ref_train.elbo_loss(None, None, r3[1], text, r3[2], r3[3], lambda_image=lam_i, lambda_text=lam_t, annealing_factor=beta)
        loss = j + i + t
        loss.backward()
        return loss, (j, i, t), (r1, r2, r3)

    # the three noise draws the reference makes, replayed (mnist/model.py:32: one [B,L] normal_ per forward)
    torch.manual_seed(77)
    noises = [torch.empty(B, L).normal_() for _ in range(3)]
    for mode, tag in ((True, "train"), (False, "eval")):
        loss, terms, rr = step(mode)
        out[f"mnist_{tag}_loss"] = np.float64(loss.item())
        out[f"mnist_{tag}_terms"] = np.array([t.item() for t in terms], np.float64)
        for pi, r in enumerate(rr):
            out[f"mnist_{tag}_mu{pi}"] = r[2].detach().numpy()
            out[f"mnist_{tag}_logvar{pi}"] = r[3].detach().numpy()
            out[f"mnist_{tag}_recon_text{pi}"] = r[1].detach().numpy()
            out[f"mnist_{tag}_recon_image{pi}_digest"] = tensor_digest(r[0])
            out[f"mnist_{tag}_recon_image{pi}_row0"] = r[0][0].detach().numpy()
        for k, v in model.named_parameters():
            out[f"mnist_{tag}_grad_digest/{k}"] = tensor_digest(v.grad)
            out[f"mnist_{tag}_grad_head/{k}"] = v.grad.detach().reshape(-1)[:32].numpy().copy()
            if v.grad.numel() <= 1024:
                out[f"mnist_{tag}_grad_full/{k}"] = v.grad.detach().numpy().copy()
    out["mnist_image"] = image.numpy(); out["mnist_text"] = text.numpy()
    out["mnist_noises"] = torch.stack(noises).numpy()
    out["mnist_hyper"] = np.array([lam_i, lam_t, beta], np.float64)
    # one Adam step from the train-mode gradients (mnist/train.py:168,219)
    step(True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    opt.step()
    for k, v in model.named_parameters():
        out[f"mnist_adam1_digest/{k}"] = tensor_digest(v)
        out[f"mnist_adam1_head/{k}"] = v.detach().reshape(-1)[:32].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "mnist_golden.npz"), **out)

    # ---------------------------------------------------------------- fashionmnist (conv enc/dec)
    fds = types.ModuleType("datasets"); fds.FashionMNIST = object
    ref_fm = load_ref("fashionmnist", "model", "ref_fashion_model")
    Lf, Bf = 64, 4
    fparams = O.make_params(O.fashion_param_shapes(Lf), seed=0)
    fmodel = ref_fm.MVAE(Lf)
    assert [k for k in fmodel.state_dict().keys()] == [k for k, _ in O.fashion_param_shapes(Lf)], list(fmodel.state_dict().keys())
    fmodel.load_state_dict(fparams)
    rs = np.random.RandomState(4321)
    fimage = torch.from_numpy(rs.uniform(0, 1, size=(Bf, 1, 28, 28)).astype(np.float32))
    ftext = torch.from_numpy(rs.randint(0, 10, size=(Bf,)).astype(np.int64))
    fout = {}
    torch.manual_seed(78)
    fnoises = [torch.empty(Bf, Lf).normal_() for _ in range(3)]
    fmodel.train(True); fmodel.zero_grad(); torch.manual_seed(78)
    r1 = fmodel(fimage, ftext); r2 = fmodel(fimage); r3 = fmodel(text=ftext)
    fj = ref_train.elbo_loss(r1[0], fimage, r1[1], ftext, r1[2], r1[3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    fi = ref_train.elbo_loss(r2[0], fimage, None, None, r2[2], r2[3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    ft = ref_train.elbo_loss(None, None, r3[1], ftext, r3[2], r3[3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    (fj + fi + ft).backward()
    fout["image"] = fimage.numpy(); fout["text"] = ftext.numpy(); fout["noises"] = torch.stack(fnoises).numpy()
    fout["terms"] = np.array([fj.item(), fi.item(), ft.item()], np.float64)
    for pi, r in enumerate((r1, r2, r3)):
        fout[f"mu{pi}"] = r[2].detach().numpy(); fout[f"logvar{pi}"] = r[3].detach().numpy()
        fout[f"recon_text{pi}"] = r[1].detach().numpy(); fout[f"recon_image{pi}"] = r[0].detach().numpy()
    for k, v in fmodel.named_parameters():
        fout[f"grad_digest/{k}"] = tensor_digest(v.grad)
        fout[f"grad_head/{k}"] = v.grad.detach().reshape(-1)[:64].numpy().copy()
        if v.grad.numel() <= 2048:
            fout[f"grad_full/{k}"] = v.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "fashion_golden.npz"), **fout)

    # ---------------------------------------------------------------- celeba (conv + BatchNorm + Dropout, PoE variant B)
    from oracle import celeba_oracle as CO
    cds = types.ModuleType("datasets"); cds.N_ATTRS = 18; cds.CelebAttributes = object
    tqm = types.ModuleType("tqdm"); tqm.tqdm = lambda x=None, **k: x
    ref_cm = load_ref("celeba", "model", "ref_celeba_model2", {"datasets": cds})
    ref_ct = load_ref("celeba", "train", "ref_celeba_train2", {"datasets": cds, "model": ref_cm, "tqdm": sys.modules.get("tqdm", tqm)})
    Lc, Bc = 100, 4
    cstate = CO.make_celeba_state(Lc, seed=0)
    cmodel = ref_cm.MVAE(Lc)
    assert list(cmodel.state_dict().keys()) == [k for k, _ in CO.celeba_state_shapes(Lc)]
    cmodel.load_state_dict(cstate)
    rs = np.random.RandomState(2468)
    cimage = torch.from_numpy(rs.uniform(0, 1, size=(Bc, 3, 64, 64)).astype(np.float32))
    cattrs = torch.from_numpy(rs.randint(0, 2, size=(Bc, 18)).astype(np.float32))
    cout = {"image": cimage.numpy(), "attrs": cattrs.numpy()}
    # replay of the reference's RNG draws of one train-mode step: dropout mask then noise per model() call
    torch.manual_seed(79)
    m1 = torch.empty(Bc, 512).bernoulli_(0.9); n1 = torch.empty(Bc, Lc).normal_()
    m2 = torch.empty(Bc, 512).bernoulli_(0.9); n2 = torch.empty(Bc, Lc).normal_()
    n3 = torch.empty(Bc, Lc).normal_()
    captured = []
    hook = cmodel.image_encoder.classifier[2].register_forward_hook(lambda mod, inp, out: captured.append((out != 0).float()))
    for mode, tag in ((True, "train"), (False, "eval")):
        cmodel.load_state_dict(cstate)
        cmodel.train(mode); cmodel.zero_grad(); torch.manual_seed(79); captured.clear()
        r1 = cmodel(cimage, cattrs); r2 = cmodel(cimage); r3 = cmodel(attrs=cattrs)
        cj = ref_ct.elbo_loss(r1[0], cimage, r1[1], cattrs, r1[2], r1[3], lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
        ci = ref_ct.elbo_loss(r2[0], cimage, None, None, r2[2], r2[3], lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
        ca = ref_ct.elbo_loss(None, None, r3[1], cattrs, r3[2], r3[3], lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
        (cj + ci + ca).backward()
        if mode:
            assert torch.equal(captured[0], m1) and torch.equal(captured[1], m2), "dropout RNG replay does not match the reference"
        cout[f"{tag}_terms"] = np.array([cj.item(), ci.item(), ca.item()], np.float64)
        for pi, r in enumerate((r1, r2, r3)):
            cout[f"{tag}_mu{pi}"] = r[2].detach().numpy(); cout[f"{tag}_logvar{pi}"] = r[3].detach().numpy()
            cout[f"{tag}_recon_attrs{pi}"] = r[1].detach().numpy()
            cout[f"{tag}_recon_image{pi}_digest"] = tensor_digest(r[0]); cout[f"{tag}_recon_image{pi}_head"] = r[0].detach().reshape(-1)[:256].numpy().copy()
        for k, v in cmodel.named_parameters():
            cout[f"{tag}_grad_digest/{k}"] = tensor_digest(v.grad)
            cout[f"{tag}_grad_head/{k}"] = v.grad.detach().reshape(-1)[:64].numpy().copy()
        if mode:
            for k, v in cmodel.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    cout[f"train_buffer/{k}"] = v.numpy().copy()
    hook.remove()
    cout["noises"] = torch.stack([n1, n2, n3]).numpy(); cout["drop_masks"] = torch.stack([m1, m2]).numpy()
    np.savez_compressed(os.path.join(HERE, "celeba_golden.npz"), **cout)

    # ---------------------------------------------------------------- celeba19 (19 experts, sampled ELBO terms)
    from oracle import celeba19_oracle as O19
    import tqdm as _tq, torchvision as _tv  # noqa: F401  (real modules must be imported before the stubs below)
    ref19_m = load_ref("celeba19", "model", "ref_celeba19_model", {"datasets": cds})
    ref19_t = load_ref("celeba19", "train", "ref_celeba19_train", {"datasets": cds, "model": ref19_m})
    L9, B9 = 100, 3
    st9 = O19.make_celeba19_state(L9, seed=0)
    m9 = ref19_m.MVAE(L9)
    assert list(m9.state_dict().keys()) == [k for k, _ in O19.celeba19_state_shapes(L9)]
    m9.load_state_dict(st9)
    rs = np.random.RandomState(1357)
    image9 = torch.from_numpy(rs.uniform(0, 1, size=(B9, 3, 64, 64)).astype(np.float32))
    attrs9 = torch.from_numpy(rs.randint(0, 2, size=(B9, 18)).astype(np.float32))
    combos = np.zeros((2, 19), dtype=bool); combos[0, [0, 3, 6, 12]] = True; combos[1, [1, 8]] = True
    passes = O19.pass_list(combos)
    torch.manual_seed(80)
    noises9, masks9 = [], []
    for present, _ in passes:
        if present[0]:
            masks9.append(torch.empty(B9, 512).bernoulli_(0.9))
        noises9.append(torch.empty(B9, L9).normal_())
    m9.train(True); m9.zero_grad(); torch.manual_seed(80)
    alist = ref19_t.tensor_2d_to_list(attrs9)
    total = 0; terms9 = []
    recon_image, recon_attrs, mu, logvar = m9(image9, alist)
    t = ref19_t.elbo_loss([recon_image] + recon_attrs, [image9] + alist, mu, logvar, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    total = total + t; terms9.append(t.item())
    recon_image, _, mu, logvar = m9(image=image9)
    t = ref19_t.elbo_loss([recon_image], [image9], mu, logvar, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    total = total + t; terms9.append(t.item())
    for ix in range(18):
        _, recon_attrs, mu, logvar = m9(attrs=[alist[k] if k == ix else None for k in range(18)])
        t = ref19_t.elbo_loss([recon_attrs[ix]], [alist[ix]], mu, logvar, annealing_factor=0.5)
        total = total + t; terms9.append(t.item())
    for combo in combos:
        ac = combo[1:]
        recon_image, recon_attrs, mu, logvar = m9(image=image9 if combo[0] else None, attrs=[alist[ix] if ac[ix] else None for ix in range(18)])
        if combo[0]:
            t = ref19_t.elbo_loss([recon_image] + [recon_attrs[ix] for ix in range(18) if ac[ix]], [image9] + [alist[ix] for ix in range(18) if ac[ix]], mu, logvar, annealing_factor=0.5)
        else:
            t = ref19_t.elbo_loss([recon_attrs[ix] for ix in range(18) if ac[ix]], [alist[ix] for ix in range(18) if ac[ix]], mu, logvar, annealing_factor=0.5)
        total = total + t; terms9.append(t.item())
    total.backward()
    o9 = {"image": image9.numpy(), "attrs": attrs9.numpy(), "combos": combos, "noises": torch.stack(noises9).numpy(),
          "drop_masks": torch.stack(masks9).numpy(), "terms": np.array(terms9, np.float64), "total": np.float64(total.item())}
    for k, v in m9.named_parameters():
        o9[f"grad_digest/{k}"] = tensor_digest(v.grad); o9[f"grad_head/{k}"] = v.grad.detach().reshape(-1)[:32].numpy().copy()
    for k, v in m9.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            o9[f"buffer/{k}"] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "celeba19_golden.npz"), **o9)

    # ------------------------------------------------------- element-wise KATs
    ew = {}
    poeA = ref_model.ProductOfExperts()
    celeba_ds = types.ModuleType("datasets"); celeba_ds.N_ATTRS = 18; celeba_ds.CelebAttributes = object
    ref_celeba_model = load_ref("celeba", "model", "ref_celeba_model", {"datasets": celeba_ds})
    tq = types.ModuleType("tqdm"); tq.tqdm = lambda x, **k: x
    ref_celeba_train = load_ref("celeba", "train", "ref_celeba_train",
                                {"datasets": celeba_ds, "model": ref_celeba_model, "tqdm": sys.modules.get("tqdm", tq)})
    poeB = ref_celeba_model.ProductOfExperts()
    rs = np.random.RandomState(99)
    mu = torch.from_numpy(rs.standard_normal((4, 6, 16)).astype(np.float32)); mu[0] = 0
    lv = torch.from_numpy((0.7 * rs.standard_normal((4, 6, 16))).astype(np.float32)); lv[0] = 0
    for tag, poe in (("A", poeA), ("B", poeB)):
        m, l = poe(mu, lv)
        ew[f"poe{tag}_mu"] = m.numpy(); ew[f"poe{tag}_logvar"] = l.numpy()
        m2, l2 = poe(mu[:2], lv[:2])
        ew[f"poe{tag}_mu_2"] = m2.numpy(); ew[f"poe{tag}_logvar_2"] = l2.numpy()
    ew["poe_in_mu"] = mu.numpy(); ew["poe_in_logvar"] = lv.numpy()
    # SURVEY G1/G2/G3
    g_mu = torch.tensor([0., 2., -1.]).view(3, 1, 1); g_lv = torch.tensor([0., -2., 1.]).view(3, 1, 1)
    ew["G2"] = np.array([v.item() for v in poeA(g_mu, g_lv)]); ew["G3"] = np.array([v.item() for v in poeB(g_mu, g_lv)])
    ew["G1"] = np.array([v.item() for v in poeA(torch.tensor([0., 1.]).view(2, 1, 1), torch.zeros(2, 1, 1))])
    x = torch.tensor([-30., -2., 0., 0.5, 3., 30., 100.]); t = torch.tensor([0., 1., 0.5, 0.25, 1., 0., 1.])
    ew["G4_x"] = x.numpy(); ew["G4_t"] = t.numpy()
    ew["G4"] = ref_train.binary_cross_entropy_with_logits(x, t).numpy()
    lg = torch.zeros(2, 10); lg[0, :3] = torch.tensor([1., 2., 3.])
    ew["G5"] = ref_train.cross_entropy(lg, torch.tensor([2, 7])).sum(1).numpy()
    xr = torch.from_numpy((3 * rs.standard_normal((5, 10))).astype(np.float32)); tr = torch.from_numpy(rs.randint(0, 10, 5))
    ew["ce_x"] = xr.numpy(); ew["ce_t"] = tr.numpy(); ew["ce_out"] = ref_train.cross_entropy(xr, tr).numpy()
    # celeba elbo (attrs column loop) on random tensors
    Bc = 4
    ri = torch.from_numpy(rs.standard_normal((Bc, 3, 64, 64)).astype(np.float32)); im = torch.from_numpy(rs.uniform(0, 1, (Bc, 3, 64, 64)).astype(np.float32))
    ra = torch.from_numpy(rs.standard_normal((Bc, 18)).astype(np.float32)); at = torch.from_numpy(rs.randint(0, 2, (Bc, 18)).astype(np.float32))
    mu_c = torch.from_numpy(rs.standard_normal((Bc, 100)).astype(np.float32)); lv_c = torch.from_numpy((0.5 * rs.standard_normal((Bc, 100))).astype(np.float32))
    ew["celeba_recon_attrs"] = ra.numpy(); ew["celeba_attrs"] = at.numpy(); ew["celeba_mu"] = mu_c.numpy(); ew["celeba_logvar"] = lv_c.numpy()
    ew["celeba_recon_image_seed"] = np.int64(99)
    ew["celeba_elbo_attrs_only"] = np.float64(ref_celeba_train.elbo_loss(None, None, ra, at, mu_c, lv_c, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.25).item())
    ew["celeba_elbo_joint"] = np.float64(ref_celeba_train.elbo_loss(ri, im, ra, at, mu_c, lv_c, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.25).item())
    ew["celeba_recon_image"] = ri.numpy().astype(np.float16)  # compact; test upcasts the same way
    ew["celeba_image"] = im.numpy().astype(np.float16)
    ew["celeba_elbo_joint_f16in"] = np.float64(ref_celeba_train.elbo_loss(
        torch.from_numpy(ew["celeba_recon_image"].astype(np.float32)), torch.from_numpy(ew["celeba_image"].astype(np.float32)),
        ra, at, mu_c, lv_c, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.25).item())
    del ew["celeba_elbo_joint"]
    np.savez_compressed(os.path.join(HERE, "elementwise_golden.npz"), **ew)
    print("wrote", os.path.join(HERE, "mnist_golden.npz"), os.path.getsize(os.path.join(HERE, "mnist_golden.npz")))
    print("wrote", os.path.join(HERE, "elementwise_golden.npz"), os.path.getsize(os.path.join(HERE, "elementwise_golden.npz")))


if __name__ == "__main__":
    main()
