"""GPU parity of the drop-in module surface (mnist/model.py + mnist/train.py names) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import mvae_oracle as O

pytestmark = pytest.mark.gpu


def _setup(B=24, L=64, seed=2):
    from multimodal_vae_public_b200.mnist import model as M
    rs = np.random.RandomState(seed)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64))
    p32 = O.make_params(O.mnist_param_shapes(L), seed=seed)
    m = M.MVAE(L)
    m.load_state_dict(p32)
    return m.cuda(), image, text, p32


def test_state_dict_and_surface():
    from multimodal_vae_public_b200.mnist import model as M, train as T
    m = M.MVAE(64)
    assert list(m.state_dict().keys()) == [k for k, _ in O.mnist_param_shapes(64)]
    for name in ("MVAE", "ImageEncoder", "ImageDecoder", "TextEncoder", "TextDecoder", "ProductOfExperts", "Swish", "prior_expert"):
        assert hasattr(M, name)
    for name in ("elbo_loss", "binary_cross_entropy_with_logits", "cross_entropy", "AverageMeter", "save_checkpoint", "load_checkpoint"):
        assert hasattr(T, name)
    mu, lv = M.prior_expert((1, 5, 64))
    assert mu.shape == (1, 5, 64) and float(mu.abs().sum() + lv.abs().sum()) == 0.0


def test_eval_forward_and_elbo_backward_match_oracle():
    from multimodal_vae_public_b200.mnist import train as T
    m, image, text, p32 = _setup()
    m.eval()
    L = 64
    ic, tc = image.cuda(), text.cuda()
    outs = [m(ic, tc), m(ic), m(text=tc)]
    j = T.elbo_loss(outs[0][0], ic, outs[0][1], tc, outs[0][2], outs[0][3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    i = T.elbo_loss(outs[1][0], ic, None, None, outs[1][2], outs[1][3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    t = T.elbo_loss(None, None, outs[2][1], tc, outs[2][2], outs[2][3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    loss = j + i + t
    loss.backward()
    p64 = {k: v.double() for k, v in p32.items()}
    ref, terms, grads, aux = O.mnist_step_grads(p64, image.double(), text, L, [None] * 3, 1.0, 10.0, 0.5)
    assert abs(loss.item() - ref.item()) <= 5e-6 * abs(ref.item())
    for got, want in zip((j, i, t), terms):
        assert abs(got.item() - want.item()) <= 5e-6 * abs(want.item()) + 1e-5
    for pi in range(3):
        np.testing.assert_allclose(outs[pi][2].detach().cpu().numpy(), aux["mu"][pi].detach().numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(outs[pi][3].detach().cpu().numpy(), aux["logvar"][pi].detach().numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(outs[pi][0].detach().cpu().numpy(), aux["recon_image"][pi].detach().numpy(), rtol=1e-3, atol=5e-5)
        np.testing.assert_allclose(outs[pi][1].detach().cpu().numpy(), aux["recon_text"][pi].detach().numpy(), rtol=1e-3, atol=5e-5)
    for k, v in m.named_parameters():
        g = grads[k]
        err = (v.grad.cpu().double() - g).abs().max().item() / max(g.abs().max().item(), 1e-12)
        assert err < 3e-4, (k, err)


def test_train_mode_reparam_and_poe_module():
    from multimodal_vae_public_b200 import functional as F
    from multimodal_vae_public_b200.mnist import model as M
    m, image, text, _ = _setup(B=16)
    m.train()
    r1 = m(image.cuda(), text.cuda()); r2 = m(image.cuda(), text.cuda())
    assert not torch.equal(r1[0], r2[0]) and torch.equal(r1[2], r2[2])       # fresh noise, same posterior
    # explicit-noise reparametrisation and its gradient
    mu = torch.randn(16, 64, device="cuda", requires_grad=True); lv = torch.randn(16, 64, device="cuda", requires_grad=True)
    nz = torch.randn(16, 64, device="cuda")
    z = F.reparametrize(mu, lv, noise=nz)
    w = torch.randn_like(z)
    (z * w).sum().backward()
    zr = O.reparametrize(mu.detach().cpu().double(), lv.detach().cpu().double(), nz.cpu().double())
    np.testing.assert_allclose(z.detach().cpu().numpy(), zr.numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(mu.grad.cpu().numpy(), w.cpu().numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(lv.grad.cpu().numpy(), (w * nz * 0.5 * torch.exp(0.5 * lv.detach())).cpu().numpy(), rtol=2e-5, atol=1e-6)
    # ProductOfExperts on an explicit [M,B,D] stack (prior as row 0, like the reference's infer)
    poe = M.ProductOfExperts()
    smu = torch.randn(3, 8, 64, device="cuda"); slv = 0.5 * torch.randn(3, 8, 64, device="cuda"); smu[0] = 0; slv[0] = 0
    smu.requires_grad_(True); slv.requires_grad_(True)
    pm, pl = poe(smu, slv)
    (pm.sum() + 2 * pl.sum()).backward()
    tm = smu.detach().cpu().double().requires_grad_(True); tl = slv.detach().cpu().double().requires_grad_(True)
    rm, rl = O.product_of_experts(tm, tl, variant="A")
    (rm.sum() + 2 * rl.sum()).backward()
    np.testing.assert_allclose(pm.detach().cpu().numpy(), rm.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(pl.detach().cpu().numpy(), rl.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(smu.grad.cpu().numpy(), tm.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(slv.grad.cpu().numpy(), tl.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_errors_like_reference():
    from multimodal_vae_public_b200.mnist import train as T
    with pytest.raises(ValueError):
        T.binary_cross_entropy_with_logits(torch.zeros(3, 4, device="cuda"), torch.zeros(3, 5, device="cuda"))
    with pytest.raises(ValueError):
        T.cross_entropy(torch.zeros(3, 10, device="cuda"), torch.zeros(4, dtype=torch.long, device="cuda"))


def test_elementwise_loss_functions_match_oracle():
    from multimodal_vae_public_b200.mnist import train as T
    rs = np.random.RandomState(3)
    x = torch.from_numpy((3 * rs.standard_normal((5, 7))).astype(np.float32)).cuda().requires_grad_(True)
    t = torch.from_numpy(rs.uniform(0, 1, (5, 7)).astype(np.float32)).cuda()
    out = T.binary_cross_entropy_with_logits(x, t)
    w = torch.from_numpy(rs.standard_normal((5, 7)).astype(np.float32)).cuda()
    (out * w).sum().backward()
    xr = x.detach().cpu().double().requires_grad_(True)
    ro = O.bce_with_logits(xr, t.cpu().double()); (ro * w.cpu().double()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ro.detach().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.numpy(), rtol=2e-5, atol=1e-6)
    lg = torch.from_numpy((2 * rs.standard_normal((6, 10))).astype(np.float32)).cuda().requires_grad_(True)
    tg = torch.from_numpy(rs.randint(0, 10, 6)).cuda()
    ce = T.cross_entropy(lg, tg)
    ce.sum(dim=1).mean().backward()
    lr = lg.detach().cpu().double().requires_grad_(True)
    rce = O.cross_entropy_rows(lr, tg.cpu()); rce.sum(dim=1).mean().backward()
    np.testing.assert_allclose(ce.detach().cpu().numpy(), rce.detach().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(lg.grad.cpu().numpy(), lr.grad.numpy(), rtol=2e-5, atol=1e-7)


def test_fashion_modules_match_oracle():
    """Drop-in fashionmnist/model.py surface: state_dict keys, eval forward of the three passes, ELBO, every gradient."""
    from multimodal_vae_public_b200.fashionmnist import model as FM, train as FT
    L, B = 64, 12
    m = FM.MVAE(L)
    assert list(m.state_dict().keys()) == [k for k, _ in O.fashion_param_shapes(L)]
    p32 = O.make_params(O.fashion_param_shapes(L), seed=4)
    m.load_state_dict(p32)
    m = m.cuda().eval()
    rs = np.random.RandomState(8)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32)); text = torch.from_numpy(rs.randint(0, 10, B))
    ic, tc = image.cuda(), text.cuda()
    outs = [m(ic, tc), m(ic), m(text=tc)]
    j = FT.elbo_loss(outs[0][0], ic, outs[0][1], tc, outs[0][2], outs[0][3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    i = FT.elbo_loss(outs[1][0], ic, None, None, outs[1][2], outs[1][3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    t = FT.elbo_loss(None, None, outs[2][1], tc, outs[2][2], outs[2][3], lambda_image=1.0, lambda_text=10.0, annealing_factor=0.5)
    (j + i + t).backward()
    ref, terms, grads, aux = O.fashion_step_grads({k: v.double() for k, v in p32.items()}, image.double(), text, L,
                                                  [None] * 3, 1.0, 10.0, 0.5)
    assert abs((j + i + t).item() - ref.item()) <= 5e-6 * abs(ref.item())
    assert outs[0][0].shape == (B, 1, 28, 28)
    np.testing.assert_allclose(outs[0][0].detach().cpu().numpy(), aux["recon_image"][0].detach().numpy(), rtol=1e-3, atol=1e-4)
    for k, v in m.named_parameters():
        g = grads[k]
        err = (v.grad.cpu().double() - g).abs().max().item() / max(g.abs().max().item(), 1e-12)
        assert err < 5e-4, (k, err)
    assert FT.annealing_factor(1, 0, 600, 200) == pytest.approx(601.0 / (200 * 600))


def test_celeba_modules_match_oracle():
    """Drop-in celeba/model.py surface: eval-mode three-pass objective + gradients, and the train-mode BatchNorm path
    (batch statistics, running-stat updates) of the encoders, against the oracle."""
    from oracle import celeba_oracle as CO
    from multimodal_vae_public_b200.celeba import model as CM, train as CT
    L, B = 100, 6
    st = CO.make_celeba_state(L, seed=6)
    m = CM.MVAE(L)
    assert list(m.state_dict().keys()) == [k for k, _ in CO.celeba_state_shapes(L)]
    m.load_state_dict(st)
    m = m.cuda().eval()
    rs = np.random.RandomState(10)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 3, 64, 64)).astype(np.float32))
    attrs = torch.from_numpy(rs.randint(0, 2, (B, 18)).astype(np.float32))
    ic, ac = image.cuda(), attrs.cuda()
    outs = [m(ic, ac), m(ic), m(attrs=ac)]
    assert outs[0][0].shape == (B, 3, 64, 64) and outs[0][1].shape == (B, 18)
    j = CT.elbo_loss(outs[0][0], ic, outs[0][1], ac, outs[0][2], outs[0][3], lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    i = CT.elbo_loss(outs[1][0], ic, None, None, outs[1][2], outs[1][3], lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    a = CT.elbo_loss(None, None, outs[2][1], ac, outs[2][2], outs[2][3], lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    (j + i + a).backward()
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in st.items()}
    ref, terms, grads, _, aux = CO.step_grads(st64, image.double(), attrs.double(), L, [None] * 3, [None] * 2, 1.0, 10.0, 0.5,
                                              training=False)
    assert abs((j + i + a).item() - ref.item()) <= 1e-5 * abs(ref.item())
    np.testing.assert_allclose(outs[0][0].detach().cpu().numpy(), aux["recon_image"][0].detach().numpy(), rtol=2e-3, atol=2e-4)
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, v in m.named_parameters():
        g = grads[k]
        err = (v.grad.cpu().double() - g).abs().max().item() / max(g.abs().max().item(), 1e-4 * gmax)
        assert err < 2e-3, (k, err)
    # ---- train mode: BatchNorm batch statistics + running-stat update (Dropout disabled: p = 0)
    m.zero_grad(); m.train()
    m.image_encoder.classifier[2].p = 0.0
    mu_i, lv_i = m.image_encoder(ic)
    mu_a, lv_a = m.attrs_encoder(ac)
    st_ref = {k: (v.double().clone() if v.dtype.is_floating_point else v.clone()) for k, v in st.items()}
    keep = torch.full((B, 512), 0.9, dtype=torch.float64)          # h * 0.9 / 0.9 == h in the oracle
    rmu_i, rlv_i = CO.image_encoder(st_ref, image.double(), L, True, keep)
    rmu_a, rlv_a = CO.attrs_encoder(st_ref, attrs.double(), L, True)
    np.testing.assert_allclose(mu_i.detach().cpu().numpy(), rmu_i.numpy(), rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(lv_a.detach().cpu().numpy(), rlv_a.numpy(), rtol=2e-3, atol=2e-4)
    sd = m.state_dict()
    for k in ("image_encoder.features.6.running_mean", "image_encoder.features.9.running_var", "attrs_encoder.net.4.running_var"):
        np.testing.assert_allclose(sd[k].cpu().numpy(), st_ref[k].numpy(), rtol=1e-4, atol=1e-5)
    assert int(sd["image_encoder.features.3.num_batches_tracked"]) == 1


def test_celeba19_modules_match_oracle():
    """Drop-in celeba19/model.py + train.py surface: eval-mode objective of the full 22-term step (joint, image-only,
    18 singles, 2 sampled subsets) built from MVAE.forward + the list-based elbo_loss, and its gradients, vs the oracle."""
    from oracle import celeba19_oracle as O19
    from multimodal_vae_public_b200.celeba19 import model as M19, train as T19
    L, B = 100, 4
    st = O19.make_celeba19_state(L, seed=3)
    m = M19.MVAE(L)
    assert list(m.state_dict().keys()) == [k for k, _ in O19.celeba19_state_shapes(L)]
    m.load_state_dict(st)
    m = m.cuda().eval()
    rs = np.random.RandomState(12)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 3, 64, 64)).astype(np.float32))
    attrs = torch.from_numpy(rs.randint(0, 2, (B, 18)).astype(np.float32))
    combos = np.zeros((2, 19), dtype=bool); combos[0, [0, 2, 17]] = True; combos[1, [4, 5, 6]] = True
    ic = image.cuda(); alist = T19.tensor_2d_to_list(attrs.cuda())
    total = 0
    ri, ra, mu, lv = m(ic, alist)
    assert ri.shape == (B, 3, 64, 64) and len(ra) == 18 and ra[0].shape == (B,)
    total = total + T19.elbo_loss([ri] + ra, [ic] + alist, mu, lv, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    ri, _, mu, lv = m(image=ic)
    total = total + T19.elbo_loss([ri], [ic], mu, lv, lambda_image=1.0, lambda_attrs=10.0, annealing_factor=0.5)
    for i in range(18):
        _, ra, mu, lv = m(attrs=[alist[k] if k == i else None for k in range(18)])
        total = total + T19.elbo_loss([ra[i]], [alist[i]], mu, lv, annealing_factor=0.5)
    for c in combos:
        ri, ra, mu, lv = m(image=ic if c[0] else None, attrs=[alist[k] if c[1 + k] else None for k in range(18)])
        rec = ([ri] if c[0] else []) + [ra[k] for k in range(18) if c[1 + k]]
        dat = ([ic] if c[0] else []) + [alist[k] for k in range(18) if c[1 + k]]
        total = total + T19.elbo_loss(rec, dat, mu, lv, annealing_factor=0.5)
    total.backward()
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in st.items()}
    ref, terms, grads, _ = O19.step_grads(st64, image.double(), attrs.double(), L, [None] * 22, [], combos, 1.0, 10.0, 0.5,
                                          training=False)
    assert abs(total.item() - ref.item()) <= 1e-5 * abs(ref.item())
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, v in m.named_parameters():
        g = grads[k]
        err = (v.grad.cpu().double() - g).abs().max().item() / max(g.abs().max().item(), 1e-4 * gmax)
        assert err < 2e-3, (k, err)
    # the subset sampler consumes numpy's global RNG like the reference's pool-based one
    np.random.seed(21); a = T19.sample_combinations(T19.enumerate_combinations(19), 3)
    np.random.seed(21); b = O19.sample_combinations_fast(19, 3, np.random)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", ["prior", "image", "text", "both"])
def test_generate_matches_oracle(mode):
    """mnist/sample.py:66-112 as a function: the latent Gaussian of the four modes, n samples decoded by both decoders,
    sigmoid / log_softmax outputs -- against the fp64 oracle with the same noise."""
    from multimodal_vae_public_b200.mnist.sample import generate
    m, image, text, p32 = _setup(B=4)
    L, n = 64, 12
    rs = np.random.RandomState(9)
    noise = torch.from_numpy(rs.standard_normal((n, L)).astype(np.float32))
    img = image[1:2] if mode in ("image", "both") else None
    txt = int(text[1]) if mode in ("text", "both") else None
    m.train()                                     # generate() must switch to eval itself and restore the mode
    img_recon, txt_logp, z = generate(m, n, image=img, text=txt, noise=noise)
    assert m.training and img_recon.shape == (n, 1, 28, 28) and txt_logp.shape == (n, 10)
    p64 = {k: v.double() for k, v in p32.items()}
    if mode == "prior":
        mu, std = torch.zeros(1, L, dtype=torch.float64), torch.ones(1, L, dtype=torch.float64)
    else:
        mu, lv = O.mnist_infer(p64, None if img is None else img.double(), None if txt is None else torch.tensor([txt]), L)
        std = (0.5 * lv).exp()
    zr = noise.double() * std + mu
    np.testing.assert_allclose(z.cpu().numpy(), zr.numpy(), rtol=1e-4, atol=2e-5)
    ri = torch.sigmoid(O.mnist_decoder(p64, "image_decoder", zr)); rt = torch.log_softmax(O.mnist_decoder(p64, "text_decoder", zr), 1)
    np.testing.assert_allclose(img_recon.reshape(n, 784).cpu().numpy(), ri.numpy(), rtol=1e-3, atol=5e-5)
    np.testing.assert_allclose(txt_logp.cpu().numpy(), rt.numpy(), rtol=1e-3, atol=1e-4)
