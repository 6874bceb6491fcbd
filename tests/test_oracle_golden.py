"""Pin the oracle (oracle/) to the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import elementwise_np as EN
from oracle import mvae_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def mn():
    return dict(np.load(os.path.join(G, "mnist_golden.npz")))


@pytest.fixture(scope="module")
def ew():
    return dict(np.load(os.path.join(G, "elementwise_golden.npz")))


def _digest(t):
    a = t.detach().double().reshape(-1)
    return np.array([a.sum().item(), a.abs().sum().item(), a.pow(2).sum().sqrt().item()])


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_mnist_step_matches_reference(mn, mode):
    L = 64
    p = O.make_params(O.mnist_param_shapes(L), seed=0)
    image = torch.from_numpy(mn["mnist_image"]); text = torch.from_numpy(mn["mnist_text"])
    lam_i, lam_t, beta = mn["mnist_hyper"]
    noises = [torch.from_numpy(n) for n in mn["mnist_noises"]] if mode == "train" else [None] * 3
    loss, terms, grads, aux = O.mnist_step_grads(p, image, text, L, noises, lam_i, lam_t, beta)
    assert abs(loss.item() - mn[f"mnist_{mode}_loss"]) <= 2e-6 * abs(mn[f"mnist_{mode}_loss"])
    np.testing.assert_allclose([t.item() for t in terms], mn[f"mnist_{mode}_terms"], rtol=2e-6)
    for pi in range(3):
        np.testing.assert_allclose(aux["mu"][pi].detach().numpy(), mn[f"mnist_{mode}_mu{pi}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(aux["logvar"][pi].detach().numpy(), mn[f"mnist_{mode}_logvar{pi}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(aux["recon_text"][pi].detach().numpy(), mn[f"mnist_{mode}_recon_text{pi}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(aux["recon_image"][pi][0].detach().numpy(), mn[f"mnist_{mode}_recon_image{pi}_row0"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_digest(aux["recon_image"][pi]), mn[f"mnist_{mode}_recon_image{pi}_digest"], rtol=1e-5, atol=1e-5)
    for k, g in grads.items():
        np.testing.assert_allclose(g.reshape(-1)[:32].numpy(), mn[f"mnist_{mode}_grad_head/{k}"], rtol=2e-4, atol=1e-7)
        d = mn[f"mnist_{mode}_grad_digest/{k}"]
        np.testing.assert_allclose(_digest(g)[1:], d[1:], rtol=1e-4)
        if f"mnist_{mode}_grad_full/{k}" in mn:
            np.testing.assert_allclose(g.numpy(), mn[f"mnist_{mode}_grad_full/{k}"], rtol=2e-4, atol=1e-7)


def test_adam_matches_torch_optim(mn):
    L = 64
    p = O.make_params(O.mnist_param_shapes(L), seed=0)
    image = torch.from_numpy(mn["mnist_image"]); text = torch.from_numpy(mn["mnist_text"])
    lam_i, lam_t, beta = mn["mnist_hyper"]
    noises = [torch.from_numpy(n) for n in mn["mnist_noises"]]
    _, _, grads, _ = O.mnist_step_grads(p, image, text, L, noises, lam_i, lam_t, beta)
    O.adam_update(p, grads, {}, step=1, lr=1e-3)
    for k, v in p.items():
        np.testing.assert_allclose(v.reshape(-1)[:32].numpy(), mn[f"mnist_adam1_head/{k}"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(_digest(v)[1:], mn[f"mnist_adam1_digest/{k}"][1:], rtol=1e-5)


def test_poe_variants(ew):
    mu = torch.from_numpy(ew["poe_in_mu"]); lv = torch.from_numpy(ew["poe_in_logvar"])
    for v in "AB":
        m, l = O.product_of_experts(mu, lv, variant=v)
        np.testing.assert_allclose(m.numpy(), ew[f"poe{v}_mu"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(l.numpy(), ew[f"poe{v}_logvar"], rtol=1e-6, atol=1e-7)
        m, l = O.product_of_experts(mu[:2], lv[:2], variant=v)
        np.testing.assert_allclose(m.numpy(), ew[f"poe{v}_mu_2"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(l.numpy(), ew[f"poe{v}_logvar_2"], rtol=1e-6, atol=1e-7)
    g_mu = torch.tensor([0., 2., -1.]).view(3, 1, 1); g_lv = torch.tensor([0., -2., 1.]).view(3, 1, 1)
    np.testing.assert_allclose([t.item() for t in O.product_of_experts(g_mu, g_lv, variant="A")], ew["G2"], rtol=1e-6)
    np.testing.assert_allclose([t.item() for t in O.product_of_experts(g_mu, g_lv, variant="B")], ew["G3"], rtol=1e-6)
    np.testing.assert_allclose(ew["G1"], [0.5, -0.693147182], rtol=1e-6)


def test_bce_ce_known_answers(ew):
    out = O.bce_with_logits(torch.from_numpy(ew["G4_x"]), torch.from_numpy(ew["G4_t"]))
    np.testing.assert_allclose(out.numpy(), ew["G4"], rtol=1e-6, atol=1e-7)
    lg = torch.zeros(2, 10); lg[0, :3] = torch.tensor([1., 2., 3.])
    np.testing.assert_allclose(O.cross_entropy_rows(lg, torch.tensor([2, 7])).sum(1).numpy(), ew["G5"], rtol=1e-6)
    np.testing.assert_allclose(O.cross_entropy_rows(torch.from_numpy(ew["ce_x"]), torch.from_numpy(ew["ce_t"])).numpy(),
                               ew["ce_out"], rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        O.bce_with_logits(torch.zeros(3, 4), torch.zeros(3, 5))
    with pytest.raises(ValueError):
        O.cross_entropy_rows(torch.zeros(3, 10), torch.zeros(4, dtype=torch.long))


def test_celeba_elbo(ew):
    ra = torch.from_numpy(ew["celeba_recon_attrs"]); at = torch.from_numpy(ew["celeba_attrs"])
    mu = torch.from_numpy(ew["celeba_mu"]); lv = torch.from_numpy(ew["celeba_logvar"])
    v = O.elbo_loss_celeba(None, None, ra, at, mu, lv, 1.0, 10.0, 0.25)
    assert abs(v.item() - ew["celeba_elbo_attrs_only"]) <= 2e-6 * abs(ew["celeba_elbo_attrs_only"])
    ri = torch.from_numpy(ew["celeba_recon_image"].astype(np.float32)); im = torch.from_numpy(ew["celeba_image"].astype(np.float32))
    v = O.elbo_loss_celeba(ri, im, ra, at, mu, lv, 1.0, 10.0, 0.25)
    assert abs(v.item() - ew["celeba_elbo_joint_f16in"]) <= 2e-6 * abs(ew["celeba_elbo_joint_f16in"])


def test_annealing_schedules():
    # mnist/train.py:180-186 vs fashionmnist/train.py:182
    assert O.annealing_factor(1, 0, 600, 200, "mnist") == pytest.approx(1.0 / (200 * 600))
    assert O.annealing_factor(1, 0, 600, 200, "fashionmnist") == pytest.approx(601.0 / (200 * 600))
    assert O.annealing_factor(200, 5, 600, 200, "mnist") == 1.0


# ---- the analytic numpy formulas (what the CUDA kernels implement) vs autograd of the oracle ----

@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("train", [True, False])
def test_poe_multipass_analytic_vs_autograd(variant, train):
    rs = np.random.RandomState(5)
    E, B, L = 3, 5, 8
    masks = [0b111, 0b001, 0b110, 0b010]
    mu_e = rs.standard_normal((E, B, L)); lv_e = 0.8 * rs.standard_normal((E, B, L))
    noise = rs.standard_normal((len(masks), B, L)) if train else None
    dz = rs.standard_normal((len(masks), B, L)); beta = 0.37
    mu, lv, z, kl = EN.poe_multipass_fwd(mu_e, lv_e, masks, noise, beta, variant)
    dmu, dlv = EN.poe_multipass_bwd(mu_e, lv_e, masks, noise, beta, dz, variant)
    tm = torch.tensor(mu_e, requires_grad=True); tl = torch.tensor(lv_e, requires_grad=True)
    total = 0
    for p, m in enumerate(masks):
        idx = [e for e in range(E) if (m >> e) & 1]
        pm, pl = O.prior_expert((1, B, L), torch.float64)
        ms = torch.cat([pm] + [tm[e:e + 1] for e in idx], 0); ls = torch.cat([pl] + [tl[e:e + 1] for e in idx], 0)
        fm, fl = O.product_of_experts(ms, ls, variant=variant)
        zz = O.reparametrize(fm, fl, None if noise is None else torch.tensor(noise[p]))
        np.testing.assert_allclose(fm.detach().numpy(), mu[p], rtol=1e-10); np.testing.assert_allclose(zz.detach().numpy(), z[p], rtol=1e-10)
        klp = beta * O.kl_rows(fm, fl).mean()
        assert abs(klp.item() - kl[p]) < 1e-10
        total = total + (zz * torch.tensor(dz[p])).sum() + klp
    total.backward()
    np.testing.assert_allclose(dmu, tm.grad.numpy(), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(dlv, tl.grad.numpy(), rtol=1e-8, atol=1e-12)


def test_recon_analytic_vs_autograd():
    rs = np.random.RandomState(6)
    x = 3 * rs.standard_normal((7, 33)); t = rs.uniform(0, 1, (7, 33)); scale = 1.7 / 7
    loss, dx = EN.bce_logits_fwd_bwd(x, t, scale)
    tx = torch.tensor(x, requires_grad=True)
    l = (O.bce_with_logits(tx, torch.tensor(t)).sum() * scale); l.backward()
    assert abs(l.item() - loss) < 1e-10; np.testing.assert_allclose(dx, tx.grad.numpy(), rtol=1e-9, atol=1e-12)
    x = 3 * rs.standard_normal((9, 10)); tg = rs.randint(0, 10, 9)
    loss, dx = EN.ce_fwd_bwd(x, tg, scale)
    tx = torch.tensor(x, requires_grad=True)
    l = O.cross_entropy_rows(tx, torch.tensor(tg)).sum() * scale; l.backward()
    assert abs(l.item() - loss) < 1e-10; np.testing.assert_allclose(dx, tx.grad.numpy(), rtol=1e-9, atol=1e-12)
    xs = rs.standard_normal(100)
    txs = torch.tensor(xs, requires_grad=True); O.swish(txs).sum().backward()
    np.testing.assert_allclose(EN.swish_grad(xs), txs.grad.numpy(), rtol=1e-10)


def test_fashion_step_matches_reference():
    """Conv flavour (fashionmnist/model.py): oracle vs the unmodified reference on the B=4 fixture."""
    fa = dict(np.load(os.path.join(G, "fashion_golden.npz")))
    L = 64
    p = O.make_params(O.fashion_param_shapes(L), seed=0)
    image = torch.from_numpy(fa["image"]); text = torch.from_numpy(fa["text"])
    noises = [torch.from_numpy(n) for n in fa["noises"]]
    loss, terms, grads, aux = O.fashion_step_grads(p, image, text, L, noises, 1.0, 10.0, 0.5)
    np.testing.assert_allclose([t.item() for t in terms], fa["terms"], rtol=3e-6)
    for pi in range(3):
        np.testing.assert_allclose(aux["mu"][pi].detach().numpy(), fa[f"mu{pi}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(aux["recon_image"][pi].detach().numpy(), fa[f"recon_image{pi}"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(aux["recon_text"][pi].detach().numpy(), fa[f"recon_text{pi}"], rtol=1e-4, atol=1e-6)
    for k, g in grads.items():
        np.testing.assert_allclose(g.reshape(-1)[:64].numpy(), fa[f"grad_head/{k}"], rtol=3e-4, atol=1e-7)
        np.testing.assert_allclose(_digest(g)[1:], fa[f"grad_digest/{k}"][1:], rtol=1e-4)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_celeba_step_matches_reference(mode):
    """CelebA flavour (conv + BatchNorm + Dropout, PoE variant B): oracle vs the unmodified reference, B=4 fixture;
    dropout masks / noise are the reference's own RNG draws (replayed by make_golden.py)."""
    from oracle import celeba_oracle as CO
    ce = dict(np.load(os.path.join(G, "celeba_golden.npz")))
    L = 100
    st = CO.make_celeba_state(L, seed=0)
    image = torch.from_numpy(ce["image"]); attrs = torch.from_numpy(ce["attrs"])
    noises = [torch.from_numpy(n) for n in ce["noises"]]; masks = [torch.from_numpy(m) for m in ce["drop_masks"]]
    loss, terms, grads, bufs, aux = CO.step_grads(st, image, attrs, L, noises, masks, 1.0, 10.0, 0.5, training=(mode == "train"))
    np.testing.assert_allclose([t.item() for t in terms], ce[f"{mode}_terms"], rtol=5e-6)
    for pi in range(3):
        np.testing.assert_allclose(aux["mu"][pi].detach().numpy(), ce[f"{mode}_mu{pi}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(aux["recon_attrs"][pi].detach().numpy(), ce[f"{mode}_recon_attrs{pi}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(aux["recon_image"][pi].detach().reshape(-1)[:256].numpy(), ce[f"{mode}_recon_image{pi}_head"], rtol=1e-4, atol=1e-5)
    for k, g in grads.items():
        np.testing.assert_allclose(g.reshape(-1)[:64].numpy(), ce[f"{mode}_grad_head/{k}"], rtol=2e-3, atol=2e-6)
        # biases feeding a BatchNorm have a mathematically zero gradient (pure rounding noise ~1e-7): absolute floor
        np.testing.assert_allclose(_digest(g)[2], ce[f"{mode}_grad_digest/{k}"][2], rtol=5e-4, atol=2e-6)
    if mode == "train":
        for k, v in bufs.items():
            np.testing.assert_allclose(v.numpy(), ce[f"train_buffer/{k}"], rtol=1e-5, atol=1e-6)


def test_celeba19_step_matches_reference():
    """CelebA-19 (19 experts, 22 ELBO terms incl. two sampled combinations): oracle vs the unmodified reference."""
    from oracle import celeba19_oracle as O19
    c9 = dict(np.load(os.path.join(G, "celeba19_golden.npz")))
    L = 100
    st = O19.make_celeba19_state(L, seed=0)
    image = torch.from_numpy(c9["image"]); attrs = torch.from_numpy(c9["attrs"])
    noises = [torch.from_numpy(n) for n in c9["noises"]]; masks = [torch.from_numpy(m) for m in c9["drop_masks"]]
    total, terms, grads, bufs = O19.step_grads(st, image, attrs, L, noises, masks, c9["combos"], 1.0, 10.0, 0.5, training=True)
    np.testing.assert_allclose([t.item() for t in terms], c9["terms"], rtol=5e-6)
    assert abs(total.item() - c9["total"]) <= 5e-6 * abs(c9["total"])
    for k, g in grads.items():
        np.testing.assert_allclose(g.reshape(-1)[:32].numpy(), c9[f"grad_head/{k}"], rtol=3e-3, atol=3e-6)
        np.testing.assert_allclose(_digest(g)[2], c9[f"grad_digest/{k}"][2], rtol=1e-3, atol=1e-5)
    for k, v in bufs.items():
        np.testing.assert_allclose(v.numpy(), c9[f"buffer/{k}"], rtol=1e-5, atol=1e-6)


def test_celeba19_sampler_distribution_and_unranking():
    from itertools import combinations
    from oracle import celeba19_oracle as O19
    for n, k in ((6, 3), (19, 2), (7, 6)):
        all_c = list(combinations(range(n), k))
        for idx in (0, 1, len(all_c) // 2, len(all_c) - 1):
            assert tuple(O19.unrank_combination(n, k, idx)) == all_c[idx]
    rng = np.random.RandomState(0)
    rows = O19.sample_combinations_fast(19, 64, rng)
    assert rows.shape == (64, 19) and rows.dtype == bool
    sizes = rows.sum(1)
    assert sizes.min() >= 2 and sizes.max() <= 18
