"""The reference's known-answer vectors (tests/golden/elementwise_golden.npz: values produced
by the UNMODIFIED reference functions) pushed through the CUDA kernels via the C ABI -- not only through the oracle.

Covers what random 4-sigma inputs never reach: saturated logits (|x| = 30, 88, 100), exact 0/1 targets, the
sigmoid-gradient at both ends (the BCE kernels use MUFU ex2/lg2/rcp), PoE with and without the implicit prior expert,
and the cross-entropy row outputs of the reference's own function."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    from multimodal_vae_public_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def ew():
    return dict(np.load(os.path.join(G, "elementwise_golden.npz")))


def _bce(ops, x, t, width):
    """Run x/t (1-D numpy) through mvae_bce_logits_fwd_bwd laid out as one row of `width` columns (repeated to fill it):
    width % 4 == 0 takes the 128-bit streaming kernel, otherwise the narrow-row kernel."""
    n = len(x)
    reps = (width + n - 1) // n
    xr = np.tile(x, reps)[:width].astype(np.float32); tr = np.tile(t, reps)[:width].astype(np.float32)
    xd = torch.from_numpy(xr).cuda().view(1, width); td = torch.from_numpy(tr).cuda().view(1, width)
    dx = torch.empty_like(xd); le = torch.empty_like(xd)
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.bce_logits_fwd_bwd(xd, td, dx, 1.0, acc, seg_rows=0, loss_elem=le)
    return le.cpu().numpy()[0], dx.cpu().numpy()[0], acc.item(), xr, tr


@pytest.mark.parametrize("width", [7, 28, 784])
def test_G4_bce_known_answers(ops, ew, width):
    le, dx, acc, xr, tr = _bce(ops, ew["G4_x"], ew["G4_t"], width)
    ref = np.tile(ew["G4"], (width + 6) // 7)[:width]
    # tolerance: 2e-6 relative + 3e-7 absolute (MUFU lg2 on the log1p term; G4 itself is fp32 output of the reference)
    np.testing.assert_allclose(le, ref, rtol=2e-6, atol=3e-7)
    assert abs(acc - ref.astype(np.float64).sum()) <= 2e-6 * ref.astype(np.float64).sum()
    sig = 1.0 / (1.0 + np.exp(-xr.astype(np.float64)))
    np.testing.assert_allclose(dx, sig - tr, rtol=3e-6, atol=1e-7)


@pytest.mark.parametrize("stacked", [False, True])
def test_bce_saturated_logits_and_binary_targets(ops, stacked):
    xs = np.array([-100., -88.5, -88., -30., -17., -1e-3, 0., 1e-3, 17., 30., 88., 88.5, 100., 1e4, -1e4, 1e-30],
                  np.float32)
    x = np.repeat(xs, 2); t = np.tile(np.array([0., 1.], np.float32), len(xs))        # every logit against t = 0 and t = 1
    x = np.tile(x, 8)[:256 - 256 % 4]; t = np.tile(t, 8)[:len(x)]
    R = 2 if stacked else 1   # stacked: two passes share the target (the image term of the joint + image-only pass)
    xd = torch.from_numpy(np.tile(x, (R, 1))).cuda().contiguous(); td = torch.from_numpy(t).cuda().view(1, -1)
    dx = torch.empty_like(xd)
    acc = torch.zeros(R, dtype=torch.float64, device="cuda")
    ops.bce_logits_fwd_bwd(xd, td, dx, 1.0, acc, seg_rows=1)
    x64 = x.astype(np.float64); t64 = t.astype(np.float64)
    ref_loss = np.maximum(x64, 0) - x64 * t64 + np.log1p(np.exp(-np.abs(x64)))
    with np.errstate(over="ignore"):
        ref_g = np.where(x64 >= 0, 1.0 / (1.0 + np.exp(-x64)), np.exp(x64) / (1.0 + np.exp(x64))) - t64
    g = dx.cpu().numpy()
    assert np.isfinite(g).all() and np.isfinite(acc.cpu().numpy()).all()
    for r in range(R):
        np.testing.assert_allclose(g[r], ref_g, rtol=3e-6, atol=1e-7)
        assert abs(acc[r].item() - ref_loss.sum()) <= 2e-6 * ref_loss.sum()
    # the ends: sigma(-100) - 0 = 3.8e-44 (a denormal, as torch.sigmoid gives), sigma(100) - 1 == 0 exactly; no NaN from inf * 0
    assert 0.0 <= g[0][(x == -100) & (t == 0)].max() <= 1e-40 and g[0][(x == 100) & (t == 1)].max() == 0.0


def test_G5_and_ce_rows(ops, ew):
    lg = np.zeros((2, 16), np.float32); lg[0, :3] = [1., 2., 3.]
    x = torch.from_numpy(lg).cuda()[:, :10]
    tg = torch.tensor([2, 7], device="cuda")
    rows = torch.zeros(2, 16, device="cuda"); dx = torch.zeros(2, 16, device="cuda")
    acc = torch.zeros(2, dtype=torch.float64, device="cuda")
    ops.ce_fwd_bwd(x, tg, dx[:, :10], 10, 1.0, acc, seg_rows=1, loss_rows=rows[:, :10])
    np.testing.assert_allclose(rows[:, :10].sum(1).cpu().numpy(), ew["G5"], rtol=1e-6)
    np.testing.assert_allclose(acc.cpu().numpy(), ew["G5"], rtol=1e-6)
    # the reference function's own [N, K] output on random logits
    cx = np.zeros((5, 16), np.float32); cx[:, :10] = ew["ce_x"]
    x = torch.from_numpy(cx).cuda()[:, :10]
    rows = torch.zeros(5, 16, device="cuda"); dx = torch.zeros(5, 16, device="cuda")
    ops.ce_fwd_bwd(x, torch.from_numpy(ew["ce_t"]).cuda(), dx[:, :10], 10, 1.0, None, seg_rows=0, loss_rows=rows[:, :10])
    np.testing.assert_allclose(rows[:, :10].cpu().numpy(), ew["ce_out"], rtol=2e-6, atol=2e-7)
    sm = torch.softmax(torch.from_numpy(ew["ce_x"]).double(), 1).numpy()
    sm[np.arange(5), ew["ce_t"]] -= 1.0
    np.testing.assert_allclose(dx[:, :10].cpu().numpy(), sm, rtol=2e-5, atol=1e-6)


def _poe(ops, mu, lv, variant, no_prior, L):
    """mu, lv: [E, B, L] numpy -> fused (mu, logvar) [B, L] from mvae_poe_fwd (eval mode: z = mu)."""
    E, B = mu.shape[0], mu.shape[1]
    enc = [torch.from_numpy(np.concatenate([mu[e], lv[e]], 1).astype(np.float32)).cuda() for e in range(E)]
    z = torch.empty(B, L, device="cuda"); mo = torch.empty(B, L, device="cuda"); lo = torch.empty(B, L, device="cuda")
    ops.poe_fwd([e[:, :L] for e in enc], [e[:, L:] for e in enc], [(1 << E) - 1], B, L, z, variant=variant | (2 if no_prior else 0),
                training=False, mu_out=mo, lv_out=lo)
    assert torch.equal(z, mo)
    return mo.cpu().numpy(), lo.cpu().numpy()


@pytest.mark.parametrize("variant", [0, 1])
def test_poe_reference_vectors(ops, ew, variant):
    v = "AB"[variant]
    mu, lv = ew["poe_in_mu"], ew["poe_in_logvar"]          # [4 experts, 6, 16]: ProductOfExperts.forward(mu, logvar)
    m, l = _poe(ops, mu, lv, variant, True, 16)
    np.testing.assert_allclose(m, ew[f"poe{v}_mu"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(l, ew[f"poe{v}_logvar"], rtol=2e-6, atol=2e-7)
    m, l = _poe(ops, mu[:2], lv[:2], variant, True, 16)
    np.testing.assert_allclose(m, ew[f"poe{v}_mu_2"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(l, ew[f"poe{v}_logvar_2"], rtol=2e-6, atol=2e-7)
    # G2 / G3: three scalar experts mu=(0,2,-1), logvar=(0,-2,1) (broadcast to one 4-wide row)
    gm = np.array([0., 2., -1.], np.float32).reshape(3, 1, 1) * np.ones((1, 1, 4), np.float32)
    gl = np.array([0., -2., 1.], np.float32).reshape(3, 1, 1) * np.ones((1, 1, 4), np.float32)
    m, l = _poe(ops, gm, gl, variant, True, 4)
    ref = ew["G2"] if variant == 0 else ew["G3"]
    np.testing.assert_allclose([m[0, 0], l[0, 0]], ref, rtol=1e-6)


def test_G1_prior_times_unit_gaussian(ops, ew):
    # experts {N(0,1) prior (implicit in the kernel), N(1,1)} -> mu = 0.5, logvar = -log 2   (variant A)
    m, l = _poe(ops, np.ones((1, 1, 4), np.float32), np.zeros((1, 1, 4), np.float32), 0, False, 4)
    np.testing.assert_allclose([m[0, 0], l[0, 0]], ew["G1"], rtol=1e-6)
