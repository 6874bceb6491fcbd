"""GPU parity of the fused MNIST-MVAE training step (trainer.MnistMVAETrainer, through the C ABI)
against (a) the golden fixture produced by the unmodified reference and (b) the CPU oracle on seeded
inputs, plus size-independent properties at the benchmark batch size."""
import os

import numpy as np
import pytest
import torch

from oracle import mvae_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")

# stated tolerances (relative): ELBO 1e-4 is the north-star bar; measured headroom is far larger in 3xTF32.
ELBO_RTOL = {1: 2e-6, 0: 1e-4}
GRAD_RTOL = {1: 2e-4, 0: 3e-2}   # max|g - g_ref| / max|g_ref| per tensor, fp64 oracle


def _trainer(B, prec, graph=False, L=64, **kw):
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    return MnistMVAETrainer(n_latents=L, batch_size=B, precision=prec, use_graph=graph, **kw)


@pytest.fixture(scope="module")
def mn():
    return dict(np.load(os.path.join(G, "mnist_golden.npz")))


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_step_matches_reference_golden(mn, prec, mode):
    """B=8 batch of the fixture: the three ELBO terms, their sum and the parameter gradients."""
    L = 64
    tr = _trainer(8, prec)
    tr.load_state_dict(O.make_params(O.mnist_param_shapes(L), seed=0))
    lam_i, lam_t, beta = mn["mnist_hyper"]
    tr.lam_i, tr.lam_t = float(lam_i), float(lam_t)
    image = torch.from_numpy(mn["mnist_image"]); text = torch.from_numpy(mn["mnist_text"])
    noise = torch.from_numpy(mn["mnist_noises"]) if mode == "train" else None
    tr.step(image, text, annealing_factor=float(beta), noise=noise, training=(mode == "train"), update=False)
    ls = tr.losses()
    ref_terms = mn[f"mnist_{mode}_terms"]
    for name, ref in zip(("joint", "image", "text"), ref_terms):
        assert abs(ls[name] - ref) <= ELBO_RTOL[prec] * abs(ref) + 1e-5, (name, ls[name], ref)
    assert abs(ls["total"] - mn[f"mnist_{mode}_loss"]) <= ELBO_RTOL[prec] * abs(mn[f"mnist_{mode}_loss"]) + 1e-5
    for k in tr.grads:
        g = tr.grads[k].cpu()
        head = mn[f"mnist_{mode}_grad_head/{k}"]
        scale = max(np.abs(head).max(), mn[f"mnist_{mode}_grad_digest/{k}"][2] / np.sqrt(g.numel()))
        assert np.abs(g.reshape(-1)[:32].numpy() - head).max() <= GRAD_RTOL[prec] * scale + 1e-7, k
        if f"mnist_{mode}_grad_full/{k}" in mn:
            full = mn[f"mnist_{mode}_grad_full/{k}"]
            assert np.abs(g.numpy() - full).max() <= GRAD_RTOL[prec] * np.abs(full).max() + 1e-7, k
        d = mn[f"mnist_{mode}_grad_digest/{k}"]
        assert abs(g.double().norm().item() - d[2]) <= 5 * GRAD_RTOL[prec] * d[2] + 1e-7, k


@pytest.mark.parametrize("prec,B", [(1, 64), (0, 64), (1, 200), (0, 200), (1, 4096), (1, 512)])
def test_step_matches_oracle_fp64(prec, B):
    """Seeded batch; every gradient tensor and the Adam-updated parameters against the fp64 oracle.  B = 4096 is the
    BASELINE.json batch (configs[1]: multi-wave tile schedules, split-K sized by B), B = 512 the per-GPU batch of the
    8-GPU strong-scaling run (small-M tile shapes)."""
    L = 64
    rs = np.random.RandomState(B)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32))
    p32 = O.make_params(O.mnist_param_shapes(L), seed=3)
    p64 = {k: v.double() for k, v in p32.items()}
    beta = 0.5
    loss, terms, grads, _ = O.mnist_step_grads(p64, image.double(), text, L, [n.double() for n in noise], 1.0, 10.0, beta)
    tr = _trainer(B, prec)
    tr.load_state_dict(p32)
    got = tr.step(image, text, annealing_factor=beta, noise=noise, update=True)
    assert abs(got - loss.item()) <= ELBO_RTOL[prec] * abs(loss.item())
    ls = tr.losses()
    for name, ref in zip(("joint", "image", "text"), terms):
        assert abs(ls[name] - ref.item()) <= ELBO_RTOL[prec] * abs(ref.item()) + 1e-5
    for k, gref in grads.items():
        g = tr.grads[k].cpu().double()
        err = (g - gref).abs().max().item() / max(gref.abs().max().item(), 1e-12)
        assert err <= GRAD_RTOL[prec], (k, err)
    # one Adam step from those gradients (update=True above)
    O.adam_update(p64, grads, {}, step=1, lr=1e-3)
    for k, wref in p64.items():
        w = tr.params[k].cpu().double()
        # first Adam step moves every weight by lr * g/(|g| + eps): for |g| ~ eps = 1e-8 the update is sensitive to
        # the relative gradient error, so the bound is one full step lr = 1e-3 (exact Adam arithmetic is pinned separately
        # in test_kernels_gpu.py::test_adam_matches_oracle with identical gradients).
        assert (w - wref).abs().max().item() <= (1e-3 if prec == 1 else 2.5e-3), k
    assert tr.step_count.item() == 1


def test_graph_replay_equals_eager_and_philox_noise_changes():
    B, L = 256, 64
    rs = np.random.RandomState(1)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 784)).astype(np.float32)); text = torch.from_numpy(rs.randint(0, 10, B))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32))
    a, b = _trainer(B, 1, graph=False), _trainer(B, 1, graph=True)
    b.load_state_dict(a.state_dict())
    for it in range(3):
        la = a.step(image, text, annealing_factor=0.1 * (it + 1), noise=noise)
        lb = b.step(image, text, annealing_factor=0.1 * (it + 1), noise=noise)
        assert abs(la - lb) <= 1e-6 * abs(la), (it, la, lb)      # same kernels; only atomics order differs
    # (Adam turns rounding-level differences of near-zero gradients -- the order of the fp32 atomics differs between the
    # two runs -- into differences of a fraction of lr = 1e-3 per step; 3 steps here)
    for k in a.params:
        assert (a.params[k] - b.params[k]).abs().max().item() <= 5e-5, k
    # Philox path under graph replay: fresh noise every step, loss finite and decreasing over a few steps
    l0 = b.step(image, text, annealing_factor=1.0)
    n0 = b.noise.clone()
    l1 = b.step(image, text, annealing_factor=1.0)
    assert not torch.equal(n0, b.noise)
    for _ in range(20):
        l2 = b.step(image, text, annealing_factor=1.0)
    assert np.isfinite([l0, l1, l2]).all() and l2 < l0


def test_full_size_properties():
    """configs[1] size (B=4096): linearity in the loss weights and data-parallel shard additivity --
    size-independent properties checked on the GPU path only."""
    B, L = 4096, 64
    g = torch.Generator().manual_seed(0)
    image = torch.rand(B, 784, generator=g); text = torch.randint(0, 10, (B,), generator=g)
    noise = torch.randn(3, B, L, generator=g)
    tr = _trainer(B, 1)
    tr.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    full = {k: v.clone() for k, v in tr.grads.items()}; lfull = tr.losses()["total"]
    # eval-mode ELBO is deterministic and grads finite
    assert all(torch.isfinite(v).all() for v in full.values())
    # shard additivity: two half-batch "ranks" with world_size=2 scaling sum to the full-batch gradient
    acc = {k: torch.zeros_like(v) for k, v in full.items()}; lsum = 0.0
    half = _trainer(B // 2, 1)
    half.world = 2  # scales by the global batch; no process group is used because we never call all_reduce here
    half.load_state_dict(tr.state_dict())
    import multimodal_vae_public_b200.trainer as T
    for r in range(2):
        sl = slice(r * B // 2, (r + 1) * B // 2)
        half.set_inputs(image[sl], text[sl], noise[:, sl], 0.5)
        with torch.cuda.stream(half._stream):
            half.grad_bucket.zero_(); half.zero_region.zero_()
            if half.presplit:
                T.ops.split_lo(half.flat_params, half.params_lo)
            half._enqueue_forward(True, True)
            half._enqueue_loss_and_backward(True, B)
            T.ops.elbo_finalize(half.acc[0:3], half.acc[3:6], half.acc[6:9], 3, half.lam_i, half.lam_t, 1.0, 1.0 / B,
                                half.loss_out, beta_dev=half.beta_dev)
        half.synchronize()
        lsum += half.loss_out[0].item()
        for k in acc:
            acc[k] += half.grads[k]
    assert abs(lsum - lfull) <= 2e-6 * abs(lfull)
    for k in full:
        err = (acc[k] - full[k]).abs().max().item() / max(full[k].abs().max().item(), 1e-12)
        assert err <= 2e-5, (k, err)


def test_pipelined_host_fed_steps_equal_synchronous_steps():
    B, L = 128, 64
    rs = np.random.RandomState(2)
    batches = [(torch.from_numpy(rs.uniform(0, 1, (B, 784)).astype(np.float32)).pin_memory(),
                torch.from_numpy(rs.randint(0, 10, B)).pin_memory()) for _ in range(5)]
    a, b = _trainer(B, 1, graph=True), _trainer(B, 1, graph=True)
    b.load_state_dict(a.state_dict())
    # eval-mode steps (no noise) so both trainers see identical inputs; updates on
    ref = [a.step(im, tx, annealing_factor=0.5, training=False) for im, tx in batches]
    got = []
    for im, tx in batches:
        v = b.step_pipelined(im, tx, annealing_factor=0.5, training=False)
        if v is not None:
            got.append(v)
    got.append(b.flush())
    assert len(got) == len(ref)
    for x, y in zip(ref, got):
        assert abs(x - y) <= 2e-6 * abs(x)
    for k in a.params:   # (several Adam steps on atomically-summed gradients: equal up to a small fraction of lr per step)
        assert (a.params[k] - b.params[k]).abs().max().item() <= 1e-4, k


@pytest.mark.parametrize("flavour", ["mnist", "fashion"])
def test_label_table_mode_equals_per_sample_label_encoder(flavour):
    """The label encoder evaluated once per class (default) and evaluated on all B rows (MVAE_LABEL_TABLE=0 path) are the
    same function: same losses, same gradients up to summation order / 3xTF32-vs-FMA rounding."""
    B, L = 300, 64
    rs = np.random.RandomState(8)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28)).astype(np.float32))
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32))
    if flavour == "mnist":
        from multimodal_vae_public_b200.trainer import MnistMVAETrainer as T
    else:
        from multimodal_vae_public_b200.trainer_fashion import FashionMVAETrainer as T
    a = T(n_latents=L, batch_size=B, use_graph=False, label_table=True)
    b = T(n_latents=L, batch_size=B, use_graph=False, label_table=False)
    b.load_state_dict(a.state_dict())
    la = a.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    lb = b.step(image, text, annealing_factor=0.5, noise=noise, update=False)
    assert abs(la - lb) <= 2e-6 * abs(la)
    for k in a.grads:
        ref = b.grads[k]
        err = (a.grads[k] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        assert err <= 1e-4, (k, err)
