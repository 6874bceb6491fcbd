"""GPU parity of the CelebA-19 trainer (image + 18 attribute experts, 20 + approx_m ELBO terms per step) against the
golden fixture produced by the unmodified reference (celeba19/model.py, celeba19/train.py) and the fp64 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import celeba19_oracle as O19

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
L = 100


def _trainer(B, approx_m, **kw):
    from multimodal_vae_public_b200.trainer_celeba19 import CelebA19MVAETrainer
    return CelebA19MVAETrainer(n_latents=L, batch_size=B, approx_m=approx_m, precision=1, **kw)


def test_names_and_state_dict_round_trip():
    from multimodal_vae_public_b200 import trainer_celeba19 as T
    assert T.celeba19_param_shapes(L) == O19.celeba19_param_shapes(L)
    tr = _trainer(4, 1)
    st = O19.make_celeba19_state(L, seed=5)
    tr.load_state_dict(st)
    sd = tr.state_dict()
    assert list(sd.keys()) == [k for k, _ in O19.celeba19_state_shapes(L)]
    for k, v in st.items():
        assert torch.equal(sd[k].cpu(), v), k


def test_sampler_matches_oracle_sampler():
    from multimodal_vae_public_b200.trainer_celeba19 import sample_combinations
    a = sample_combinations(19, 5, np.random.RandomState(11))
    b = O19.sample_combinations_fast(19, 5, np.random.RandomState(11))
    assert np.array_equal(a, b)


def test_celeba19_step_matches_reference_golden():
    ce = dict(np.load(os.path.join(G, "celeba19_golden.npz")))
    combos = ce["combos"]
    tr = _trainer(3, len(combos))
    tr.load_state_dict(O19.make_celeba19_state(L, seed=0))
    image = torch.from_numpy(ce["image"]); attrs = torch.from_numpy(ce["attrs"])
    noise = torch.from_numpy(ce["noises"]); masks = torch.from_numpy(ce["drop_masks"])
    total = tr.step(image, attrs, annealing_factor=0.5, noise=noise, drop_masks=masks, combos=combos, update=False)
    ls = tr.losses()
    assert len(ls["terms"]) == len(ce["terms"]) == 22
    for i, (mine, ref) in enumerate(zip(ls["terms"], ce["terms"])):
        assert abs(mine - ref) <= 1e-5 * abs(ref) + 1e-4, (i, mine, ref)
    assert abs(total - ce["total"]) <= 1e-5 * abs(ce["total"])
    grads = tr.export_grads()
    for k, g in grads.items():
        g = g.cpu()
        head = ce[f"grad_head/{k}"]
        d = ce[f"grad_digest/{k}"]
        scale = max(np.abs(head).max(), d[2] / np.sqrt(g.numel()), 1e-5)
        n = len(head)
        assert np.abs(g.reshape(-1)[:n].numpy() - head).max() <= 2e-3 * scale + 1e-5, k
        assert abs(g.double().norm().item() - d[2]) <= 5e-3 * d[2] + 5e-5, k
    sd = tr.state_dict()
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(sd[k].cpu().numpy(), ce[f"buffer/{k}"], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("case,B", [("mixed", 8), ("no_image", 8), ("all_image", 8), ("mixed", 64), ("mixed", 512)])
def test_celeba19_step_matches_oracle_fp64(case, B):
    # B = 512: BASELINE.json configs[4]; B = 64: its per-GPU batch at 8 GPUs (the fp64 oracle step takes ~30 s of CPU)
    rs = np.random.RandomState(4)
    combos = np.zeros((3, 19), dtype=bool)
    if case == "mixed":
        combos[0, [0, 1, 2, 18]] = True; combos[1, [5, 6]] = True; combos[2, [0, 9]] = True
    elif case == "no_image":
        combos[0, [1, 2]] = True; combos[1, [3, 4, 5, 6, 7, 8, 9, 10]] = True; combos[2, 1:] = True
    else:
        combos[0, [0, 1]] = True; combos[1, :18] = True; combos[2, [0, 7, 18]] = True
    passes = O19.pass_list(combos)
    n_img = sum(1 for p, _ in passes if p[0])
    image = torch.from_numpy(rs.uniform(0, 1, (B, 3, 64, 64)).astype(np.float32))
    attrs = torch.from_numpy(rs.randint(0, 2, (B, 18)).astype(np.float32))
    noise = torch.from_numpy(rs.standard_normal((len(passes), B, L)).astype(np.float32))
    masks = torch.from_numpy((rs.uniform(0, 1, (n_img, B, 512)) > 0.1).astype(np.float32))
    st = O19.make_celeba19_state(L, seed=2)
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in st.items()}
    total, terms, grads, bufs = O19.step_grads(st64, image.double(), attrs.double(), L, [n.double() for n in noise],
                                               [m.double() for m in masks], combos, 1.0, 10.0, 0.5, training=True)
    tr = _trainer(B, 3)
    tr.load_state_dict(st)
    got = tr.step(image, attrs, annealing_factor=0.5, noise=noise, drop_masks=masks, combos=combos, update=False)
    assert abs(got - total.item()) <= 5e-6 * abs(total.item())
    for mine, ref in zip(tr.losses()["terms"], terms):
        assert abs(mine - ref.item()) <= 5e-6 * abs(ref.item()) + 1e-5
    mine = tr.export_grads()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, gref in grads.items():
        err = (mine[k].cpu().double() - gref).abs().max().item() / max(gref.abs().max().item(), 1e-4 * gmax)
        assert err <= 2e-3, (k, err)
    sd = tr.state_dict()
    for k, v in bufs.items():
        np.testing.assert_allclose(sd[k].cpu().numpy(), v.numpy(), rtol=1e-4, atol=1e-5)


def test_celeba19_training_decreases_loss_and_samples_combos():
    """Default path: device-generated noise / dropout, host-sampled modality subsets, Adam updates."""
    B = 16
    rs = np.random.RandomState(0)
    np.random.seed(5)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 3, 64, 64)).astype(np.float32))
    attrs = torch.from_numpy(rs.randint(0, 2, (B, 18)).astype(np.float32))
    tr = _trainer(B, 2, lr=1e-3)
    tr.step(image, attrs, annealing_factor=0.0)
    first = tr.losses()["terms"][0]          # the joint term (the total varies with the sampled subsets)
    for _ in range(15):
        tr.step(image, attrs, annealing_factor=0.0)
    last = tr.losses()["terms"][0]
    assert np.isfinite(first) and np.isfinite(last) and last < first
    sd = tr.state_dict()
    assert int(sd["image_decoder.hallucinate.1.num_batches_tracked"]) == 16 * 22


def test_celeba19_eval_mode_matches_oracle_fp64():
    """training=False: z = mu, Dropout off, BatchNorm uses running statistics (which therefore must not change)."""
    B = 8
    rs = np.random.RandomState(6)
    combos = np.zeros((1, 19), dtype=bool); combos[0, [0, 4, 11]] = True
    image = torch.from_numpy(rs.uniform(0, 1, (B, 3, 64, 64)).astype(np.float32))
    attrs = torch.from_numpy(rs.randint(0, 2, (B, 18)).astype(np.float32))
    st = O19.make_celeba19_state(L, seed=8)
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in st.items()}
    total, terms, grads, bufs = O19.step_grads(st64, image.double(), attrs.double(), L, [None] * 21, [], combos, 1.0, 10.0,
                                               0.25, training=False)
    tr = _trainer(B, 1)
    tr.load_state_dict(st)
    got = tr.step(image, attrs, annealing_factor=0.25, combos=combos, training=False, update=False)
    assert abs(got - total.item()) <= 5e-6 * abs(total.item())
    mine = tr.export_grads()
    gmax = max(g.abs().max().item() for g in grads.values())
    for k, gref in grads.items():
        err = (mine[k].cpu().double() - gref).abs().max().item() / max(gref.abs().max().item(), 1e-4 * gmax)
        assert err <= 2e-3, (k, err)
    sd = tr.state_dict()
    for k, v in bufs.items():
        assert torch.equal(sd[k].cpu(), st[k]), k


def test_pipelined_host_fed_steps_equal_synchronous_steps():
    """step_pipelined (upload of batch i+1 on a copy stream, loss read one call late) == step on the same batches and the
    same sampled subsets (numpy's global stream, re-seeded), eval mode (no noise / masks)."""
    B = 8
    rs = np.random.RandomState(12)
    batches = [(torch.from_numpy(rs.uniform(0, 1, (B, 3, 64, 64)).astype(np.float32)).pin_memory(),
                torch.from_numpy(rs.randint(0, 2, (B, 18)).astype(np.float32)).pin_memory()) for _ in range(4)]
    a = _trainer(B, 2)
    b = _trainer(B, 2)
    b.load_state_dict(a.state_dict())
    np.random.seed(5)
    ref = [a.step(im, at, annealing_factor=0.5, training=False) for im, at in batches]
    np.random.seed(5)
    got = []
    for im, at in batches:
        v = b.step_pipelined(im, at, annealing_factor=0.5, training=False)
        if v is not None:
            got.append(v)
    got.append(b.flush())
    assert len(got) == len(ref)
    for x, y in zip(ref, got):
        assert abs(x - y) <= 2e-6 * abs(x), (ref, got)
    for k in a.params:   # (several Adam steps on atomically-summed gradients: equal up to a small fraction of lr per step)
        assert (a.params[k] - b.params[k]).abs().max().item() <= 1e-4, k
