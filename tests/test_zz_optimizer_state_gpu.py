"""Optimizer-state interchange of the fused trainer: the Adam moments leave and enter the flat
arena in ``torch.optim.Adam.state_dict()`` format (reference checkpoint key 'optimizer', mnist/train.py:263-268), and a
trainer resumed from (state_dict, optimizer_state_dict) continues like the original."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(B, seed):
    rs = np.random.RandomState(seed)
    return (torch.from_numpy(rs.uniform(0, 1, (B, 784)).astype(np.float32)).cuda(),
            torch.from_numpy(rs.randint(0, 10, B).astype(np.int64)).cuda(),
            torch.from_numpy(rs.standard_normal((3, B, 64)).astype(np.float32)).cuda())


def test_mnist_trainer_resumes_from_torch_format_optimizer_state():
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    B = 64
    a = MnistMVAETrainer(64, B, use_graph=False, seed=3)
    for s in range(2):
        im, tx, nz = _batch(B, s)
        a.step(im, tx, annealing_factor=0.5, noise=nz)
    sd, osd = a.state_dict(), a.optimizer_state_dict()
    # the dict is a valid torch.optim.Adam state for parameters of the reference's shapes, in the reference's order
    plist = [torch.nn.Parameter(v.detach().cpu().clone()) for v in sd.values()]
    topt = torch.optim.Adam(plist, lr=1.0)
    topt.load_state_dict(osd)
    assert topt.param_groups[0]["lr"] == a.lr
    assert all(int(st["step"]) == 2 for st in topt.state_dict()["state"].values())
    for p, st in zip(plist, (topt.state_dict()["state"][i] for i in range(len(plist)))):
        assert st["exp_avg"].shape == p.shape and st["exp_avg_sq"].shape == p.shape
    # resume in a fresh trainer and take the same third step
    b = MnistMVAETrainer(64, B, use_graph=False, seed=99)
    b.load_state_dict(sd)
    b.load_optimizer_state_dict(osd)
    assert int(b.step_count.item()) == 2
    im, tx, nz = _batch(B, 7)
    la = a.step(im, tx, annealing_factor=0.9, noise=nz)
    lb = b.step(im, tx, annealing_factor=0.9, noise=nz)
    assert abs(la - lb) <= 1e-6 * abs(la)
    for k in a.params:
        d = (a.params[k] - b.params[k]).abs().max().item()
        assert d <= 1e-4, (k, d)      # same moments, same step count: only the order of fp32 atomics may differ
                                      # (Adam turns a 1e-7 relative gradient difference into at most ~lr for
                                      # near-zero gradients, so the bound is a few Adam steps' worth, as in test_dp_gpu)


def test_device_resident_dataset_matches_host_fed_steps():
    """batches gathered on the device from the uint8 dataset (/255 and label lookup fused)
    give the same training steps as host-fed float batches."""
    from multimodal_vae_public_b200 import ops
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    rs = np.random.RandomState(5)
    N, B = 1000, 64
    images = torch.from_numpy(rs.randint(0, 256, (N, 1, 28, 28)).astype(np.uint8))
    labels = torch.from_numpy(rs.randint(0, 10, N).astype(np.int64))
    # the gather kernel alone: ToTensor semantics, exact
    idx = torch.from_numpy(rs.permutation(N)[:B].astype(np.int64)).cuda()
    out = torch.full((B, 784), -1.0, device="cuda"); lab = torch.full((B,), -1, dtype=torch.int64, device="cuda")
    ops.gather_batch_u8(images.reshape(N, 784).cuda(), labels.cuda(), idx, out, lab)
    assert torch.equal(out.cpu(), images.reshape(N, 784)[idx.cpu()].float().div(255))
    assert torch.equal(lab.cpu(), labels[idx.cpu()])
    # two trainers, same seed (same Philox noise): device-resident vs host-fed
    a = MnistMVAETrainer(64, B, use_graph=False, seed=4)
    b = MnistMVAETrainer(64, B, use_graph=False, seed=4)
    a.attach_dataset(images, labels)
    perm = a.epoch_permutation(seed=1)
    assert sorted(perm.cpu().tolist()) == list(range(N))
    for it in range(3):
        sel = perm[it * B:(it + 1) * B]
        la = a.step_from_dataset(sel, annealing_factor=0.5)
        lb = b.step(images[sel.cpu()].float().div(255), labels[sel.cpu()], annealing_factor=0.5)
        assert abs(la - lb) <= 1e-6 * abs(lb), (it, la, lb)
    for k in a.params:
        assert (a.params[k] - b.params[k]).abs().max().item() <= 1e-4, k   # fp32 atomic order only (see above)
