"""Checkpoint compatibility of the optimizer state: the conversion between named Adam
moments and ``torch.optim.Adam.state_dict()`` -- what the reference stores under checkpoint['optimizer']
(mnist/train.py:263-268) -- must round-trip through a real torch optimizer and resume bit-exactly.  Pure host logic."""
import copy

import torch

from multimodal_vae_public_b200.trainer import adam_state_from_torch, adam_state_to_torch


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))


def _step(model, opt, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(4, 5, generator=g)
    opt.zero_grad()
    model(x).pow(2).sum().backward()
    opt.step()


def test_round_trip_through_torch_adam_resumes_exactly():
    a = _model(); oa = torch.optim.Adam(a.parameters(), lr=1e-3)
    for s in range(3):
        _step(a, oa, s)
    names = [k for k, _ in a.named_parameters()]
    m, v, step, lr = adam_state_from_torch(oa.state_dict(), names)
    assert step == 3 and lr == 1e-3 and set(m) == set(names)
    for k, p in a.named_parameters():
        assert m[k].shape == p.shape and v[k].shape == p.shape
    sd = adam_state_to_torch(names, m, v, step, lr)
    # a fresh optimizer on a copy of the model, loaded from OUR dict, continues exactly like the original
    b = copy.deepcopy(a); ob = torch.optim.Adam(b.parameters(), lr=123.0)
    ob.load_state_dict(sd)
    assert ob.param_groups[0]["lr"] == 1e-3
    _step(a, oa, 99); _step(b, ob, 99)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)


def test_unstepped_optimizer_and_legacy_int_step():
    a = _model(); oa = torch.optim.Adam(a.parameters(), lr=1e-4)
    names = [k for k, _ in a.named_parameters()]
    m, v, step, lr = adam_state_from_torch(oa.state_dict(), names)      # never stepped: no state entries
    assert step == 0 and all(t is None for t in m.values()) and lr == 1e-4
    assert adam_state_to_torch(names, {}, {}, 0, lr)["state"] == {}
    _step(a, oa, 1)
    sd = oa.state_dict()
    for st in sd["state"].values():                                      # 2018-era checkpoints store step as an int
        st["step"] = int(st["step"].item())
    assert adam_state_from_torch(sd, names)[2] == 1
    try:
        adam_state_from_torch(sd, names[:-1])
        raise AssertionError("length mismatch not detected")
    except ValueError:
        pass
