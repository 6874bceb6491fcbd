"""world_size-2 gloo test (CPU) of the data-parallel contract the trainer implements on NCCL:
every rank scales its loss terms by 1/B_global, gradients are SUM-all-reduced over one flat bucket, and the
result equals the single-process full-batch gradient (mnist/train.py:57 is a batch mean)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mvae_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _inputs(B, L):
    rs = np.random.RandomState(7)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 1, 28, 28))); text = torch.from_numpy(rs.randint(0, 10, B))
    noise = torch.from_numpy(rs.standard_normal((3, B, L)))
    return image, text, noise


def _worker(rank, world, port, B, L, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    image, text, noise = _inputs(B, L)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    p = {k: v.double() for k, v in O.make_params(O.mnist_param_shapes(L), seed=5).items()}
    loss, _, grads, _ = O.mnist_step_grads(p, image[sl], text[sl], L, [n[sl] for n in noise], 1.0, 10.0, 0.5)
    # local batch mean -> contribution to the global mean
    w = (B // world) / B
    names = [k for k, _ in O.mnist_param_shapes(L)]
    flat = torch.cat([grads[k].reshape(-1) * w for k in names] + [loss.reshape(1) * w])   # one flat bucket
    dist.all_reduce(flat)                                                                  # SUM
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_two_rank_sum_allreduce_equals_full_batch(tmp_path):
    B, L, world = 16, 64, 2
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, L, out), nprocs=world, join=True)
    flat = torch.load(out)
    image, text, noise = _inputs(B, L)
    p = {k: v.double() for k, v in O.make_params(O.mnist_param_shapes(L), seed=5).items()}
    loss, _, grads, _ = O.mnist_step_grads(p, image, text, L, list(noise), 1.0, 10.0, 0.5)
    ref = torch.cat([grads[k].reshape(-1) for k, _ in O.mnist_param_shapes(L)] + [loss.reshape(1)])
    assert torch.allclose(flat, ref, rtol=1e-9, atol=1e-12)
