"""Two data-parallel ranks of the fused trainer (graph(fwd+bwd) -> all-reduce -> graph(Adam)) must reproduce the
single-process full-batch step.  Runs on ONE GPU with the gloo backend (both ranks share cuda:0): it exercises the
trainer's multi-rank logic end to end; NCCL itself is covered by `bench.py --gpus N` on a multi-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _data(B, L):
    rs = np.random.RandomState(11)
    return (torch.from_numpy(rs.uniform(0, 1, (B, 784)).astype(np.float32)), torch.from_numpy(rs.randint(0, 10, B)),
            torch.from_numpy(rs.standard_normal((3, B, L)).astype(np.float32)))


def _worker(rank, world, port, B, L, out, use_graph):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    image, text, noise = _data(B, L)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    tr = MnistMVAETrainer(L, B // world, world_size=world, rank=rank, seed=0, use_graph=use_graph)
    losses = []
    for it in range(3):
        losses.append(tr.step(image[sl], text[sl], annealing_factor=0.3 * (it + 1), noise=noise[:, sl]))
    if rank == 0:
        torch.save({"losses": losses, "params": {k: v.cpu() for k, v in tr.params.items()}}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_ranks_match_single_process(tmp_path, use_graph):
    B, L, world = 256, 64, 2
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, L, out, use_graph), nprocs=world, join=True)
    got = torch.load(out)
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    image, text, noise = _data(B, L)
    ref = MnistMVAETrainer(L, B, seed=0, use_graph=False)
    for it in range(3):
        l = ref.step(image, text, annealing_factor=0.3 * (it + 1), noise=noise)
        assert abs(l - got["losses"][it]) <= 2e-6 * abs(l), (it, l, got["losses"][it])
    for k, v in ref.params.items():
        assert (v.cpu() - got["params"][k]).abs().max().item() <= 1e-4, k   # 3 Adam steps, lr 1e-3, fp32 summation order
