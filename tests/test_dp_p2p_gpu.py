"""The data-parallel exchange as one fused kernel per rank over NVLink peer memory (mvae_allreduce_adam_p2p: gradient
reduce-scatter by peer loads -> Adam on the rank's slice -> parameter all-gather by peer stores) must reproduce the
single-process full-batch trajectory, exactly like the NCCL + flat-Adam path does.  Needs two GPUs with peer access."""
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


def _worker(rank, world, port, B, L, out, use_graph, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    image, text, noise = _data(B, L)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    tr = MnistMVAETrainer(L, B // world, device=f"cuda:{rank}", world_size=world, rank=rank, seed=0, use_graph=use_graph,
                          dp_mode=mode)
    assert tr.dp_mode == mode, f"requested {mode}, got {tr.dp_mode}"
    losses = []
    for it in range(4):
        losses.append(tr.step(image[sl], text[sl], annealing_factor=0.25 * (it + 1), noise=noise[:, sl]))
    # a gradient-only step (no update) must still return the global loss
    losses.append(tr.step(image[sl], text[sl], annealing_factor=1.0, noise=noise[:, sl], update=False))
    err = int(tr._symm["flags"][2 * world].item()) if mode == "p2p" else 0
    torch.save({"losses": losses, "params": {k: v.cpu() for k, v in tr.params.items()}, "err": err}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_graph", [False, True])
def test_p2p_exchange_matches_single_process_and_nccl(tmp_path, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    B, L, world = 256, 64, 2
    res = {}
    for mode in ("p2p", "nccl"):
        out = str(tmp_path / f"dp_{mode}.pt")
        mp.spawn(_worker, args=(world, _free_port(), B, L, out, use_graph, mode), nprocs=world, join=True)
        res[mode] = [torch.load(f"{out}.{r}") for r in range(world)]
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    image, text, noise = _data(B, L)
    ref = MnistMVAETrainer(L, B, seed=0, use_graph=False)
    ref_losses = [ref.step(image, text, annealing_factor=0.25 * (it + 1), noise=noise) for it in range(4)]
    ref_losses.append(ref.step(image, text, annealing_factor=1.0, noise=noise, update=False))
    for mode in ("p2p", "nccl"):
        r0, r1 = res[mode]
        assert r0["err"] == 0 and r1["err"] == 0
        for it, l in enumerate(ref_losses):
            assert abs(l - r0["losses"][it]) <= 2e-6 * abs(l), (mode, it, l, r0["losses"][it])
            assert r0["losses"][it] == r1["losses"][it]          # every rank reports the same global loss
        for k, v in ref.params.items():
            assert torch.equal(r0["params"][k], r1["params"][k]), k       # replicas stay bit-identical
            assert (v.cpu() - r0["params"][k]).abs().max().item() <= 1e-4, (mode, k)
