"""Per-launch time of every GEMM launch of one MNIST step (eager, CUDA events) next to its unit count
(unit = one 128x128x32 k-block tile) -> cycles per unit per SM.   python tools/chain_times.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402
from multimodal_vae_public_b200.trainer import MnistMVAETrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tr = MnistMVAETrainer(64, B, use_graph=False)
g = torch.Generator().manual_seed(0)
im = torch.rand(B, 784, generator=g).cuda(); tx = torch.randint(0, 10, (B,), generator=g).cuda()
for _ in range(5):
    tr.step(im, tx)
recs = []
orig_chain, orig_batch = ops.gemm_chain, ops.gemm_batch


def units(descs):
    u = t = 0
    for d in descs:
        bn = 128 if d.N >= 128 else d.N
        tiles = -(-d.M // 128) * -(-d.N // 128)
        kb = -(-d.K // 32)
        u += tiles * kb * (min(d.N, 128) / 128 if d.N < 128 else 1)
        t += tiles * max(1, d.split_k)
    return u, t


def wrap(fn, name):
    def f(descs, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(descs, *a, **k); e.record()
        recs.append((name, len(descs), units(descs), s, e))
        return r
    return f


ops.gemm_chain, ops.gemm_batch = wrap(orig_chain, "chain"), wrap(orig_batch, "batch")
reps = 10
for _ in range(reps):
    tr.step(im, tx)
torch.cuda.synchronize()
n = len(recs) // reps
sm, mhz = 148, 1965.0
tot = 0.0
for i in range(n):
    ms = sum(recs[j * n + i][3].elapsed_time(recs[j * n + i][4]) for j in range(reps)) / reps
    name, np_, (u, t), _, _ = recs[i]
    tot += ms
    print(f"{name} #{i}: {np_:2d} problems, {t:5d} tiles, {u:9.0f} units -> {ms * 1e3:7.1f} us, "
          f"{ms * 1e-3 * mhz * 1e6 * sm / u:7.0f} cycles/unit/SM")
print(f"total {tot * 1e3:.1f} us")
