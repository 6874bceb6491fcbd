"""Gantt analysis of the chained GEMM launches of one MNIST step (MVAE_DBG_TILELOG).
   python tools/tile_gantt.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402
from multimodal_vae_public_b200.trainer import MnistMVAETrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tr = MnistMVAETrainer(64, B, use_graph=False)
g = torch.Generator().manual_seed(0)
im = torch.rand(B, 784, generator=g).cuda(); tx = torch.randint(0, 10, (B,), generator=g).cuda()
for _ in range(5):
    tr.step(im, tx)
torch.cuda.synchronize()
logs = []
orig = ops.gemm_chain
pair = os.environ.get("MVAE_PAIR", "0") != "0"


def wrapped(descs, deps, ws, prec):
    buf = torch.full((148 * 64 * 4,), -1, dtype=torch.int64, device="cuda")
    os.environ["MVAE_DBG_TILELOG"] = str(buf.data_ptr())
    try:
        orig(descs, deps, ws, prec)
    finally:
        del os.environ["MVAE_DBG_TILELOG"]
    tm = 256 if pair else 128
    tiles, names = [], []
    for i, d in enumerate(descs):
        bn = 128 if d.N >= 128 else d.N
        n = -(-d.M // tm) * -(-d.N // bn) * max(1, d.split_k if d.split_k <= -(-d.K // 32) else -(-d.K // 32))
        tiles.append(n)
        kind = "wgrad" if d.a_mn_major else ("dgrad" if d.b_mn_major else "fwd")
        names.append(f"p{i}:{kind} {d.M}x{d.N}x{d.K}" + (f" dep{deps[i]}" if deps[i] >= 0 else ""))
    logs.append((buf, tiles, names))


ops.gemm_chain = wrapped
tr.step(im, tx)
torch.cuda.synchronize()
for ci, (buf, tiles, names) in enumerate(logs):
    a = buf.cpu().numpy().reshape(148, 64, 4)
    valid = a[:, :, 0] >= 0
    t0 = a[:, :, 1][valid].min()
    end = a[:, :, 3][valid].max()
    print(f"=== chain {ci}: {valid.sum()} tiles logged, span {(end - t0) / 1e3:.1f} us")
    starts = np.cumsum([0] + tiles)
    tidx = a[:, :, 0]
    for pi, nm in enumerate(names):
        m = valid & (tidx >= starts[pi]) & (tidx < starts[pi + 1])
        if not m.any():
            print(f"  {nm}: no tiles logged"); continue
        b, r, e = a[:, :, 1][m], a[:, :, 2][m], a[:, :, 3][m]
        print(f"  {nm:34s} tiles {m.sum():4d}  begin {(b.min() - t0) / 1e3:6.1f}..{(b.max() - t0) / 1e3:6.1f} us  end "
              f"{(e.min() - t0) / 1e3:6.1f}..{(e.max() - t0) / 1e3:6.1f} us  dep-wait mean {(r - b).mean() / 1e3:5.2f} max "
              f"{(r - b).max() / 1e3:5.1f} us  tile (deps->epilogue end) mean {(e - r).mean() / 1e3:5.1f} us")
    # per-CTA: time of last epilogue end, number of tiles
    last = np.where(valid, a[:, :, 3], 0).max(axis=1)
    ncta = (valid.sum(axis=1) > 0).sum()
    lastv = last[valid.sum(axis=1) > 0]
    print(f"  CTAs {ncta}: finish time min {(lastv.min() - t0) / 1e3:.1f} / mean {(lastv.mean() - t0) / 1e3:.1f} / max "
          f"{(lastv.max() - t0) / 1e3:.1f} us; total dep-wait {((a[:, :, 2] - a[:, :, 1])[valid]).sum() / 1e3 / ncta:.1f} us per CTA")
