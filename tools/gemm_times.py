"""Per-launch time of every GEMM launch (and every other library call) of one step of a trainer, eager, CUDA events.
   python tools/gemm_times.py mnist|fashion|celeba|celeba19 [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "celeba"
B = int(sys.argv[2]) if len(sys.argv) > 2 else {"mnist": 4096, "fashion": 4096, "celeba": 1024, "celeba19": 512}[wl]
g = torch.Generator().manual_seed(0)
if wl in ("mnist", "fashion"):
    if wl == "mnist":
        from multimodal_vae_public_b200.trainer import MnistMVAETrainer as T
    else:
        from multimodal_vae_public_b200.trainer_fashion import FashionMVAETrainer as T
    tr = T(64, B, use_graph=False)
    args = (torch.rand(B, 1, 28, 28, generator=g).cuda(), torch.randint(0, 10, (B,), generator=g).cuda())
elif wl == "celeba":
    from multimodal_vae_public_b200.trainer_celeba import CelebAMVAETrainer as T
    tr = T(100, B, use_graph=False)
    args = (torch.rand(B, 3, 64, 64, generator=g).cuda(), torch.randint(0, 2, (B, 18), generator=g).float().cuda())
else:
    from multimodal_vae_public_b200.trainer_celeba19 import CelebA19MVAETrainer as T
    np.random.seed(1)
    tr = T(100, B, approx_m=1)
    args = (torch.rand(B, 3, 64, 64, generator=g).cuda(), torch.randint(0, 2, (B, 18), generator=g).float().cuda())
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
recs = []
names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and n not in
         ("gemm_desc", "chain_workspace", "GemmDesc", "Optional", "Sequence")]
orig = {n: getattr(ops, n) for n in names}


def describe(n, a):
    if n in ("gemm_batch", "gemm_chain"):
        out = []
        for d in a[0]:
            kind = "wgrad" if d.a_mn_major else ("dgrad" if d.b_mn_major else "fwd")
            out.append(f"{kind} {d.M}x{d.N}x{d.K}" + (f"/s{d.split_k}" if d.split_k > 1 else ""))
        return "; ".join(out), sum(2.0 * d.M * d.N * d.K for d in a[0])
    shp = [tuple(t.shape) for t in a if isinstance(t, torch.Tensor)][:2]
    return str(shp), 0.0


def wrap(n):
    def f(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig[n](*a, **k); e.record()
        recs.append((n, describe(n, a), s, e))
        return r
    return f


for n in names:
    if isinstance(orig[n], type(wrap)):
        setattr(ops, n, wrap(n))
tr.step(*args)
torch.cuda.synchronize()
tot = {}
for n, (desc, fl), s, e in recs:
    ms = s.elapsed_time(e)
    tot[n] = tot.get(n, 0.0) + ms
    if ms > float(os.environ.get("MVAE_TIMES_MIN_MS", "0.05")):
        print(f"{n:22s} {ms * 1e3:8.1f} us  {fl / ms / 1e9 if fl else 0:7.1f} TFLOP/s  {desc[:150]}")
print({k: round(v, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
