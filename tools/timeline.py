"""Per-role clock64 timeline of the first CTAs of one GEMM launch (debug aid; MVAE_DBG_TIMELINE)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
kind = sys.argv[2] if len(sys.argv) > 2 else "fwd"      # fwd | dgrad | wgrad
M, N, K = 8192, 512, 512
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda"); h = torch.empty(M, N, device="cuda")
dy = torch.randn(M, N, device="cuda"); dx = torch.empty(M, K, device="cuda"); dw = torch.zeros(N, K, device="cuda")


def run():
    if kind == "fwd":
        ops.linear_fwd(x, w, b, y, h, precision=prec)
    elif kind == "dgrad":
        ops.linear_dgrad(dy, w, dx, a_prev=x, precision=prec)
    else:
        ops.linear_wgrad(dy, x, dw, split_k=16, precision=prec)


for _ in range(20):
    run()
torch.cuda.synchronize()
dbg = torch.zeros(8 * 6 * 64, dtype=torch.int64, device="cuda")
os.environ["MVAE_DBG_TIMELINE"] = str(dbg.data_ptr())
run()
torch.cuda.synchronize()
del os.environ["MVAE_DBG_TIMELINE"]
d = dbg.cpu().view(8, 6, 64)
for blk in (0, 1):
    if blk >= 8:
        continue
    t0 = int(d[blk][d[blk] > 0].min())
    for role, name in enumerate(("producer", "mma", "epilogue", "epi-chunk", "splitter")):
        ev = [int(v) - t0 for v in d[blk, role] if v > 0]
        print(f"block {blk} {name:9s} n={len(ev):2d}:", " ".join(str(e) for e in ev))
        if name == "splitter":   # per k-block: [loop top, full seen, A in TMEM issued, B lo written, st waited + fenced]
            rows = [ev[i:i + 5] for i in range(0, len(ev) - 4, 5)]
            print("   per k-block (wait full | split A | split B | wait st+fence | to next):",
                  " ; ".join("/".join(str(b - a) for a, b in zip(r, r[1:])) for r in rows[:12]))
