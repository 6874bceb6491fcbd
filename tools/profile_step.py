"""Run a few eager (non-graph) training steps / single kernels for ncu captures.
   python tools/profile_step.py step [--steps N] [--batch B] [--prec 0|1]
   python tools/profile_step.py gemm  fwd|dgrad|wgrad  prec  M N K [iters]
   python tools/profile_step.py bce R D [iters]
   python tools/profile_step.py poe B L [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402


def main():
    mode = sys.argv[1]
    if mode == "step":
        import argparse
        ap = argparse.ArgumentParser(); ap.add_argument("mode"); ap.add_argument("--steps", type=int, default=4)
        ap.add_argument("--batch", type=int, default=4096); ap.add_argument("--prec", type=int, default=1)
        a = ap.parse_args()
        from multimodal_vae_public_b200.trainer import MnistMVAETrainer
        tr = MnistMVAETrainer(64, a.batch, precision=a.prec, use_graph=False)
        g = torch.Generator().manual_seed(0)
        im = torch.rand(a.batch, 784, generator=g).cuda(); tx = torch.randint(0, 10, (a.batch,), generator=g).cuda()
        for i in range(a.steps):
            loss = tr.step(im, tx, annealing_factor=0.5)
        print("loss", loss, "launches/step", tr.launches_per_step)
    elif mode == "gemm":
        kind, prec, M, N, K = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
        iters = int(sys.argv[7]) if len(sys.argv) > 7 else 10
        x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
        y = torch.empty(M, N, device="cuda"); h = torch.empty(M, N, device="cuda")
        dy = torch.randn(M, N, device="cuda"); dw = torch.zeros(N, K, device="cuda"); dx = torch.empty(M, K, device="cuda")
        for _ in range(iters):
            if kind == "fwd": ops.linear_fwd(x, w, b, y, h, precision=prec)
            elif kind == "dgrad": ops.linear_dgrad(dy, w, dx, a_prev=x, precision=prec)
            else: ops.linear_wgrad(dy, x, dw, split_k=8, precision=prec)
        torch.cuda.synchronize()
    elif mode == "bce":
        R, D = int(sys.argv[2]), int(sys.argv[3]); iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
        x = torch.randn(R, D, device="cuda"); t = torch.rand(R // 2, D, device="cuda"); dx = torch.empty_like(x)
        acc = torch.zeros(2, dtype=torch.float64, device="cuda")
        for _ in range(iters):
            ops.bce_logits_fwd_bwd(x, t, dx, 1e-3, acc, seg_rows=R // 2)
        torch.cuda.synchronize()
    elif mode == "poe":
        # roofline-size run of the fused PoE + reparametrise + KL kernels (MNIST pass structure): poe B L [iters]
        B, L = int(sys.argv[2]), int(sys.argv[3]); iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
        enc = [torch.randn(B, 2 * L, device="cuda") * 0.5 for _ in range(2)]
        mu_e = [e[:, :L] for e in enc]; lv_e = [e[:, L:] for e in enc]
        z = torch.empty(3 * B, L, device="cuda"); nz = torch.empty(3 * B, L, device="cuda"); dz = torch.randn(3 * B, L, device="cuda")
        d_enc = [torch.empty(B, 2 * L, device="cuda") for _ in range(2)]
        kl = torch.zeros(3, dtype=torch.float64, device="cuda"); stepc = torch.zeros(1, dtype=torch.int32, device="cuda")
        for _ in range(iters):
            ops.poe_fwd(mu_e, lv_e, (1, 3, 2), B, L, z, variant=0, training=True, noise=None, noise_out=nz, seed=1, step_dev=stepc,
                        kl_acc=kl)
            ops.poe_bwd(mu_e, lv_e, (1, 3, 2), B, L, dz, [d[:, :L] for d in d_enc], [d[:, L:] for d in d_enc], kl_scale=1.0 / B,
                        variant=0, training=True, noise=nz)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
