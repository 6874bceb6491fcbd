"""Where do the N = 64 sub-pixel GEMMs lose their time?  CUDA-event A/B of one ConvT(128 -> 64, 7x7 -> 14x14) layer at the
FashionMNIST decoder size (2B = 8192 images) in several forms that each remove one suspect:

  dense64      4 dense problems 401408 x 64 x 512 (plain row-major A, no view, no row map)        -> cost of N = 64 tiles as such
  dense128     2 dense problems 401408 x 128 x 512 (same FLOPs)                                   -> the N = 128 yardstick
  view64       4 problems with the 2x2 stride-1 im2col A view + tap-split B, rows stored densely   -> cost of the implicit A
  subpixel     the product form (view + tap split + row map), as a chain like the trainer
  subpixel_b   the same through mvae_gemm_batch
  + each of the above with MVAE_DBG_EPI=1 (global stores skipped) when run as  diag_subpixel.py noepi

   python tools/diag_subpixel.py [B2] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multimodal_vae_public_b200 import ops  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    args = [a for a in sys.argv[1:] if a != "noepi"]
    n = int(args[0]) if args else 8192
    reps = int(args[1]) if len(args) > 1 else 5
    IH = IW = 7; Cx, Cy = 128, 64
    M = n * IH * IW
    dev = "cuda"
    x = torch.randn(n, IH, IW, Cx, device=dev)
    w = torch.randn(16 * Cy, Cx, device=dev) * 0.05            # Wt [(kh,kw,cy)][cx]
    out = torch.empty(n * 4 * IH * IW, Cy, device=dev); out2 = torch.empty_like(out)
    xd = torch.randn(M, 4 * Cx, device=dev)                    # a dense A of the same K
    wd64 = torch.randn(Cy, 4 * Cx, device=dev) * 0.05; wd128 = torch.randn(2 * Cy, 4 * Cx, device=dev) * 0.05
    o64 = [torch.empty(M, Cy, device=dev) for _ in range(8)]
    o128 = [torch.empty(M, 2 * Cy, device=dev) for _ in range(4)]
    ws = ops.chain_workspace(dev)
    P = ops.PREC_3XTF32
    fl = 2.0 * M * 4 * Cy * 4 * Cx
    E = ops.EPI_BIAS_SWISH

    def dense64():
        ops.gemm_batch([ops.gemm_desc(xd, wd64, o64[i], M, Cy, 4 * Cx, out2=o64[4 + i], epilogue=E) for i in range(4)], P)

    def dense128():
        ops.gemm_batch([ops.gemm_desc(xd, wd128, o128[i], M, 2 * Cy, 4 * Cx, out2=o128[2 + i], epilogue=E) for i in range(2)], P)

    def view64():
        descs = ops.subpixel_k4s2p1(x, w, out, n, IH, IW, Cx, Cy, out2=out2, epilogue=E)
        for i, d in enumerate(descs):          # drop the row map: class i stores its rows densely
            d.rowmap_IH = d.rowmap_IW = d.rowmap_s = d.rowmap_py = d.rowmap_px = 0
            d.C, d.ldc, d.out2, d.ldout2 = o64[i].data_ptr(), Cy, o64[4 + i].data_ptr(), Cy
        ops.gemm_batch(descs, P)

    def subpixel():
        ops.gemm_chain(ops.subpixel_k4s2p1(x, w, out, n, IH, IW, Cx, Cy, out2=out2, epilogue=E), [-1] * 4, ws, P)

    def subpixel_b():
        ops.gemm_batch(ops.subpixel_k4s2p1(x, w, out, n, IH, IW, Cx, Cy, out2=out2, epilogue=E), P)

    def subpixel_store():
        ops.gemm_batch(ops.subpixel_k4s2p1(x, w, out, n, IH, IW, Cx, Cy, epilogue=ops.EPI_STORE), P)

    def dense64_tf32():
        ops.gemm_batch([ops.gemm_desc(xd, wd64, o64[i], M, Cy, 4 * Cx, out2=o64[4 + i], epilogue=E) for i in range(4)],
                       ops.PREC_TF32)

    def subpixel_tf32():
        ops.gemm_batch(ops.subpixel_k4s2p1(x, w, out, n, IH, IW, Cx, Cy, out2=out2, epilogue=E), ops.PREC_TF32)

    print("MVAE_DBG_EPI =", os.environ.get("MVAE_DBG_EPI"), " M =", M)
    for name, fn in (("dense64", dense64), ("dense128", dense128), ("view64", view64), ("subpixel", subpixel),
                     ("subpixel_b", subpixel_b), ("subpixel_store", subpixel_store), ("dense64_tf32", dense64_tf32),
                     ("subpixel_tf32", subpixel_tf32)):
        us = timed(fn, reps)
        print(f"{name:16s} {us:9.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s  {us * 1.9e3 / (M / 128 * 4 / 148) / 16:7.0f} cyc/k-block @1.9GHz")

    # clock64 stamps per warp role of CTA 0 (MVAE_DBG_TIMELINE): per-k-block cadence of the TMA producer and the MMA issuer,
    # epilogue wait / work spans
    for name, fn in (("dense64", dense64), ("dense128", dense128), ("view64", view64), ("subpixel_b", subpixel_b)):
        dbg = torch.zeros(8 * 6 * 64, dtype=torch.int64, device="cuda")
        os.environ["MVAE_DBG_TIMELINE"] = str(dbg.data_ptr())
        fn(); torch.cuda.synchronize()
        del os.environ["MVAE_DBG_TIMELINE"]
        d = dbg.cpu().view(8, 6, 64)[0]
        t0 = int(d[d > 0].min())
        print(f"--- timeline {name} (CTA 0, cycles from first stamp)")
        for role, rn in enumerate(("producer", "mma", "epilogue", "epi-chunk", "splitter")):
            ev = [int(v) - t0 for v in d[role] if v > 0]
            if rn in ("producer", "mma"):
                print(f"  {rn:9s} n={len(ev):2d} deltas:", " ".join(str(b - a) for a, b in zip(ev, ev[1:])))
            elif rn == "splitter":
                rows = [ev[i:i + 5] for i in range(0, len(ev) - 4, 5)]
                print("  splitter per k-block (wait full | split A | split B | wait st+fence | to next):",
                      " ; ".join("/".join(str(b - a) for a, b in zip(r, r[1:])) for r in rows[:12]))
            else:
                print(f"  {rn:9s} n={len(ev):2d}:", " ".join(str(e) for e in ev))


if __name__ == "__main__":
    main()
