"""GPU probe: run one named check of the raw kernels and print a JSON line (never asserts).
Used under gpurun to learn as much as possible from one call:  python tools/probe.py <case> | all"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gemm_case(kind, M, N, K, prec, seed=0, split_k=1):
    import torch
    from multimodal_vae_public_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    if kind == "fwd":      # y[M,N] = x[M,K] w[N,K]^T + b ; h = swish(y)
        x = torch.randn(M, K, device=dev, generator=g); w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
        b = torch.randn(N, device=dev, generator=g)
        ldn = (N + 3) // 4 * 4
        y = torch.full((M, ldn), float("nan"), device=dev)[:, :N]; h = torch.full((M, ldn), float("nan"), device=dev)[:, :N]
        ops.linear_fwd(x, w, b, y, h, precision=prec)
        torch.cuda.synchronize()
        ref = x.double() @ w.double().t() + b.double()
        refh = ref * torch.sigmoid(ref)
        out = {"y": (y.double() - ref).abs().max().item() / ref.abs().max().item(),
               "h": (h.double() - refh).abs().max().item() / refh.abs().max().item()}
    elif kind == "dgrad":  # dx[M,K] = dy[M,N] w[N,K] * swish'(a[M,K])
        ldn = (N + 3) // 4 * 4
        dy = torch.randn(M, ldn, device=dev, generator=g)[:, :N]; w = torch.randn(N, K, device=dev, generator=g) / N ** 0.5
        a = torch.randn(M, K, device=dev, generator=g)
        dx = torch.full((M, K), float("nan"), device=dev)
        ops.linear_dgrad(dy, w, dx, a_prev=a, precision=prec)
        torch.cuda.synchronize()
        s = torch.sigmoid(a.double())
        ref = (dy.double() @ w.double()) * (s * (1 + a.double() * (1 - s)))
        out = {"dx": (dx.double() - ref).abs().max().item() / ref.abs().max().item()}
    elif kind == "wgrad":  # dw[N,K] += dy[M,N]^T x[M,K]
        ldn = (N + 3) // 4 * 4
        dy = torch.randn(M, ldn, device=dev, generator=g)[:, :N]; x = torch.randn(M, K, device=dev, generator=g)
        dw = torch.zeros(N, K, device=dev)
        ops.linear_wgrad(dy, x, dw, split_k=split_k, precision=prec)
        torch.cuda.synchronize()
        ref = dy.double().t() @ x.double()
        out = {"dw": (dw.double() - ref).abs().max().item() / ref.abs().max().item()}
    return out


def trunc_case():
    """Does kind::tf32 truncate or round fp32 operands?  Compare raw vs pre-truncated vs pre-rounded inputs."""
    import torch
    from multimodal_vae_public_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 256, 128, 256
    x = torch.randn(M, K, device="cuda", generator=g); w = torch.randn(N, K, device="cuda", generator=g)
    b = torch.zeros(N, device="cuda")

    def run(xx, ww):
        y = torch.empty(M, N, device="cuda")
        ops.linear_fwd(xx, ww, b, y, None, precision=ops.PREC_TF32)
        torch.cuda.synchronize()
        return y

    def trunc(t):
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

    def rna(t):
        i = t.view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    y_raw = run(x, w); y_tr = run(trunc(x), trunc(w)); y_rn = run(rna(x), rna(w))
    return {"raw_eq_trunc": bool(torch.equal(y_raw, y_tr)), "raw_eq_rna": bool(torch.equal(y_raw, y_rn)),
            "max_raw_vs_trunc": (y_raw - y_tr).abs().max().item(), "max_raw_vs_rna": (y_raw - y_rn).abs().max().item()}


def time_case(M, N, K, prec, kind="fwd", iters=20):
    import torch
    from multimodal_vae_public_b200 import ops
    dev = "cuda"
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev); h = torch.empty(M, N, device=dev)
    dy = torch.randn(M, N, device=dev); dw = torch.zeros(N, K, device=dev); dx = torch.empty(M, K, device=dev)

    def go():
        if kind == "fwd": ops.linear_fwd(x, w, b, y, h, precision=prec)
        elif kind == "dgrad": ops.linear_dgrad(dy, w, dx, a_prev=x, precision=prec)
        else: ops.linear_wgrad(dy, x, dw, split_k=max(1, 148 // (((N + 127) // 128) * ((K + 127) // 128))), precision=prec)
    import time as _t
    t0 = _t.time()
    while _t.time() - t0 < 0.5:   # let the clocks ramp up before timing
        for _ in range(50): go()
        torch.cuda.synchronize()
    iters = 200
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): go()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    return {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}


CASES = {}
for prec, pn in ((0, "tf32"), (1, "3x")):
    CASES[f"fwd_min_{pn}"] = lambda prec=prec: gemm_case("fwd", 128, 128, 32, prec)
    CASES[f"fwd_k2_{pn}"] = lambda prec=prec: gemm_case("fwd", 128, 128, 64, prec)
    CASES[f"fwd_mnist1_{pn}"] = lambda prec=prec: gemm_case("fwd", 512, 512, 784, prec)
    CASES[f"fwd_small_n_{pn}"] = lambda prec=prec: gemm_case("fwd", 300, 10, 512, prec)
    CASES[f"fwd_n64_{pn}"] = lambda prec=prec: gemm_case("fwd", 1000, 128, 512, prec)
    CASES[f"fwd_big_{pn}"] = lambda prec=prec: gemm_case("fwd", 8192, 512, 512, prec)
    CASES[f"dgrad_min_{pn}"] = lambda prec=prec: gemm_case("dgrad", 128, 128, 128, prec)
    CASES[f"dgrad_mnist_{pn}"] = lambda prec=prec: gemm_case("dgrad", 1024, 784, 512, prec)
    CASES[f"dgrad_n10_{pn}"] = lambda prec=prec: gemm_case("dgrad", 512, 10, 512, prec)
    CASES[f"dgrad_k64_{pn}"] = lambda prec=prec: gemm_case("dgrad", 512, 512, 64, prec)
    CASES[f"wgrad_min_{pn}"] = lambda prec=prec: gemm_case("wgrad", 32, 128, 128, prec)
    CASES[f"wgrad_mnist_{pn}"] = lambda prec=prec: gemm_case("wgrad", 2048, 512, 784, prec, split_k=5)
    CASES[f"wgrad_n10_{pn}"] = lambda prec=prec: gemm_case("wgrad", 1000, 10, 512, prec, split_k=3)
    CASES[f"wgrad_k64_{pn}"] = lambda prec=prec: gemm_case("wgrad", 4096, 512, 64, prec, split_k=8)
    for kind in ("fwd", "dgrad", "wgrad"):
        CASES[f"time_{kind}_{pn}"] = lambda prec=prec, kind=kind: time_case(8192, 512, 512, prec, kind)
CASES["trunc"] = trunc_case


MN_SWEEP = {
    "default": {},
    "swap_lbo_sbo": {"MVAE_DBG_MN_LBO": "512", "MVAE_DBG_MN_SBO": "4096"},
    "flip8b": {"MVAE_DBG_MN_SWIZZLE": "5"},
    "sw128_16B_layout2": {"MVAE_DBG_MN_SWIZZLE": "3", "MVAE_DBG_MN_LAYOUT": "2", "MVAE_DBG_MN_SBO": "1024"},
    "sw128_16B_layout1": {"MVAE_DBG_MN_SWIZZLE": "3", "MVAE_DBG_MN_LAYOUT": "1"},
}


def mn_sweep():
    for tag, env in MN_SWEEP.items():
        for case in ("dgrad_min_tf32", "wgrad_min_tf32", "dgrad_mnist_tf32", "wgrad_mnist_tf32"):
            e = dict(os.environ); e.update(env)
            try:
                r = subprocess.run([sys.executable, __file__, case], capture_output=True, text=True, timeout=120, env=e)
                line = (r.stdout.strip().splitlines() or ["<no output> " + r.stderr[-300:]])[-1]
            except subprocess.TimeoutExpired:
                line = "TIMEOUT"
            print("mnsweep", tag, line, flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "mnsweep":
        return mn_sweep()
    if which == "all":
        only = sys.argv[2] if len(sys.argv) > 2 else ""
        for name in CASES:
            if only and only not in name:
                continue
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
                line = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                if r.returncode != 0:
                    line = json.dumps({"case": name, "rc": r.returncode, "stderr": r.stderr[-600:]})
            except subprocess.TimeoutExpired:
                line = json.dumps({"case": name, "error": "TIMEOUT (hang)"})
            print(line, f"[{time.time() - t0:.1f}s]", flush=True)
        return
    try:
        out = CASES[which]()
        print(json.dumps({"case": which, **out}))
    except Exception as ex:  # noqa: BLE001
        print(json.dumps({"case": which, "error": repr(ex)[:500]}))


if __name__ == "__main__":
    main()
