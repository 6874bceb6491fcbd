#!/usr/bin/env python
"""Benchmark of the MVAE training-step hot path (BASELINE.json metric: MVAE train samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mnist|fashion|celeba|celeba19]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one iteration of the reference training loop (mnist/train.py:196-219) on one synthetic batch: 3 forwards,
3 ELBO terms, backward, Adam.  Headline workload = BASELINE.json configs[1] (MNIST MVAE, n_latents=64, batch 4096); at
every N the same invocation ALSO runs configs[2] (FashionMNIST, batch 4096 -- the configuration the strong-scaling target
is quoted on) and reports it under "extra" in the same JSON line.  Prints ONE JSON line (rank 0).

--impl reference times the reference's own CPU implementation of the step on the host cores (oracle/_ref: byte-identical
copy of the reference modules driven by oracle/ref_harness.py, kind "reference"; the oracle port, kind "port", only if
that copy is missing).  It is the reported CPU baseline, never the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LATENTS = 64
BATCH = 4096
POOL = 16                      # rotating pool of distinct input batches (> L2: 16 x 12.9 MB = 206 MB)
LAMBDA_IMAGE, LAMBDA_TEXT = 1.0, 10.0
ANNEAL_EPOCHS, N_MINI = 200, 15  # KL annealing schedule of mnist/train.py:180-186 with 60000/4096 ~ 15 batches/epoch

# DRAM traffic per launch from the committed round-2 ncu captures (profiles/r02_*_raw_key_metrics.txt): the four chained GEMM
# launches of one MNIST B=4096 step moved 16.1 + 202.6 + 492.9 + 54.9 MB; the roofline-size BCE launch 308.3 MB read +
# 152.3 MB written; the roofline-size PoE launches 268.5 + 352.8 MB (forward) and 671.1 + 247.7 MB (backward)
GEMM_CHAIN_DRAM_BYTES_PER_LAUNCH = (16.14e6 + 202.64e6 + 492.93e6 + 54.89e6) / 4
BCE_DRAM_BYTES_PER_LAUNCH = 308.29e6 + 152.33e6
POE_FWD_DRAM_BYTES_PER_LAUNCH = 268.48e6 + 352.78e6
POE_BWD_DRAM_BYTES_PER_LAUNCH = 671.12e6 + 247.67e6

# algorithmic work per sample per step
FLOP_PER_SAMPLE_REFERENCE = 32.4e6     # everything the reference executes (incl. dead decoder passes, duplicate encoders)
ELEMENTWISE_BYTES_PER_SAMPLE = 30512   # K1 fwd/bwd + K2 image/label terms
POE_FWD_BYTES_PER_SAMPLE, POE_BWD_BYTES_PER_SAMPLE = 4352, 7168     # SURVEY 8d, MNIST passes M = (2, 1, 1), L = 64

# per workload: BASELINE.json config, default global batch, n_latents, Adam lr, untimed clock-ramp steps, CPU sample batch
WORKLOADS = {
    "mnist": dict(cfg="configs[1]", batch=4096, L=64, lr=1e-3, ramp=400, cpu_batch=4096,
                  name="MNIST MVAE (image 28x28x1 + label one-of-10)", passes="3"),
    "fashion": dict(cfg="configs[2]", batch=4096, L=64, lr=1e-3, ramp=60, cpu_batch=512,
                    name="FashionMNIST (conv enc/dec) MVAE (image 28x28x1 + label one-of-10)", passes="3"),
    "celeba": dict(cfg="configs[3]", batch=1024, L=100, lr=1e-4, ramp=40, cpu_batch=64,
                   name="CelebA MVAE (image 64x64x3 + 18 attrs, conv+BatchNorm+Dropout)", passes="3"),
    "celeba19": dict(cfg="configs[4]", batch=512, L=100, lr=1e-4, ramp=20, cpu_batch=16,
                     name="CelebA-19 MVAE (image 64x64x3 + 18 single-attribute experts, 20 + approx_m=1 ELBO terms per step, "
                          "subsets re-sampled every step)", passes="21"),
}


def workload_string(wl: str, b_global: int) -> str:
    """The SAME text in both arms (ours / reference), so the driver's same-config check compares like with like."""
    w = WORKLOADS[wl]
    return (f"{w['name']}, n_latents={w['L']}, global batch {b_global}, full train step ({w['passes']} passes + ELBO + "
            f"backward + Adam), BASELINE.json {w['cfg']}")


def executed_gemm_flops_per_sample(L: int = N_LATENTS, label_table: bool = True) -> float:
    """2*MAC of the GEMMs this implementation launches per sample (fwd + dgrad + wgrad), see DESIGN.md.  With the label
    encoder evaluated on its 10-row class table (the default) its GEMMs are gone: O(10) rows, CUDA cores."""
    enc_i = 784 * 512 + 512 * 512 + 512 * 2 * L
    enc_t = 0 if label_table else 512 * 512 + 512 * 2 * L   # embedding row gather is not a GEMM
    dec_i = L * 512 + 512 * 512 * 2 + 512 * 784
    dec_t = L * 512 + 512 * 512 * 2 + 512 * 10
    fwd = enc_i + enc_t + 2 * (dec_i + dec_t)              # each decoder sees 2 of the 3 passes
    wgrad = fwd
    dgrad = (enc_i - 784 * 512) + (enc_t) + 2 * (dec_i + dec_t)   # no dgrad into the image; text enc fc2 dgrad kept
    return 2.0 * (fwd + wgrad + dgrad)


def fashion_gemm_flops_per_sample(L: int = N_LATENTS, label_table: bool = True) -> float:
    """2*MAC of the GEMMs the FashionMNIST-flavour trainer launches per sample (fwd + wgrad + dgrad)."""
    enc_i = 196 * 64 * 16 + 49 * 128 * 1024 + 6272 * 512 + 512 * 2 * L
    enc_t = 0 if label_table else 512 * 512 + 512 * 2 * L
    dec_i = L * 512 + 512 * 6272 + 49 * 128 * 1024 + 196 * 64 * 16
    dec_t = L * 512 + 512 * 512 * 2 + 512 * 10
    fwd = enc_i + enc_t + 2 * (dec_i + dec_t)
    dgrad = (enc_i - 196 * 64 * 16) + enc_t + 2 * (dec_i + dec_t)
    return 2.0 * (2 * fwd + dgrad)


def celeba_gemm_flops_per_sample(L: int = 100) -> float:
    """2*MAC of the GEMMs the CelebA-flavour trainer launches per sample: encoders once, decoders forward on the three
    passes, backward on the two live passes of each decoder."""
    enc_conv = 1024 * 32 * 48 + 256 * 64 * 512 + 64 * 128 * 1024 + 25 * 256 * 2048
    enc_i = enc_conv + 6400 * 512 + 2 * 512 * 2 * L
    enc_a = 20 * 512 + 512 * 512 + 512 * 2 * L
    dec_i = L * 6400 + 25 * 2048 * 256 + 64 * 1024 * 128 + 256 * 512 * 64 + 1024 * 48 * 32
    dec_a = L * 512 + 2 * 512 * 512 + 512 * 18
    fwd = enc_i + enc_a + 3 * (dec_i + dec_a)
    bwd = 2 * (enc_i + enc_a) - 1024 * 32 * 48 - 20 * 512 + 2 * 2 * (dec_i + dec_a)
    return 2.0 * (fwd + bwd)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed regions run (every ~2 ms: a timed
    region of 20 MNIST steps is only ~14 ms long)."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synth_host_batches(wl: str, n: int, batch: int, seed: int = 0):
    """Synthetic data of the workload's shape: image ~ U[0,1) fp32; label ~ randint(0,10) int64 or
    attrs ~ randint(0,2) float [B,18]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    if wl in ("mnist", "fashion"):
        return [(torch.rand(batch, 1, 28, 28, generator=g), torch.randint(0, 10, (batch,), generator=g)) for _ in range(n)]
    return [(torch.rand(batch, 3, 64, 64, generator=g), torch.randint(0, 2, (batch, 18), generator=g).float())
            for _ in range(n)]


def synth_batches(n: int, batch: int, seed: int = 0):
    return synth_host_batches("mnist", n, batch, seed)


def annealing(step: int) -> float:
    epoch, idx = 1 + step // N_MINI, step % N_MINI
    return min(1.0, float(idx + (epoch - 1) * N_MINI + 1) / float(ANNEAL_EPOCHS * N_MINI))


# ------------------------------------------------------------------------------------------ CPU / eager-GPU baselines
def _reference_stepper(wl: str, device: str = "cpu"):
    """(step(image, other, beta) -> loss, kind): the reference's own modules when oracle/_ref is there, else the port."""
    import torch
    w = WORKLOADS[wl]
    from oracle import ref_harness
    if ref_harness.available():
        rs = ref_harness.RefStep(wl, w["L"], w["lr"], LAMBDA_IMAGE, LAMBDA_TEXT, device=device, approx_m=1, seed=0)
        return (lambda im, ot, beta: rs.step(im, ot, beta)), "reference"
    if wl != "mnist" or device != "cpu":
        return None, "unavailable"
    from oracle.mvae_oracle import MnistCpuBaseline
    m = MnistCpuBaseline(w["L"], seed=0)
    return (lambda im, ot, beta: torch.tensor(m.step(im, ot, LAMBDA_IMAGE, LAMBDA_TEXT, beta))), "port"


_THREADS = {}


def cpu_model() -> str:
    """CPU model name of the box (the CPU baseline states core count and CPU model)."""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    import platform
    return platform.processor() or "unknown CPU"


def pick_threads(wl: str, step, data) -> tuple:
    """"All the host threads it can use": the box may expose 128 logical CPUs behind a much smaller cgroup quota, where
    128 intra-op threads are 100x SLOWER than 8.  Time the REAL step at the REAL sample batch (1 warm-up + 2 timed steps
    per candidate thread count) and keep the fastest; both arms (cpu_baseline of the GPU line, --impl reference) use this
    same procedure, so they agree up to run-to-run noise."""
    import torch
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if wl in _THREADS:
        return _THREADS[wl], avail
    cands = sorted({c for c in (4, 8, 16, 32, 64, avail) if c <= avail}) or [1]
    best, cores = None, cands[0]
    for c in cands:
        torch.set_num_threads(c)
        step(*data[0], 0.5)
        t0 = time.perf_counter()
        step(*data[1 % len(data)], 0.5); step(*data[0], 0.5)
        dt = (time.perf_counter() - t0) / 2
        if best is None or dt < best:
            best, cores = dt, c
        if dt > 3 * best:
            break
    _THREADS[wl] = (cores, cands)
    return _THREADS[wl], avail


def cpu_baseline(wl: str, batch: int, budget_s: float, steps: int | None = None):
    """Time the reference's CPU implementation of the step (kind "reference") on the host cores, on a bounded sample."""
    import torch
    step, kind = _reference_stepper(wl, "cpu")
    if step is None:
        return None
    data = synth_host_batches(wl, 2, batch, seed=1)
    (cores, cands), avail = pick_threads(wl, step, data)
    torch.set_num_threads(cores)
    t0 = time.perf_counter(); step(*data[0], 0.5); warm = time.perf_counter() - t0
    n = steps if steps is not None else max(3, min(200, int(budget_s / max(warm, 1e-3))))
    t0 = time.perf_counter()
    for i in range(n):
        step(*data[i % 2], annealing(i))
    dt = time.perf_counter() - t0
    return {"value": batch * n / dt, "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": f"{n} steps of B={batch} of the same workload in {dt:.1f}s, torch {torch.__version__} CPU fp32, "
                      f"{cores} intra-op threads (fastest of {cands} on the real step; affinity mask {avail} CPUs; {cpu_model()})",
            "ms_per_step": 1e3 * dt / n, "batch": batch}


def gpu_eager_baseline(wl: str, batch: int, dev, steps: int = 10):
    """Informative: the reference's only existing GPU path -- the same unmodified modules under stock PyTorch eager
    (cuBLAS / cuDNN / ATen, fp32 matmul, default TF32 conv policy) on this B200 (`python train.py --cuda`)."""
    import torch
    step, kind = _reference_stepper(wl, "cuda")
    if step is None:
        return None
    data = [(im.to(dev), ot.to(dev)) for im, ot in synth_host_batches(wl, 2, batch, seed=1)]
    for i in range(3):
        step(*data[i % 2], 0.5)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        loss = step(*data[i % 2], annealing(i))
        float(loss)                      # the reference reads the loss every step (train_loss.data[0])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    return {"value": batch / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "kind": kind, "batch": batch,
            "what": f"unmodified reference modules + torch.optim.Adam under stock torch {torch.__version__} eager on this GPU, "
                    f"inputs resident, loss read every step, {steps} steps"}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    wl = args.workload
    w = WORKLOADS[wl]
    steps = args.steps
    r = cpu_baseline(wl, w["cpu_batch"], budget_s=60.0, steps=None if steps <= 0 else min(steps, 200))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref missing and no port for this workload"}), flush=True)
        return
    line = {"impl": "reference", "metric": "mvae_train_samples_per_sec", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_string(wl, args.global_batch or w["batch"]),
                       "cpu_sample_batch": r["batch"], "impl_note": "reference step body on the host CPU (torch CPU fp32)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def exchange_entry(m: dict) -> dict:
    """Cost of the data-parallel exchange at this N: step time minus the time of the same per-GPU batch on one GPU."""
    return {"mode": m["dp_mode"], "step_ms": m["ms_per_step"], "compute_only_ms": m["compute_only_ms"],
            "exchange_ms": m["ms_per_step"] - m["compute_only_ms"], "per_gpu_batch": m["b_local"],
            "how": "compute_only_ms = the same per-GPU batch stepped by a world-size-1 trainer (no exchange, plain fused Adam) "
                   "on rank 0's GPU, CUDA events; exchange_ms = step_ms - compute_only_ms (includes the rank skew the "
                   "exchange has to absorb)"}


def make_trainer(wl: str, b_local: int, dev, prec, world: int, rank: int, no_graph: bool):
    w = WORKLOADS[wl]
    kw = dict(device=dev, lr=w["lr"], lambda_image=LAMBDA_IMAGE, precision=prec, world_size=world, rank=rank, seed=0)
    if wl == "celeba19":
        import numpy as np
        from multimodal_vae_public_b200.trainer_celeba19 import CelebA19MVAETrainer
        np.random.seed(1234)             # the modality subsets of every step come from numpy's global RNG (as in the reference)
        return CelebA19MVAETrainer(w["L"], b_local, approx_m=1, lambda_attrs=LAMBDA_TEXT, **kw)
    if wl == "celeba":
        from multimodal_vae_public_b200.trainer_celeba import CelebAMVAETrainer
        return CelebAMVAETrainer(w["L"], b_local, lambda_attrs=LAMBDA_TEXT, use_graph=not no_graph, **kw)
    if wl == "fashion":
        from multimodal_vae_public_b200.trainer_fashion import FashionMVAETrainer as T
    else:
        from multimodal_vae_public_b200.trainer import MnistMVAETrainer as T
    return T(w["L"], b_local, lambda_text=LAMBDA_TEXT, use_graph=not no_graph, **kw)


def bench_workload(args, wl: str, steps: int, dev, prec, rank: int, local_rank: int, world: int, strong: bool,
                   with_rooflines: bool, micro: bool = True):
    """Timed regions of one workload on this process group; returns the measurements (every rank) -- rank 0 prints."""
    import torch
    import torch.distributed as dist
    from multimodal_vae_public_b200 import _lib
    w = WORKLOADS[wl]
    batch = args.global_batch if (args.global_batch and wl == args.workload) else w["batch"]
    b_local = batch // world if strong else batch
    b_global = b_local * world
    celeba, c19 = wl in ("celeba", "celeba19"), wl == "celeba19"
    tr = make_trainer(wl, b_local, dev, prec, world, rank, args.no_graph)
    npool = POOL if not celeba else 4
    host = [(im.pin_memory(), ot.pin_memory()) for im, ot in synth_host_batches(wl, npool, b_local, seed=100 + rank)]
    pool = [(im.to(dev), ot.to(dev)) for im, ot in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(tr._stream)
        for i in range(n):
            fn(i)
        e.record(tr._stream)
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def step_resident(i):
        im, ot = pool[i % npool]
        tr.step(im, ot, annealing_factor=annealing(i), sync=False)

    def log(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {wl}: {msg}", file=sys.stderr, flush=True)

    e2e_losses = []

    def step_e2e(i):
        # public host-fed API: pinned host batch -> (copy stream) -> device, step, loss read back on the host every
        # call (one step lagged, as in any asynchronous training loop); flush() at the end of the timed region
        im, ot = host[i % npool]
        v = tr.step_pipelined(im, ot, annealing_factor=annealing(i))
        if v is not None:
            e2e_losses.append(v)

    # warm-up (also captures the CUDA graph), then a FIXED number of extra untimed steps for the clocks to ramp.  (Never
    # "as many as fit in a second": every step of a multi-GPU run contains an exchange, and ranks that run different
    # numbers of them fall out of lockstep.)
    log("warm-up / graph capture")
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    tr.synchronize()
    for i in range(w["ramp"]):
        step_resident(i)
    tr.synchronize()

    log("timed region (device-resident inputs)")
    sampler = ClockSampler(local_rank); sampler.start()
    n0 = _lib.launch_count()
    ms = timed(step_resident, steps)
    launches = tr.launches_per_step * steps if tr.use_graph else _lib.launch_count() - n0
    value = b_global * steps / (ms * 1e-3)

    log("timed region (end to end from pinned host memory)")
    for i in range(3):
        step_e2e(i)
    tr.flush()

    def e2e_region(i):
        step_e2e(i)
        if i == steps - 1:
            e2e_losses.append(tr.flush())
    ms_e2e = timed(e2e_region, steps)
    clocks = sampler.stop()
    e2e = b_global * steps / (ms_e2e * 1e-3)
    loss = float(tr.loss19.item()) if c19 else float(tr.loss_host[0])
    out = {"workload": wl, "b_local": b_local, "b_global": b_global, "value": value, "ms_per_step": ms / steps,
           "e2e_value": e2e, "e2e_ms_per_step": ms_e2e / steps, "launches": int(launches), "loss_last": loss,
           "losses_read": len(e2e_losses), "clocks": clocks, "npool": npool, "trainer": tr, "steps": steps,
           "dp_mode": tr.dp_mode, "cuda_graph": bool(tr.use_graph), "chain": bool(getattr(tr, "chain", False)),
           "h2d_bytes_per_step": (b_local * (12288 + 18) * 4 + 4) if celeba else (b_local * (784 * 4 + 8) + 4)}
    if world > 1 and not c19:
        # What the exchange costs: the same per-GPU batch stepped by a world-size-1 trainer on this GPU (no exchange, plain
        # fused Adam), timed locally -- no collective in this block, every rank does the same, any failure only drops
        # the entry.  exchange_ms = step time at N GPUs - this.
        try:
            tr1 = make_trainer(wl, b_local, dev, prec, 1, 0, args.no_graph)
            for i in range(max(args.warmup, 3) + 5):
                tr1.step(pool[i % npool][0], pool[i % npool][1], annealing_factor=annealing(i), sync=False)
            tr1.synchronize()
            n1 = max(10, min(steps, 100))
            s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s1.record(tr1._stream)
            for i in range(n1):
                tr1.step(pool[i % npool][0], pool[i % npool][1], annealing_factor=annealing(i), sync=False)
            e1.record(tr1._stream)
            tr1.synchronize()
            out["compute_only_ms"] = s1.elapsed_time(e1) / n1
            del tr1
        except Exception as ex:  # noqa: BLE001  (informative entry only)
            log(f"compute-only comparison skipped: {ex!r}")
    if with_rooflines:
        log("per-kernel roofline pass")
        out["roof"] = measure_rooflines(tr, dev, prec, wl, world, micro=micro)
    log("done")
    return out


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    from multimodal_vae_public_b200 import ops

    if os.environ.get("MVAE_DIST_BACKEND", "nccl") != "nccl":
        local_rank = local_rank % max(torch.cuda.device_count(), 1)   # functional test: ranks may share a GPU
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        import datetime
        backend = os.environ.get("MVAE_DIST_BACKEND", "nccl")   # "gloo" only for single-GPU functional tests
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
        else:
            dist.init_process_group(backend, timeout=datetime.timedelta(seconds=90))
    strong = args.scaling == "strong"
    prec = ops.PREC_3XTF32 if args.precision == "3xtf32" else ops.PREC_TF32
    wl = args.workload
    main = bench_workload(args, wl, args.steps, dev, prec, rank, local_rank, world, strong, with_rooflines=True)
    tr = main.pop("trainer")
    extras = {}
    if wl == "mnist" and not args.no_extras:
        # configs[2] (FashionMNIST, global batch 4096): the configuration BASELINE.json quotes the >= 6x strong-scaling
        # target on -- measured in the same invocation at every N
        ex = bench_workload(args, "fashion", max(10, args.steps // 2), dev, prec, rank, local_rank, world, strong,
                            with_rooflines=True, micro=False)
        ex.pop("trainer")
        extras["fashion"] = ex

    line = None
    if rank == 0:
        peaks = load_peaks()
        tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
        roof = main["roof"]
        gemm = roof["gemm"]
        ach = gemm["algorithmic_flops_per_step"] / (gemm["ms_per_step"] * 1e-3) / 1e12
        b_local, b_global, npool = main["b_local"], main["b_global"], main["npool"]
        celeba = wl in ("celeba", "celeba19")
        line = {
            "metric": "mvae_train_samples_per_sec", "value": main["value"], "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "fp32 (3xTF32 tensor-core split products, fp32 accumulate)" if prec == ops.PREC_3XTF32
                     else "tf32 (fp32 storage/accumulate)",
            "data": "synthetic",
            "config": {"workload": workload_string(wl, b_global),
                       "parallelism": f"dp{world}", "global_batch": b_global, "per_gpu_batch": b_local,
                       "exchange": {"none": "single GPU", "nccl": "one ncclAllReduce of the flat gradient bucket per step",
                                    "p2p": "one fused peer-memory kernel per rank per step (NVLink reduce-scatter + Adam + "
                                           "all-gather, inside the step's CUDA graph)"}[main["dp_mode"]],
                       "l2": (f"rotating pool of {npool} distinct input batches and a multi-GB per-step working set, larger "
                              "than the 126 MB L2" if wl != "mnist" else
                              f"rotating pool of {POOL} distinct input batches ({POOL * b_local * 3144 / 1e6:.0f} MB) and a "
                              f"~{working_set_mb(b_local):.0f} MB per-step working set, both larger than the 126 MB L2"),
                       "cuda_graph": main["cuda_graph"], "loss_last": main["loss_last"],
                       **({"gemm_chain": main["chain"]} if wl == "mnist" else {})},
            "clocks": main["clocks"],
            "e2e": {"value": main["e2e_value"], "unit": "samples/s", "ms_per_step": main["e2e_ms_per_step"],
                    "h2d_bytes_per_step": main["h2d_bytes_per_step"], "d2h_bytes_per_step": 4 if wl == "celeba19" else 16,
                    "losses_read": main["losses_read"],
                    "note": ("trainer.step_pipelined(pinned host image, pinned host labels/attrs): H2D of batch i+1 on a copy "
                             "stream overlaps step i; the loss is copied D2H and read on the host every step (one step lagged)")},
            "gpu_launches": main["launches"],
            "roofline": {"bound": "tensor", "kernel": "gemm_kernel (tcgen05 kind::tf32; all GEMM launches of a step)",
                         "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                         "traffic": GEMM_CHAIN_DRAM_BYTES_PER_LAUNCH if (wl == "mnist" and main["chain"] and
                                                                         b_local == BATCH) else None,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum averaged over the "
                                           "4 chained GEMM launches of one B=4096 step (profiles/r02_gemm_chain_raw_key_metrics.txt)",
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained/2 (tf32 rate = half bf16) [{peaks['source']}]",
                         "algorithmic_flops_per_launch_avg": gemm["algorithmic_flops_per_step"] / max(gemm["launches"], 1),
                         "launches_per_step": gemm["launches"], "avg_launch_ms": gemm["ms_per_step"] / max(gemm["launches"], 1),
                         "executed_tensor_flops_factor": 3 if prec == ops.PREC_3XTF32 else 1,
                         "share_of_step": gemm["ms_per_step"] / main["ms_per_step"],
                         "share_note": "GEMM launch times come from a separate eager pass with CUDA events around every launch "
                                       "(cold start per launch); inside the replayed graph the same launches run back to back, so "
                                       "the share can exceed 1 by a few per cent when the step is almost all GEMM"},
            "roofline_hbm": roof["hbm"],
            "kernel_breakdown_ms": roof["breakdown"],
        }
        if "compute_only_ms" in main:
            line["exchange"] = exchange_entry(main)
        for k in ("hbm_poe_fwd", "hbm_poe_bwd"):
            if k in roof:
                line["roofline_" + k] = roof[k]
        if extras:
            line["extra"] = {}
            for name, ex in extras.items():
                g2 = ex["roof"]["gemm"]
                a2 = g2["algorithmic_flops_per_step"] / (g2["ms_per_step"] * 1e-3) / 1e12
                line["extra"][name] = {
                    "workload": workload_string(name, ex["b_global"]), "value": ex["value"], "unit": "samples/s",
                    "ms_per_step": ex["ms_per_step"], "steps": ex["steps"], "per_gpu_batch": ex["b_local"],
                    "e2e": {"value": ex["e2e_value"], "ms_per_step": ex["e2e_ms_per_step"],
                            "h2d_bytes_per_step": ex["h2d_bytes_per_step"], "d2h_bytes_per_step": 16},
                    "gpu_launches": ex["launches"], "loss_last": ex["loss_last"], "clocks": ex["clocks"],
                    "roofline": {"bound": "tensor", "achieved": a2, "peak": tf32_peak, "unit": "TFLOP/s", "frac": a2 / tf32_peak,
                                 "share_of_step": g2["ms_per_step"] / ex["ms_per_step"], "launches_per_step": g2["launches"]},
                    "kernel_breakdown_ms": ex["roof"]["breakdown"]}
                if "compute_only_ms" in ex:
                    line["extra"][name]["exchange"] = exchange_entry(ex)
        if world == 1 and not args.no_cpu_baseline:
            w = WORKLOADS[wl]
            cpu = cpu_baseline(wl, w["cpu_batch"], budget_s=12.0)
            if cpu is not None:
                line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            others = {}
            if wl == "mnist":
                c0 = cpu_baseline("mnist", 64, budget_s=4.0)      # configs[0]: the reference's own CPU-runnable case
                if c0 is not None:
                    others["mnist_b64_configs0"] = {k: c0[k] for k in ("value", "unit", "cores", "kind", "sample")}
                for name in extras:
                    cx = cpu_baseline(name, WORKLOADS[name]["cpu_batch"], budget_s=10.0)
                    if cx is not None:
                        others[name] = {k: cx[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if others:
                line["cpu_baseline_other"] = others
            try:
                eg = gpu_eager_baseline(wl, b_global if wl != "celeba19" else 64, dev)
                if eg is not None:
                    line["gpu_eager_baseline"] = eg
                for name in extras:
                    e2 = gpu_eager_baseline(name, extras[name]["b_global"], dev, steps=5)
                    if e2 is not None:
                        line.setdefault("gpu_eager_baseline_other", {})[name] = e2
            except Exception as exc:  # noqa: BLE001  (informative leg: never lose the line over it)
                line["gpu_eager_baseline"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def working_set_mb(b: int) -> float:
    rows512 = 4 * b + 3 * b + 12 * 2 * b + 4 * 2 * b + 4 * b      # 512-wide activation / gradient buffers
    return (rows512 * 512 + 2 * b * 784 + b * 784) * 4 / 1e6 + 4 * 10.4


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        d["source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measure_rooflines(tr, dev, prec, wl, world, micro=True):
    """Eager (non-graph) passes with CUDA events around every library call on the launching stream."""
    import torch
    from multimodal_vae_public_b200 import ops
    st = tr._stream
    names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and
             n not in ("gemm_desc", "chain_workspace", "GemmDesc", "Optional", "Sequence", "allreduce_adam_p2p")]
    records = []
    orig = {n: getattr(ops, n) for n in names}

    def wrap(n):
        def f(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(st); r = orig[n](*a, **k); e.record(st)
            records.append((n, s, e))
            return r
        return f
    reps = 5
    try:
        for n in names:
            if isinstance(orig[n], type(wrap)):
                setattr(ops, n, wrap(n))
        with torch.cuda.stream(st):
            for _ in range(reps):
                # forward + backward (+ the local flat Adam at N=1); the exchange step of a multi-GPU run is timed by
                # the step itself, not here (every rank would have to enter it in lockstep)
                if world == 1:
                    tr._enqueue_step(True, False, True)
                elif hasattr(tr, "_last_plan"):
                    tr._enqueue(tr._last_plan, True, False, False, 1.0, False)
                else:
                    tr._enqueue_fwd_bwd(True, False)
        st.synchronize()
    finally:
        for n in names:
            setattr(ops, n, orig[n])
    per = {}
    cnt = {}
    for n, s, e in records:
        per[n] = per.get(n, 0.0) + s.elapsed_time(e) / reps
        cnt[n] = cnt.get(n, 0) + 1
    gemm_names = ("gemm_batch", "gemm_chain", "linear_fwd")
    gemm_ms = sum(per.get(n, 0.0) for n in gemm_names)
    gemm_launches = sum(cnt.get(n, 0) for n in gemm_names) // reps
    flops = {"mnist": lambda L: executed_gemm_flops_per_sample(L, tr.label_table),
             "fashion": lambda L: fashion_gemm_flops_per_sample(L, tr.label_table),
             "celeba": celeba_gemm_flops_per_sample,
             "celeba19": lambda L: tr.gemm_flops_last_step() / tr.B}[wl](tr.L) * tr.B
    out = {"gemm": {"ms_per_step": gemm_ms, "launches": gemm_launches, "algorithmic_flops_per_step": flops},
           "breakdown": {k: round(v, 5) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}}
    if not micro:
        return out
    peaks = load_peaks()
    # ---- HBM roofline of the fused reconstruction-loss kernel at roofline size (inputs >> L2)
    R, D = 65536, 784
    x = torch.randn(R, D, device=dev); t = torch.rand(R // 2, D, device=dev); dx = torch.empty_like(x)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)

    def time_launches(fn, iters=10):
        with torch.cuda.stream(st):
            for _ in range(3):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(st)
            for _ in range(iters):
                fn()
            e.record(st)
        st.synchronize()
        return s.elapsed_time(e) / iters
    ms = time_launches(lambda: ops.bce_logits_fwd_bwd(x, t, dx, 1e-3, acc, seg_rows=R // 2))
    alg = (2 * R * D + (R // 2) * D) * 4.0   # read x once, write dx once, read the shared target once = 10 B per logit
    out["hbm"] = {"bound": "hbm", "kernel": "bce_stacked_kernel (fused BCE-with-logits loss + gradient, two passes sharing "
                                            f"one target), roofline-size run R={R} D={D}",
                  "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                  "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": BCE_DRAM_BYTES_PER_LAUNCH,
                  "traffic_source": "ncu --set full of this launch (profiles/r02_bce_raw_key_metrics.txt): below the "
                                    "algorithmic bytes because part of the gradient write-back is still in L2 when the kernel ends",
                  "algorithmic_bytes_per_launch": alg, "avg_launch_ms": ms, "peak_source": peaks["source"]}
    del x, t, dx
    # ---- the fused PoE + reparametrise + KL kernels at roofline size (MNIST pass structure, L = 64, B = 262,144)
    Bp, L = 262144, 64
    enc = [torch.randn(Bp, 2 * L, device=dev) * 0.5 for _ in range(2)]
    mu_e = [e[:, :L] for e in enc]; lv_e = [e[:, L:] for e in enc]
    z = torch.empty(3 * Bp, L, device=dev); nz = torch.empty(3 * Bp, L, device=dev); dz = torch.randn(3 * Bp, L, device=dev)
    d_enc = [torch.empty(Bp, 2 * L, device=dev) for _ in range(2)]
    kl = torch.zeros(3, dtype=torch.float64, device=dev)
    stepc = torch.zeros(1, dtype=torch.int32, device=dev)
    masks = (0b01, 0b11, 0b10)
    ms_f = time_launches(lambda: ops.poe_fwd(mu_e, lv_e, masks, Bp, L, z, variant=0, training=True, noise=None, noise_out=nz,
                                             seed=1, step_dev=stepc, kl_acc=kl))
    ms_b = time_launches(lambda: ops.poe_bwd(mu_e, lv_e, masks, Bp, L, dz, [d[:, :L] for d in d_enc],
                                             [d[:, L:] for d in d_enc], kl_scale=1.0 / Bp, variant=0, training=True, noise=nz))
    # bytes the fused launches really move per sample: fwd reads 2 experts x (mu, logvar), writes z and the noise of 3
    # passes; bwd reads the experts, dz and the noise, writes 2 experts x (dmu, dlogvar).  SURVEY 8d's algorithmic
    # figure counts every pass separately (4,352 / 7,168 B per sample) -- both fractions are reported.
    moved_f, moved_b = (4 * L + 6 * L) * 4, (4 * L + 6 * L + 4 * L) * 4
    for key, ms_k, moved, alg_b, name, dram in (
            ("hbm_poe_fwd", ms_f, moved_f, POE_FWD_BYTES_PER_SAMPLE, "poe_fwd_fast_kernel", POE_FWD_DRAM_BYTES_PER_LAUNCH),
            ("hbm_poe_bwd", ms_b, moved_b, POE_BWD_BYTES_PER_SAMPLE, "poe_bwd_fast_kernel", POE_BWD_DRAM_BYTES_PER_LAUNCH)):
        out[key] = {"bound": "hbm", "kernel": f"{name} (PoE + reparametrise + KL, 3 passes in one launch), roofline-size run "
                                              f"B={Bp} L={L}",
                    "achieved": moved * Bp / (ms_k * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": moved * Bp / (ms_k * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "achieved_algorithmic": alg_b * Bp / (ms_k * 1e-3) / 1e9,
                    "frac_algorithmic": alg_b * Bp / (ms_k * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "bytes_moved_per_launch": moved * Bp, "algorithmic_bytes_per_launch": alg_b * Bp,
                    "avg_launch_ms": ms_k, "traffic": dram,
                    "traffic_source": "ncu --set full of this launch (profiles/r02_poe_raw_key_metrics.txt)",
                    "peak_source": peaks["source"]}
    del enc, z, nz, dz, d_enc
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--precision", choices=["3xtf32", "tf32"], default="3xtf32")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="override the global batch (experiments only; the default is the BASELINE.json configuration)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the FashionMNIST (configs[2]) run that rides along")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--workload", choices=list(WORKLOADS), default="mnist",
                    help="mnist = BASELINE.json configs[1] (default, the headline); fashion = conv flavour (configs[2]); "
                         "celeba = conv+BatchNorm flavour, global batch 1024 (configs[3]); celeba19 = 19 experts, "
                         "approx_m=1, global batch 512 (configs[4])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            import subprocess
            port = 29500 + os.getpid() % 1000
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
