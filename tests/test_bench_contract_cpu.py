"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) times the reference step (its own
modules from oracle/_ref, or the oracle port when that copy is absent) on the host cores and prints ONE JSON line with the agreed keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mvae_train_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import ref_harness
    # the reference's own modules when oracle/_ref (byte-identical copy made by oracle/make_ref.py) is there, else the port
    assert cb["kind"] == ("reference" if ref_harness.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_exchange_entry_shape():
    """bench.py's per-N exchange cost entry (N > 1): step time minus the same per-GPU batch on a world-size-1 trainer."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    e = bench.exchange_entry({"dp_mode": "p2p", "ms_per_step": 0.42, "compute_only_ms": 0.32, "b_local": 512})
    assert e["mode"] == "p2p" and e["per_gpu_batch"] == 512
    assert abs(e["exchange_ms"] - 0.10) < 1e-9 and e["step_ms"] == 0.42 and e["compute_only_ms"] == 0.32
    assert "Intel" in bench.cpu_model() or len(bench.cpu_model()) > 0
