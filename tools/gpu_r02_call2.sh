#!/bin/bash
# round-2 call 2: full GPU suite with the new parity cases + the restructured bench (MNIST headline + Fashion extra,
# reference-arm, eager-GPU baseline)
mkdir -p gpurun_out
O=gpurun_out/r2c2
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -6 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --verbose > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -3 ${O}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --workload celeba --steps 10 --no-cpu-baseline > ${O}_bench_celeba.json 2> ${O}_bench_celeba.err; echo "celeba rc=$?"; tail -2 ${O}_bench_celeba.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c2_bench.json","gpurun_out/r2c2_bench_ref.json","gpurun_out/r2c2_bench_celeba.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d.get("e2e",{}).get("value"), d.get("cpu_baseline"), d.get("gpu_eager_baseline"))
        for k in ("roofline_hbm_poe_fwd","roofline_hbm_poe_bwd"):
            if k in d: print(k, d[k]["frac"], d[k]["frac_algorithmic"], d[k]["avg_launch_ms"])
        if "extra" in d: print("extra", {k:(v["value"], v["ms_per_step"]) for k,v in d["extra"].items()})
        print(d.get("cpu_baseline_other")); print(d.get("gpu_eager_baseline_other")); print(d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
