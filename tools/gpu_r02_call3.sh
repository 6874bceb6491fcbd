#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c3
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -3 ${O}_bench.err
MVAE_LABEL_TABLE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > ${O}_bench_nolt.json 2> ${O}_bench_nolt.err
timeout 200 python bench.py --global-batch 512 --steps 100 --no-cpu-baseline > ${O}_bench_b512.json 2> ${O}_bench_b512.err
export MVAE_TIMES_MIN_MS=0.003
timeout 200 python tools/gemm_times.py mnist 4096 > ${O}_times_mnist_4096.txt 2>&1
timeout 200 python tools/gemm_times.py mnist 512 > ${O}_times_mnist_512.txt 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c3_bench.json","gpurun_out/r2c3_bench_nolt.json","gpurun_out/r2c3_bench_b512.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d.get("e2e",{}).get("value"), d["roofline"]["frac"])
        print("  ", d["kernel_breakdown_ms"])
        for k in ("roofline_hbm_poe_fwd","roofline_hbm_poe_bwd"):
            if k in d: print("  ", k, d[k]["frac"], d[k]["frac_algorithmic"], d[k]["avg_launch_ms"])
        if "extra" in d: print("  extra", {k:(v["value"], v["ms_per_step"], v["kernel_breakdown_ms"]) for k,v in d["extra"].items()})
    except Exception as e: print(f, "ERR", e)
PY
cat ${O}_times_mnist_4096.txt | tail -25
