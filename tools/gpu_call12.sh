#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 240 ) > gpurun_out/c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c14_pytest.log; tail -6 gpurun_out/c14_pytest.log
for w in fashion celeba celeba19; do
timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/c14_bench_$w.json 2> gpurun_out/c14_bench_$w.err
done
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/c14_bench_mnist.json 2> gpurun_out/c14_bench_mnist.err
for f in gpurun_out/c14_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("kernel_breakdown_ms"))
except Exception as e: print("ERR", e)
PY
done
