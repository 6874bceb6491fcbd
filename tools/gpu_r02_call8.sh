#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c8
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
timeout 300 python bench.py --workload fashion --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_fashion.json 2> ${O}_bench_fashion.err
timeout 300 python bench.py --workload celeba --steps 10 --warmup 5 --no-cpu-baseline > ${O}_bench_celeba.json 2> ${O}_bench_celeba.err
MVAE_IMPLICIT_CONV=0 timeout 300 python bench.py --workload celeba --steps 10 --warmup 5 --no-cpu-baseline > ${O}_bench_celeba_noimp.json 2> ${O}_bench_celeba_noimp.err
timeout 300 python bench.py --workload celeba19 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_bench_celeba19.json 2> ${O}_bench_celeba19.err
export MVAE_TIMES_MIN_MS=0.02
timeout 200 python tools/gemm_times.py celeba 1024 > ${O}_times_celeba.txt 2>&1
python - <<'PY'
import json
for f in ("fashion","celeba","celeba_noimp","celeba19"):
    try:
        d=json.loads(open(f"gpurun_out/r2c8_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -5 ${O}_bench_celeba.err; tail -45 ${O}_times_celeba.txt
