#!/bin/bash
# full GPU suite, default bench + other workloads, ncu launch list + full captures of the two roofline kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 240 ) > gpurun_out/ev_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ev_pytest.log; tail -4 gpurun_out/ev_pytest.log
timeout 300 python bench.py > gpurun_out/ev_bench_mnist.json 2> gpurun_out/ev_bench_mnist.err; tail -2 gpurun_out/ev_bench_mnist.err
timeout 300 python bench.py --workload fashion --steps 30 --no-cpu-baseline > gpurun_out/ev_bench_fashion.json 2> gpurun_out/ev_bench_fashion.err
timeout 300 python bench.py --workload celeba --steps 20 --no-cpu-baseline > gpurun_out/ev_bench_celeba.json 2> gpurun_out/ev_bench_celeba.err
timeout 300 python bench.py --workload celeba19 --steps 10 --no-cpu-baseline > gpurun_out/ev_bench_celeba19.json 2> gpurun_out/ev_bench_celeba19.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/ev_bench_reference.json 2> gpurun_out/ev_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ev_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 8 -c 4 -o gpurun_out/ev_gemm_chain -f python tools/profile_step.py step --steps 4 > gpurun_out/ev_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bce -s 2 -c 1 -o gpurun_out/ev_bce -f python tools/profile_step.py bce 65536 784 4 > gpurun_out/ev_ncu_bce.log 2>&1
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/ev_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("kernel_breakdown_ms"))
except Exception as e: print("ERR", e)
PY
done
