#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c7
( timeout 900 python -m pytest tests/test_implicit_conv_gpu.py tests/test_conv_small_gpu.py tests/test_fashion_step_gpu.py tests/test_gemm_chain_gpu.py tests/test_mnist_step_gpu.py -m gpu -q --timeout 600 ) > ${O}_pytest.log 2>&1; tail -8 ${O}_pytest.log
timeout 300 python tools/profile_conv_small.py 4096 > ${O}_conv_small_times.txt 2>&1; cat ${O}_conv_small_times.txt
timeout 300 python bench.py --workload fashion --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_fashion.json 2> ${O}_bench_fashion.err
MVAE_IMPLICIT_CONV=0 timeout 300 python bench.py --workload fashion --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_fashion_noimp.json 2> ${O}_bench_fashion_noimp.err
timeout 300 python bench.py --workload fashion --global-batch 512 --steps 50 --warmup 5 --no-cpu-baseline > ${O}_bench_f512.json 2> ${O}_bench_f512.err
export MVAE_TIMES_MIN_MS=0.003
timeout 200 python tools/gemm_times.py fashion 4096 > ${O}_times_fashion_4096.txt 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c7_bench_fashion.json","gpurun_out/r2c7_bench_fashion_noimp.json","gpurun_out/r2c7_bench_f512.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
cat ${O}_times_fashion_4096.txt | tail -42
