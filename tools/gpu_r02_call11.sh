#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c11
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
MVAE_PRESPLIT=1 timeout 600 python -m pytest tests/test_fashion_step_gpu.py tests/test_mnist_step_gpu.py -m gpu -q --timeout 500 > ${O}_pytest_presplit.log 2>&1; tail -3 ${O}_pytest_presplit.log
MVAE_SUBPIXEL=1 timeout 600 python -m pytest tests/test_fashion_step_gpu.py -m gpu -q --timeout 500 > ${O}_pytest_subpixel.log 2>&1; tail -3 ${O}_pytest_subpixel.log
for tag in "base:" "subpixel:MVAE_SUBPIXEL=1"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --workload fashion --steps 30 --warmup 5 --no-cpu-baseline > ${O}_fashion_${name}.json 2> ${O}_fashion_${name}.err
done
export MVAE_TIMES_MIN_MS=0.05
MVAE_SUBPIXEL=1 timeout 200 python tools/gemm_times.py fashion 4096 > ${O}_times_fashion_subpixel.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c11_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c11_")[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
grep "gemm_chain\|gemm_batch" ${O}_times_fashion_subpixel.txt | cut -c1-200
