#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c13
( time timeout 900 python -m pytest tests/test_celeba_step_gpu.py tests/test_celeba19_step_gpu.py tests/test_implicit_conv_gpu.py -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
for tag in "base:MVAE_SUBPIXEL=0" "subpixel:MVAE_SUBPIXEL=1"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --workload celeba --steps 10 --warmup 5 --no-cpu-baseline > ${O}_celeba_${name}.json 2> ${O}_celeba_${name}.err
done
export MVAE_TIMES_MIN_MS=0.05
timeout 200 python tools/gemm_times.py celeba 1024 > ${O}_times_celeba_subpixel.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c13_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c13_")[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 ${O}_celeba_subpixel.err; grep "gemm_chain" ${O}_times_celeba_subpixel.txt | cut -c1-260
