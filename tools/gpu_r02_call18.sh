#!/bin/bash
# A/B session 18: bias-gradient column sums on the side stream (beside the GEMM chains) -- tests + same-box A/B
mkdir -p gpurun_out
O=gpurun_out/r2c18
( time timeout 900 python -m pytest tests/test_mnist_step_gpu.py tests/test_fashion_step_gpu.py tests/test_dp_gpu.py tests/test_zz_optimizer_state_gpu.py -m gpu -q -x --timeout 600 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
for rep in 1 2; do
for tag in "side:" "serial:MVAE_OVERLAP=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras > ${O}_mnist_${name}${rep}.json 2> ${O}_mnist_${name}${rep}.err
  env $envs timeout 300 python bench.py --workload fashion --steps 30 --warmup 5 --no-cpu-baseline > ${O}_fashion_${name}${rep}.json 2> ${O}_fashion_${name}${rep}.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c18_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c18_")[1], {k:d.get(k) for k in ("value","ms_per_step")}, round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
