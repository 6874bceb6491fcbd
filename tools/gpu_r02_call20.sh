#!/bin/bash
# A/B session 20: programmatic dependent launch (PDL) for the kernels of the MNIST-shape step -- full GPU suite, then same-box
# A/B (MVAE_PDL=0 turns the launch attribute off)
mkdir -p gpurun_out
O=gpurun_out/r2c20
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
for rep in 1 2; do
for tag in "pdl:MVAE_PDL=1" "nopdl:MVAE_PDL=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras > ${O}_mnist_${name}${rep}.json 2> ${O}_mnist_${name}${rep}.err
  env $envs timeout 300 python bench.py --global-batch 512 --steps 200 --warmup 5 --no-cpu-baseline --no-extras > ${O}_m512_${name}${rep}.json 2> ${O}_m512_${name}${rep}.err
done; done
for tag in "pdl:MVAE_PDL=1" "nopdl:MVAE_PDL=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --workload fashion --steps 30 --warmup 5 --no-cpu-baseline > ${O}_fashion_${name}.json 2> ${O}_fashion_${name}.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c20_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c20_")[1], {k:d.get(k) for k in ("value","ms_per_step")}, round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["config"]["loss_last"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 ${O}_mnist_pdl1.err
