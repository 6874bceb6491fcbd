#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c14
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
for tag in "overlap:" "nooverlap:MVAE_OVERLAP=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras > ${O}_mnist_${name}.json 2> ${O}_mnist_${name}.err
  env $envs timeout 300 python bench.py --workload fashion --steps 30 --warmup 5 --no-cpu-baseline > ${O}_fashion_${name}.json 2> ${O}_fashion_${name}.err
done
timeout 300 python bench.py --workload celeba --steps 10 --warmup 5 --no-cpu-baseline > ${O}_celeba.json 2> ${O}_celeba.err
timeout 300 python bench.py --workload celeba19 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_celeba19.json 2> ${O}_celeba19.err
MVAE_SUBPIXEL=0 timeout 300 python bench.py --workload celeba19 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_celeba19_nosub.json 2> ${O}_celeba19_nosub.err
timeout 300 python bench.py --global-batch 512 --steps 100 --no-cpu-baseline --no-extras > ${O}_mnist_b512.json 2> ${O}_mnist_b512.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c14_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c14_")[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 ${O}_celeba19.err
