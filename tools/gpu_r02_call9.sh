#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c9
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > ${O}_pytest.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest.log; tail -8 ${O}_pytest.log
# same-box A/Bs (MNIST headline): pre-split weights, label table
for tag in "base:" "nopresplit:MVAE_PRESPLIT=0" "nolt:MVAE_LABEL_TABLE=0" "nolt_nopre:MVAE_LABEL_TABLE=0 MVAE_PRESPLIT=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras > ${O}_mnist_${name}.json 2> ${O}_mnist_${name}.err
done
for tag in "base:" "subpixel:MVAE_SUBPIXEL=1" "nopresplit:MVAE_PRESPLIT=0"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout 300 python bench.py --workload fashion --steps 30 --warmup 5 --no-cpu-baseline > ${O}_fashion_${name}.json 2> ${O}_fashion_${name}.err
done
MVAE_SUBPIXEL=1 timeout 300 python bench.py --workload fashion --global-batch 512 --steps 50 --warmup 5 --no-cpu-baseline > ${O}_f512_subpixel.json 2> ${O}_f512_subpixel.err
timeout 300 python bench.py --workload fashion --global-batch 512 --steps 50 --warmup 5 --no-cpu-baseline > ${O}_f512_base.json 2> ${O}_f512_base.err
export MVAE_TIMES_MIN_MS=0.003
MVAE_SUBPIXEL=1 timeout 200 python tools/gemm_times.py fashion 4096 > ${O}_times_fashion_subpixel.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c9_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2c9_")[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -45 ${O}_times_fashion_subpixel.txt
